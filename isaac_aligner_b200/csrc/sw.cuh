// Smith-Waterman score set and CIGAR word helper shared by the kernels.  (The scalar one-alignment-per-thread restatement of
// BandedSmithWaterman::align that used to live here is kept, uncompiled, in tests/cuda/legacy_sw_kernels.cuh; the product
// runs the packed 16x2 form of sw2.cuh.)
//
// Arithmetic: the reference works in wrapping int16.  With gapExtend <= gapOpen and the constructor's overflow guard
// (BandedSmithWaterman.cpp:47-53) no intermediate leaves the int16 range, so plain int arithmetic is bit-identical;
// the host rejects any other score set (isaac_ext_create).
#pragma once
#include "device_types.cuh"

namespace isaac_b200
{

struct SwScores
{
    int match, mismatch, open, ext;   // open/ext positive (BandedSmithWaterman.hh:44-46)
    int init;                         // -32768 + open (BandedSmithWaterman.cpp:44)
};

__host__ __device__ __forceinline__ uint32_t cigarWord(uint32_t length, uint32_t op) { return (length << 4) | op; }

} // namespace isaac_b200
