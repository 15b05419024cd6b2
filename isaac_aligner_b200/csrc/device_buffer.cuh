// Grow-only device allocation reused across calls (no per-call cudaMalloc on the hot path once warmed up).
#pragma once
#include <algorithm>
#include <cstddef>
#include <cuda_runtime.h>

namespace isaac_b200
{

template <class T> struct DeviceBuffer
{
    T *p = nullptr; size_t capacity = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= capacity) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; capacity = 0;
        const cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) capacity = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; capacity = 0; }
};

} // namespace isaac_b200
