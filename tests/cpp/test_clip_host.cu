// kernels_clip.cuh on the CPU: clipTemplateEndsOfCluster -- what every thread of clipTemplateEndsKernel runs: SemialignedEndsClipper
// and OverlappingEndsClipper on one template -- over a reference and a read set packed on the host in the layouts of
// device_types.cuh.  tests/test_clip_host.py compares it with the reference's own clippers (oracle_build_templates with the clip
// flags) and with the literals of the reference's testOverlappingEndsClipper.cpp.  Host code of an nvcc-compiled shared library,
// no CUDA call.  TEST CODE, not a product path.
#include <cstring>
#include <vector>

#include "../../isaac_aligner_b200/csrc/sw.cuh"
#include "../../isaac_aligner_b200/csrc/kernels_clip.cuh"

using namespace isaac_b200;

/// contigs: ASCII ACGTN, back to back at contigOffsetIn; reads: the tile's BCL bytes.  fragments are clipped in place, their
/// CIGARs land in cigarsOut (room for cigarWords + 4 per fragment).  \return 0, or 5 when a CIGAR does not fit the clippers
extern "C" int clip_templates_host(uint32_t contigCount, const char *bases, const uint64_t *contigBegin, const isaac_ext_reads_t *r,
                                   uint32_t clipFlags, const isaac_ext_template_t *templates, isaac_ext_fragment_t *fragments,
                                   const uint32_t *cigarsIn, uint32_t *cigarsOut)
{
    // ---- reference: every contig on a 128-base boundary, 2 bits per base + N mask (ReferenceView)
    std::vector<uint64_t> offset(contigCount), length(contigCount);
    uint64_t total = 0;
    for (uint32_t c = 0; c < contigCount; ++c) { offset[c] = total; length[c] = contigBegin[c + 1] - contigBegin[c]; total += (length[c] + 127) / 128 * 128; }
    total += 128;
    std::vector<uint32_t> refBases2(total / 16 + 1, 0), refNmask(total / 32 + 1, 0);
    for (uint32_t c = 0; c < contigCount; ++c)
        for (uint64_t i = 0; i < length[c]; ++i)
        {
            const char b = bases[contigBegin[c] + i];
            const uint64_t g = offset[c] + i;
            const unsigned code = b == 'A' ? 0u : b == 'C' ? 1u : b == 'G' ? 2u : b == 'T' ? 3u : 4u;
            if (code < 4) refBases2[g >> 4] |= code << ((g & 15) * 2); else refNmask[g >> 5] |= 1u << (g & 31);
        }
    ReferenceView ref{};
    ref.bases2 = refBases2.data(); ref.nmask = refNmask.data(); ref.contigOffset = offset.data(); ref.contigLength = length.data();
    ref.contigCount = contigCount; ref.totalBases = total;
    // ---- reads: forward strand, 2 bits per base + n mask + qualities (ReadSetView; Read::decodeBcl: BCL N = 'n', quality 2)
    const uint32_t rc = r->readCount, len[2] = {r->readLength[0], rc > 1 ? r->readLength[1] : 0};
    const uint32_t maxLen = len[0] > len[1] ? len[0] : len[1];
    const uint32_t wordsN = (maxLen + 31) / 32, words2 = wordsN * 2, qualityStride = wordsN * 32;
    const size_t readTotal = size_t(r->clusterCount) * rc;
    std::vector<uint32_t> bases2(readTotal * words2, 0), nmask(readTotal * wordsN, 0);
    std::vector<uint8_t> quality(readTotal * qualityStride, 0);
    for (size_t c = 0; c < r->clusterCount; ++c)
        for (uint32_t k = 0; k < rc; ++k)
        {
            const uint8_t *bcl = r->bcl + c * (len[0] + len[1]) + (k ? len[0] : 0);
            const size_t id = c * rc + k;
            for (uint32_t i = 0; i < len[k]; ++i)
            {
                const uint8_t b = bcl[i];
                if (b & 0xFC) { bases2[id * words2 + (i >> 4)] |= uint32_t(b & 3) << ((i & 15) * 2); quality[id * qualityStride + i] = b >> 2; }
                else { nmask[id * wordsN + (i >> 5)] |= 1u << (i & 31); quality[id * qualityStride + i] = 2; }
            }
        }
    ReadSetView reads{};
    reads.bases2 = bases2.data(); reads.nmask = nmask.data(); reads.quality = quality.data();
    reads.words2 = words2; reads.wordsN = wordsN; reads.qualityStride = qualityStride; reads.readCount = rc;
    reads.readLength[0] = len[0]; reads.readLength[1] = len[1]; reads.readTotal = uint32_t(readTotal);
    int rcode = 0;
    for (uint32_t c = 0; c < r->clusterCount; ++c)
        if (!clipTemplateEndsOfCluster(ref, reads, c, clipFlags, templates, fragments, cigarsIn, cigarsOut)) rcode = 5;
    return rcode;
}
