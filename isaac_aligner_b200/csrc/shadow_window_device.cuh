// R1 of the rescue pass for host AND device code: the window of reference positions ShadowAligner::rescueShadow scans for the mate of
// an orphan (calculateShadowRescueRange, ShadowAligner.cpp:119-149; the clamps of rescueShadow, :170-198) from the mate model of the
// template length statistics (TemplateLengthStatistics::mateOrientation / mateMinPosition / mateMaxPosition,
// TemplateLengthStatistics.cpp:186-238) as a function a one-thread-per-request kernel calls (planWriteKernel behind plan_device.cuh,
// shadowWindowsKernel for requests that come from the caller), so that requests never visit the host.
// tests/test_template_worker.py checks it on the CPU against the reference's own calculateShadowRescueRange.
#pragma once
#include <cstdint>
#include "../../include/isaac_ext.h"

#ifndef ISAAC_HD
#ifdef __CUDACC__
#define ISAAC_HD __host__ __device__
#else
#define ISAAC_HD
#endif
#endif

namespace isaac_b200
{

struct ShadowWindowModel
{
    uint32_t mateMin, mateMax, models[2];
};

ISAAC_HD inline ShadowWindowModel makeShadowWindowModel(const isaac_ext_tls_t &t)
{
    ShadowWindowModel m;
    m.mateMin = -1 == t.mateDriftRange ? t.min : t.median - uint32_t(t.mateDriftRange);        // TemplateLengthStatistics.hh:205-214
    m.mateMax = -1 == t.mateDriftRange ? t.max : t.median + uint32_t(t.mateDriftRange);
    m.models[0] = t.bestModel[0]; m.models[1] = t.bestModel[1];
    return m;
}
ISAAC_HD inline unsigned shadowAlignmentClass(const unsigned model) { return model < 4 ? model : ((~model) & 3u); }
/// TemplateLengthStatistics::isCoherent (TemplateLengthStatistics.hh:137-151): rescueShadow gives up without it (ShadowAligner.cpp:164-168)
ISAAC_HD inline bool shadowModelCoherent(const ShadowWindowModel &m)
{
    return m.models[0] < 8 && m.models[1] < 8 && m.models[0] != m.models[1] && shadowAlignmentClass(m.models[0]) == shadowAlignmentClass(m.models[1]);
}
ISAAC_HD inline bool shadowValidModel(const ShadowWindowModel &m, const bool reverse, const unsigned readIndex)
{
    const unsigned shift = (readIndex + 1) % 2;
    return unsigned(reverse) == ((m.models[0] >> shift) & 1u) || unsigned(reverse) == ((m.models[1] >> shift) & 1u);
}
ISAAC_HD inline bool shadowFirstFragment(const ShadowWindowModel &m, const bool reverse, const unsigned readIndex)
{
    const unsigned shift = (readIndex + 1) % 2;
    for (unsigned i = 0; i < 2; ++i) if (unsigned(reverse) == ((m.models[i] >> shift) & 1u)) return ((m.models[i] >> 2) & 1u) == readIndex;
    return false;
}
ISAAC_HD inline bool shadowMateOrientation(const ShadowWindowModel &m, const unsigned readIndex, const bool reverse)
{
    const unsigned shift = (readIndex + 1) % 2;
    for (unsigned i = 0; i < 2; ++i) if (unsigned(reverse) == ((m.models[i] >> shift) & 1u)) return (m.models[i] >> readIndex) & 1u;
    return (m.models[0] >> readIndex) & 1u;
}
ISAAC_HD inline long shadowMateMinPosition(const ShadowWindowModel &m, const unsigned readIndex, const bool reverse, const long position, const uint32_t *len)
{
    if (!shadowValidModel(m, reverse, readIndex)) return position;
    return shadowFirstFragment(m, reverse, readIndex) ? position + long(m.mateMin) - long(len[(readIndex + 1) % 2]) : position - long(m.mateMax) + long(len[readIndex]);
}
ISAAC_HD inline long shadowMateMaxPosition(const ShadowWindowModel &m, const unsigned readIndex, const bool reverse, const long position, const uint32_t *len)
{
    if (!shadowValidModel(m, reverse, readIndex)) return position;
    return shadowFirstFragment(m, reverse, readIndex) ? position + long(m.mateMax) - long(len[(readIndex + 1) % 2]) : position - long(m.mateMin) + long(len[readIndex]);
}

/// calculateShadowRescueRange (ShadowAligner.cpp:119-149): first / second of the pair it returns
ISAAC_HD inline void shadowRescueRange(const ShadowWindowModel &m, const isaac_ext_rescue_request_t &q, const uint32_t *len, long &first, long &second)
{
    const unsigned orphanReadIndex = q.orphanReadId % 2, shadowReadIndex = (orphanReadIndex + 1) % 2;
    const bool orphanReverse = q.orphanContigStrand & 1u;
    long shadowMin = shadowMateMinPosition(m, orphanReadIndex, orphanReverse, q.orphanPosition, len);
    long shadowMax = shadowMateMaxPosition(m, orphanReadIndex, orphanReverse, q.orphanPosition, len) + long(len[shadowReadIndex]) - 1;
    if (q.bestTemplateLength)
    {
        const long fStrand = q.orphanPosition;                                                      // FragmentMetadata.hh:90-95
        const long end = q.orphanPosition + long(q.orphanObservedLength);
        const long rStrand = (end > 1L ? end : 1L) - 1;                                             // :97-103
        if (shadowMin < fStrand) { const long wide = rStrand - q.bestTemplateLength; if (wide < shadowMin) shadowMin = wide; }
        if (shadowMax > fStrand) { const long wide = fStrand + q.bestTemplateLength; if (wide > shadowMax) shadowMax = wide; }
    }
    first = shadowMin - 10; second = shadowMax + 10;                                                // :147
}

/// the scan window of a request: Task is any struct with windowBegin, windowEnd, shadowReadId, contigStrand (ShadowTask, kernels_shadow.cuh)
template <class Task>
ISAAC_HD inline void shadowWindowOf(const ShadowWindowModel &m, const isaac_ext_rescue_request_t &q, const uint32_t *len, const long contigLength, Task &task)
{
    const unsigned orphanReadIndex = q.orphanReadId % 2, shadowReadIndex = (orphanReadIndex + 1) % 2;
    const bool orphanReverse = q.orphanContigStrand & 1u;
    long first, second;
    shadowRescueRange(m, q, len, first, second);
    task.shadowReadId = q.orphanReadId - orphanReadIndex + shadowReadIndex;
    task.contigStrand = ((q.orphanContigStrand >> 1) << 1) | (shadowMateOrientation(m, orphanReadIndex, orphanReverse) ? 1u : 0u);
    task.windowBegin = first > 0 ? first : 0;                                                       // ShadowAligner.cpp:194
    task.windowEnd = contigLength < second + 1 ? contigLength : second + 1;                         // :197
    if (second < first || second + 1 + long(len[shadowReadIndex]) < 0) task.windowEnd = task.windowBegin;   // :179-190
}

} // namespace isaac_b200
