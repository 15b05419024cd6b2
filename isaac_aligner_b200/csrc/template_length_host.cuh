// alignment::TemplateLengthDistribution (TemplateLengthStatistics.hh:262-340, .cpp:95-160,266-357) on the flat fragment records:
// the sequential bookkeeping isaac_ext_determine_template_length keeps on the host (isaac_ext_tls.cuh).  Plain host C++;
// tests/cpp/test_tls_host.cpp replays the reference's own unit test (testTemplateLengthStatistics.cpp) on it on the CPU.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <vector>

#include "../../include/isaac_ext.h"

namespace
{

/// alignment::TemplateLengthDistribution (TemplateLengthStatistics.hh:262-340)
struct TemplateLengthDistributionHost
{
    static const unsigned UPDATE_FREQUENCY = 10000;                 // :317
    static const unsigned TEMPLATE_LENGTH_THRESHOLD = 50000;        // :220
    static const unsigned INVALID_MODEL = 8;                        // InvalidAlignmentModel (:59)
    int mateDriftRange;
    unsigned min = -1U, max = -1U, median = -1U, lowStdDev = -1U, highStdDev = -1U, bestModels[2] = {INVALID_MODEL, INVALID_MODEL};
    bool stable = false;
    unsigned templateCount = 0, uniqueCount = 0, count = 0;
    std::vector<unsigned> histograms[INVALID_MODEL], lengthList;
    double lowerPercent, upperPercent, lowerPercent1z, upperPercent1z;

    explicit TemplateLengthDistributionHost(int drift) : mateDriftRange(drift)
    {
        // TemplateLengthStatistics.cpp:31-38 (boost::math::erf; STANDARD_DEVIATIONS_MAX = 3.0)
        const double interval = std::erf(3.0 / std::sqrt(2.0)), interval1z = std::erf(1.0 / std::sqrt(2.0));
        lowerPercent = (1.0 - interval) / 2.0; upperPercent = (1.0 + interval) / 2.0;
        lowerPercent1z = (1.0 - interval1z) / 2.0; upperPercent1z = (1.0 + interval1z) / 2.0;
    }

    struct Snapshot { unsigned min, median, max, low, high, m0, m1; };
    Snapshot snapshot() const { return Snapshot{min, median, max, lowStdDev, highStdDev, bestModels[0], bestModels[1]}; }
    bool sameNumbers(const Snapshot &o) const { return o.min == min && o.median == median && o.max == max && o.low == lowStdDev && o.high == highStdDev; }

    static unsigned alignmentClass(unsigned model) { return model < 4 ? model : ((~model) & 3); }

    /// updateStatistics (TemplateLengthStatistics.cpp:105-160)
    void updateStatistics()
    {
        const Snapshot old = snapshot();
        bestModels[0] = histograms[1].size() <= histograms[0].size() ? 0u : 1u;
        bestModels[1] = (bestModels[0] + 1) % 2;
        for (unsigned i = 2; i < INVALID_MODEL; ++i)
        {
            if (histograms[i].size() > histograms[bestModels[0]].size()) { bestModels[1] = bestModels[0]; bestModels[0] = i; }
            else if (histograms[i].size() > histograms[bestModels[1]].size()) bestModels[1] = i;
        }
        lengthList.clear();
        lengthList.insert(lengthList.end(), histograms[bestModels[0]].begin(), histograms[bestModels[0]].end());
        lengthList.insert(lengthList.end(), histograms[bestModels[1]].begin(), histograms[bestModels[1]].end());
        std::sort(lengthList.begin(), lengthList.end());
        const size_t n = lengthList.size();
        min = n ? lengthList[unsigned(n * lowerPercent)] : 0;
        median = n ? lengthList[unsigned(n * 0.5)] : TEMPLATE_LENGTH_THRESHOLD / 2;
        max = n ? lengthList[unsigned(n * upperPercent)] : TEMPLATE_LENGTH_THRESHOLD;
        lowStdDev = n ? median - lengthList[unsigned(n * lowerPercent1z)] : median;
        highStdDev = n ? lengthList[unsigned(n * upperPercent1z)] - median : median;
        if (sameNumbers(old) && old.m0 == bestModels[0] && old.m1 == bestModels[1]) stable = true;
    }

    /// addTemplate (:266-340) on the final fragment lists of the two reads of one cluster
    bool addTemplate(const isaac_ext_fragment_t *f0, size_t n0, const isaac_ext_fragment_t *f1, size_t n1, const uint32_t *cigars)
    {
        if (!n0 || !n1) return stable;
        ++templateCount;
        if (n0 > 1 || n1 > 1) return stable;
        ++uniqueCount;
        if (f0->contigId != f1->contigId) return stable;
        const isaac_ext_fragment_t *f[2] = {f0, f1};
        for (unsigned i = 0; i < 2; ++i)
        {
            const uint32_t firstOp = cigars[f[i]->cigarOffset], lastOp = cigars[f[i]->cigarOffset + f[i]->cigarLength - 1];
            if ((firstOp & 0xFu) == ISAAC_EXT_CIGAR_INSERT || (lastOp & 0xFu) == ISAAC_EXT_CIGAR_INSERT) return stable;
        }
        // TemplateLengthStatistics::getLength (TemplateLengthStatistics.hh:165-176)
        const unsigned long length = f0->position < f1->position
            ? (unsigned long)std::max<long>(f1->position + long(f1->observedLength) - f0->position, long(f0->observedLength))
            : (unsigned long)std::max<long>(f0->position + long(f0->observedLength) - f1->position, long(f1->observedLength));
        if (length > TEMPLATE_LENGTH_THRESHOLD) return stable;
        // alignmentModel (:153-163); the contigs are equal here
        const unsigned model = (f0->position <= f1->position ? 0u : 4u) | (f0->reverse ? 2u : 0u) | (f1->reverse ? 1u : 0u);
        histograms[model].push_back(unsigned(length));
        ++count;
        if (0 == count % UPDATE_FREQUENCY)
        {
            const Snapshot old = snapshot();
            updateStatistics();
            if (sameNumbers(old)) stable = true;
        }
        return stable;
    }

    /// finalize (:342-357)
    bool finalize()
    {
        const Snapshot old = snapshot();
        updateStatistics();
        if (sameNumbers(old)) stable = true;
        return stable;
    }
};

} // namespace
