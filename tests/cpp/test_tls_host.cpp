// template_length_host.cuh on the CPU: the reference's testTemplateLengthStatistics.cpp (addTemplates :108-146) is replayed on it by
// tests/test_tls_host.py, one addTemplate call per hand-made pair of fragments.  TEST CODE, not a product path.
#include "../../isaac_aligner_b200/csrc/template_length_host.cuh"

extern "C" void *tls_host_new(int mateDriftRange) { return new TemplateLengthDistributionHost(mateDriftRange); }
extern "C" void tls_host_free(void *h) { delete static_cast<TemplateLengthDistributionHost *>(h); }
/// count addTemplate calls on the pair (f0, f1), the position of 'moving' (0 or 1) growing by one after each call;
/// \return how many of the calls returned true
extern "C" unsigned tls_host_add(void *h, isaac_ext_fragment_t *f0, isaac_ext_fragment_t *f1, const uint32_t *cigars, unsigned count, int moving)
{
    TemplateLengthDistributionHost &d = *static_cast<TemplateLengthDistributionHost *>(h);
    unsigned stable = 0;
    for (unsigned i = 0; i < count; ++i)
    {
        stable += d.addTemplate(f0, 1, f1, 1, cigars) ? 1u : 0u;
        if (moving == 0) ++f0->position; else if (moving == 1) ++f1->position;
    }
    return stable;
}
extern "C" void tls_host_get(void *h, unsigned *out)
{
    const TemplateLengthDistributionHost &d = *static_cast<TemplateLengthDistributionHost *>(h);
    out[0] = d.min; out[1] = d.median; out[2] = d.max; out[3] = d.lowStdDev; out[4] = d.highStdDev; out[5] = d.bestModels[0]; out[6] = d.bestModels[1];
    out[7] = d.stable ? 1u : 0u;
}
