// One instantiation of the sweep kernel (FAST and EXACT arithmetic) for K = BRTA_INST_K.  The library
// compiles this file once per supported K, in parallel (basicrta_b200/_cabi.py: build).
#include "brta_sweep.cuh"

#ifndef BRTA_INST_K
#error "compile with -DBRTA_INST_K=<ncomp>"
#endif
#define BRTA_CAT2(a, b) a##b
#define BRTA_CAT(a, b) BRTA_CAT2(a, b)

namespace brta {
Variant BRTA_CAT(variant_k, BRTA_INST_K)()
{
    return Variant{BRTA_INST_K, gibbs_sweep_kernel<BRTA_INST_K, false>, gibbs_sweep_kernel<BRTA_INST_K, true>,
                   three_cta_kernel<BRTA_INST_K>()};
}
}  // namespace brta

#ifdef BRTA_PHASE_TIMING
extern "C" int brta_debug_set_phase_buffer(void* dev_ptr)
{
    unsigned long long* p = (unsigned long long*)dev_ptr;
    return (int)cudaMemcpyToSymbol(brta::g_phase, &p, sizeof(p));
}
#endif
