// contract18_fused_bwd.cu -- the backward half of the fused 18-way kernels: 128-thread tiles, 3 CTAs per SM
// (see contract18_fused_impl.cuh).
#ifndef CCN_KTHREADS  /* overridable for A/B builds (profiles/build_variants.sh) */
#define CCN_KTHREADS 128
#endif
#define CCN_FUSED_BACKWARD 1
#include "contract18_fused_impl.cuh"
