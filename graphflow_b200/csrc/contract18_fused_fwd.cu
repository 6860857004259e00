// contract18_fused_fwd.cu -- the forward half of the fused 18-way kernels: 256-thread tiles, 2 CTAs per SM
// (see contract18_fused_impl.cuh).
#ifndef CCN_KTHREADS  /* overridable for A/B builds (profiles/build_variants.sh) */
#define CCN_KTHREADS 256
#endif
#define CCN_FUSED_FORWARD 1
#include "contract18_fused_impl.cuh"
