// contract18_fused_fwd.cu -- the forward half of the fused 18-way kernels: 256-thread tiles, 2 CTAs per SM
// (see contract18_fused_impl.cuh).
#define CCN_KTHREADS 256
#define CCN_FUSED_FORWARD 1
#include "contract18_fused_impl.cuh"
