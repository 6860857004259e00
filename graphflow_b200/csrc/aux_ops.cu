// aux_ops.cu -- the remaining operators on the CCN hot path, sm_100a:
//
//   promotion   Q_w = X f_{l-1}[w] X^T with X the 0/1 selection matrix of init_permutation_matrix (SMP_beta.h:446-459),
//               i.e. MatTensorMul::forward (MatTensorMul.h:47-65) + TensorMatMul::forward (TensorMatMul.h:46-64) as wired
//               at SMP_beta.h:588-594, followed by StackTensor3D: a gather with zero fill straight into the stacked
//               T[a] (and an atomic scatter-add back for the two backward passes, MatTensorMul.h:67-85).
//   TensorMul   per-channel [R x K] . [K x Cc] product (TensorMul.h:48-86), the feature mix of SMP_2D v1-5.
//   transposes  for CustomMatMulTensor (CustomMatMulTensor.h:47-85): the same GEMM as the feature mix with K stored
//               [C_out, 18 C]; the C-ABI transposes the small weight matrix and reuses the mix kernels.
//
// All three are HBM-bound copies / small products: coalesced over the channel index, one thread per output element.
#include <cmath>
#include "contract18_kernels.cuh"

namespace ccn {

namespace {

constexpr int kThreads = 256;

// One CTA per (slab a, instance): CTAs of slabs beyond the instance's n exit at once, and the threads stride over the
// slab's real n*n*C elements (no padded work for ragged batches).
__global__ void __launch_bounds__(kThreads) k_promote_fwd(PromoteArgs a) {
    const int inst = blockIdx.y, slab = blockIdx.x;
    const int n = a.n ? a.n[inst] : a.n_max, C = a.C;
    if (slab >= n) return;
    const int64_t s = (int64_t)inst * a.n_max + slab;
    const int *pos = a.pos + s * a.n_max;
    const int m = a.m[s];
    const float *src = a.f + a.f_off[s];
    float *dst = a.T + inst * a.stride_T + (int64_t)slab * n * n * C;
    const int total = n * n * C;
    for (int idx = threadIdx.x; idx < total; idx += kThreads) {
        const int c = idx % C, ij = idx / C, j = ij % n, i = ij / n;
        const int pi = pos[i], pj = pos[j];
        dst[idx] = (pi >= 0 && pj >= 0) ? src[((int64_t)pi * m + pj) * C + c] : 0.f;
    }
}

__global__ void __launch_bounds__(kThreads) k_promote_bwd(PromoteArgs a) {
    const int inst = blockIdx.y, slab = blockIdx.x;
    const int n = a.n ? a.n[inst] : a.n_max, C = a.C;
    if (slab >= n) return;
    const int64_t s = (int64_t)inst * a.n_max + slab;
    const int *pos = a.pos + s * a.n_max;
    const int m = a.m[s];
    float *dstf = a.f + a.f_off[s];
    const float *g = a.T + inst * a.stride_T + (int64_t)slab * n * n * C;
    const int total = n * n * C;
    for (int idx = threadIdx.x; idx < total; idx += kThreads) {
        const int c = idx % C, ij = idx / C, j = ij % n, i = ij / n;
        const int pi = pos[i], pj = pos[j];
        // the same f_{l-1}[w] is promoted into the stack of every vertex whose field contains w: accumulate atomically
        if (pi >= 0 && pj >= 0) atomicAdd(dstf + ((int64_t)pi * m + pj) * C + c, g[idx]);
    }
}

// float4 versions (C % 4 == 0, C <= 128, 16-byte aligned buffers): C/4 lanes cover one (i, j) cell with one 16-byte
// access each, a warp covers 32 / (C/4) cells per step; the index arithmetic is per cell, not per element, and the
// backward uses one vector atomic (red.global.add.v4.f32) per 16 bytes.
template <bool BACKWARD>
__global__ void __launch_bounds__(kThreads) k_promote_v4(PromoteArgs a) {
    const int inst = blockIdx.y, slab = blockIdx.x;
    const int n = a.n ? a.n[inst] : a.n_max, C = a.C;
    if (slab >= n) return;
    const int64_t s = (int64_t)inst * a.n_max + slab;
    const int *pos = a.pos + s * a.n_max;
    const int m = a.m[s];
    float *fbase = a.f + a.f_off[s];
    float *tbase = a.T + inst * a.stride_T + (int64_t)slab * n * n * C;
    const int lpc = C >> 2;                      // lanes per cell
    const int cpw = 32 / lpc;                    // cells per warp step (lanes beyond cpw * lpc idle)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / lpc, c4 = lane - sub * lpc;
    if (sub >= cpw) return;
    const int cells = n * n, step = (kThreads / 32) * cpw;
    for (int cell = warp * cpw + sub; cell < cells; cell += step) {
        const int i = cell / n, j = cell - i * n;
        const int pi = pos[i], pj = pos[j];
        float4 *t = reinterpret_cast<float4 *>(tbase + (int64_t)cell * C) + c4;
        if (pi >= 0 && pj >= 0) {
            float4 *f = reinterpret_cast<float4 *>(fbase + ((int64_t)pi * m + pj) * C) + c4;
            if (BACKWARD) atomicAdd(f, *t);
            else *t = *f;
        } else if (!BACKWARD) {
            *t = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

struct TMulArgs {
    const float *A, *B, *g;
    float *out, *gA, *gB;
    int R, K, Cc, D;
    float beta;
};

// mode 0: out[i,j,d] = sum_k A[i,k,d] B[k,j,d];  1: gA[i,k,d] = beta gA + sum_j g[i,j,d] B[k,j,d];
// mode 2: gB[k,j,d] = beta gB + sum_i g[i,j,d] A[i,k,d]
template <int MODE>
__global__ void __launch_bounds__(kThreads) k_tensor_mul(TMulArgs a) {
    const int64_t inst = blockIdx.y;
    const int R = a.R, K = a.K, Cc = a.Cc, D = a.D;
    const int d1 = MODE == 0 ? R : (MODE == 1 ? R : K), d2 = MODE == 0 ? Cc : (MODE == 1 ? K : Cc);
    const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (idx >= (int64_t)d1 * d2 * D) return;
    const int d = (int)(idx % D), y = (int)((idx / D) % d2), x = (int)(idx / ((int64_t)D * d2));
    const float *A = a.A + inst * (int64_t)R * K * D, *B = a.B + inst * (int64_t)K * Cc * D;
    float acc = 0.f;
    if (MODE == 0) {
        for (int k = 0; k < K; ++k) acc = fmaf(A[((int64_t)x * K + k) * D + d], B[((int64_t)k * Cc + y) * D + d], acc);
        a.out[inst * (int64_t)R * Cc * D + idx] = acc;
    } else {
        const float *g = a.g + inst * (int64_t)R * Cc * D;
        if (MODE == 1) {
            for (int j = 0; j < Cc; ++j) acc = fmaf(g[((int64_t)x * Cc + j) * D + d], B[((int64_t)y * Cc + j) * D + d], acc);
            float *dst = a.gA + inst * (int64_t)R * K * D + idx;
            *dst = a.beta != 0.f ? fmaf(a.beta, *dst, acc) : acc;
        } else {
            for (int i = 0; i < R; ++i) acc = fmaf(g[((int64_t)i * Cc + y) * D + d], A[((int64_t)i * K + x) * D + d], acc);
            float *dst = a.gB + inst * (int64_t)K * Cc * D + idx;
            *dst = a.beta != 0.f ? fmaf(a.beta, *dst, acc) : acc;
        }
    }
}

// dst[c, r] = beta dst[c, r] + src[r, c]   (src is [rows, cols])
__global__ void __launch_bounds__(kThreads) k_transpose_add(const float *__restrict__ src, float *__restrict__ dst, int rows, int cols,
                                                            float beta) {
    const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (idx >= (int64_t)rows * cols) return;
    const int c = (int)(idx % cols), r = (int)(idx / cols);
    float *d = dst + (int64_t)c * rows + r;
    *d = beta != 0.f ? fmaf(beta, *d, src[idx]) : src[idx];
}

inline unsigned blocks_for(int64_t elems) { return (unsigned)((elems + kThreads - 1) / kThreads); }

}  // namespace

cudaError_t launch_promote(bool backward, const PromoteArgs &a, int batch, cudaStream_t st, LaunchLog *log) {
    dim3 grid(a.n_max, batch);
    const bool v4 = (a.C % 4) == 0 && a.C <= 128 && ((reinterpret_cast<uintptr_t>(a.f) | reinterpret_cast<uintptr_t>(a.T)) & 15u) == 0 &&
                    (a.stride_T % 4) == 0 && a.v4_offsets_ok;
    if (v4) {
        if (!backward)
            CCN_LAUNCH(log, K_PROMOTE_FWD, st, k_promote_v4<false><<<grid, kThreads, 0, st>>>(a));
        else
            CCN_LAUNCH(log, K_PROMOTE_BWD, st, k_promote_v4<true><<<grid, kThreads, 0, st>>>(a));
        return cudaGetLastError();
    }
    if (!backward)
        CCN_LAUNCH(log, K_PROMOTE_FWD, st, k_promote_fwd<<<grid, kThreads, 0, st>>>(a));
    else
        CCN_LAUNCH(log, K_PROMOTE_BWD, st, k_promote_bwd<<<grid, kThreads, 0, st>>>(a));
    return cudaGetLastError();
}

cudaError_t launch_tensor_mul_forward(const float *A, const float *B, float *out, int R, int K, int Cc, int D, int batch,
                                      cudaStream_t st, LaunchLog *log) {
    TMulArgs a{A, B, nullptr, out, nullptr, nullptr, R, K, Cc, D, 0.f};
    CCN_LAUNCH(log, K_TENSOR_MUL, st, (k_tensor_mul<0><<<dim3(blocks_for((int64_t)R * Cc * D), batch), kThreads, 0, st>>>(a)));
    return cudaGetLastError();
}

cudaError_t launch_tensor_mul_backward(const float *A, const float *B, const float *g, float *gA, float *gB, int R, int K, int Cc,
                                       int D, int batch, float beta, cudaStream_t st, LaunchLog *log) {
    TMulArgs a{A, B, g, nullptr, gA, gB, R, K, Cc, D, beta};
    if (gA) CCN_LAUNCH(log, K_TENSOR_MUL, st, (k_tensor_mul<1><<<dim3(blocks_for((int64_t)R * K * D), batch), kThreads, 0, st>>>(a)));
    if (gB) CCN_LAUNCH(log, K_TENSOR_MUL, st, (k_tensor_mul<2><<<dim3(blocks_for((int64_t)K * Cc * D), batch), kThreads, 0, st>>>(a)));
    return cudaGetLastError();
}

// Slab dropout (RisiContraction_18_dropout.h:113-131,467-472) around the 18-way kernels: zero the dropped slabs of `out`
// in place (and scale the kept ones in test mode), or copy `gout` with the dropped slabs zeroed.
// One thread per (cell, channel); consecutive threads = consecutive channels.
__global__ void __launch_bounds__(kThreads) k_slab_mask_inplace(float *out, int64_t stride, Batch b, uint32_t keep, float scale, int S) {
    const int inst = blockIdx.y, n = b.n_of(inst), C = b.C;
    const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (idx >= (int64_t)n * n * C) return;
    const int64_t cell = idx / C;
    const int f = (int)(idx - cell * C);
    float *o = out + inst * stride + cell * ((int64_t)S * C) + f;
    for (int k = 0; k < S; ++k) {
        if (!((keep >> k) & 1u)) o[(int64_t)k * C] = 0.f;
        else if (scale != 1.f) o[(int64_t)k * C] *= scale;
    }
}
__global__ void __launch_bounds__(kThreads) k_slab_mask_copy(const float *__restrict__ g, int64_t stride_g, float *__restrict__ dst,
                                                             int64_t stride_d, Batch b, uint32_t keep, int S) {
    const int inst = blockIdx.y, n = b.n_of(inst), C = b.C;
    const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (idx >= (int64_t)n * n * S * C) return;
    const int k = (int)((idx / C) % S);
    dst[inst * stride_d + idx] = ((keep >> k) & 1u) ? g[inst * stride_g + idx] : 0.f;
}

cudaError_t launch_slab_mask_inplace(float *out, int64_t stride, Batch b, uint32_t keep, float scale, int S, cudaStream_t st,
                                     LaunchLog *log) {
    dim3 grid(blocks_for((int64_t)b.n_max * b.n_max * b.C), b.count);
    CCN_LAUNCH(log, K_TRANSPOSE, st, k_slab_mask_inplace<<<grid, kThreads, 0, st>>>(out, stride, b, keep, scale, S));
    return cudaGetLastError();
}
cudaError_t launch_slab_mask_copy(const float *g, int64_t stride_g, float *dst, int64_t stride_d, Batch b, uint32_t keep, int S,
                                  cudaStream_t st, LaunchLog *log) {
    dim3 grid(blocks_for((int64_t)b.n_max * b.n_max * S * b.C), b.count);
    CCN_LAUNCH(log, K_TRANSPOSE, st, k_slab_mask_copy<<<grid, kThreads, 0, st>>>(g, stride_g, dst, stride_d, b, keep, S));
    return cudaGetLastError();
}

// Adam::Learn (Adam.h:76-137).  per_element: the `Learn(alpha, nBatch)` overload multiplies beta1_t / beta2_t by beta once
// per ELEMENT (:123,127), so element i of the flat vector, after `before` earlier element updates, is corrected with
// beta^(before + i + 1); the `Learn(alpha)` overload (:82-83) advances them once per call: beta^(before + 1).
// Moments and parameters are fp32; the correction powers are evaluated in fp64 (exp(k log beta), rel. error ~ k eps).
__global__ void __launch_bounds__(kThreads) k_adam_step(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                                        float *__restrict__ v, int64_t n, float alpha, float beta1, float beta2,
                                                        float eps, float inv_batch, double log_b1, double log_b2, double before,
                                                        int per_element) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i] * inv_batch;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const double k = before + (per_element ? (double)(i + 1) : 1.0);
    const float c1 = (float)(1.0 - exp(k * log_b1)), c2 = (float)(1.0 - exp(k * log_b2));
    p[i] -= alpha * (mi / c1) / (sqrtf(vi / c2) + eps);
}

// Momentum::Learn (Momentum.h:51-67); gamma = 0 is SGD::Learn (SGD.h:36-50).
__global__ void __launch_bounds__(kThreads) k_momentum_step(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ mom,
                                                            int64_t n, float lr, float gamma, float inv_batch) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const float mi = gamma * mom[i] + lr * g[i] * inv_batch;
    mom[i] = mi;
    p[i] -= mi;
}

cudaError_t launch_adam_step(float *p, const float *g, float *m, float *v, int64_t n, double alpha, double beta1, double beta2,
                             double eps, double inv_batch, int64_t updates_before, bool per_element, cudaStream_t st, LaunchLog *log) {
    CCN_LAUNCH(log, K_OPTIMIZER, st,
               k_adam_step<<<blocks_for(n), kThreads, 0, st>>>(p, g, m, v, n, (float)alpha, (float)beta1, (float)beta2, (float)eps,
                                                               (float)inv_batch, std::log(beta1), std::log(beta2), (double)updates_before,
                                                               per_element ? 1 : 0));
    return cudaGetLastError();
}

cudaError_t launch_momentum_step(float *p, const float *g, float *mom, int64_t n, double lr, double gamma, double inv_batch,
                                 cudaStream_t st, LaunchLog *log) {
    CCN_LAUNCH(log, K_OPTIMIZER, st, k_momentum_step<<<blocks_for(n), kThreads, 0, st>>>(p, g, mom, n, (float)lr, (float)gamma, (float)inv_batch));
    return cudaGetLastError();
}

cudaError_t launch_transpose_add(const float *src, float *dst, int rows, int cols, float beta, cudaStream_t st, LaunchLog *log) {
    CCN_LAUNCH(log, K_TRANSPOSE, st, k_transpose_add<<<blocks_for((int64_t)rows * cols), kThreads, 0, st>>>(src, dst, rows, cols, beta));
    return cudaGetLastError();
}

}  // namespace ccn
