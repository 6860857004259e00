// contract18_generic.cu -- shape-agnostic CUDA path for StackTensor3D + RisiContraction_18 (any n, any C), plus the
// adjacency-table kernel shared with the fast path.
//
// Replaces GraphFlow/RisiContraction_18.h:73-560 (and the reference kernels
// GraphFlow_gpu/RisiContraction_18_gpu.h:49-379, 541-685) using the closed forms of SURVEY.md section 8(a):
// every slab is (a partial sum / diagonal of T over its own indices) x (A, its row sums r, its total sA or its
// trace tr), so the reference's nnz(adj)*N^3*C*5 updates become one N^3*C pass plus O(N^2 * nnz-per-row * C).
// This file favours clarity: one thread per plane / output element, coalesced over the channel index.  The
// TMA-streamed kernels for the benchmark shapes are in contract18_fast.cu.
#include <algorithm>

#include "contract18_kernels.cuh"

namespace ccn {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ const float *slab_ptr(const TensorRef &t, int inst, int a, int n, int n_max, int C) {
    return t.slabs ? t.slabs[(int64_t)inst * n_max + a] : t.base + inst * t.stride + (int64_t)a * n * n * C;
}

// ---------------------------------------------------------------------------------------------------------------
// Adjacency table.  One CTA per instance.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_adj_prepare(const float *__restrict__ adj, int64_t stride_adj, Batch b,
                                                     int positive_part, float *__restrict__ adjtab, int words) {
    const int inst = blockIdx.x;
    const int n = b.n_of(inst);
    const int nm = b.n_max;
    const float *A = adj + inst * stride_adj;
    float *tab = adjtab + (int64_t)inst * words;
    AdjTabLayout L{nm};
    float *Ae = tab + L.A();
    float *r = tab + L.r();
    float *scal = tab + L.scal();
    int *rowptr = reinterpret_cast<int *>(tab + L.rowptr());
    int *colptr = reinterpret_cast<int *>(tab + L.colptr());
    int *rowidx = reinterpret_cast<int *>(tab + L.rowidx());
    float *rowval = tab + L.rowval();
    int *colidx = reinterpret_cast<int *>(tab + L.colidx());
    float *colval = tab + L.colval();

    for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
        float v = A[i];
        if (positive_part && !(v > 0.0f)) v = 0.0f;  // RisiContraction_18.h:90 `if (adj_value > 0)`
        Ae[i] = v;
    }
    __syncthreads();
    for (int d = threadIdx.x; d < n; d += blockDim.x) {
        float s = 0.0f;
        int rc = 0, cc = 0;
        for (int e = 0; e < n; ++e) {
            const float v = Ae[d * n + e];
            s += v;
            rc += (v != 0.0f);
            cc += (Ae[e * n + d] != 0.0f);
        }
        r[d] = s;
        rowptr[d + 1] = rc;
        colptr[d + 1] = cc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        rowptr[0] = 0;
        colptr[0] = 0;
        float sA = 0.0f, tr = 0.0f;
        for (int d = 0; d < n; ++d) {
            rowptr[d + 1] += rowptr[d];
            colptr[d + 1] += colptr[d];
            sA += r[d];
            tr += Ae[d * n + d];
        }
        scal[0] = sA;
        scal[1] = tr;
        scal[2] = (float)rowptr[n];
        scal[3] = 0.0f;
    }
    __syncthreads();
    for (int d = threadIdx.x; d < n; d += blockDim.x) {
        int j = rowptr[d], k = colptr[d];
        for (int e = 0; e < n; ++e) {
            const float v = Ae[d * n + e];
            if (v != 0.0f) {
                rowidx[j] = e;
                rowval[j] = v;
                ++j;
            }
            const float w = Ae[e * n + d];
            if (w != 0.0f) {
                colidx[k] = e;
                colval[k] = w;
                ++k;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Generic forward.  scratch per instance: planes P,Q,W6,W10,D1,D2 [6][nm*nm*C], sums S2,S4,S8,S11 [4][nm*C],
// totals tot5,tot14,tot15,tot18 [4][C].
// ---------------------------------------------------------------------------------------------------------------
struct GenFwdScratch {
    int64_t plane, sums, tots, words;
    __host__ __device__ GenFwdScratch(int nm, int C) {
        plane = (int64_t)nm * nm * C;
        sums = 6 * plane;
        tots = sums + 4 * (int64_t)nm * C;
        words = (tots + 4 * C + 3) & ~(int64_t)3;
    }
};

__global__ void __launch_bounds__(kThreads) k_gen_fwd_planes(Contract18Fwd a) {
    const int inst = blockIdx.y;
    const int n = a.b.n_of(inst), C = a.b.C, nm = a.b.n_max;
    const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (idx >= (int64_t)n * n * C) return;
    const uint32_t i32 = (uint32_t)idx;  // n*n*C < 2^31 (checked by the C-ABI): 32-bit divisions, not 64-bit ones
    const int f = (int)(i32 % (uint32_t)C);
    const int y = (int)((i32 / (uint32_t)C) % (uint32_t)n);
    const int x = (int)(i32 / ((uint32_t)C * (uint32_t)n));
    const AdjView av = adj_view(a.adjtab + (int64_t)inst * a.adjtab_words, nm);
    const GenFwdScratch S(nm, C);
    float *sc = a.scratch + inst * a.scratch_words;

    const float *Tx = slab_ptr(a.T, inst, x, n, nm, C);
    float P = 0.f, W6 = 0.f;
    for (int c = 0; c < n; ++c) {
        const float t = Tx[((int64_t)y * n + c) * C + f];
        P += t;
        W6 = fmaf(t, av.r[c], W6);
    }
    float Q = 0.f, W10 = 0.f;
    for (int s = 0; s < n; ++s) {
        const float t = slab_ptr(a.T, inst, s, n, nm, C)[((int64_t)x * n + y) * C + f];
        Q += t;
        W10 = fmaf(av.r[s], t, W10);
    }
    sc[0 * S.plane + idx] = P;
    sc[1 * S.plane + idx] = Q;
    sc[2 * S.plane + idx] = W6;
    sc[3 * S.plane + idx] = W10;
    sc[4 * S.plane + idx] = Tx[((int64_t)y * n + y) * C + f];  // D1[x,y] = T[x,y,y]
    sc[5 * S.plane + idx] = Tx[((int64_t)y * n + x) * C + f];  // D2[x,y] = T[x,y,x]
}

__global__ void __launch_bounds__(kThreads) k_gen_fwd_sums(Contract18Fwd a) {
    const int inst = blockIdx.x;
    const int n = a.b.n_of(inst), C = a.b.C, nm = a.b.n_max;
    const GenFwdScratch S(nm, C);
    float *sc = a.scratch + inst * a.scratch_words;
    const float *P = sc, *D1 = sc + 4 * S.plane, *D2 = sc + 5 * S.plane;
    float *sums = sc + S.sums;
    const int64_t vec = (int64_t)nm * C;
    for (int i = threadIdx.x; i < n * C; i += blockDim.x) {
        const int row = i / C, f = i % C;
        float s2 = 0.f, s4 = 0.f, s8 = 0.f, s11 = 0.f;
        for (int j = 0; j < n; ++j) {
            s2 += P[((int64_t)row * n + j) * C + f];
            s4 += P[((int64_t)j * n + row) * C + f];
            s8 += D1[((int64_t)row * n + j) * C + f];
            s11 += D2[((int64_t)j * n + row) * C + f];
        }
        sums[0 * vec + i] = s2;
        sums[1 * vec + i] = s4;
        sums[2 * vec + i] = s8;
        sums[3 * vec + i] = s11;
    }
    __syncthreads();
    float *tots = sc + S.tots;
    for (int f = threadIdx.x; f < C; f += blockDim.x) {
        float t5 = 0.f, t14 = 0.f, t15 = 0.f, t18 = 0.f;
        for (int j = 0; j < n; ++j) {
            t5 += sums[0 * vec + j * C + f];
            t15 += sums[2 * vec + j * C + f];
            t14 += P[((int64_t)j * n + j) * C + f];
            t18 += D1[((int64_t)j * n + j) * C + f];
        }
        tots[0 * C + f] = t5;
        tots[1 * C + f] = t14;
        tots[2 * C + f] = t15;
        tots[3 * C + f] = t18;
    }
}

__global__ void __launch_bounds__(kThreads) k_gen_fwd_out(Contract18Fwd a) {
    const int inst = blockIdx.y;
    const int n = a.b.n_of(inst), C = a.b.C, nm = a.b.n_max;
    const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (idx >= (int64_t)n * n * C) return;
    const uint32_t i32 = (uint32_t)idx;  // n*n*C < 2^31 (checked by the C-ABI): 32-bit divisions, not 64-bit ones
    const int f = (int)(i32 % (uint32_t)C);
    const int y = (int)((i32 / (uint32_t)C) % (uint32_t)n);
    const int x = (int)(i32 / ((uint32_t)C * (uint32_t)n));
    const AdjView av = adj_view(a.adjtab + (int64_t)inst * a.adjtab_words, nm);
    const GenFwdScratch S(nm, C);
    const float *sc = a.scratch + inst * a.scratch_words;
    const float *P = sc, *Q = sc + S.plane, *W6 = sc + 2 * S.plane, *W10 = sc + 3 * S.plane, *D1 = sc + 4 * S.plane,
                *D2 = sc + 5 * S.plane;
    const float *sums = sc + S.sums, *tots = sc + S.tots;
    const int64_t vec = (int64_t)nm * C;
    const float sA = av.scal[0], tr = av.scal[1];
    const float ry = av.r[y], Axy = av.A[x * n + y];

    float o9 = 0.f, o12 = 0.f, o13 = 0.f, o16 = 0.f, o17 = 0.f;
    for (int j = av.rowptr[y]; j < av.rowptr[y + 1]; ++j) {  // sum_e A[y,e] * (...)
        const int e = av.rowidx[j];
        const float w = av.rowval[j];
        o9 = fmaf(w, P[((int64_t)x * n + e) * C + f], o9);
        o12 = fmaf(w, P[((int64_t)e * n + x) * C + f], o12);
        o13 = fmaf(w, Q[((int64_t)x * n + e) * C + f], o13);
        o16 = fmaf(w, D1[((int64_t)x * n + e) * C + f], o16);
        o17 = fmaf(w, D2[((int64_t)e * n + x) * C + f], o17);
    }
    float *o = a.out + inst * a.stride_out + ((int64_t)x * n + y) * (kSlabs * C) + f;
    const float p = P[idx];
    o[0 * C] = sA * p;                          // case 1  (RisiContraction_18.h:102)
    o[1 * C] = ry * sums[0 * vec + x * C + f];  // case 2  (:106)
    o[2 * C] = sA * Q[idx];                     // case 3  (:110)
    o[3 * C] = ry * sums[1 * vec + x * C + f];  // case 4  (:114)
    o[4 * C] = Axy * tots[0 * C + f];           // case 5  (:118)
    o[5 * C] = W6[idx];                         // case 6  (:133)
    o[6 * C] = tr * p;                          // case 7  (:149)
    o[7 * C] = ry * sums[2 * vec + x * C + f];  // case 8  (:165)
    o[8 * C] = o9;                              // case 9  (:180)
    o[9 * C] = W10[idx];                        // case 10 (:195)
    o[10 * C] = ry * sums[3 * vec + x * C + f]; // case 11 (:211)
    o[11 * C] = o12;                            // case 12 (:226)
    o[12 * C] = o13;                            // case 13 (:241)
    o[13 * C] = Axy * tots[1 * C + f];          // case 14 (:256)
    o[14 * C] = Axy * tots[2 * C + f];          // case 15 (:271)
    o[15 * C] = o16;                            // case 16 (:290)
    o[16 * C] = o17;                            // case 17 (:304)
    o[17 * C] = Axy * tots[3 * C + f];          // case 18 (:318)
}

// ---------------------------------------------------------------------------------------------------------------
// Generic backward.  scratch per instance: planes U,V,E1,E2 [4][nm*nm*C], u2,u4,u8,u11 [4][nm*C],
// s5,s14,s15,s18 [4][C].  Formulas: SURVEY.md section 8(a) "Backward".
// ---------------------------------------------------------------------------------------------------------------
struct GenBwdScratch {
    int64_t plane, uvec, scal, words;
    __host__ __device__ GenBwdScratch(int nm, int C) {
        plane = (int64_t)nm * nm * C;
        uvec = 4 * plane;
        scal = uvec + 4 * (int64_t)nm * C;
        words = (scal + 4 * C + 3) & ~(int64_t)3;
    }
};

__device__ __forceinline__ float g_at(const float *g, int n, int C, int x, int y, int k, int f) {
    return g[((int64_t)x * n + y) * (kSlabs * C) + k * C + f];
}

__global__ void __launch_bounds__(kThreads) k_gen_bwd_vectors(Contract18Bwd a) {
    const int inst = blockIdx.x;
    const int n = a.b.n_of(inst), C = a.b.C, nm = a.b.n_max;
    const AdjView av = adj_view(a.adjtab + (int64_t)inst * a.adjtab_words, nm);
    const GenBwdScratch S(nm, C);
    float *sc = a.scratch + inst * a.scratch_words;
    const float *g = a.gout + inst * a.stride_gout;
    const int64_t vec = (int64_t)nm * C;
    for (int i = threadIdx.x; i < n * C; i += blockDim.x) {
        const int row = i / C, f = i % C;
        float u2 = 0.f, u4 = 0.f, u8 = 0.f, u11 = 0.f;
        for (int d = 0; d < n; ++d) {
            const float rd = av.r[d];
            u2 = fmaf(rd, g_at(g, n, C, row, d, 1, f), u2);
            u4 = fmaf(rd, g_at(g, n, C, row, d, 3, f), u4);
            u8 = fmaf(rd, g_at(g, n, C, row, d, 7, f), u8);
            u11 = fmaf(rd, g_at(g, n, C, row, d, 10, f), u11);
        }
        sc[S.uvec + 0 * vec + i] = u2;
        sc[S.uvec + 1 * vec + i] = u4;
        sc[S.uvec + 2 * vec + i] = u8;
        sc[S.uvec + 3 * vec + i] = u11;
    }
    for (int f = threadIdx.x; f < C; f += blockDim.x) {
        float s5 = 0.f, s14 = 0.f, s15 = 0.f, s18 = 0.f;
        for (int d = 0; d < n; ++d)
            for (int j = av.rowptr[d]; j < av.rowptr[d + 1]; ++j) {
                const int e = av.rowidx[j];
                const float w = av.rowval[j];
                s5 = fmaf(w, g_at(g, n, C, d, e, 4, f), s5);
                s14 = fmaf(w, g_at(g, n, C, d, e, 13, f), s14);
                s15 = fmaf(w, g_at(g, n, C, d, e, 14, f), s15);
                s18 = fmaf(w, g_at(g, n, C, d, e, 17, f), s18);
            }
        sc[S.scal + 0 * C + f] = s5;
        sc[S.scal + 1 * C + f] = s14;
        sc[S.scal + 2 * C + f] = s15;
        sc[S.scal + 3 * C + f] = s18;
    }
}

__global__ void __launch_bounds__(kThreads) k_gen_bwd_planes(Contract18Bwd a) {
    const int inst = blockIdx.y;
    const int n = a.b.n_of(inst), C = a.b.C, nm = a.b.n_max;
    const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (idx >= (int64_t)n * n * C) return;
    const uint32_t i32 = (uint32_t)idx;  // n*n*C < 2^31 (checked by the C-ABI): 32-bit divisions, not 64-bit ones
    const int f = (int)(i32 % (uint32_t)C);
    const int y = (int)((i32 / (uint32_t)C) % (uint32_t)n);
    const int x = (int)(i32 / ((uint32_t)C * (uint32_t)n));
    const AdjView av = adj_view(a.adjtab + (int64_t)inst * a.adjtab_words, nm);
    const GenBwdScratch S(nm, C);
    float *sc = a.scratch + inst * a.scratch_words;
    const float *g = a.gout + inst * a.stride_gout;
    const int64_t vec = (int64_t)nm * C;
    const float sA = av.scal[0], tr = av.scal[1];
    const float *u2 = sc + S.uvec, *u4 = u2 + vec, *u8 = u4 + vec, *u11 = u8 + vec;
    const float *s5 = sc + S.scal, *s14 = s5 + C, *s15 = s14 + C, *s18 = s15 + C;

    // column y of A: sum_d A[d,y] * g_k[x,d]  (k = 9, 13, 16, 17);  column x of A: sum_d A[d,x] * g12[y,d]
    float a9 = 0.f, a13 = 0.f, a16 = 0.f, a17 = 0.f, a12 = 0.f;
    for (int j = av.colptr[y]; j < av.colptr[y + 1]; ++j) {
        const int d = av.colidx[j];
        const float w = av.colval[j];
        a9 = fmaf(w, g_at(g, n, C, x, d, 8, f), a9);
        a13 = fmaf(w, g_at(g, n, C, x, d, 12, f), a13);
        a16 = fmaf(w, g_at(g, n, C, x, d, 15, f), a16);
        a17 = fmaf(w, g_at(g, n, C, x, d, 16, f), a17);
    }
    for (int j = av.colptr[x]; j < av.colptr[x + 1]; ++j)
        a12 = fmaf(av.colval[j], g_at(g, n, C, y, av.colidx[j], 11, f), a12);

    const float diag = (x == y) ? 1.f : 0.f;
    // U[a=x, b=y]
    sc[0 * S.plane + idx] = sA * g_at(g, n, C, x, y, 0, f) + tr * g_at(g, n, C, x, y, 6, f) + a9 + a12 + u2[x * C + f] +
                            u4[y * C + f] + s5[f] + diag * s14[f];
    // V[b=x, c=y]
    sc[1 * S.plane + idx] = sA * g_at(g, n, C, x, y, 2, f) + a13;
    // E1[a=x, b=y]  (applies where c == b)
    sc[2 * S.plane + idx] = u8[x * C + f] + s15[f] + a16 + diag * s18[f];
    // E2[b=x, a=y]  (applies where c == a)
    sc[3 * S.plane + idx] = u11[x * C + f] + a17;
}

// Grid (x, instance) with a strided loop over the instance's real n^3 C elements: ragged batches launch the same number of
// CTAs per instance whatever n is, instead of n_max^3 C / 256 mostly empty ones.
__global__ void __launch_bounds__(kThreads) k_gen_bwd_scatter(Contract18Bwd a) {
    const int inst = blockIdx.y;
    const int n = a.b.n_of(inst), C = a.b.C, nm = a.b.n_max;
    const int64_t slab = (int64_t)n * n * C;
    const uint32_t total = (uint32_t)(slab * n);  // n^3 C < 2^31, checked by the C-ABI
    const AdjView av = adj_view(a.adjtab + (int64_t)inst * a.adjtab_words, nm);
    const GenBwdScratch S(nm, C);
    const float *sc = a.scratch + inst * a.scratch_words;
    const float *g = a.gout + inst * a.stride_gout;
    for (uint32_t idx = blockIdx.x * kThreads + threadIdx.x; idx < total; idx += gridDim.x * kThreads) {
        const int s = (int)(idx / (uint32_t)slab);  // a
        const uint32_t rem = idx - (uint32_t)s * (uint32_t)slab;
        const int f = (int)(rem % (uint32_t)C);
        const int c = (int)((rem / (uint32_t)C) % (uint32_t)n);
        const int bb = (int)(rem / ((uint32_t)C * (uint32_t)n));
        const int64_t ab = ((int64_t)s * n + bb) * C + f, bc = ((int64_t)bb * n + c) * C + f;
        float v = sc[0 * S.plane + ab] + sc[1 * S.plane + bc] + g_at(g, n, C, s, bb, 5, f) * av.r[c] +
                  av.r[s] * g_at(g, n, C, bb, c, 9, f);
        if (bb == c) v += sc[2 * S.plane + ab];
        if (s == c) v += sc[3 * S.plane + ((int64_t)bb * n + s) * C + f];
        float *dst = (a.gT.slabs ? a.gT.slabs[(int64_t)inst * nm + s] : a.gT.base + inst * a.gT.stride + (int64_t)s * slab) + rem;
        *dst = (a.beta != 0.f) ? fmaf(a.beta, *dst, v) : v;
    }
}

inline unsigned blocks_for(int64_t elems) { return (unsigned)((elems + kThreads - 1) / kThreads); }

}  // namespace

cudaError_t launch_adj_prepare(const float *adj, int64_t stride_adj, Batch b, int adj_mode, float *adjtab,
                               cudaStream_t st, LaunchLog *log) {
    AdjTabLayout L{b.n_max};
    CCN_LAUNCH(log, K_ADJ_PREPARE, st,
               k_adj_prepare<<<b.count, 128, 0, st>>>(adj, stride_adj, b, adj_mode == 0 ? 1 : 0, adjtab, L.words()));
    return cudaGetLastError();
}

int64_t generic_fwd_scratch_words(int n_max, int C) { return GenFwdScratch(n_max, C).words; }
int64_t generic_bwd_scratch_words(int n_max, int C) { return GenBwdScratch(n_max, C).words; }

// blockIdx.y carries the instance, so a launch covers at most 65535 instances; the C-ABI chunks above that.
cudaError_t launch_generic_forward(const Contract18Fwd &a, cudaStream_t st, LaunchLog *log) {
    const int64_t plane = (int64_t)a.b.n_max * a.b.n_max * a.b.C;
    dim3 grid(blocks_for(plane), a.b.count);
    CCN_LAUNCH(log, K_GEN_FWD_PLANES, st, k_gen_fwd_planes<<<grid, kThreads, 0, st>>>(a));
    CCN_LAUNCH(log, K_GEN_FWD_SUMS, st, k_gen_fwd_sums<<<a.b.count, kThreads, 0, st>>>(a));
    CCN_LAUNCH(log, K_GEN_FWD_OUT, st, k_gen_fwd_out<<<grid, kThreads, 0, st>>>(a));
    return cudaGetLastError();
}

cudaError_t launch_generic_backward(const Contract18Bwd &a, cudaStream_t st, LaunchLog *log) {
    const int64_t plane = (int64_t)a.b.n_max * a.b.n_max * a.b.C;
    dim3 grid(blocks_for(plane), a.b.count);
    const unsigned per_inst = std::min(64u, std::max(1u, blocks_for(plane * a.b.n_max) / 8));
    dim3 grid3(per_inst, a.b.count);
    CCN_LAUNCH(log, K_GEN_BWD_VECTORS, st, k_gen_bwd_vectors<<<a.b.count, kThreads, 0, st>>>(a));
    CCN_LAUNCH(log, K_GEN_BWD_PLANES, st, k_gen_bwd_planes<<<grid, kThreads, 0, st>>>(a));
    CCN_LAUNCH(log, K_GEN_BWD_SCATTER, st, k_gen_bwd_scatter<<<grid3, kThreads, 0, st>>>(a));
    return cudaGetLastError();
}

}  // namespace ccn
