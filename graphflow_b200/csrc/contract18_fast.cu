// contract18_fast.cu -- TMA-streamed StackTensor3D + RisiContraction_18 forward/backward for the benchmark shapes
// (n <= 32, C in {32, 64, 128}); sm_100a only.
//
// Replaces GraphFlow/StackTensor3D.h:54-90 + GraphFlow/RisiContraction_18.h:73-560 and the reference kernels
// GraphFlow_gpu/RisiContraction_18_gpu.h:49-379 (forward_job) / :541-685 (backward_job).
//
// Design.  Work unit = one *row tile*: (instance, TB consecutive values of T's middle index b),
// TB = 256 / C, one 256-thread CTA, thread (f, bl) owns channel f of row b = b0 + bl for the whole kernel.
//
//   k_fast_fwd_stream   for a = 0..n-1 the contiguous chunk T[a, b0:b0+TB, :, :] (TB*n*C floats, 32 KiB at n=32) is
//                       pulled by one cp.async.bulk (TMA engine) into a 3-stage shared-memory ring guarded by
//                       mbarriers.  Each thread walks its row's n cells once: Q[b,c] += t, W10[b,c] += r[a] t stay in
//                       registers across the a loop (sum over a); P[a,b] = sum_c t and W6[a,b] = sum_c r[c] t finish
//                       inside the step (sum over c).  No cross-thread reduction is needed anywhere.  Writes slabs
//                       1,3,4,6,7,10,11,13 and parks P, D1 = T[a,b,b], D2 = T[a,b,a] (3 N^2 C planes) for the finisher.
//   k_fast_fwd_finish   row i of the parked planes -> slabs 2,5,8,9,12,14,15,16,17,18 (the ones that need a sum over
//                       T's middle index, which is split across CTAs in the stream kernel).
//   k_fast_bwd_planes   gout -> five N^2 C planes (U_a, U_b^T, E1, E2, V), all products with A done sparsely.
//   k_fast_bwd_stream   gT[a,b,c] = U[a,b] + V[b,c] + g6[a,b] r[c] + r[a] g10[b,c] + diagonal terms: pure broadcast,
//                       V and g10 rows live in registers, 128-byte coalesced streaming stores of gT.
//
// The five [n x n] * A^T products per channel (cases 9, 12, 13, 16, 17) run over the CSR/CSC lists of A built by
// k_adj_prepare: molecular adjacency has ~3 non-zeros per row, so they cost ~10x less than a dense product and
// nothing is gained by reshaping them for the tensor cores; a dense A takes the same code path (longer lists).
#include "contract18_kernels.cuh"

namespace ccn {

namespace {

constexpr int NMAX = 32;       // largest receptive field handled by the register-resident row accumulators
constexpr int kThreads = 256;  // one thread per (channel, row-in-tile)
constexpr int kStages = 3;     // TMA ring depth (3 x 32 KiB in flight per CTA, 2 CTAs per SM)

__host__ __device__ inline int tiles_of(int n, int C) { return (n + (kThreads / C) - 1) / (kThreads / C); }

struct FastFwdScratch {  // planes P, D1, D2 then per-tile partial totals [tiles][4][C]
    int64_t plane, partials, words;
    __host__ __device__ FastFwdScratch(int nm, int C) {
        plane = (int64_t)nm * nm * C;
        partials = 3 * plane;
        words = (partials + (int64_t)tiles_of(nm, C) * 4 * C + 3) & ~(int64_t)3;
    }
};

struct FastBwdScratch {  // planes U_a[a,b], U_b^T[b,a], E1[a,b], E2[b,a], V[b,c]
    int64_t plane, words;
    __host__ __device__ FastBwdScratch(int nm, int C) {
        plane = (int64_t)nm * nm * C;
        words = 5 * plane;
    }
};

__device__ __forceinline__ float *slab_ptr(const TensorRef &t, int inst, int a, int n, int n_max, int C) {
    return t.slabs ? t.slabs[(int64_t)inst * n_max + a] : t.base + inst * t.stride + (int64_t)a * n * n * C;
}

constexpr int kStageFloats = kThreads * NMAX;  // TB * NMAX * C
constexpr size_t kFwdStreamSmem = (size_t)(kStages * kStageFloats + NMAX + 4 * kThreads) * 4 + kStages * 8;
constexpr size_t kRowBufSmem = (size_t)NMAX * kThreads * 4;

// One row (fixed a, b, f) of the staged chunk: n cells, stride C floats.
template <int C, bool FULL>
__device__ __forceinline__ void consume_row(const float *__restrict__ st, int n, const float *__restrict__ r_s,
                                            float ra, float (&Q)[NMAX], float (&W10)[NMAX], float &p, float &w6) {
#pragma unroll
    for (int c = 0; c < NMAX; ++c) {
        if (FULL || c < n) {
            const float t = st[c * C];
            Q[c] += t;                     // case 3 / 13 source: sum over a
            W10[c] = fmaf(ra, t, W10[c]);  // case 10: sum_a r[a] T[a,b,c]
            p += t;                        // cases 1,2,4,5,7,9,12,14 source: sum over c
            w6 = fmaf(r_s[c], t, w6);      // case 6: sum_c T[a,b,c] r[c]
        }
    }
}

template <int C>
__global__ void __launch_bounds__(kThreads, 2) k_fast_fwd_stream(Contract18Fwd a) {
    constexpr int TB = kThreads / C;
    extern __shared__ __align__(128) float smem[];
    float *ring = smem;
    float *r_s = ring + kStages * kStageFloats;
    float *red = r_s + NMAX;
    uint64_t *full = reinterpret_cast<uint64_t *>(red + 4 * kThreads);

    const int nm = a.b.n_max;
    const int tiles = tiles_of(nm, C);
    const int inst = blockIdx.x / tiles, tile = blockIdx.x - inst * tiles;
    const int n = a.b.n_of(inst);
    const int b0 = tile * TB;
    if (b0 >= n) return;
    const int tb = min(TB, n - b0);
    const int tid = threadIdx.x, f = tid % C, bl = tid / C, b = b0 + bl;
    const bool active = bl < tb;
    const AdjView av = adj_view(a.adjtab + (int64_t)inst * a.adjtab_words, nm);
    if (tid < NMAX) r_s[tid] = tid < n ? av.r[tid] : 0.f;
    const float sA = av.scal[0], tr = av.scal[1];
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    const uint32_t bytes = (uint32_t)(tb * n * C) * 4u;
    const int64_t row_off = (int64_t)b0 * n * C;
    if (tid == 0) {
        for (int s = 0; s < kStages && s < n; ++s) {
            mbar_arrive_expect_tx(&full[s], bytes);
            bulk_g2s(ring + s * kStageFloats, slab_ptr(a.T, inst, s, n, nm, C) + row_off, bytes, &full[s]);
        }
    }

    float Q[NMAX], W10[NMAX];
#pragma unroll
    for (int c = 0; c < NMAX; ++c) Q[c] = W10[c] = 0.f;
    float S4 = 0.f, S11 = 0.f, S15 = 0.f, t14 = 0.f, t18 = 0.f;

    float *outi = a.out + inst * a.stride_out;
    float *sc = a.scratch + inst * a.scratch_words;
    const FastFwdScratch S(nm, C);

    for (int s = 0; s < n; ++s) {  // s = T's first index a
        const int slot = s % kStages;
        mbar_wait(&full[slot], (uint32_t)(s / kStages) & 1u);
        if (active) {
            const float *st = ring + slot * kStageFloats + (bl * n) * C + f;
            float p = 0.f, w6 = 0.f;
            const float ra = r_s[s];
            if (n == NMAX)
                consume_row<C, true>(st, n, r_s, ra, Q, W10, p, w6);
            else
                consume_row<C, false>(st, n, r_s, ra, Q, W10, p, w6);
            const float d1 = st[b * C];  // T[a,b,b]
            const float d2 = st[s * C];  // T[a,b,a]
            const int64_t ab = (int64_t)s * n + b;
            float *o = outi + ab * (kSlabs * C) + f;
            o[0 * C] = sA * p;  // case 1 (RisiContraction_18.h:102)
            o[5 * C] = w6;      // case 6 (:133)
            o[6 * C] = tr * p;  // case 7 (:149)
            sc[ab * C + f] = p;
            sc[S.plane + ab * C + f] = d1;
            sc[2 * S.plane + ab * C + f] = d2;
            S4 += p;
            S11 += d2;
            S15 += d1;
            if (s == b) {
                t14 = p;
                t18 = d1;
            }
        }
        __syncthreads();  // every thread is done with `slot`
        if (tid == 0 && s + kStages < n) {
            mbar_arrive_expect_tx(&full[slot], bytes);
            bulk_g2s(ring + slot * kStageFloats, slab_ptr(a.T, inst, s + kStages, n, nm, C) + row_off, bytes,
                     &full[slot]);
        }
    }

    // Row-owned outputs.  The ring is idle now; each thread parks its Q row in a private column of it so the sparse
    // product can index it dynamically.
    if (active) {
        float *qs = ring + (bl * NMAX) * C + f;
#pragma unroll
        for (int c = 0; c < NMAX; ++c) {
            if (c < n) {
                float *o = outi + ((int64_t)b * n + c) * (kSlabs * C) + f;
                o[2 * C] = sA * Q[c];  // case 3 (:110)
                o[9 * C] = W10[c];     // case 10 (:195)
                qs[c * C] = Q[c];
            }
        }
        for (int d = 0; d < n; ++d) {
            float acc = 0.f;
            for (int j = av.rowptr[d]; j < av.rowptr[d + 1]; ++j) acc = fmaf(av.rowval[j], qs[av.rowidx[j] * C], acc);
            float *o = outi + ((int64_t)b * n + d) * (kSlabs * C) + f;
            const float rd = r_s[d];
            o[3 * C] = rd * S4;    // case 4 (:114)
            o[10 * C] = rd * S11;  // case 11 (:211)
            o[12 * C] = acc;       // case 13 (:241)
        }
    }
    red[0 * kThreads + tid] = active ? S4 : 0.f;   // -> total of T        (case 5)
    red[1 * kThreads + tid] = active ? t14 : 0.f;  // -> sum_a P[a,a]      (case 14)
    red[2 * kThreads + tid] = active ? S15 : 0.f;  // -> sum_{a,b} T[a,b,b] (case 15)
    red[3 * kThreads + tid] = active ? t18 : 0.f;  // -> sum_a T[a,a,a]    (case 18)
    __syncthreads();
    if (tid < C) {
        float *part = sc + S.partials + (int64_t)tile * 4 * C;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float v = 0.f;
            for (int j = 0; j < TB; ++j) v += red[k * kThreads + j * C + tid];
            part[k * C + tid] = v;
        }
    }
}

// Sparse row product: acc = sum_{j in list} val[j] * buf[idx[j]] with buf a thread-private shared-memory column.
__device__ __forceinline__ float sparse_dot(const int *__restrict__ ptr, const int *__restrict__ idx,
                                            const float *__restrict__ val, int row, const float *buf) {
    float acc = 0.f;
    for (int j = ptr[row]; j < ptr[row + 1]; ++j) acc = fmaf(val[j], buf[idx[j] * kThreads], acc);
    return acc;
}

template <int C>
__global__ void __launch_bounds__(kThreads) k_fast_fwd_finish(Contract18Fwd a) {
    constexpr int TB = kThreads / C;
    extern __shared__ __align__(128) float smem[];
    const int nm = a.b.n_max;
    const int tiles = tiles_of(nm, C);
    const int inst = blockIdx.x / tiles, tile = blockIdx.x - inst * tiles;
    const int n = a.b.n_of(inst);
    const int tid = threadIdx.x, f = tid % C, il = tid / C, i = tile * TB + il;
    if (i >= n) return;
    const AdjView av = adj_view(a.adjtab + (int64_t)inst * a.adjtab_words, nm);
    const FastFwdScratch S(nm, C);
    const float *sc = a.scratch + inst * a.scratch_words;
    const float *P = sc, *D1 = sc + S.plane, *D2 = sc + 2 * S.plane;
    float *buf = smem + tid;  // private column: buf[e * kThreads]
    float *orow = a.out + inst * a.stride_out + ((int64_t)i * n) * (kSlabs * C) + f;  // + d*18C + k*C

    float tot[4] = {0.f, 0.f, 0.f, 0.f};
    const int tiles_n = tiles_of(n, C);
    for (int t = 0; t < tiles_n; ++t) {
        const float *part = sc + S.partials + (int64_t)t * 4 * C;
#pragma unroll
        for (int k = 0; k < 4; ++k) tot[k] += part[k * C + f];
    }

    float s = 0.f;
    for (int e = 0; e < n; ++e) {  // row i of P
        const float v = P[((int64_t)i * n + e) * C + f];
        buf[e * kThreads] = v;
        s += v;
    }
    for (int d = 0; d < n; ++d) {
        float *o = orow + (int64_t)d * (kSlabs * C);
        o[1 * C] = av.r[d] * s;                                             // case 2 (:106)
        o[8 * C] = sparse_dot(av.rowptr, av.rowidx, av.rowval, d, buf);     // case 9 (:180)
    }
    s = 0.f;
    for (int e = 0; e < n; ++e) {  // row i of D1 = T[i,e,e]
        const float v = D1[((int64_t)i * n + e) * C + f];
        buf[e * kThreads] = v;
        s += v;
    }
    for (int d = 0; d < n; ++d) {
        float *o = orow + (int64_t)d * (kSlabs * C);
        o[7 * C] = av.r[d] * s;                                             // case 8 (:165)
        o[15 * C] = sparse_dot(av.rowptr, av.rowidx, av.rowval, d, buf);    // case 16 (:290)
    }
    for (int e = 0; e < n; ++e) buf[e * kThreads] = P[((int64_t)e * n + i) * C + f];  // column i of P
    for (int d = 0; d < n; ++d)
        orow[(int64_t)d * (kSlabs * C) + 11 * C] = sparse_dot(av.rowptr, av.rowidx, av.rowval, d, buf);  // case 12 (:226)
    for (int e = 0; e < n; ++e) buf[e * kThreads] = D2[((int64_t)e * n + i) * C + f];  // column i of D2 = T[e,i,e]
    for (int d = 0; d < n; ++d)
        orow[(int64_t)d * (kSlabs * C) + 16 * C] = sparse_dot(av.rowptr, av.rowidx, av.rowval, d, buf);  // case 17 (:304)
    for (int e = 0; e < n; ++e) {  // slabs indexed (d, e) = (i, e)
        const float w = av.A[i * n + e];
        float *o = orow + (int64_t)e * (kSlabs * C);
        o[4 * C] = w * tot[0];   // case 5 (:118)
        o[13 * C] = w * tot[1];  // case 14 (:256)
        o[14 * C] = w * tot[2];  // case 15 (:271)
        o[17 * C] = w * tot[3];  // case 18 (:318)
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Backward
// ---------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(kThreads) k_fast_bwd_planes(Contract18Bwd a) {
    constexpr int TB = kThreads / C;
    extern __shared__ __align__(128) float smem[];
    __shared__ float s_sh[4 * C];
    const int nm = a.b.n_max;
    const int tiles = tiles_of(nm, C);
    const int inst = blockIdx.x / tiles, tile = blockIdx.x - inst * tiles;
    const int n = a.b.n_of(inst);
    if (tile * TB >= n) return;
    const int tid = threadIdx.x, f = tid % C, il = tid / C, i = tile * TB + il;
    const AdjView av = adj_view(a.adjtab + (int64_t)inst * a.adjtab_words, nm);
    const float *g = a.gout + inst * a.stride_gout;
    const int64_t cell = (int64_t)kSlabs * C;

    // s5, s14, s15, s18: sum_{d,e} A[d,e] g_k[d,e]   (slabs indexed by (d,e)); split over the TB row-threads of f
    for (int k = il; k < 4; k += TB) {
        const int slab = (k == 0) ? 4 : (k == 1) ? 13 : (k == 2) ? 14 : 17;
        float acc = 0.f;
        for (int d = 0; d < n; ++d)
            for (int j = av.rowptr[d]; j < av.rowptr[d + 1]; ++j)
                acc = fmaf(av.rowval[j], g[((int64_t)d * n + av.rowidx[j]) * cell + slab * C + f], acc);
        s_sh[k * C + f] = acc;
    }
    __syncthreads();
    if (i >= n) return;
    const float s5 = s_sh[f], s14 = s_sh[C + f], s15 = s_sh[2 * C + f], s18 = s_sh[3 * C + f];
    const float sA = av.scal[0], tr = av.scal[1];
    const float *gi = g + ((int64_t)i * n) * cell + f;  // row i: gi[d*cell + k*C]
    float *buf = smem + tid;
    const FastBwdScratch S(nm, C);
    float *sc = a.scratch + inst * a.scratch_words;
    float *Ua = sc, *UbT = sc + S.plane, *E1 = sc + 2 * S.plane, *E2 = sc + 3 * S.plane, *V = sc + 4 * S.plane;
    const int64_t rowi = ((int64_t)i * n) * C + f;

    float u2 = 0.f, u4 = 0.f, u8 = 0.f, u11 = 0.f;
    for (int d = 0; d < n; ++d) {
        const float rd = av.r[d];
        const float *gd = gi + (int64_t)d * cell;
        u2 = fmaf(rd, gd[1 * C], u2);     // case 2
        u4 = fmaf(rd, gd[3 * C], u4);     // case 4
        u8 = fmaf(rd, gd[7 * C], u8);     // case 8
        u11 = fmaf(rd, gd[10 * C], u11);  // case 11
        buf[d * kThreads] = gd[8 * C];    // g9[i, d]
    }
    for (int b = 0; b < n; ++b) {  // U_a[a=i, b]
        const float *gb = gi + (int64_t)b * cell;
        float v = sA * gb[0] + tr * gb[6 * C] + sparse_dot(av.colptr, av.colidx, av.colval, b, buf) + u2 + s5;
        if (b == i) v += s14;
        Ua[rowi + (int64_t)b * C] = v;
    }
    for (int d = 0; d < n; ++d) buf[d * kThreads] = gi[(int64_t)d * cell + 15 * C];  // g16[i, d]
    for (int b = 0; b < n; ++b) {  // E1[a=i, b]
        float v = u8 + s15 + sparse_dot(av.colptr, av.colidx, av.colval, b, buf);
        if (b == i) v += s18;
        E1[rowi + (int64_t)b * C] = v;
    }
    for (int d = 0; d < n; ++d) buf[d * kThreads] = gi[(int64_t)d * cell + 11 * C];  // g12[i, d]
    for (int s = 0; s < n; ++s)  // U_b[a=s, b=i], stored transposed
        UbT[rowi + (int64_t)s * C] = u4 + sparse_dot(av.colptr, av.colidx, av.colval, s, buf);
    for (int d = 0; d < n; ++d) buf[d * kThreads] = gi[(int64_t)d * cell + 16 * C];  // g17[i, d]
    for (int s = 0; s < n; ++s)  // E2[b=i, a=s]
        E2[rowi + (int64_t)s * C] = u11 + sparse_dot(av.colptr, av.colidx, av.colval, s, buf);
    for (int d = 0; d < n; ++d) buf[d * kThreads] = gi[(int64_t)d * cell + 12 * C];  // g13[i, d]
    for (int c = 0; c < n; ++c)  // V[b=i, c]
        V[rowi + (int64_t)c * C] = sA * gi[(int64_t)c * cell + 2 * C] + sparse_dot(av.colptr, av.colidx, av.colval, c, buf);
}

template <int C, bool ACCUM, bool FULL>
__device__ __forceinline__ void emit_row(float *__restrict__ dst, int n, int b, int s, float ua, float g6, float ra,
                                         float e1, float e2, float beta, const float *__restrict__ r_s,
                                         const float (&V)[NMAX], const float (&G10)[NMAX]) {
#pragma unroll
    for (int c = 0; c < NMAX; ++c) {
        if (FULL || c < n) {
            float v = ua + V[c];
            v = fmaf(g6, r_s[c], v);
            v = fmaf(ra, G10[c], v);
            if (c == b) v += e1;
            if (c == s) v += e2;
            if (ACCUM) v = fmaf(beta, dst[c * C], v);
            __stcs(dst + c * C, v);
        }
    }
}

template <int C, bool ACCUM>
__global__ void __launch_bounds__(kThreads, 2) k_fast_bwd_stream(Contract18Bwd a) {
    constexpr int TB = kThreads / C;
    __shared__ float r_s[NMAX];
    const int nm = a.b.n_max;
    const int tiles = tiles_of(nm, C);
    const int inst = blockIdx.x / tiles, tile = blockIdx.x - inst * tiles;
    const int n = a.b.n_of(inst);
    if (tile * TB >= n) return;
    const int tid = threadIdx.x, f = tid % C, bl = tid / C, b = tile * TB + bl;
    const AdjView av = adj_view(a.adjtab + (int64_t)inst * a.adjtab_words, nm);
    if (tid < NMAX) r_s[tid] = tid < n ? av.r[tid] : 0.f;
    __syncthreads();
    if (b >= n) return;
    const float *g = a.gout + inst * a.stride_gout;
    const int64_t cell = (int64_t)kSlabs * C;
    const FastBwdScratch S(nm, C);
    const float *sc = a.scratch + inst * a.scratch_words;
    const float *Ua = sc, *UbT = sc + S.plane, *E1 = sc + 2 * S.plane, *E2 = sc + 3 * S.plane, *Vp = sc + 4 * S.plane;

    float V[NMAX], G10[NMAX];
#pragma unroll
    for (int c = 0; c < NMAX; ++c) {
        const bool ok = c < n;
        V[c] = ok ? Vp[((int64_t)b * n + c) * C + f] : 0.f;
        G10[c] = ok ? g[((int64_t)b * n + c) * cell + 9 * C + f] : 0.f;  // case 10
    }
    const int64_t rowb = ((int64_t)b * n) * C + f;  // [b, s] planes
    // software pipeline: values of step s+1 are fetched while step s is written out
    float ua = Ua[(int64_t)b * C + f] + UbT[rowb];
    float g6 = g[(int64_t)b * cell + 5 * C + f];
    float e1 = E1[(int64_t)b * C + f];
    float e2 = E2[rowb];
    for (int s = 0; s < n; ++s) {
        float ua_n = 0.f, g6_n = 0.f, e1_n = 0.f, e2_n = 0.f;
        if (s + 1 < n) {
            const int64_t ab = ((int64_t)(s + 1) * n + b);
            ua_n = Ua[ab * C + f] + UbT[rowb + (int64_t)(s + 1) * C];
            g6_n = g[ab * cell + 5 * C + f];  // case 6
            e1_n = E1[ab * C + f];
            e2_n = E2[rowb + (int64_t)(s + 1) * C];
        }
        float *dst = slab_ptr(a.gT, inst, s, n, nm, C) + ((int64_t)b * n) * C + f;
        if (n == NMAX)
            emit_row<C, ACCUM, true>(dst, n, b, s, ua, g6, r_s[s], e1, e2, a.beta, r_s, V, G10);
        else
            emit_row<C, ACCUM, false>(dst, n, b, s, ua, g6, r_s[s], e1, e2, a.beta, r_s, V, G10);
        ua = ua_n;
        g6 = g6_n;
        e1 = e1_n;
        e2 = e2_n;
    }
}

template <int C>
cudaError_t configure_for() {
    cudaError_t e = cudaFuncSetAttribute(k_fast_fwd_stream<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kFwdStreamSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_fast_fwd_finish<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowBufSmem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_fast_bwd_planes<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowBufSmem);
}

template <int C>
cudaError_t forward_for(const Contract18Fwd &a, cudaStream_t st, LaunchLog *log) {
    const unsigned grid = (unsigned)(a.b.count * tiles_of(a.b.n_max, C));
    CCN_LAUNCH(log, K_FWD_STREAM, st, k_fast_fwd_stream<C><<<grid, kThreads, kFwdStreamSmem, st>>>(a));
    CCN_LAUNCH(log, K_FWD_FINISH, st, k_fast_fwd_finish<C><<<grid, kThreads, kRowBufSmem, st>>>(a));
    return cudaGetLastError();
}

template <int C>
cudaError_t backward_for(const Contract18Bwd &a, cudaStream_t st, LaunchLog *log) {
    const unsigned grid = (unsigned)(a.b.count * tiles_of(a.b.n_max, C));
    CCN_LAUNCH(log, K_BWD_PLANES, st, k_fast_bwd_planes<C><<<grid, kThreads, kRowBufSmem, st>>>(a));
    if (a.beta != 0.f)
        CCN_LAUNCH(log, K_BWD_STREAM, st, (k_fast_bwd_stream<C, true><<<grid, kThreads, 0, st>>>(a)));
    else
        CCN_LAUNCH(log, K_BWD_STREAM, st, (k_fast_bwd_stream<C, false><<<grid, kThreads, 0, st>>>(a)));
    return cudaGetLastError();
}

}  // namespace

bool fast_path_supported(int n_max, int C) { return n_max >= 1 && n_max <= NMAX && (C == 32 || C == 64 || C == 128); }
int64_t fast_fwd_scratch_words(int n_max, int C) { return FastFwdScratch(n_max, C).words; }
int64_t fast_bwd_scratch_words(int n_max, int C) { return FastBwdScratch(n_max, C).words; }

cudaError_t fast_path_configure() {
    cudaError_t e = configure_for<32>();
    if (e != cudaSuccess) return e;
    e = configure_for<64>();
    if (e != cudaSuccess) return e;
    return configure_for<128>();
}

cudaError_t launch_fast_forward(const Contract18Fwd &a, cudaStream_t st, LaunchLog *log) {
    switch (a.b.C) {
        case 32: return forward_for<32>(a, st, log);
        case 64: return forward_for<64>(a, st, log);
        case 128: return forward_for<128>(a, st, log);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_fast_backward(const Contract18Bwd &a, cudaStream_t st, LaunchLog *log) {
    switch (a.b.C) {
        case 32: return backward_for<32>(a, st, log);
        case 64: return backward_for<64>(a, st, log);
        case 128: return backward_for<128>(a, st, log);
    }
    return cudaErrorInvalidValue;
}

}  // namespace ccn
