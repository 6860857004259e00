// contract18_fused.cu -- single-kernel StackTensor3D + RisiContraction_18 forward and backward for the benchmark
// shapes (n <= 32, C in {32, 64, 128}); sm_100a only.
//
// Replaces GraphFlow/StackTensor3D.h:54-90 + GraphFlow/RisiContraction_18.h:73-560 and the reference kernels
// GraphFlow_gpu/RisiContraction_18_gpu.h:49-379 (forward_job) / :541-685 (backward_job).
//
// Design.  Work unit = one *row tile*: (instance, TB consecutive values of T's middle index b),
// TB = 256 / C, one 256-thread CTA, thread (f, bl) owns channel f of row b = b0 + bl for the whole kernel.  The
// tiles of one instance (its *siblings*) are handed out back to back through a ticket counter, so they run at the
// same time on different SMs; they exchange their few N^2-sized intermediates through an L2-resident scratch slot
// and one arrival counter per instance, and every byte of `out` / `gT` is written exactly once, by the tile that
// owns its row.  DRAM sees T (or gout) once and out (or gT) once.
//
//   k_fwd_fused   stream:  for a = 0..n-1 the contiguous chunk T[a, b0:b0+TB, :, :] (TB*n*C floats, 32 KiB at n=32) is
//                          pulled by one cp.async.bulk (TMA engine, L2 evict-first) into a 3-stage shared-memory ring
//                          guarded by mbarriers.  Each thread walks its row's n cells once: Q[b,c] += t and
//                          W10[b,c] += r[a] t stay in registers across the a loop (sum over a); P[a,b] = sum_c t,
//                          W6[a,b] = sum_c r[c] t, D1 = T[a,b,b], D2 = T[a,b,a] finish inside the step and are parked
//                          in the instance's scratch slot.  No cross-thread reduction anywhere.
//                 pass A:  the seven slabs that only need the tile's own rows (cases 3,4,10,11,12,13,17) -> out[b, :].
//                 pass B:  after all siblings have arrived: the eleven slabs that need whole rows of P / D1 / W6 or the
//                          instance totals (cases 1,2,5,6,7,8,9,14,15,16,18) -> out[x, :] for the tile's rows x.
//   k_bwd_fused   phase 1: from the tile's own rows of gout: u2, u8 and the partial sums s5,s14,s15,s18 (published to
//                          the scratch slot), V[b,:] and g10[b,:] (registers), and the per-a coefficients U, E2 (shared
//                          memory planes); after the siblings have arrived: the remaining terms of U and E1.
//                 phase 2: gT[a,b,c] = U[a,b] + V[b,c] + g6[a,b] r[c] + r[a] g10[b,c] + [c==b] E1[a,b] + [c==a] E2[b,a]:
//                          pure broadcast, 128-byte coalesced streaming stores, no global loads except g6.
//
// The five [n x n] * A^T products per channel (cases 9, 12, 13, 16, 17) run over per-row / per-column lists of the
// non-zeros of A built in shared memory in the CTA prologue: molecular adjacency has ~3 non-zeros per row, so they
// cost ~10x less than a dense product and nothing is gained by reshaping them for the tensor cores; a dense A takes
// the same code path (longer lists).
#include "contract18_kernels.cuh"

namespace ccn {

namespace {

constexpr int NMAX = 32;       // largest receptive field handled by the register-resident row accumulators
constexpr int kThreads = 256;  // one thread per (channel, row-in-tile)
constexpr int kStages = 3;     // TMA ring depth (3 x 32 KiB in flight per CTA, 2 CTAs per SM)
constexpr int kStageFloats = kThreads * NMAX;  // TB * NMAX * C
constexpr int kColFloats = kThreads * NMAX;    // one thread-private column: col[e * kThreads + tid]
constexpr long long kSpinLimit = 4000000000ll; // ~2 s of SM clocks: a sibling that never arrives is a bug, not a wait

__host__ __device__ inline int tiles_of(int n, int C) { return (n + (kThreads / C) - 1) / (kThreads / C); }

// control block (ints): [0] ticket, [1] error flag, [2..3] unused, then per slot {arrive, finish, gen_done, pad}
__host__ __device__ inline int ctl_words(int slots) { return 4 + 4 * slots; }

struct FwdScratch {  // per slot: planes P, W6, D1, D2 [n*n*C] then per-tile partial totals [tiles][4][C]
    int64_t plane, partials, words;
    __host__ __device__ FwdScratch(int nm, int C) {
        plane = (int64_t)nm * nm * C;
        partials = 4 * plane;
        words = (partials + (int64_t)tiles_of(nm, C) * 4 * C + 31) & ~(int64_t)31;
    }
};

struct BwdScratch {  // per slot: u2, u8 [n][C] then per-tile partial sums s5,s14,s15,s18 [tiles][4][C]
    int64_t vec, partials, words;
    __host__ __device__ BwdScratch(int nm, int C) {
        vec = (int64_t)nm * C;
        partials = 2 * vec;
        words = (partials + (int64_t)tiles_of(nm, C) * 4 * C + 31) & ~(int64_t)31;
    }
};

// Adjacency of one instance in shared memory: dense effective A (row stride n), row sums, sA, tr and the non-zeros
// of every row (forward) or column (backward) as lists stored entry-major: entry j of list l is val[j*NMAX + l],
// idx[j*NMAX + l], so the j-th entries of eight consecutive lists are one 32-byte / one 8-byte shared-memory load.
struct AdjShared {
    float A[NMAX * NMAX];
    float val[NMAX * NMAX];
    unsigned char idx[NMAX * NMAX];  // unused tail entries are 0 (a valid index) and never contribute
    int cnt[NMAX];
    float r[NMAX];
    float sA, tr;
    int maxcnt;
    int pad;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// BY_COLUMN = false: list d holds (e, A[d,e]) of row d.  BY_COLUMN = true: list b holds (d, A[d,b]) of column b.
template <bool BY_COLUMN>
__device__ __forceinline__ void build_adjacency(AdjShared &S, const float *__restrict__ adj, int n, bool positive_part) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < n * n; i += kThreads) {
        float v = adj[i];
        if (positive_part && !(v > 0.0f)) v = 0.0f;  // RisiContraction_18.h:90 `if (adj_value > 0)`
        S.A[i] = v;
    }
    for (int i = tid; i < NMAX * NMAX / 4; i += kThreads) reinterpret_cast<uint32_t *>(S.idx)[i] = 0u;
    if (tid < NMAX) S.cnt[tid] = 0;
    __syncthreads();
    for (int l = warp; l < n; l += kThreads / 32) {
        const float row_v = lane < n ? S.A[l * n + lane] : 0.0f;
        const float v = BY_COLUMN ? (lane < n ? S.A[lane * n + l] : 0.0f) : row_v;
        const bool nz = v != 0.0f;
        const unsigned m = __ballot_sync(0xffffffffu, nz);
        if (nz) {
            const int j = __popc(m & ((1u << lane) - 1u));
            S.val[j * NMAX + l] = v;
            S.idx[j * NMAX + l] = (unsigned char)lane;
        }
        const float rs = warp_sum(row_v);
        if (lane == 0) {
            S.cnt[l] = __popc(m);
            S.r[l] = rs;
        }
    }
    for (int l = n + tid; l < NMAX; l += kThreads) S.r[l] = 0.0f;
    __syncthreads();
    if (warp == 0) {
        const float sA = warp_sum(lane < n ? S.r[lane] : 0.0f);
        const float tr = warp_sum(lane < n ? S.A[lane * n + lane] : 0.0f);
        int mc = S.cnt[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mc = max(mc, __shfl_xor_sync(0xffffffffu, mc, o));
        if (lane == 0) {
            S.sA = sA;
            S.tr = tr;
            S.maxcnt = mc;
        }
    }
    __syncthreads();
}

// Eight sparse dot products at once, against NC thread-private shared-memory columns (stride kThreads):
//   acc[q][k] = sum over entries (i, w) of list l0+k of  w * cols[q][i]        l0 a multiple of 8.
// The lists are walked entry-position by entry-position, so the eight (x NC) dependent index -> value chains run
// in parallel instead of one after the other.
template <int NC>
__device__ __forceinline__ void list_dot8(const AdjShared &S, int l0, const float *const (&cols)[NC], float (&acc)[NC][8]) {
    int cnt[8];
    {
        const int4 c0 = *reinterpret_cast<const int4 *>(S.cnt + l0), c1 = *reinterpret_cast<const int4 *>(S.cnt + l0 + 4);
        cnt[0] = c0.x, cnt[1] = c0.y, cnt[2] = c0.z, cnt[3] = c0.w, cnt[4] = c1.x, cnt[5] = c1.y, cnt[6] = c1.z, cnt[7] = c1.w;
    }
#pragma unroll
    for (int q = 0; q < NC; ++q)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[q][k] = 0.f;
    int m = max(max(max(cnt[0], cnt[1]), max(cnt[2], cnt[3])), max(max(cnt[4], cnt[5]), max(cnt[6], cnt[7])));
    for (int j = 0; j < m; ++j) {
        const float4 w0 = *reinterpret_cast<const float4 *>(S.val + j * NMAX + l0);
        const float4 w1 = *reinterpret_cast<const float4 *>(S.val + j * NMAX + l0 + 4);
        const uint2 id = *reinterpret_cast<const uint2 *>(S.idx + j * NMAX + l0);
        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const unsigned word = k < 4 ? id.x : id.y;
            const int e = (word >> (8 * (k & 3))) & 0xffu;
            const bool on = j < cnt[k];
#pragma unroll
            for (int q = 0; q < NC; ++q) {
                const float x = cols[q][e * kThreads];
                acc[q][k] = on ? fmaf(w[k], x, acc[q][k]) : acc[q][k];
            }
        }
    }
}

__device__ __forceinline__ float *slab_ptr(const TensorRef &t, int64_t inst, int a, int n, int n_max, int C) {
    return t.slabs ? t.slabs[inst * n_max + a] : t.base + inst * t.stride + (int64_t)a * n * n * C;
}

__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---- sibling protocol ------------------------------------------------------------------------------------------
// Tickets make the start order of the tiles equal to their work order, so a tile only ever waits for tiles that
// have already started (its siblings' later tickets are taken by the CTAs that retire next; at least 2 * #SMs
// CTAs are resident, far more than the <= 16 tiles of one instance).
struct Slot {
    int *arrive, *finish, *gen_done, *error;
};

__device__ __forceinline__ Slot slot_of(int *ctl, int slot) { return Slot{ctl + 4 + 4 * slot, ctl + 5 + 4 * slot, ctl + 6 + 4 * slot, ctl + 1}; }

__device__ __forceinline__ void spin_until(const int *p, int target, int *error) {
    const long long t0 = clock64();
    while (ld_acquire(p) < target) {
        __nanosleep(100);
        if (clock64() - t0 > kSpinLimit) {
            atomicExch(error, 1);
            break;
        }
    }
}

// all threads call; returns after the slot's previous generation has been fully consumed
__device__ __forceinline__ void slot_acquire(const Slot &s, int gen) {
    if (gen > 0) {
        if (threadIdx.x == 0) spin_until(s.gen_done, gen, s.error);
        __syncthreads();
    }
}
// all threads call after their scratch stores
__device__ __forceinline__ void slot_publish(const Slot &s) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(s.arrive, 1);
}
// all threads call before reading the siblings' scratch
__device__ __forceinline__ void slot_wait_siblings(const Slot &s, int tiles_n) {
    if (threadIdx.x == 0) {
        spin_until(s.arrive, tiles_n, s.error);
        __threadfence();
    }
    __syncthreads();
}
// all threads call after their last scratch read
__device__ __forceinline__ void slot_release(const Slot &s, int tiles_n) {
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(s.finish, 1) == tiles_n - 1) {
            *s.arrive = 0;
            *s.finish = 0;
            __threadfence();
            atomicAdd(s.gen_done, 1);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Forward
// ---------------------------------------------------------------------------------------------------------------
struct FwdSmem {
    AdjShared adj;
    uint64_t full[kStages];
    int work;
};
constexpr size_t kFwdSmem = (size_t)kStages * kStageFloats * 4 + sizeof(FwdSmem);

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
__global__ void __launch_bounds__(kThreads, 2) k_fwd_fused(Fused18Fwd a) {
    constexpr int TB = kThreads / C;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *ring = reinterpret_cast<float *>(smem_raw);
    FwdSmem &S = *reinterpret_cast<FwdSmem *>(smem_raw + (size_t)kStages * kStageFloats * 4);

    const int tid = threadIdx.x;
    if (tid == 0) S.work = atomicAdd(a.ctl, 1);
    __syncthreads();
    const int nm = a.b.n_max;
    const int tiles = tiles_of(nm, C);
    const int64_t inst = S.work / tiles;
    const int tile = S.work - (int)inst * tiles;
    const int n = a.b.n_of((int)inst);
    const int b0 = tile * TB;
    if (b0 >= n) return;
    const int tiles_n = tiles_of(n, C);
    const int tb = min(TB, n - b0);
    const int f = tid % C, bl = tid / C, b = b0 + bl;
    const bool active = bl < tb;
    const Slot slot = slot_of(a.ctl, (int)(inst % a.slots));

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&S.full[s], 1);
        fence_mbar_init();
    }
    build_adjacency<false>(S.adj, a.adj + inst * a.stride_adj, n, a.positive_part != 0);  // ends with __syncthreads

    const uint32_t bytes = (uint32_t)(tb * n * C) * 4u;
    const int64_t row_off = (int64_t)b0 * n * C;
    uint64_t policy = 0;
    if (tid == 0) {
        policy = l2_evict_first_policy();
        for (int s = 0; s < kStages && s < n; ++s) {
            mbar_arrive_expect_tx(&S.full[s], bytes);
            bulk_g2s_hint(ring + s * kStageFloats, slab_ptr(a.T, inst, s, n, nm, C) + row_off, bytes, &S.full[s], policy);
        }
    }
    slot_acquire(slot, (int)(inst / a.slots));

    const float *r_s = S.adj.r;
    const float sA = S.adj.sA, tr = S.adj.tr;
    const FwdScratch L(n, C);
    float *sc = a.scratch + (inst % a.slots) * a.scratch_words;
    float *Pp = sc, *W6p = sc + L.plane, *D1p = sc + 2 * L.plane, *D2p = sc + 3 * L.plane;

    float Q[NMAX], W10[NMAX];
#pragma unroll
    for (int c = 0; c < NMAX; ++c) Q[c] = W10[c] = 0.f;
    float S4 = 0.f, S11 = 0.f, S15 = 0.f, t14 = 0.f, t18 = 0.f;

    for (int s = 0; s < n; ++s) {  // s = T's first index a
        const int st_i = s % kStages;
        mbar_wait(&S.full[st_i], (uint32_t)(s / kStages) & 1u);
        if (active) {
            const float *st = ring + st_i * kStageFloats + (bl * n) * C + f;
            float p = 0.f, w6 = 0.f;
            const float ra = r_s[s];
            if (n == NMAX)
                consume_row<C, true>(st, n, r_s, ra, Q, W10, p, w6);
            else
                consume_row<C, false>(st, n, r_s, ra, Q, W10, p, w6);
            const float d1 = st[b * C];  // T[a,b,b]
            const float d2 = st[s * C];  // T[a,b,a]
            const int64_t ab = ((int64_t)s * n + b) * C + f;
            Pp[ab] = p;
            W6p[ab] = w6;
            D1p[ab] = d1;
            D2p[ab] = d2;
            S4 += p;
            S11 += d2;
            S15 += d1;
            if (s == b) {
                t14 = p;
                t18 = d1;
            }
        }
        __syncthreads();  // every thread is done with this stage
        if (tid == 0 && s + kStages < n) {
            mbar_arrive_expect_tx(&S.full[st_i], bytes);
            bulk_g2s_hint(ring + st_i * kStageFloats, slab_ptr(a.T, inst, s + kStages, n, nm, C) + row_off, bytes,
                          &S.full[st_i], policy);
        }
    }

    // per-tile partial totals (the ring is idle from here on and is reused as plain shared memory)
    ring[0 * kThreads + tid] = active ? S4 : 0.f;   // -> total of T            (case 5)
    ring[1 * kThreads + tid] = active ? t14 : 0.f;  // -> sum_a P[a,a]          (case 14)
    ring[2 * kThreads + tid] = active ? S15 : 0.f;  // -> sum_{a,b} T[a,b,b]    (case 15)
    ring[3 * kThreads + tid] = active ? t18 : 0.f;  // -> sum_a T[a,a,a]        (case 18)
    __syncthreads();
    if (tid < C) {
        float *part = sc + L.partials + (int64_t)tile * 4 * C;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float v = 0.f;
            for (int j = 0; j < TB; ++j) v += ring[k * kThreads + j * C + tid];
            part[k * C + tid] = v;
        }
    }
    slot_publish(slot);  // fence + barrier + arrive: also orders the reads of ring[] above before the writes below

    float *outi = a.out + inst * a.stride_out;
    const int64_t cell = (int64_t)kSlabs * C;
    float *col0 = ring + tid, *col1 = col0 + kColFloats, *col2 = col1 + kColFloats;

    // ---- pass A: slabs that only need this tile's rows ------------------------------------------------------------
    if (active) {
        float *orow = outi + ((int64_t)b * n) * cell + f;  // + y*cell + k*C
#pragma unroll
        for (int c = 0; c < NMAX; ++c) {
            if (c < n) {
                __stcs(orow + c * cell + 2 * C, sA * Q[c]);  // case 3  (RisiContraction_18.h:110)
                __stcs(orow + c * cell + 9 * C, W10[c]);     // case 10 (:195)
                col0[c * kThreads] = Q[c];
            }
        }
        for (int e = 0; e < n; ++e) {  // columns b of P and D2: this thread's own stores
            const int64_t eb = ((int64_t)e * n + b) * C + f;
            col1[e * kThreads] = __ldcg(Pp + eb);
            col2[e * kThreads] = __ldcg(D2p + eb);
        }
        const float *const colsA[3] = {col1, col0, col2};
        for (int d0 = 0; d0 < n; d0 += 8) {
            float acc[3][8];
            list_dot8<3>(S.adj, d0, colsA, acc);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int d = d0 + k;
                if (d < n) {
                    float *o = orow + d * cell;
                    const float rd = r_s[d];
                    __stcs(o + 3 * C, rd * S4);      // case 4  (:114)
                    __stcs(o + 10 * C, rd * S11);    // case 11 (:211)
                    __stcs(o + 11 * C, acc[0][k]);   // case 12 (:226)  sum_e A[d,e] P[e,b]
                    __stcs(o + 12 * C, acc[1][k]);   // case 13 (:241)  sum_e A[d,e] Q[b,e]
                    __stcs(o + 16 * C, acc[2][k]);   // case 17 (:304)  sum_e A[d,e] T[e,b,e]
                }
            }
        }
    }

    // ---- pass B: slabs that need the siblings' rows ----------------------------------------------------------------
    slot_wait_siblings(slot, tiles_n);
    if (active) {
        const int x = b;
        float tot[4] = {0.f, 0.f, 0.f, 0.f};
        for (int t = 0; t < tiles_n; ++t) {
            const float *part = sc + L.partials + (int64_t)t * 4 * C;
#pragma unroll
            for (int k = 0; k < 4; ++k) tot[k] += __ldcg(part + k * C + f);
        }
        const int64_t xrow = ((int64_t)x * n) * C + f;
        float s2 = 0.f, s8 = 0.f;
        for (int e = 0; e < n; ++e) {
            const float pv = __ldcg(Pp + xrow + (int64_t)e * C);
            const float dv = __ldcg(D1p + xrow + (int64_t)e * C);
            col0[e * kThreads] = pv;
            col1[e * kThreads] = dv;
            s2 += pv;
            s8 += dv;
        }
        float *orow = outi + ((int64_t)x * n) * cell + f;
        const float *const colsB[2] = {col0, col1};
        for (int d0 = 0; d0 < n; d0 += 8) {
            float w6[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) w6[k] = (d0 + k < n) ? __ldcg(W6p + xrow + (int64_t)(d0 + k) * C) : 0.f;
            float acc[2][8];
            list_dot8<2>(S.adj, d0, colsB, acc);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int d = d0 + k;
                if (d < n) {
                    float *o = orow + d * cell;
                    const float pv = col0[d * kThreads];
                    const float rd = r_s[d];
                    const float axd = S.adj.A[x * n + d];
                    __stcs(o + 0 * C, sA * pv);        // case 1  (:102)
                    __stcs(o + 1 * C, rd * s2);        // case 2  (:106)
                    __stcs(o + 4 * C, axd * tot[0]);   // case 5  (:118)
                    __stcs(o + 5 * C, w6[k]);          // case 6  (:133)
                    __stcs(o + 6 * C, tr * pv);        // case 7  (:149)
                    __stcs(o + 7 * C, rd * s8);        // case 8  (:165)
                    __stcs(o + 8 * C, acc[0][k]);      // case 9  (:180)  sum_e A[d,e] P[x,e]
                    __stcs(o + 13 * C, axd * tot[1]);  // case 14 (:256)
                    __stcs(o + 14 * C, axd * tot[2]);  // case 15 (:271)
                    __stcs(o + 15 * C, acc[1][k]);     // case 16 (:290)  sum_e A[d,e] T[x,e,e]
                    __stcs(o + 17 * C, axd * tot[3]);  // case 18 (:318)
                }
            }
        }
    }
    slot_release(slot, tiles_n);
}

// ---------------------------------------------------------------------------------------------------------------
// Backward.  Formulas: SURVEY.md section 8(a) "Backward" (exact transpose of RisiContraction_18.h:333-560).
// ---------------------------------------------------------------------------------------------------------------
struct BwdSmem {
    AdjShared adj;
    float red[4 * kThreads];
    int work;
};
constexpr int kBlk = 16;  // rows whose cells are fetched together in the column passes of the backward
constexpr size_t kBwdSmem = (size_t)3 * kColFloats * 4 + sizeof(BwdSmem);

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

// One sweep over row b of gout: kCells cells at a time, all loads of a batch issued before any is consumed.
//   acc: u2, u4, u8, u11 (x r[d]) and s5, s14, s15, s18 (x A[b,d]);  V[d] = sA g3[b,d];  columns c13, c12, c17.
template <int C, bool FULL>
__device__ __forceinline__ void sweep_own_row(const float *__restrict__ grow, int n, const float *__restrict__ r_s,
                                              const float *__restrict__ Arow, float sA, float *c13, float *c12,
                                              float *c17, float (&V)[NMAX], float (&acc)[8]) {
    constexpr int kCells = 4;
    constexpr int kSl = 12;
    constexpr int slab[kSl] = {1, 3, 7, 10, 4, 13, 14, 17, 2, 12, 11, 16};  // 0-based slab index k-1 of case k
    const int64_t cell = (int64_t)kSlabs * C;
#pragma unroll
    for (int d0 = 0; d0 < NMAX; d0 += kCells) {
        float t[kCells][kSl];
#pragma unroll
        for (int i = 0; i < kCells; ++i)
#pragma unroll
            for (int k = 0; k < kSl; ++k)
                t[i][k] = (FULL || d0 + i < n) ? ld_stream(grow + (d0 + i) * cell + slab[k] * C) : 0.f;
#pragma unroll
        for (int i = 0; i < kCells; ++i) {
            const int d = d0 + i;
            const bool ok = FULL || d < n;
            const float rd = ok ? r_s[d] : 0.f;
            const float w = ok ? Arow[d] : 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] = fmaf(rd, t[i][k], acc[k]);          // cases 2, 4, 8, 11
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[4 + k] = fmaf(w, t[i][4 + k], acc[4 + k]);  // cases 5, 14, 15, 18
            V[d] = sA * t[i][8];  // case 3 (the A^T g13 term is added by the caller)
            if (ok) {
                c13[d * kThreads] = t[i][9];   // case 13
                c12[d * kThreads] = t[i][10];  // case 12
                c17[d * kThreads] = t[i][11];  // case 17
            }
        }
    }
}

template <int C, bool ACCUM>
__global__ void __launch_bounds__(kThreads, 2) k_bwd_fused(Fused18Bwd a) {
    constexpr int TB = kThreads / C;
    constexpr int kG6Ahead = 4;  // register prefetch distance for g6[a,b]
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *planes = reinterpret_cast<float *>(smem_raw);
    BwdSmem &S = *reinterpret_cast<BwdSmem *>(smem_raw + (size_t)3 * kColFloats * 4);

    const int tid = threadIdx.x;
    if (tid == 0) S.work = atomicAdd(a.ctl, 1);
    __syncthreads();
    const int nm = a.b.n_max;
    const int tiles = tiles_of(nm, C);
    const int64_t inst = S.work / tiles;
    const int tile = S.work - (int)inst * tiles;
    const int n = a.b.n_of((int)inst);
    const int b0 = tile * TB;
    if (b0 >= n) return;
    const int tiles_n = tiles_of(n, C);
    const int f = tid % C, bl = tid / C, b = b0 + bl;
    const bool active = b < n;
    const Slot slot = slot_of(a.ctl, (int)(inst % a.slots));

    build_adjacency<true>(S.adj, a.adj + inst * a.stride_adj, n, a.positive_part != 0);
    slot_acquire(slot, (int)(inst / a.slots));

    const float *r_s = S.adj.r;
    const float sA = S.adj.sA, tr = S.adj.tr;
    const BwdScratch L(n, C);
    float *sc = a.scratch + (inst % a.slots) * a.scratch_words;
    float *u2p = sc, *u8p = sc + L.vec;
    const float *g = a.gout + inst * a.stride_gout;
    const int64_t cell = (int64_t)kSlabs * C;

    // ---- phase 1a: one sweep over the tile's own rows of gout (12 of the 18 slabs of every cell (b, d)) -------------
    // Staging columns (thread-private, stride kThreads): c13 = g13[b,:], c12 = g12[b,:], c17 = g17[b,:].  The three
    // shared-memory planes are reused as soon as their staging content is dead:
    //   region 0: c13 -> E2 plane      region 1: c12 -> E1 plane      region 2: c17 -> U plane
    float *reg0 = planes + tid, *reg1 = reg0 + kColFloats, *reg2 = reg1 + kColFloats;
    float *E2s = reg0, *E1s = reg1, *Us = reg2;  // [a * kThreads]
    const float *grow = g + ((int64_t)(active ? b : b0) * n) * cell + f;  // row b: grow[d*cell + k*C]
    float V[NMAX];
    float u4 = 0.f, u11 = 0.f;
    {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // u2, u4, u8, u11, s5, s14, s15, s18
        if (active) {
            if (n == NMAX)
                sweep_own_row<C, true>(grow, n, r_s, S.adj.A + b * n, sA, reg0, reg1, reg2, V, acc);
            else
                sweep_own_row<C, false>(grow, n, r_s, S.adj.A + b * n, sA, reg0, reg1, reg2, V, acc);
            u2p[(int64_t)b * C + f] = acc[0];
            u8p[(int64_t)b * C + f] = acc[2];
        } else {
#pragma unroll
            for (int c = 0; c < NMAX; ++c) V[c] = 0.f;
        }
        u4 = acc[1];
        u11 = acc[3];
        float *red = S.red;
#pragma unroll
        for (int k = 0; k < 4; ++k) red[k * kThreads + tid] = acc[4 + k];
        __syncthreads();
        if (tid < C) {
            float *part = sc + L.partials + (int64_t)tile * 4 * C;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float v = 0.f;
                for (int j = 0; j < TB; ++j) v += red[k * kThreads + j * C + tid];
                part[k * C + tid] = v;
            }
        }
        slot_publish(slot);
    }

    if (active) {
        const float *c13 = reg0, *c12 = reg1, *c17 = reg2;
        // V[b,c] = sA g3[b,c] + sum_d A[d,c] g13[b,d]
        {
            const float *const cols1[1] = {c13};
#pragma unroll
            for (int c0 = 0; c0 < NMAX; c0 += 8) {
                if (c0 < n) {
                    float acc[1][8];
                    list_dot8<1>(S.adj, c0, cols1, acc);
#pragma unroll
                    for (int k = 0; k < 8; ++k) V[c0 + k] += acc[0][k];  // lists >= n are empty
                }
            }
        }
        // E2[b, a] = u11[b] + sum_d A[d,a] g17[b,d]                  (region 0; c13 is dead)
        {
            const float *const cols1[1] = {c17};
            for (int s0 = 0; s0 < n; s0 += 8) {
                float acc[1][8];
                list_dot8<1>(S.adj, s0, cols1, acc);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (s0 + k < n) E2s[(s0 + k) * kThreads] = u11 + acc[0][k];
            }
        }
        // U[a,b], own-row part: u4[b] + sum_d A[d,a] g12[b,d]         (region 2; c17 is dead)
        {
            const float *const cols1[1] = {c12};
            for (int s0 = 0; s0 < n; s0 += 8) {
                float acc[1][8];
                list_dot8<1>(S.adj, s0, cols1, acc);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (s0 + k < n) Us[(s0 + k) * kThreads] = u4 + acc[0][k];
            }
        }
        // U[a,b] += sA g1[a,b] + tr g7[a,b];  E1[a,b] = 0              (region 1; c12 is dead)
        // The cells (a, b) of the other rows are read kBlk at a time so kBlk * 2 loads are in flight per thread.
        const float *gcol = g + (int64_t)b * cell + f;  // cell (a, b): gcol[a*n*cell + k*C]
        const int64_t astep = (int64_t)n * cell;
        for (int s0 = 0; s0 < n; s0 += kBlk) {
            float t1[kBlk], t7[kBlk];
#pragma unroll
            for (int k = 0; k < kBlk; ++k) {
                const bool ok = s0 + k < n;
                t1[k] = ok ? ld_stream(gcol + (s0 + k) * astep) : 0.f;
                t7[k] = ok ? ld_stream(gcol + (s0 + k) * astep + 6 * C) : 0.f;
            }
#pragma unroll
            for (int k = 0; k < kBlk; ++k) {
                if (s0 + k < n) {
                    Us[(s0 + k) * kThreads] += fmaf(sA, t1[k], tr * t7[k]);
                    E1s[(s0 + k) * kThreads] = 0.f;
                }
            }
        }
        // U[a,b] += sum_d A[d,b] g9[a,d];  E1[a,b] += sum_d A[d,b] g16[a,d]: for every non-zero (d, b) of column b the
        // cells (a, d) of all rows a.  (The sibling terms are added in phase 1b.)
        const int cb = S.adj.cnt[b];
        for (int j = 0; j < cb; ++j) {
            const float w = S.adj.val[j * NMAX + b];
            const float *gd = g + (int64_t)S.adj.idx[j * NMAX + b] * cell + f;  // cell (a, d): gd[a*n*cell + k*C]
            for (int s0 = 0; s0 < n; s0 += kBlk) {
                float t9[kBlk], t16[kBlk];
#pragma unroll
                for (int k = 0; k < kBlk; ++k) {
                    const bool ok = s0 + k < n;
                    t9[k] = ok ? gd[(s0 + k) * astep + 8 * C] : 0.f;
                    t16[k] = ok ? gd[(s0 + k) * astep + 15 * C] : 0.f;
                }
#pragma unroll
                for (int k = 0; k < kBlk; ++k) {
                    if (s0 + k < n) {
                        Us[(s0 + k) * kThreads] = fmaf(w, t9[k], Us[(s0 + k) * kThreads]);
                        E1s[(s0 + k) * kThreads] = fmaf(w, t16[k], E1s[(s0 + k) * kThreads]);
                    }
                }
            }
        }
    }

    // ---- phase 1b: terms that need the siblings ------------------------------------------------------------------
    slot_wait_siblings(slot, tiles_n);
    float G10[NMAX];
    if (active) {
        float tot[4] = {0.f, 0.f, 0.f, 0.f};  // s5, s14, s15, s18
        for (int t = 0; t < tiles_n; ++t) {
            const float *part = sc + L.partials + (int64_t)t * 4 * C;
#pragma unroll
            for (int k = 0; k < 4; ++k) tot[k] += __ldcg(part + k * C + f);
        }
        for (int s0 = 0; s0 < n; s0 += kBlk) {
            float t2[kBlk], t8[kBlk];
#pragma unroll
            for (int k = 0; k < kBlk; ++k) {
                const bool ok = s0 + k < n;
                t2[k] = ok ? __ldcg(u2p + (int64_t)(s0 + k) * C + f) : 0.f;
                t8[k] = ok ? __ldcg(u8p + (int64_t)(s0 + k) * C + f) : 0.f;
            }
#pragma unroll
            for (int k = 0; k < kBlk; ++k) {
                const int s = s0 + k;
                if (s < n) {
                    float u = Us[s * kThreads] + t2[k] + tot[0];
                    float e1 = E1s[s * kThreads] + t8[k] + tot[2];
                    if (s == b) {
                        u += tot[1];
                        e1 += tot[3];
                    }
                    Us[s * kThreads] = u;
                    E1s[s * kThreads] = e1;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < NMAX; ++c) G10[c] = (c < n) ? grow[c * cell + 9 * C] : 0.f;  // case 10
    }
    slot_release(slot, tiles_n);  // last scratch read is above; the stream below touches only gout and gT
    if (!active) return;

    // ---- phase 2: stream gT ------------------------------------------------------------------------------------------
    // g6[a,b] is fetched kG6Ahead steps ahead into a register queue; the a loop is unrolled by the queue length so
    // the queue is renamed statically (shifting it would wait for the newest load every step).
    const float *g6p = g + (int64_t)b * cell + 5 * C + f;  // g6[a, b] at g6p[a*n*cell]
    const int64_t astep = (int64_t)n * cell;
    float g6q[kG6Ahead];
#pragma unroll
    for (int k = 0; k < kG6Ahead; ++k) g6q[k] = (k < n) ? g6p[k * astep] : 0.f;
    for (int s0 = 0; s0 < n; s0 += kG6Ahead) {
#pragma unroll
        for (int k = 0; k < kG6Ahead; ++k) {
            const int s = s0 + k;
            if (s < n) {
                const float g6 = g6q[k];
                g6q[k] = (s + kG6Ahead < n) ? g6p[(s + kG6Ahead) * astep] : 0.f;
                float *dst = slab_ptr(a.gT, inst, s, n, nm, C) + ((int64_t)b * n) * C + f;
                const float ua = Us[s * kThreads], e1 = E1s[s * kThreads], e2 = E2s[s * kThreads];
                if (n == NMAX)
                    emit_row<C, ACCUM, true>(dst, n, b, s, ua, g6, r_s[s], e1, e2, a.beta, r_s, V, G10);
                else
                    emit_row<C, ACCUM, false>(dst, n, b, s, ua, g6, r_s[s], e1, e2, a.beta, r_s, V, G10);
            }
        }
    }
}

template <int C>
cudaError_t configure_for() {
    cudaError_t e = cudaFuncSetAttribute(k_fwd_fused<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_bwd_fused<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_bwd_fused<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem);
}

template <int C>
cudaError_t forward_for(const Fused18Fwd &a, cudaStream_t st, LaunchLog *log) {
    const unsigned grid = (unsigned)(a.b.count * tiles_of(a.b.n_max, C));
    CCN_LAUNCH(log, K_FWD_FUSED, st, k_fwd_fused<C><<<grid, kThreads, kFwdSmem, st>>>(a));
    return cudaGetLastError();
}

template <int C>
cudaError_t backward_for(const Fused18Bwd &a, cudaStream_t st, LaunchLog *log) {
    const unsigned grid = (unsigned)(a.b.count * tiles_of(a.b.n_max, C));
    if (a.beta != 0.f)
        CCN_LAUNCH(log, K_BWD_FUSED, st, (k_bwd_fused<C, true><<<grid, kThreads, kBwdSmem, st>>>(a)));
    else
        CCN_LAUNCH(log, K_BWD_FUSED, st, (k_bwd_fused<C, false><<<grid, kThreads, kBwdSmem, st>>>(a)));
    return cudaGetLastError();
}

}  // namespace

bool fused_path_supported(int n_max, int C) { return n_max >= 1 && n_max <= NMAX && (C == 32 || C == 64 || C == 128); }
int fused_tiles(int n_max, int C) { return tiles_of(n_max, C); }
int fused_ctl_words(int slots) { return ctl_words(slots); }
int64_t fused_fwd_scratch_words(int n_max, int C) { return FwdScratch(n_max, C).words; }
int64_t fused_bwd_scratch_words(int n_max, int C) { return BwdScratch(n_max, C).words; }

cudaError_t fused_path_configure() {
    cudaError_t e = configure_for<32>();
    if (e != cudaSuccess) return e;
    e = configure_for<64>();
    if (e != cudaSuccess) return e;
    return configure_for<128>();
}

cudaError_t launch_fused_forward(const Fused18Fwd &a, cudaStream_t st, LaunchLog *log) {
    cudaError_t e = cudaMemsetAsync(a.ctl, 0, (size_t)ctl_words(a.slots) * sizeof(int), st);
    if (e != cudaSuccess) return e;
    switch (a.b.C) {
        case 32: return forward_for<32>(a, st, log);
        case 64: return forward_for<64>(a, st, log);
        case 128: return forward_for<128>(a, st, log);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_fused_backward(const Fused18Bwd &a, cudaStream_t st, LaunchLog *log) {
    cudaError_t e = cudaMemsetAsync(a.ctl, 0, (size_t)ctl_words(a.slots) * sizeof(int), st);
    if (e != cudaSuccess) return e;
    switch (a.b.C) {
        case 32: return backward_for<32>(a, st, log);
        case 64: return backward_for<64>(a, st, log);
        case 128: return backward_for<128>(a, st, log);
    }
    return cudaErrorInvalidValue;
}

}  // namespace ccn
