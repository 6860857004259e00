// contract18_fused_impl.cuh -- single-kernel StackTensor3D + RisiContraction_18 forward and backward for the benchmark
// shapes (n <= 32, C in {8, 16, 32, 64, 128}); sm_100a only.
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
//
// This file is compiled twice: contract18_fused_fwd.cu (CCN_KTHREADS = 256: 2 CTAs of 256 threads per SM, TB = 256 / C rows per
// tile) defines the forward entry points and the backward WITH fused promotion (scatter), contract18_fused_bwd.cu
// (CCN_KTHREADS = 128: 3 CTAs of 128 threads per SM) the plain backward.  Measured at N = 32, C = 64
// (profiles/r02_kernel_experiments.md): forward 1.219 ms per 512 instances with 256-thread tiles vs 1.250 with 128 and 1.432
// with 64; backward 1.335 / 1.272 / 1.504 -- the backward's latency-bound phase 1 gains from a third independent tile per SM,
// the forward's bandwidth-bound stream from the deeper ring of the larger tile; the scatter backward is bound by the L2
// reduction stream and prefers the larger tile (1.81 vs 2.04 ms).
#include "contract18_kernels.cuh"

#if !defined(CCN_FUSED_FORWARD) && !defined(CCN_FUSED_BACKWARD)
#error "include through contract18_fused_fwd.cu / contract18_fused_bwd.cu"
#endif

namespace ccn {

namespace {

constexpr int NMAX = 32;       // largest receptive field handled by the register-resident row accumulators
constexpr int kThreads = CCN_KTHREADS;  // one thread per (channel, row-in-tile)
constexpr int kMinCtas = CCN_KTHREADS >= 256 ? 2 : 3;  // resident CTAs per SM the kernels are compiled for
#ifndef CCN_KSTAGES
#define CCN_KSTAGES 3
#endif
constexpr int kStages = CCN_KSTAGES;  // TMA ring depth (3 x 32 KiB in flight per CTA, 2 CTAs per SM)
constexpr int kStageFloats = kThreads * NMAX;  // TB * NMAX * C
constexpr int kColFloats = kThreads * NMAX;    // one thread-private column: col[e * kThreads + tid]
// A/B switches of Fused18*::variant (ccn_ctx env CCN_FUSED_VARIANT overrides the default): none at present.
// Tried in round 2, measured and removed again (profiles/r02_kernel_experiments.md):
//   - L2 prefetch of the backward tile's whole gout block at tile start: 1.55 vs 1.35 ms per 512 instances (296 resident
//     tiles x 576 KiB exceed L2: the lines are evicted before use and read twice); of only the small dependent rounds
//     (slab 7, slab 10, the sparse cells): no change; of the b-side rows during the a-side: 1.40 ms;
//   - 16-byte reductions in the fused promotion backward (quad-transposed with four shuffles so that one lane adds four
//     channels): 2.20 vs 1.81 ms -- the reduction stream is bound in L2, not by the SM's issue rate;
//   - the dense product inlined at every list_dot8 call site: backward 1.61 vs 1.33 ms, hence __noinline__ for it;
//   - code-size reductions of the backward (ncu shows ~1 warp per issue slot stalled on instruction fetch): the column
//     staging loops rolled (-10 % SASS): 1.339 vs 1.343 ms; the sparse list walk out of line (-25 % SASS): 1.54 ms.
constexpr int kDefaultVariant = 0;
constexpr long long kSpinLimit = 4000000000ll; // ~2 s of SM clocks: a sibling that never arrives is a bug, not a wait

__host__ __device__ inline int tiles_of(int n, int C) { return (n + (kThreads / C) - 1) / (kThreads / C); }

// control block (ints): [0] ticket, [1..3] unused, then per slot {arrive, finish, gen_done, pad}.  Cleared by every launch.
// The failure flag is NOT here: it is a sticky word of mapped host memory owned by the context (Fused18*::fault).
__host__ __device__ inline int ctl_words(int slots) { return 4 + 4 * slots; }

struct FwdScratch {  // per slot: planes P, W6, D1, D2 [n*n*C] then per-tile partial totals [tiles][4][C]
    int64_t plane, partials, words;
    __host__ __device__ FwdScratch(int nm, int C) {
        plane = (int64_t)nm * nm * C;
        partials = 4 * plane;
        words = (partials + (int64_t)tiles_of(nm, C) * 4 * C + 31) & ~(int64_t)31;
    }
};

struct BwdScratch {  // per slot: planes UA, EA [n*n*C] then per-tile partial sums s5,s14,s15,s18 [tiles][4][C]
    int64_t plane, partials, words;
    __host__ __device__ BwdScratch(int nm, int C) {
        plane = (int64_t)nm * nm * C;
        partials = 2 * plane;
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
    int dense;  // 1: val holds ALL entries, entry-major (val[e*NMAX + l] = weight of member e for list l); idx / cnt unused
};
// A list walk costs ~6 instructions per entry (index decode, predicate); the dense walk ~1.4.  Molecular adjacency has ~3
// entries per list, a Coulomb-style dense matrix (SMP_beta.h:521-524) n: switch when the longest list passes this many.
constexpr int kDenseThreshold = 10;

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
#ifndef CCN_NO_DENSE
            S.dense = mc > kDenseThreshold;
#else
            S.dense = 0;
#endif
        }
    }
    __syncthreads();
    if (S.dense) {  // CTA-uniform
        for (int i = tid; i < n * NMAX; i += kThreads) {
            const int e = i / NMAX, l = i - e * NMAX;
            S.val[i] = l < n ? (BY_COLUMN ? S.A[e * n + l] : S.A[l * n + e]) : 0.0f;
        }
        __syncthreads();
    }
}

// Eight sparse dot products at once, against NC thread-private shared-memory columns (stride kThreads):
//   acc[q][k] = sum over entries (i, w) of list l0+k of  w * cols[q][i]        l0 a multiple of 8.
// The lists are walked entry-position by entry-position, so the eight (x NC) dependent index -> value chains run
// in parallel instead of one after the other.
// Dense product: acc[q][k] = sum_e W[e][l0+k] * cols[q][e], eight FMAs per column load.  Deliberately NOT inlined: the fused
// kernels are ~10 000 instructions already and the list walk below is the hot path for molecular graphs; inlining this at
// every call site grew the backward by 16 % and cost it 19 % of its speed (instruction cache), measured in round 2.
template <int NC>
__device__ __noinline__ void dense_dot8(const float *__restrict__ w_l0, const float *const *cols, float *__restrict__ acc, int n) {
    float a[NC][8];
#pragma unroll
    for (int q = 0; q < NC; ++q)
#pragma unroll
        for (int k = 0; k < 8; ++k) a[q][k] = 0.f;
#pragma unroll 4
    for (int e = 0; e < n; ++e) {
        const float4 w0 = *reinterpret_cast<const float4 *>(w_l0 + e * NMAX);
        const float4 w1 = *reinterpret_cast<const float4 *>(w_l0 + e * NMAX + 4);
        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            const float x = cols[q][e * kThreads];
#pragma unroll
            for (int k = 0; k < 8; ++k) a[q][k] = fmaf(w[k], x, a[q][k]);
        }
    }
#pragma unroll
    for (int q = 0; q < NC; ++q)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[q * 8 + k] = a[q][k];
}

template <int NC>
__device__ __forceinline__ void list_dot8(const AdjShared &S, int l0, const float *const (&cols)[NC], float (&acc)[NC][8], int n) {
#ifndef CCN_NO_DENSE
    if (S.dense) {
        float tmp[NC * 8];  // the out-of-line call needs addressable results; acc itself stays in registers for the list walk
        dense_dot8<NC>(S.val + l0, cols, tmp, n);
#pragma unroll
        for (int q = 0; q < NC; ++q)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[q][k] = tmp[q * 8 + k];
        return;
    }
#endif
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

// Optional phase trace (debug / profiling aid): thread 0 of every tile records %globaltimer at up to 8 marks.
__device__ __forceinline__ void trace_mark(unsigned long long *trace, int work, int k) {
    if (trace != nullptr && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        trace[(int64_t)work * 8 + k] = t;
    }
}

// ---- sibling protocol ------------------------------------------------------------------------------------------
// Tickets make the start order of the tiles equal to their work order, so a tile only ever waits for tiles that
// have already started (its siblings' later tickets are taken by the CTAs that retire next; at least 2 * #SMs
// CTAs are resident, far more than the <= 16 tiles of one instance).
struct Slot {
    int *arrive, *finish, *gen_done, *error;
};

__device__ __forceinline__ Slot slot_of(int *ctl, int slot, int *fault) {
    return Slot{ctl + 4 + 4 * slot, ctl + 5 + 4 * slot, ctl + 6 + 4 * slot, fault};
}

__device__ __forceinline__ void spin_until(const int *p, int target, int *error) {
    const long long t0 = clock64();
    while (ld_acquire(p) < target) {
        __nanosleep(100);
        if (clock64() - t0 > kSpinLimit) {
            *reinterpret_cast<volatile int *>(error) = 1;  // mapped host memory: the next C-ABI call on the context fails
            __threadfence_system();
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

// An instance with n = 0 has no tile that does work, but its scratch slot's generation must still advance, or the next
// instance mapped to the slot would wait for it forever.  Called by all threads of its tile 0.
__device__ __forceinline__ void retire_empty_instance(const Slot &s, int gen) {
    slot_acquire(s, gen);
    slot_release(s, 1);
}

// ---- register-free staging of thread-private columns (cp.async, SASS LDGSTS) ---------------------------------------
__device__ __forceinline__ void cp_async4(float *dst_smem, const float *src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(float *dst_smem, const float *src_gmem) {  // L2 only (.cg): coherent
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
// 16-byte copy that writes zeros instead when !valid (src-size 0: nothing is read)
__device__ __forceinline__ void cp_async16_zfill(float *dst_smem, const float *src_gmem, bool valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(valid ? 16 : 0)
                 : "memory");
}
// the mbarrier receives one arrival from this thread once all of the thread's earlier cp.async copies have landed
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// col[e * kThreads] <- src[e * stride] for e < n, for the calling thread and its three right-hand neighbours at once:
// the threads with tid % 4 == 0 copy 16 bytes (4 consecutive channels = 4 consecutive threads' words).  All copies of
// a column are in flight together and use no registers.  Consume after cp_async_wait<..>() + __syncwarp().
__device__ __forceinline__ void stage_column(float *col, const float *src, int64_t stride, int n) {
    if ((threadIdx.x & 3) == 0) {
#pragma unroll
        for (int e = 0; e < NMAX; ++e)
            if (e < n) cp_async16(col + e * kThreads, src + e * stride);
    }
}

// Promotion table of one instance (GatherRef) in shared memory.
struct alignas(16) GatherShared {
    int64_t off[NMAX];            // element offset of slab a's source tensor
    int m[NMAX];                  // its side
    alignas(16) short pos[NMAX * NMAX];  // pos[a * n + i]: position of member i inside the source of slab a, or -1
};

__device__ __forceinline__ void load_gather_table(GatherShared &G, const GatherRef &g, int64_t inst, int n, int nm) {
    const int tid = threadIdx.x;
    const int32_t *pos = g.pos + inst * nm * nm;
    for (int i = tid; i < n * n; i += kThreads) {
        const int a = i / n, r = i - a * n;
        G.pos[i] = (short)pos[a * nm + r];
    }
    if (tid < n) {
        G.off[tid] = g.f_off[inst * nm + tid];
        G.m[tid] = g.m[inst * nm + tid];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Forward
// ---------------------------------------------------------------------------------------------------------------
struct FwdSmem {
    AdjShared adj;
    uint64_t full[kStages];
    float *slab[NMAX];  // the instance's slab pointers (stacked input: base + a * n^2 C), read once in the prologue
    int work;
};
struct FwdSmemGather {
    FwdSmem base;
    GatherShared g;
};

// Gathered version of the stage copy: chunk[(bl*n + c)*C + :] <- F_a[pos_a[b0+bl], pos_a[c], :] (or zeros), 16 bytes per
// cp.async, all threads; each thread then posts its arrival on the stage's mbarrier (initialised to kThreads arrivals).
// The pieces a thread copies are the same for every stage (piece q = tid + j * kThreads covers floats [4q, 4q + 4) of the
// chunk): its (row-in-tile, member) pairs are worked out once per tile, only the two position look-ups depend on the slab.
constexpr int kGatherPieces = NMAX / 4;  // per thread and stage: TB * NMAX * (C / 4) / kThreads
struct GatherPlan {
    short b[kGatherPieces], c[kGatherPieces];  // member indices (b = b0 + bl, c) of piece j, b < 0: past the chunk
};
template <int C>
__device__ __forceinline__ void gather_plan(GatherPlan &gp, int n, int b0, int tb) {
    constexpr int PPR = C / 4;  // 16-byte pieces per cell
#pragma unroll
    for (int j = 0; j < kGatherPieces; ++j) {
        const int row = (threadIdx.x + j * kThreads) / PPR;
        const int bl = row / n;
        gp.b[j] = bl < tb ? (short)(b0 + bl) : (short)-1;
        gp.c[j] = (short)(row - bl * n);
    }
}
template <int C>
__device__ __forceinline__ void gather_stage(float *stage, uint64_t *bar, const GatherShared &G, const GatherPlan &gp, const float *f,
                                             int a, int n) {
    constexpr int PPR = C / 4;
    const float *F = f + G.off[a] + (threadIdx.x % PPR) * 4;
    const int m = G.m[a];
    const short *P = G.pos + a * n;
    if (n == NMAX) {
        // full field: piece j of this thread is row r0 + j * S of the chunk, S = kThreads / PPR rows apart, so its member pair is
        // (b0 + row / 32, row % 32) by shifts, and the position look-ups repeat (at C = 64: two distinct c, four distinct b per
        // stage instead of sixteen shared-memory loads); gp.b[0] carries b0
        constexpr int S = kThreads / PPR;
        const int r0 = threadIdx.x / PPR, b0 = gp.b[0] - r0 / NMAX;
#pragma unroll
        for (int j = 0; j < kGatherPieces; ++j) {
            const int row = r0 + j * S;
            const int pb = P[b0 + row / NMAX], pc = P[row % NMAX];
            const bool ok = (pb | pc) >= 0;
            cp_async16_zfill(stage + (threadIdx.x + j * kThreads) * 4, ok ? F + (int64_t)(pb * m + pc) * C : f, ok);
        }
        cp_async_mbar_arrive(bar);
        return;
    }
#pragma unroll
    for (int j = 0; j < kGatherPieces; ++j) {
        if (gp.b[j] >= 0) {
            const int pb = P[gp.b[j]], pc = P[gp.c[j]];
            const bool ok = (pb | pc) >= 0;
            cp_async16_zfill(stage + (threadIdx.x + j * kThreads) * 4, ok ? F + (int64_t)(pb * m + pc) * C : f, ok);
        }
    }
    cp_async_mbar_arrive(bar);
}
constexpr size_t kFwdSmem = (size_t)kStages * kStageFloats * 4 + sizeof(FwdSmem);
constexpr size_t kFwdSmemGather = (size_t)kStages * kStageFloats * 4 + sizeof(FwdSmemGather);

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

template <int C, bool GATHER, bool MASKED>
__global__ void __launch_bounds__(kThreads, kMinCtas) k_fwd_fused(Fused18Fwd a) {
    constexpr int TB = kThreads / C;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *ring = reinterpret_cast<float *>(smem_raw);
    FwdSmem &S = *reinterpret_cast<FwdSmem *>(smem_raw + (size_t)kStages * kStageFloats * 4);
    // only dereferenced when GATHER (the launch then provides kFwdSmemGather bytes)
    GatherShared &GS = reinterpret_cast<FwdSmemGather *>(smem_raw + (size_t)kStages * kStageFloats * 4)->g;

    const int tid = threadIdx.x;
    if (tid == 0) S.work = atomicAdd(a.ctl, 1);
    __syncthreads();
    const int nm = a.b.n_max;
    const int tiles = tiles_of(nm, C);
    const int64_t inst = S.work / tiles;
    const int tile = S.work - (int)inst * tiles;
    const int n = a.b.n_of((int)inst);
    const int b0 = tile * TB;
    if (b0 >= n) {
        if (n <= 0 && tile == 0) retire_empty_instance(slot_of(a.ctl, (int)(inst % a.slots), a.fault), (int)(inst / a.slots));
        return;
    }
    const int tiles_n = tiles_of(n, C);
    const int tb = min(TB, n - b0);
    const int f = tid % C, bl = tid / C, b = b0 + bl;
    const bool active = bl < tb;
    // for C < 32 a warp spans several rows, some of which may be past n; for C >= 32 activity is warp-uniform
    const unsigned amask = C >= 32 ? 0xffffffffu : __ballot_sync(0xffffffffu, active);
    const Slot slot = slot_of(a.ctl, (int)(inst % a.slots), a.fault);

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&S.full[s], GATHER ? kThreads : 1);
        fence_mbar_init();
    }
    trace_mark(a.trace, S.work, 0);
    if (GATHER) load_gather_table(GS, a.G, inst, n, nm);
    else if (tid < n) S.slab[tid] = slab_ptr(a.T, inst, tid, n, nm, C);
    build_adjacency<false>(S.adj, a.adj + inst * a.stride_adj, n, a.positive_part != 0);  // ends with __syncthreads

    const uint32_t bytes = (uint32_t)(tb * n * C) * 4u;
    const int64_t row_off = (int64_t)b0 * n * C;
    uint64_t policy = 0;
    GatherPlan gp;
    if (GATHER) {
        gather_plan<C>(gp, n, b0, tb);
        for (int s = 0; s < kStages && s < n; ++s) gather_stage<C>(ring + s * kStageFloats, &S.full[s], GS, gp, a.G.f, s, n);
    } else if (tid == 0) {
        policy = l2_evict_first_policy();
        for (int s = 0; s < kStages && s < n; ++s) {
            mbar_arrive_expect_tx(&S.full[s], bytes);
            bulk_g2s_hint(ring + s * kStageFloats, S.slab[s] + row_off, bytes, &S.full[s], policy);
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

    trace_mark(a.trace, S.work, 1);
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
        if (GATHER) {
            if (s + kStages < n) gather_stage<C>(ring + st_i * kStageFloats, &S.full[st_i], GS, gp, a.G.f, s + kStages, n);
        } else if (tid == 0 && s + kStages < n) {
            mbar_arrive_expect_tx(&S.full[st_i], bytes);
            bulk_g2s_hint(ring + st_i * kStageFloats, S.slab[s + kStages] + row_off, bytes, &S.full[st_i], policy);
        }
    }

    trace_mark(a.trace, S.work, 2);
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
    // MASKED: slab dropout fused in (RisiContraction_18_dropout.h:104-478): a dropped slab is written as exact zeros, the kept
    // ones scaled (1, or nKept/18 in the operator's test mode).  The plain operator is the !MASKED instantiation (no cost).
    const uint32_t keep = a.keep;
    const float oscale = a.out_scale;
    auto put = [&](float *p, int k, float v) {
        if (MASKED) __stcs(p + k * C, ((keep >> k) & 1u) ? v * oscale : 0.f);
        else __stcs(p + k * C, v);
    };
    float *col0 = ring + tid, *col1 = col0 + kColFloats, *col2 = col1 + kColFloats;

    // ---- pass A: slabs that only need this tile's rows ------------------------------------------------------------
    constexpr int kMaxTiles = NMAX / TB;
    if (active) {
        // columns b of P and D2 (this tile's own stores, visible after the fence in slot_publish)
        stage_column(col1, Pp + (int64_t)b * C + f, (int64_t)n * C, n);
        stage_column(col2, D2p + (int64_t)b * C + f, (int64_t)n * C, n);
        cp_async_commit();
        float *orow = outi + ((int64_t)b * n) * cell + f;  // + y*cell + k*C
#pragma unroll
        for (int c = 0; c < NMAX; ++c) {
            if (c < n) {
                put(orow + c * cell, 2, sA * Q[c]);  // case 3  (RisiContraction_18.h:110)
                put(orow + c * cell, 9, W10[c]);     // case 10 (:195)
                col0[c * kThreads] = Q[c];
            }
        }
        cp_async_wait<0>();
        __syncwarp(amask);
        const float *const colsA[3] = {col1, col0, col2};
        for (int d0 = 0; d0 < n; d0 += 8) {
            float acc[3][8];
            list_dot8<3>(S.adj, d0, colsA, acc, n);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int d = d0 + k;
                if (d < n) {
                    float *o = orow + d * cell;
                    const float rd = r_s[d];
                    put(o, 3, rd * S4);      // case 4  (:114)
                    put(o, 10, rd * S11);    // case 11 (:211)
                    put(o, 11, acc[0][k]);   // case 12 (:226)  sum_e A[d,e] P[e,b]
                    put(o, 12, acc[1][k]);   // case 13 (:241)  sum_e A[d,e] Q[b,e]
                    put(o, 16, acc[2][k]);   // case 17 (:304)  sum_e A[d,e] T[e,b,e]
                }
            }
        }
    }

    // ---- pass B: slabs that need the siblings' rows ----------------------------------------------------------------
    trace_mark(a.trace, S.work, 3);
    slot_wait_siblings(slot, tiles_n);
    trace_mark(a.trace, S.work, 4);
    if (active) {
        const int x = b;
        const int64_t xrow = ((int64_t)x * n) * C + f;
        __syncwarp(amask);  // every lane is done with the pass A columns
        stage_column(col0, Pp + xrow, C, n);   // row x of P
        stage_column(col1, D1p + xrow, C, n);  // row x of D1 = T[x,e,e]
        stage_column(col2, W6p + xrow, C, n);  // row x of W6
        cp_async_commit();
        float part[kMaxTiles][4];
#pragma unroll
        for (int t = 0; t < kMaxTiles; ++t)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                part[t][k] = (t < tiles_n) ? __ldcg(sc + L.partials + ((int64_t)t * 4 + k) * C + f) : 0.f;
        float tot[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int t = 0; t < kMaxTiles; ++t)
#pragma unroll
            for (int k = 0; k < 4; ++k) tot[k] += part[t][k];
        cp_async_wait<0>();
        __syncwarp(amask);
        float s2 = 0.f, s8 = 0.f;
#pragma unroll
        for (int e = 0; e < NMAX; ++e) {
            if (e < n) {
                s2 += col0[e * kThreads];
                s8 += col1[e * kThreads];
            }
        }
        float *orow = outi + ((int64_t)x * n) * cell + f;
        const float *const colsB[2] = {col0, col1};
        for (int d0 = 0; d0 < n; d0 += 8) {
            float acc[2][8];
            list_dot8<2>(S.adj, d0, colsB, acc, n);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int d = d0 + k;
                if (d < n) {
                    float *o = orow + d * cell;
                    const float pv = col0[d * kThreads];
                    const float rd = r_s[d];
                    const float axd = S.adj.A[x * n + d];
                    put(o, 0, sA * pv);                 // case 1  (:102)
                    put(o, 1, rd * s2);                 // case 2  (:106)
                    put(o, 4, axd * tot[0]);            // case 5  (:118)
                    put(o, 5, col2[d * kThreads]);      // case 6  (:133)
                    put(o, 6, tr * pv);                 // case 7  (:149)
                    put(o, 7, rd * s8);                 // case 8  (:165)
                    put(o, 8, acc[0][k]);               // case 9  (:180)  sum_e A[d,e] P[x,e]
                    put(o, 13, axd * tot[1]);           // case 14 (:256)
                    put(o, 14, axd * tot[2]);           // case 15 (:271)
                    put(o, 15, acc[1][k]);              // case 16 (:290)  sum_e A[d,e] T[x,e,e]
                    put(o, 17, axd * tot[3]);           // case 18 (:318)
                }
            }
        }
    }
    slot_release(slot, tiles_n);
    trace_mark(a.trace, S.work, 5);
}

// ---------------------------------------------------------------------------------------------------------------
// Backward.  Formulas: SURVEY.md section 8(a) "Backward" (exact transpose of RisiContraction_18.h:333-560).
// ---------------------------------------------------------------------------------------------------------------
struct BwdSmem {
    AdjShared adj;
    float red[4 * kThreads];
    float *slab[NMAX];  // the instance's gradient slab pointers, read once in the prologue
    int work;
};
struct BwdSmemScatter {
    BwdSmem base;
    GatherShared g;
};
constexpr size_t kBwdSmem = (size_t)3 * kColFloats * 4 + sizeof(BwdSmem);
constexpr size_t kBwdSmemScatter = (size_t)3 * kColFloats * 4 + sizeof(BwdSmemScatter);

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

// Scatter form of emit_row (promotion backward fused in): row (a = s, b) of gT is ADDED into the level l-1 gradient at
// gf[f_off[a]][pos_a[b], pos_a[c], :] for the members c present in the source (MatTensorMul.h:67-85, TensorMatMul.h:66-84 with
// the 0/1 selection matrices, then StackTensor3D.h:74-90).  One f_{l-1}[w] feeds many stacks, hence reductions (SASS RED).
// shared-memory loads through a 32-bit address (register + immediate after unrolling) and a predicated reduction: see emit_row_scatter
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float x;
    asm("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(addr));
    return x;
}
__device__ __forceinline__ int lds_s16(uint32_t addr) {
    int x;
    asm("ld.shared.s16 %0, [%1];" : "=r"(x) : "r"(addr));
    return x;
}
// *addr += v unless pos < 0 (an absent member)
__device__ __forceinline__ void red_add_if(float *addr, float v, int pos) {
    asm volatile("{\n .reg .pred q;\n setp.ge.s32 q, %2, 0;\n @q red.global.add.f32 [%0], %1;\n}" ::"l"(addr), "f"(v), "r"(pos) : "memory");
}
template <int C, bool FULL>
__device__ __forceinline__ void emit_row_scatter(float *__restrict__ dst_row, const short *__restrict__ P, int n, int b, int s,
                                                 float ua, float g6, float ra, float e1, float e2,
                                                 const float *__restrict__ r_s, const float (&V)[NMAX], const float (&G10)[NMAX]) {
#ifdef CCN_SCATTER_DIAG_IN
#pragma unroll
    for (int c = 0; c < NMAX; ++c) {
        if (FULL || c < n) {
            const int pc = P[c];
            float v = ua + V[c];
            v = fmaf(g6, r_s[c], v);
            v = fmaf(ra, G10[c], v);
            if (c == b) v += e1;
            if (c == s) v += e2;
            if (pc >= 0) atomicAdd(dst_row + pc * C, v);
        }
    }
#else
    // The stream of reductions is instruction-bound (phase trace: 86 of a tile's 143 us; the same reductions alone run at three
    // times the rate, profiles/scatter_gather_probe.cu), so the two diagonal terms -- a compare and a predicated add per element
    // each -- are kept out of the element loop and added as two more reductions per row (reductions commute).
    // SASS of the plain C++ form: a branch with BSSY / BSYNC around every atomicAdd and the shared-memory window base of r_s
    // recomputed per element (13 instructions per element); with a predicated red and register + immediate shared-memory
    // addresses it is 8.
    const uint32_t rs_addr = smem_u32(r_s), p_addr = smem_u32(P);
#pragma unroll
    for (int c = 0; c < NMAX; ++c) {
        if (FULL || c < n) {
            const int pc = lds_s16(p_addr + 2 * c);
            float v = ua + V[c];
            v = fmaf(g6, lds_f32(rs_addr + 4 * c), v);
            v = fmaf(ra, G10[c], v);
            // byte address = 64-bit row base + unsigned 32-bit product: one IMAD.WIDE.U32 (unused when pc < 0)
            red_add_if(reinterpret_cast<float *>(reinterpret_cast<char *>(dst_row) + (size_t)((uint32_t)pc * (uint32_t)(C * 4))), v, pc);
        }
    }
    {
        const int pb = P[b], ps = P[s];  // cells (b, c = b) and (b, c = s) of slab s
        red_add_if(dst_row + (uint32_t)pb * (uint32_t)C, e1, pb);
        red_add_if(dst_row + (uint32_t)ps * (uint32_t)C, e2, ps);
    }
#endif
}

template <int C, bool ACCUM, bool SCATTER, bool MASKED>
__global__ void __launch_bounds__(kThreads, kMinCtas) k_bwd_fused(Fused18Bwd a) {
    constexpr int TB = kThreads / C;
    constexpr int kG6Ring = 8;  // g6[a,b] is fetched this many steps ahead of its use (cp.async into shared memory)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *planes = reinterpret_cast<float *>(smem_raw);
    BwdSmem &S = *reinterpret_cast<BwdSmem *>(smem_raw + (size_t)3 * kColFloats * 4);
    // only dereferenced when SCATTER (the launch then provides kBwdSmemScatter bytes)
    GatherShared &GS = reinterpret_cast<BwdSmemScatter *>(smem_raw + (size_t)3 * kColFloats * 4)->g;

    const int tid = threadIdx.x;
    if (tid == 0) S.work = atomicAdd(a.ctl, 1);
    __syncthreads();
    const int nm = a.b.n_max;
    const int tiles = tiles_of(nm, C);
    const int64_t inst = S.work / tiles;
    const int tile = S.work - (int)inst * tiles;
    const int n = a.b.n_of((int)inst);
    const int b0 = tile * TB;
    if (b0 >= n) {
        if (n <= 0 && tile == 0) retire_empty_instance(slot_of(a.ctl, (int)(inst % a.slots), a.fault), (int)(inst / a.slots));
        return;
    }
    const int tiles_n = tiles_of(n, C);
    const int f = tid % C, bl = tid / C, b = b0 + bl;
    const bool active = b < n;
    // for C < 32 a warp spans several rows, some of which may be past n; for C >= 32 activity is warp-uniform
    const unsigned amask = C >= 32 ? 0xffffffffu : __ballot_sync(0xffffffffu, active);
    const Slot slot = slot_of(a.ctl, (int)(inst % a.slots), a.fault);

    trace_mark(a.trace, S.work, 0);
    if (SCATTER) load_gather_table(GS, a.G, inst, n, nm);  // visible after the barriers inside build_adjacency
    else if (tid < n) S.slab[tid] = slab_ptr(a.gT, inst, tid, n, nm, C);
    build_adjacency<true>(S.adj, a.adj + inst * a.stride_adj, n, a.positive_part != 0);
    slot_acquire(slot, (int)(inst / a.slots));
    trace_mark(a.trace, S.work, 1);

    const float *r_s = S.adj.r;
    const float sA = S.adj.sA, tr = S.adj.tr;
    const BwdScratch L(n, C);
    float *sc = a.scratch + (inst % a.slots) * a.scratch_words;
    float *UAp = sc, *EAp = sc + L.plane;  // [a][b][f]: the a-side of U and E1, written by the owner of row a
    const float *g = a.gout + inst * a.stride_gout;
    const int64_t cell = (int64_t)kSlabs * C;
    // three thread-private shared-memory columns / planes (stride kThreads), reused as their content dies:
    //   region 0: base -> c13 -> E2 plane     region 1: c9 -> c12 -> E1 plane     region 2: c16 -> c17 -> U plane
    float *reg0 = planes + tid, *reg1 = reg0 + kColFloats, *reg2 = reg1 + kColFloats;
    float *E2s = reg0, *E1s = reg1, *Us = reg2;  // [a * kThreads]
    const float *grow = g + ((int64_t)(active ? b : b0) * n) * cell + f;  // row b: grow[d*cell + k*C]
    // MASKED: slab dropout fused in (RisiContraction_18_dropout.h:480-797): a dropped slab of gout is never read and counts as
    // zero.  The plain operator is the !MASKED instantiation (kept() folds to true: no predicates on the load paths).
    const uint32_t keep = a.keep;
    auto kept = [&](int k) { return !MASKED || ((keep >> k) & 1u) != 0u; };
    // gout slab k of this thread's row as a thread-private column (zeros for a dropped slab)
    auto stage_gout = [&](float *col, int k) {
        if (kept(k)) {
            stage_column(col, grow + k * C, cell, n);
        } else if ((tid & 3) == 0) {
#pragma unroll
            for (int e = 0; e < NMAX; ++e)
                if (e < n) *reinterpret_cast<float4 *>(col + e * kThreads) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };

    // Every slab of gout is read from DRAM by exactly one phase of exactly one tile: the slabs that are only
    // copied (cases 1, 9, 16 / 13, 12, 17) go straight to shared memory with cp.async, the ones that are reduced or
    // scaled go through registers, a whole row (32 cells) per slab in flight at once.
    constexpr int kMaxTiles = NMAX / TB;
    const int64_t brow = ((int64_t)b * n) * C + f;  // row b of a scratch plane

    // ---- phase 1a: a-side (this tile as the owner of rows a = b of U and E1) -------------------------------------
    //   UA[b][p] = sA g1[b,p] + tr g7[b,p] + u2[b] + sum_d A[d,p] g9[b,d]        EA[b][p] = u8[b] + sum_d A[d,p] g16[b,d]
    {
        float s4[4] = {0.f, 0.f, 0.f, 0.f};  // s5, s14, s15, s18 partial sums over this row
        if (active) {
            stage_gout(reg0, 0);   // g1[b, :]
            stage_gout(reg1, 8);   // g9[b, :]
            stage_gout(reg2, 15);  // g16[b, :]
            cp_async_commit();
            float u2 = 0.f, u8 = 0.f;
            {
                float t2[NMAX], t8[NMAX];
#pragma unroll
                for (int d = 0; d < NMAX; ++d) {
                    t2[d] = (d < n && kept(1)) ? ld_stream(grow + d * cell + 1 * C) : 0.f;  // case 2
                    t8[d] = (d < n && kept(7)) ? ld_stream(grow + d * cell + 7 * C) : 0.f;  // case 8
                }
#pragma unroll
                for (int d = 0; d < NMAX; ++d) {
                    const float rd = r_s[d];  // 0 beyond n
                    u2 = fmaf(rd, t2[d], u2);
                    u8 = fmaf(rd, t8[d], u8);
                }
            }
            float t7[NMAX];
#pragma unroll
            for (int d = 0; d < NMAX; ++d) t7[d] = (d < n && kept(6)) ? ld_stream(grow + d * cell + 6 * C) : 0.f;  // case 7
            // cases 5, 14, 15, 18: only the cells (b, d) with A[b,d] != 0 matter
            if (C >= 32) {  // b is the same for the whole warp: walk the row's non-zeros with warp-uniform control flow
                const int lane = tid & 31;
                const float wl = lane < n ? S.adj.A[b * n + lane] : 0.f;
                unsigned mask = __ballot_sync(0xffffffffu, wl != 0.f);
                while (mask) {
                    float w[8], t[8][4];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int d = mask ? __ffs(mask) - 1 : -1;
                        mask &= mask - 1;  // stays 0 once empty
                        w[k] = d >= 0 ? S.adj.A[b * n + d] : 0.f;
                        const float *gd = grow + (d >= 0 ? d : 0) * cell;
                        t[k][0] = (d >= 0 && kept(4)) ? ld_stream(gd + 4 * C) : 0.f;
                        t[k][1] = (d >= 0 && kept(13)) ? ld_stream(gd + 13 * C) : 0.f;
                        t[k][2] = (d >= 0 && kept(14)) ? ld_stream(gd + 14 * C) : 0.f;
                        t[k][3] = (d >= 0 && kept(17)) ? ld_stream(gd + 17 * C) : 0.f;
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k)
#pragma unroll
                        for (int q = 0; q < 4; ++q) s4[q] = fmaf(w[k], t[k][q], s4[q]);
                }
            } else {  // narrow channel counts: a warp covers 32 / C rows, every thread scans its own row of A
                for (int d = 0; d < n; ++d) {
                    const float w = S.adj.A[b * n + d];
                    if (w != 0.f) {
                        const float *gd = grow + d * cell;
                        if (kept(4)) s4[0] = fmaf(w, ld_stream(gd + 4 * C), s4[0]);
                        if (kept(13)) s4[1] = fmaf(w, ld_stream(gd + 13 * C), s4[1]);
                        if (kept(14)) s4[2] = fmaf(w, ld_stream(gd + 14 * C), s4[2]);
                        if (kept(17)) s4[3] = fmaf(w, ld_stream(gd + 17 * C), s4[3]);
                    }
                }
            }
            cp_async_wait<0>();
            __syncwarp(amask);
            const float *const cols2[2] = {reg1, reg2};
#pragma unroll
            for (int p0 = 0; p0 < NMAX; p0 += 8) {
                if (p0 < n) {
                    float d8[2][8];
                    list_dot8<2>(S.adj, p0, cols2, d8, n);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int pp = p0 + k;
                        if (pp < n) {
                            float t7v = 0.f;
#pragma unroll
                            for (int q = 0; q < NMAX; ++q) t7v = (q == pp) ? t7[q] : t7v;  // static index after unrolling
                            UAp[brow + (int64_t)pp * C] = fmaf(sA, reg0[pp * kThreads], fmaf(tr, t7v, u2 + d8[0][k]));
                            EAp[brow + (int64_t)pp * C] = u8 + d8[1][k];
                        }
                    }
                }
            }
        }
        float *red = S.red;
#pragma unroll
        for (int k = 0; k < 4; ++k) red[k * kThreads + tid] = s4[k];
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
    trace_mark(a.trace, S.work, 2);

    // ---- phase 1b: b-side (this tile as the owner of columns b of gT), own rows only ----------------------------------
    float V[NMAX];
    if (active) {
        const float *c13 = reg0, *c12 = reg1, *c17 = reg2;
        stage_gout(reg0, 12);  // g13[b, :]
        stage_gout(reg1, 11);  // g12[b, :]
        stage_gout(reg2, 16);  // g17[b, :]
        cp_async_commit();
        float u4 = 0.f, u11 = 0.f;
        {
            float t4[NMAX], t11[NMAX];
#pragma unroll
            for (int d = 0; d < NMAX; ++d) {
                t4[d] = (d < n && kept(3)) ? ld_stream(grow + d * cell + 3 * C) : 0.f;    // case 4
                t11[d] = (d < n && kept(10)) ? ld_stream(grow + d * cell + 10 * C) : 0.f;  // case 11
                V[d] = (d < n && kept(2)) ? ld_stream(grow + d * cell + 2 * C) : 0.f;     // case 3
            }
#pragma unroll
            for (int d = 0; d < NMAX; ++d) {
                const float rd = r_s[d];
                u4 = fmaf(rd, t4[d], u4);
                u11 = fmaf(rd, t11[d], u11);
                V[d] *= sA;
            }
        }
        cp_async_wait<0>();
        __syncwarp(amask);
        {  // V[b,c] = sA g3[b,c] + sum_d A[d,c] g13[b,d]
            const float *const cols1[1] = {c13};
#pragma unroll
            for (int c0 = 0; c0 < NMAX; c0 += 8) {
                if (c0 < n) {
                    float d8[1][8];
                    list_dot8<1>(S.adj, c0, cols1, d8, n);
#pragma unroll
                    for (int k = 0; k < 8; ++k) V[c0 + k] += d8[0][k];  // lists >= n are empty
                }
            }
        }
        {  // E2[b, a] = u11[b] + sum_d A[d,a] g17[b,d]                  (region 0; c13 is dead)
            const float *const cols1[1] = {c17};
            for (int s0 = 0; s0 < n; s0 += 8) {
                float d8[1][8];
                list_dot8<1>(S.adj, s0, cols1, d8, n);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (s0 + k < n) E2s[(s0 + k) * kThreads] = u11 + d8[0][k];
            }
        }
        {  // U[a,b], b-side: u4[b] + sum_d A[d,a] g12[b,d]               (region 2; c17 is dead)
            const float *const cols1[1] = {c12};
            for (int s0 = 0; s0 < n; s0 += 8) {
                float d8[1][8];
                list_dot8<1>(S.adj, s0, cols1, d8, n);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (s0 + k < n) Us[(s0 + k) * kThreads] = u4 + d8[0][k];
            }
        }
    } else {
#pragma unroll
        for (int c = 0; c < NMAX; ++c) V[c] = 0.f;
    }

    // ---- phase 1c: add the siblings' a-side --------------------------------------------------------------------------
    trace_mark(a.trace, S.work, 3);
    slot_wait_siblings(slot, tiles_n);
    trace_mark(a.trace, S.work, 4);
    float G10[NMAX];
    if (active) {
        const int64_t bcol = (int64_t)b * C + f;  // cell (a, b) of a scratch plane: bcol + a*n*C
        __syncwarp(amask);                              // every lane is done with c12 (region 1)
        stage_column(E1s, EAp + bcol, (int64_t)n * C, n);  // E1[a,b] <- EA[a][b]
        cp_async_commit();
        {
            float part[kMaxTiles][4], tu[NMAX];
#pragma unroll
            for (int t = 0; t < kMaxTiles; ++t)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    part[t][k] = (t < tiles_n) ? __ldcg(sc + L.partials + ((int64_t)t * 4 + k) * C + f) : 0.f;
#pragma unroll
            for (int s = 0; s < NMAX; ++s) tu[s] = s < n ? __ldcg(UAp + bcol + (int64_t)s * n * C) : 0.f;
            float tot[4] = {0.f, 0.f, 0.f, 0.f};  // s5, s14, s15, s18
#pragma unroll
            for (int t = 0; t < kMaxTiles; ++t)
#pragma unroll
                for (int k = 0; k < 4; ++k) tot[k] += part[t][k];
            cp_async_wait<0>();
            __syncwarp(amask);
#pragma unroll
            for (int s = 0; s < NMAX; ++s) {
                if (s < n) {
                    float u = Us[s * kThreads] + tu[s] + tot[0];
                    float e1 = E1s[s * kThreads] + tot[2];
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
        for (int c = 0; c < NMAX; ++c) G10[c] = (c < n && kept(9)) ? ld_stream(grow + c * cell + 9 * C) : 0.f;  // case 10
    }
    slot_release(slot, tiles_n);  // last scratch read is above; the stream below touches only gout and gT
    trace_mark(a.trace, S.work, 5);
    if (!active) return;

    // ---- phase 2: stream gT ------------------------------------------------------------------------------------------
    // g6[a,b] arrives through a kG6Ring-deep cp.async ring in the shared memory that held the dense adjacency and
    // the reduction buffer (both dead by now); every thread copies and reads only its own word, so no barrier.
    const float *g6p = g + (int64_t)b * cell + 5 * C + f;  // g6[a, b] at g6p[a*n*cell]
    const int64_t astep = (int64_t)n * cell;
    float *ringA = S.adj.A + tid, *ringB = S.red + tid;  // slots 0..3 / 4..7, stride kThreads
#pragma unroll
    for (int k = 0; k < kG6Ring; ++k) {
        if (k < n) {
            float *slot_k = (k < 4 ? ringA : ringB) + (k & 3) * kThreads;
            if (kept(5)) cp_async4(slot_k, g6p + k * astep);
            else *slot_k = 0.f;
        }
        cp_async_commit();
    }
    for (int s = 0; s < n; ++s) {
        cp_async_wait<kG6Ring - 1>();
        float *slot_s = ((s & 4) ? ringB : ringA) + (s & 3) * kThreads;
        const float g6 = *slot_s;
        const float ua = Us[s * kThreads], e1 = E1s[s * kThreads], e2 = E2s[s * kThreads];
        if (SCATTER) {
            const short *P = GS.pos + s * n;
            const int pb = P[b];
            if (pb >= 0) {  // else member b is absent from the source of slab s: the whole row is structurally zero
                float *row = a.G.f + GS.off[s] + ((int64_t)pb * GS.m[s]) * C;
                if (n == NMAX)
                    emit_row_scatter<C, true>(row + f, P, n, b, s, ua, g6, r_s[s], e1, e2, r_s, V, G10);
                else
                    emit_row_scatter<C, false>(row + f, P, n, b, s, ua, g6, r_s[s], e1, e2, r_s, V, G10);
            }
        } else {
            float *dst = S.slab[s] + ((int64_t)b * n) * C + f;
            if (n == NMAX)
                emit_row<C, ACCUM, true>(dst, n, b, s, ua, g6, r_s[s], e1, e2, a.beta, r_s, V, G10);
            else
                emit_row<C, ACCUM, false>(dst, n, b, s, ua, g6, r_s[s], e1, e2, a.beta, r_s, V, G10);
        }
        if (s + kG6Ring < n && kept(5)) cp_async4(slot_s, g6p + (s + kG6Ring) * astep);
        cp_async_commit();
    }
    cp_async_wait<0>();
    trace_mark(a.trace, S.work, 6);
}

#ifdef CCN_FUSED_FORWARD
template <int C>
cudaError_t configure_for() {
    cudaError_t e = cudaFuncSetAttribute(k_fwd_fused<C, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_fwd_fused<C, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_bwd_fused<C, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmemScatter);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_fwd_fused<C, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmemGather);
}

template <int C>
cudaError_t backward_scatter_for(const Fused18Bwd &a, cudaStream_t st, LaunchLog *log) {
    const unsigned grid = (unsigned)(a.b.count * tiles_of(a.b.n_max, C));
    CCN_LAUNCH(log, K_BWD_FUSED_SCATTER, st, (k_bwd_fused<C, false, true, false><<<grid, kThreads, kBwdSmemScatter, st>>>(a)));
    return cudaGetLastError();
}

template <int C>
cudaError_t forward_for(const Fused18Fwd &a, cudaStream_t st, LaunchLog *log) {
    const unsigned grid = (unsigned)(a.b.count * tiles_of(a.b.n_max, C));
    const bool masked = a.keep != 0x3ffffu || a.out_scale != 1.f;  // (the gather form has no masked instantiation: the level
    if (a.G.f)                                                       //  entry points never drop slabs)
        CCN_LAUNCH(log, K_FWD_FUSED_GATHER, st, (k_fwd_fused<C, true, false><<<grid, kThreads, kFwdSmemGather, st>>>(a)));
    else if (masked)
        CCN_LAUNCH(log, K_FWD_FUSED, st, (k_fwd_fused<C, false, true><<<grid, kThreads, kFwdSmem, st>>>(a)));
    else
        CCN_LAUNCH(log, K_FWD_FUSED, st, (k_fwd_fused<C, false, false><<<grid, kThreads, kFwdSmem, st>>>(a)));
    return cudaGetLastError();
}
#else
template <int C>
cudaError_t configure_for() {
    cudaError_t e = cudaFuncSetAttribute(k_bwd_fused<C, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_bwd_fused<C, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_bwd_fused<C, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_bwd_fused<C, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem);
}

template <int C>
cudaError_t backward_for(const Fused18Bwd &a, cudaStream_t st, LaunchLog *log) {
    const unsigned grid = (unsigned)(a.b.count * tiles_of(a.b.n_max, C));
    const bool masked = a.keep != 0x3ffffu;
    if (a.beta != 0.f) {
        if (masked) CCN_LAUNCH(log, K_BWD_FUSED, st, (k_bwd_fused<C, true, false, true><<<grid, kThreads, kBwdSmem, st>>>(a)));
        else CCN_LAUNCH(log, K_BWD_FUSED, st, (k_bwd_fused<C, true, false, false><<<grid, kThreads, kBwdSmem, st>>>(a)));
    } else {
        if (masked) CCN_LAUNCH(log, K_BWD_FUSED, st, (k_bwd_fused<C, false, false, true><<<grid, kThreads, kBwdSmem, st>>>(a)));
        else CCN_LAUNCH(log, K_BWD_FUSED, st, (k_bwd_fused<C, false, false, false><<<grid, kThreads, kBwdSmem, st>>>(a)));
    }
    return cudaGetLastError();
}
#endif

}  // namespace

#ifdef CCN_FUSED_FORWARD
bool fused_path_supported(int n_max, int C) {
    return n_max >= 1 && n_max <= NMAX && (C == 8 || C == 16 || C == 32 || C == 64 || C == 128);
}
int fused_ctl_words(int slots) { return ctl_words(slots); }
#define CCN_DIR(name) name##_fwd
#else
#define CCN_DIR(name) name##_bwd
#endif

int CCN_DIR(fused_tiles)(int n_max, int C) { return tiles_of(n_max, C); }
int CCN_DIR(fused_resident_ctas)() { return kMinCtas; }
#ifdef CCN_FUSED_FORWARD
int64_t fused_fwd_scratch_words(int n_max, int C) { return FwdScratch(n_max, C).words; }
int64_t fused_bwd_scatter_scratch_words(int n_max, int C) { return BwdScratch(n_max, C).words; }
#else
int64_t fused_bwd_scratch_words(int n_max, int C) { return BwdScratch(n_max, C).words; }
#endif

cudaError_t CCN_DIR(fused_path_configure)() {
    cudaError_t e = configure_for<8>();
    if (e != cudaSuccess) return e;
    e = configure_for<16>();
    if (e != cudaSuccess) return e;
    e = configure_for<32>();
    if (e != cudaSuccess) return e;
    e = configure_for<64>();
    if (e != cudaSuccess) return e;
    return configure_for<128>();
}

#ifdef CCN_FUSED_FORWARD
cudaError_t launch_fused_forward(const Fused18Fwd &a_in, cudaStream_t st, LaunchLog *log) {
    Fused18Fwd a = a_in;
    if (a.variant < 0) a.variant = kDefaultVariant;
    cudaError_t e = cudaMemsetAsync(a.ctl, 0, (size_t)ctl_words(a.slots) * sizeof(int), st);
    if (e != cudaSuccess) return e;
    switch (a.b.C) {
        case 8: return forward_for<8>(a, st, log);
        case 16: return forward_for<16>(a, st, log);
        case 32: return forward_for<32>(a, st, log);
        case 64: return forward_for<64>(a, st, log);
        case 128: return forward_for<128>(a, st, log);
    }
    return cudaErrorInvalidValue;
}

// backward with the promotion fused in (a.G.f != nullptr): 256-thread tiles, the geometry of the forward
cudaError_t launch_fused_backward_scatter(const Fused18Bwd &a_in, cudaStream_t st, LaunchLog *log) {
    Fused18Bwd a = a_in;
    if (a.variant < 0) a.variant = kDefaultVariant;
    cudaError_t e = cudaMemsetAsync(a.ctl, 0, (size_t)ctl_words(a.slots) * sizeof(int), st);
    if (e != cudaSuccess) return e;
    switch (a.b.C) {
        case 8: return backward_scatter_for<8>(a, st, log);
        case 16: return backward_scatter_for<16>(a, st, log);
        case 32: return backward_scatter_for<32>(a, st, log);
        case 64: return backward_scatter_for<64>(a, st, log);
        case 128: return backward_scatter_for<128>(a, st, log);
    }
    return cudaErrorInvalidValue;
}
#else
cudaError_t launch_fused_backward(const Fused18Bwd &a_in, cudaStream_t st, LaunchLog *log) {
    Fused18Bwd a = a_in;
    if (a.variant < 0) a.variant = kDefaultVariant;
    cudaError_t e = cudaMemsetAsync(a.ctl, 0, (size_t)ctl_words(a.slots) * sizeof(int), st);
    if (e != cudaSuccess) return e;
    switch (a.b.C) {
        case 8: return backward_for<8>(a, st, log);
        case 16: return backward_for<16>(a, st, log);
        case 32: return backward_for<32>(a, st, log);
        case 64: return backward_for<64>(a, st, log);
        case 128: return backward_for<128>(a, st, log);
    }
    return cudaErrorInvalidValue;
}
#endif

}  // namespace ccn
