// mix_tc.cu -- feature-mix forward GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM),
// sm_100a only.
//
// Replaces Reshape2D + MatMul::forward (GraphFlow/MatMul.h:48-66), the reference kernel Matrix_Multiplication_GPU
// (GraphFlow_gpu/MatMul_gpu.h:28-65) and, fused into the epilogue, VectorAddTensor::forward
// (VectorAddTensor.h:46-59) + LeakyReLU3D::forward (LeakyReLU3D.h:60-72):
//
//     Y[M, P] = X[M, K] W[K, P]        Z = lrelu(Y + bias)            M = sum n_i^2 (millions), K = 18 C, P = C_out
//
// fp32 in, fp32 out, fp32-accurate: every product is evaluated as a split-precision ("3xTF32") sum
//     x w ~= hi(x) hi(w) + lo(x) hi(w) + hi(x) lo(w),     hi = round-to-TF32, lo = x - hi (exact in fp32),
// three kind::tf32 MMAs into the same fp32 TMEM accumulator; the dropped lo*lo term is 2^-22 relative.
//
// One persistent CTA per SM, warp-specialised.  A work item is MT (1 or 2) consecutive 128-row
// tiles of X that share every W chunk brought into shared memory (W is re-streamed from L2 once per work item, so
// MT = 2 halves the L2 traffic that otherwise equals the HBM traffic and caps the kernel at the L2 throughput):
//   warp 0            TMA producer: per stage MT X boxes [128 rows x 32 k] through a 2-D tensor map (SWIZZLE_128B, zero
//                     fill past M / K, L2 evict-first) + the pre-split, pre-arranged W chunk (hi|lo) by one 1-D bulk copy.
//   warps 2..2+4MT    converters: thread <-> one row of one tile = one TMEM lane; reads its 32 floats (conflict-free
//                     through the 128B swizzle), splits hi / lo in registers and stores both into tensor memory
//                     (tcgen05.st): the A operand never goes back to shared memory.
//   warp 1            MMA issuer (one elected lane): per stage and tile 4 k-steps x 3 tcgen05.mma.kind::tf32
//                     (M=128, N=P, K=8; A from TMEM, B = W panels in shared memory, UMMA K-major no-swizzle
//                     descriptors); tcgen05.commit frees the smem stage and the TMEM A stage; accumulators are
//                     double-buffered in TMEM.
//   last 4 warps      epilogue: tcgen05.ld 32 lanes x 16 columns at a time, + bias, leaky-ReLU, row-contiguous stores.
//
// Roofline: HBM (4 M (K + P [+ P]) bytes); the tensor pipe is ~55 % busy at the HBM rate for K = 1152, P = 64.
#include <cuda.h>

#include "mix_kernels.cuh"

namespace ccn {

namespace {

constexpr int BM = 128;         // rows per tile = TMEM lanes = UMMA M
constexpr int BK = 32;          // k per stage (128 bytes per row: one 128B-swizzle atom)
constexpr int kRawBytes = BM * BK * 4;    // 16 KiB: TMA landing zone

struct TcSmemLayout {
    int P, MT, stages, stage_bytes, w_bytes, bar_off, total;
    __host__ __device__ TcSmemLayout(int P_, int MT_, int stages_) : P(P_), MT(MT_), stages(stages_) {
        w_bytes = 2 * BK * P * 4;  // hi | lo
        stage_bytes = MT * kRawBytes + w_bytes;
        bar_off = stages * stage_bytes;
        total = bar_off + 256;
    }
    __host__ __device__ int raw(int s, int t) const { return s * stage_bytes + t * kRawBytes; }
    __host__ __device__ int whi(int s) const { return s * stage_bytes + MT * kRawBytes; }
    __host__ __device__ int wlo(int s) const { return whi(s) + w_bytes / 2; }
};
constexpr int kAStages = 2;        // TMEM ring of split A operands: per stage and tile 32 hi + 32 lo columns
constexpr int kMaxSmemStages = 6;  // barrier array capacity
__host__ __device__ inline int tc_threads(int MT) { return 32 * (2 + 4 * MT + 4); }

struct TcArgs {
    const float *Wprep;  // [chunks][hi|lo][8 panels][Pn][4]
    const float *bias;
    float *Y, *Z;
    int64_t M;
    int K, P, Pn, stages;  // Pn = P rounded up to a multiple of 16 (the MMA's N); columns >= P are zero weights, never stored
    float alpha;
    int tmem_cols;
    const int *items;  // optional work-item list (mix_build_item_list): [0] = count, [1 + k] = item id; null = every item
};

// ---- PTX wrappers --------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] B[smem]: A is [128 lanes x 8 columns] of tf32 (row m in lane m, k along the columns)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 consecutive columns of the calling thread's TMEM lane <- v[0..31]
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
        "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
        "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
        "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor, K-major, no swizzle (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start address,
// leading byte offset (between the two 16-byte core-matrix columns of one K = 8 step), stride byte offset (between
// 8-row groups), all in 16-byte units; version 1 (Blackwell) at bits [46,48).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor (InstrDescriptor): D = F32, A = B = TF32, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ inline uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
    const uint32_t u = __float_as_uint(x);
    hi = __uint_as_float((u + 0x1000u) & 0xffffe000u);  // round to nearest TF32 (10 explicit mantissa bits)
    lo = x - hi;                                        // exact
}

// ---- W preparation: split hi / lo and lay out as UMMA K-major core-matrix panels, one contiguous block per chunk ----
//   Wprep[q][h][j][n][i] = part_h(W[32 q + 4 j + i][n])   (zero past K)
__global__ void __launch_bounds__(256) k_mix_prep_w(const float *__restrict__ W, float *__restrict__ Wprep, int K, int P, int Pn,
                                                    int chunks) {
    const int64_t total = (int64_t)chunks * BK * Pn;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(t & 3);
        const int n = (int)((t >> 2) % Pn);
        const int j = (int)((t >> 2) / Pn % 8);
        const int q = (int)(t / ((int64_t)BK * Pn));
        const int k = q * BK + j * 4 + i;
        const float w = (k < K && n < P) ? W[(int64_t)k * P + n] : 0.f;
        float hi, lo;
        split_tf32(w, hi, lo);
        const int64_t base = (int64_t)q * 2 * BK * Pn;
        const int64_t off = ((int64_t)j * Pn + n) * 4 + i;
        Wprep[base + off] = hi;
        Wprep[base + (int64_t)BK * Pn + off] = lo;
    }
}

template <int MT>
__global__ void __launch_bounds__(32 * (2 + 4 * MT + 4), 1) k_mix_fwd_tc(const __grid_constant__ CUtensorMap tmapX, TcArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const TcSmemLayout L(a.Pn, MT, a.stages);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bar_off);
    uint64_t *full = bars;                          // [stages]    TMA landed (MT X boxes + W chunk)
    uint64_t *empty = bars + kMaxSmemStages;        // [stages]    MMAs reading the stage's W have completed
    uint64_t *conv = bars + 2 * kMaxSmemStages;     // [kAStages]  split A stored to TMEM by the 128 MT converters
    uint64_t *a_empty = conv + kAStages;            // [kAStages]  MMAs reading the TMEM A stage have completed
    uint64_t *acc_full = a_empty + kAStages;        // [2]         accumulators ready for the epilogue
    uint64_t *acc_empty = acc_full + 2;             // [2]         accumulators drained by the 128 epilogue threads
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tiles = (a.M + BM - 1) / BM;
    const int *const ilist = a.items;
    const int64_t items = ilist ? (int64_t)ilist[0] : (tiles + MT - 1) / MT;
    auto item_at = [&](int64_t k) { return ilist ? (int64_t)ilist[1 + k] : k; };
    const int chunks = (a.K + BK - 1) / BK;
    const int stages = a.stages;
    // TMEM columns: accumulators [2 buffers][MT][P], then the A ring [kAStages][MT][hi 32 | lo 32]
    const uint32_t acc_cols = (uint32_t)(MT * a.Pn), a_ring = 2u * acc_cols;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int i = 0; i < kAStages; ++i) {
            mbar_init(&conv[i], 128 * MT);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 128);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            const uint64_t pol_stream = l2_evict_first_policy();
            uint64_t pol_keep;
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
            const uint32_t bytes = (uint32_t)(MT * kRawBytes + L.w_bytes);
            int it = 0;
            for (int64_t ik = blockIdx.x; ik < items; ik += gridDim.x) {
            const int64_t item = item_at(ik);
                for (int q = 0; q < chunks; ++q, ++it) {
                    const int s = it % stages;
                    if (it >= stages) mbar_wait(&empty[s], (uint32_t)((it / stages) - 1) & 1u);
                    mbar_arrive_expect_tx(&full[s], bytes);
#pragma unroll
                    for (int t = 0; t < MT; ++t)  // rows past M are zero-filled by the TMA unit
                        tma_load_2d(smem + L.raw(s, t), &tmapX, q * BK, (int)((item * MT + t) * BM), &full[s], pol_stream);
                    bulk_g2s_hint(smem + L.whi(s), a.Wprep + (int64_t)q * 2 * BK * a.Pn, (uint32_t)L.w_bytes, &full[s], pol_keep);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(BM, a.Pn);
            const uint32_t lbo_b = (uint32_t)a.Pn * 16u;  // W panel [Pn rows x 4 k]
            int it = 0, i_local = 0;
            for (int64_t ik = blockIdx.x; ik < items; ik += gridDim.x, ++i_local) {
            const int64_t item = item_at(ik);
                const int acc = i_local & 1;
                if (i_local >= 2) mbar_wait(&acc_empty[acc], (uint32_t)((i_local >> 1) - 1) & 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_cols;
                for (int q = 0; q < chunks; ++q, ++it) {
                    const int s = it % stages, ar = it % kAStages;
                    mbar_wait(&full[s], (uint32_t)(it / stages) & 1u);     // W chunk (async proxy)
                    mbar_wait(&conv[ar], (uint32_t)(it / kAStages) & 1u);  // split A stored to TMEM stage ar
                    tc_fence_after();
                    const uint32_t whi = smem_u32(smem + L.whi(s)), wlo = smem_u32(smem + L.wlo(s));
#pragma unroll
                    for (int t = 0; t < MT; ++t) {
                        const uint32_t a_hi = tmem_base + a_ring + (uint32_t)((ar * MT + t) * 2 * BK), a_lo = a_hi + BK;
                        const uint32_t d = d_tmem + (uint32_t)(t * a.Pn);
#pragma unroll
                        for (int k8 = 0; k8 < BK / 8; ++k8) {
                            const uint32_t bo = (uint32_t)(k8 * 2) * lbo_b;
                            const uint64_t dbh = umma_desc(whi + bo, lbo_b, 128), dbl = umma_desc(wlo + bo, lbo_b, 128);
                            umma_tf32_ts(d, a_hi + k8 * 8, dbh, idesc, (q | k8) != 0);
                            umma_tf32_ts(d, a_lo + k8 * 8, dbh, idesc, 1u);
                            umma_tf32_ts(d, a_hi + k8 * 8, dbl, idesc, 1u);
                        }
                    }
                    tc_commit(&a_empty[ar]);  // implies tcgen05.fence::before_thread_sync
                    tc_commit(&empty[s]);
                }
                tc_commit(&acc_full[acc]);
            }
        }
    } else if (warp < 2 + 4 * MT) {
        // ===== converters: thread <-> row r of tile t = TMEM lane r (warp w may touch lanes 32 (w % 4) .. + 31) =====
        const int t = (warp - 2) >> 2;
        const int r = (warp & 3) * 32 + lane;
        int it = 0;
        for (int64_t ik = blockIdx.x; ik < items; ik += gridDim.x) {
            const int64_t item = item_at(ik);
            for (int q = 0; q < chunks; ++q, ++it) {
                const int s = it % stages, ar = it % kAStages;
                mbar_wait(&full[s], (uint32_t)(it / stages) & 1u);
                const unsigned char *raw = smem + L.raw(s, t) + r * 128;
                float hi[32], lo[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 v = *reinterpret_cast<const float4 *>(raw + ((c ^ (r & 7)) << 4));
                    split_tf32(v.x, hi[4 * c + 0], lo[4 * c + 0]);
                    split_tf32(v.y, hi[4 * c + 1], lo[4 * c + 1]);
                    split_tf32(v.z, hi[4 * c + 2], lo[4 * c + 2]);
                    split_tf32(v.w, hi[4 * c + 3], lo[4 * c + 3]);
                }
                if (it >= kAStages) mbar_wait(&a_empty[ar], (uint32_t)((it / kAStages) - 1) & 1u);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + a_ring + (uint32_t)((ar * MT + t) * 2 * BK);
                tmem_st32(taddr, hi);
                tmem_st32(taddr + BK, lo);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&conv[ar]);
            }
        }
    } else {
        // ===== epilogue: warp w may touch TMEM lanes 32 (w % 4) .. + 31 =====
        const int quarter = warp & 3;
        const int row_in_tile = quarter * 32 + lane;
        int i_local = 0;
        for (int64_t ik = blockIdx.x; ik < items; ik += gridDim.x, ++i_local) {
            const int64_t item = item_at(ik);
            const int acc = i_local & 1;
            mbar_wait(&acc_full[acc], (uint32_t)(i_local >> 1) & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int t = 0; t < MT; ++t) {
                const int64_t row = (item * MT + t) * BM + row_in_tile;
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * acc_cols + (uint32_t)(t * a.Pn);
                for (int c0 = 0; c0 < a.Pn; c0 += 16) {
                    float v[16];
                    tmem_ld16(taddr + (uint32_t)c0, v);
                    if (row < a.M) {
                        if (a.Y) {  // P % 4 == 0: whole float4s are real columns or padding
                            float4 *dst = reinterpret_cast<float4 *>(a.Y + row * a.P + c0);
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (c0 + 4 * i < a.P) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                        }
                        if (a.Z) {
                            float4 *dst = reinterpret_cast<float4 *>(a.Z + row * a.P + c0);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                if (c0 + 4 * i < a.P) {
                                    const float4 b4 = __ldg(reinterpret_cast<const float4 *>(a.bias + c0) + i);
                                    float4 z = make_float4(v[4 * i] + b4.x, v[4 * i + 1] + b4.y, v[4 * i + 2] + b4.z, v[4 * i + 3] + b4.w);
                                    z.x = z.x > 0.f ? z.x : a.alpha * z.x;
                                    z.y = z.y > 0.f ? z.y : a.alpha * z.y;
                                    z.z = z.z > 0.f ? z.z : a.alpha * z.z;
                                    z.w = z.w > 0.f ? z.w : a.alpha * z.w;
                                    dst[i] = z;
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
}

// ===================================================================================================================
// grad-X on the tensor cores:  gX[M, K] = beta gX + gY[M, P] W^T,   gY = gZ * lrelu'(Y + bias)  (or gZ when bias == null)
//
// Replaces MatMul::backward's first product (MatMul.h:68-82) / kernel MatMul_backward_first (MatMul_gpu.h:71-88) with
// LeakyReLU3D::backward (LeakyReLU3D.h:74-82) and VectorAddTensor::backward's pass-through (VectorAddTensor.h:61-71)
// fused into the operand load.  The reduction dimension is tiny (P <= 64) and the output is wide (K = 18 C), so the
// roles are turned around with respect to the forward kernel: the split gY tile pair stays RESIDENT in tensor memory
// (P/32 x 2 tiles x 64 columns) for the whole work item while W^T streams through shared memory in chunks of 64 output
// columns; every chunk is 2 tiles x P/8 x 3 MMAs (M=128, N=64, K=8) into a double-buffered accumulator that the
// epilogue drains to gX.  HBM roofline: 4 M (K [+ K if beta] + 2 P) bytes, write-dominated.
// ===================================================================================================================
constexpr int GX_NC = 64;   // output columns per chunk
constexpr int GX_MT = 2;    // 128-row tiles per work item
constexpr int GX_WSTAGES = 3;
constexpr int kGxEpiWarps = 4 * GX_MT;  // one epilogue warp per (TMEM lane quarter, tile)
constexpr int kGxThreads = 32 * (2 + 4 * GX_MT + kGxEpiWarps);
constexpr int kGxStageCols = 32;        // columns transposed per pass (two passes per 64-column chunk)
constexpr int kGxRowPad = kGxStageCols + 4;  // floats per staged row: 144 bytes keeps the float4 smem accesses conflict-free

struct GxArgs {
    const float *Wtprep;  // [n-chunks][P/32][hi|lo][8 panels][64][4]
    const float *bias;    // null: gY = gZ
    float *gX;
    float *gY;            // optional [M, P] copy of gY for the grad-W kernel
    int64_t M;
    int K, P;
    float alpha, beta;
    const int *items;     // optional work-item list, as in TcArgs
};

struct GxSmem {
    int P, qn, raw_bytes, w_bytes, w_off, epi_off, bar_off, total;
    __host__ __device__ explicit GxSmem(int P_) : P(P_) {
        qn = (P + BK - 1) / BK;                // 32-column slices of gY (the last one zero-filled past P by the TMA unit)
        raw_bytes = GX_MT * 2 * kRawBytes;     // gZ and Y boxes of both tiles for one 32-column slice
        w_bytes = qn * 2 * BK * GX_NC * 4;     // one chunk of W^T: per slice hi | lo panels
        w_off = raw_bytes;
        epi_off = w_off + GX_WSTAGES * w_bytes;
        bar_off = epi_off + kGxEpiWarps * 32 * kGxRowPad * 4;
        total = bar_off + 256;
    }
};

//   Wtprep[c][q][h][j][n][i] = part_h(W[64 c + n][32 q + 4 j + i])   (zero past K)
__global__ void __launch_bounds__(256) k_mix_prep_wt(const float *__restrict__ W, float *__restrict__ Wtprep, int K, int P, int nchunks) {
    const int qn = (P + BK - 1) / BK;
    const int64_t total = (int64_t)nchunks * qn * BK * GX_NC;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(t & 3);
        const int n = (int)((t >> 2) % GX_NC);
        const int j = (int)((t >> 2) / GX_NC % 8);
        const int q = (int)(t / (BK * GX_NC) % qn);
        const int c = (int)(t / ((int64_t)BK * GX_NC * qn));
        const int k = c * GX_NC + n, pp = q * BK + j * 4 + i;
        const float w = (k < K && pp < P) ? W[(int64_t)k * P + pp] : 0.f;
        float hi, lo;
        split_tf32(w, hi, lo);
        const int64_t base = ((int64_t)c * qn + q) * 2 * BK * GX_NC;
        const int64_t off = ((int64_t)j * GX_NC + n) * 4 + i;
        Wtprep[base + off] = hi;
        Wtprep[base + (int64_t)BK * GX_NC + off] = lo;
    }
}

__global__ void __launch_bounds__(kGxThreads, 1) k_mix_gx_tc(const __grid_constant__ CUtensorMap tmapG, const __grid_constant__ CUtensorMap tmapY,
                                                             GxArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const GxSmem L(a.P);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bar_off);
    uint64_t *raw_full = bars, *raw_empty = bars + 1, *a_full = bars + 2, *a_free = bars + 3;
    uint64_t *w_full = bars + 4, *w_empty = w_full + GX_WSTAGES, *acc_full = w_empty + GX_WSTAGES, *acc_empty = acc_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tiles = (a.M + BM - 1) / BM;
    const int *const ilist = a.items;
    const int64_t items = ilist ? (int64_t)ilist[0] : (tiles + GX_MT - 1) / GX_MT;
    auto item_at = [&](int64_t k) { return ilist ? (int64_t)ilist[1 + k] : k; };
    const int nchunks = (a.K + GX_NC - 1) / GX_NC;
    const int qn = L.qn;
    const bool act = a.bias != nullptr;
    constexpr uint32_t kAccCols = GX_MT * GX_NC;  // per buffer
    constexpr uint32_t kARing = 2 * kAccCols;     // A region starts after the two accumulator buffers

    if (threadIdx.x == 0) {
        mbar_init(raw_full, 1);
        mbar_init(raw_empty, 128 * GX_MT);
        mbar_init(a_full, 128 * GX_MT);
        mbar_init(a_free, 1);
        for (int s = 0; s < GX_WSTAGES; ++s) {
            mbar_init(&w_full[s], 1);
            mbar_init(&w_empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 32 * kGxEpiWarps);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512u);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer: raw gZ / Y slices (one slot) and the W^T chunk ring =====
        if (lane == 0) {
            const uint64_t pol_stream = l2_evict_first_policy();
            uint64_t pol_keep;
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
            int ru = 0, wu = 0;
            auto load_raw = [&](int64_t item) {
                for (int q = 0; q < qn; ++q, ++ru) {
                    if (ru >= 1) mbar_wait(raw_empty, (uint32_t)(ru - 1) & 1u);
                    mbar_arrive_expect_tx(raw_full, (uint32_t)(GX_MT * (act ? 2 : 1) * kRawBytes));
                    for (int t = 0; t < GX_MT; ++t) {
                        const int row0 = (int)((item * GX_MT + t) * BM);
                        tma_load_2d(smem + (t * 2 + 0) * kRawBytes, &tmapG, q * BK, row0, raw_full, pol_stream);
                        if (act) tma_load_2d(smem + (t * 2 + 1) * kRawBytes, &tmapY, q * BK, row0, raw_full, pol_stream);
                    }
                }
            };
            if ((int64_t)blockIdx.x < items) load_raw(item_at(blockIdx.x));
            for (int64_t ik = blockIdx.x; ik < items; ik += gridDim.x) {
            const int64_t item = item_at(ik);
                for (int c = 0; c < nchunks; ++c, ++wu) {
                    const int s = wu % GX_WSTAGES;
                    if (wu >= GX_WSTAGES) mbar_wait(&w_empty[s], (uint32_t)((wu / GX_WSTAGES) - 1) & 1u);
                    mbar_arrive_expect_tx(&w_full[s], (uint32_t)L.w_bytes);
                    bulk_g2s_hint(smem + L.w_off + s * L.w_bytes, a.Wtprep + (int64_t)c * (L.w_bytes / 4), (uint32_t)L.w_bytes, &w_full[s],
                                  pol_keep);
                    // The next item's gZ / Y slices are fetched early: the converters pull them into registers and
                    // only the tensor-memory store waits for this item's last MMA.
                    if (c == 0 && ik + gridDim.x < items) load_raw(item_at(ik + gridDim.x));
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(BM, GX_NC);
            constexpr uint32_t lbo_b = GX_NC * 16u;
            int wu = 0, cc = 0, i_local = 0;
            for (int64_t ik = blockIdx.x; ik < items; ik += gridDim.x, ++i_local) {
            const int64_t item = item_at(ik);
                mbar_wait(a_full, (uint32_t)i_local & 1u);
                tc_fence_after();
                for (int c = 0; c < nchunks; ++c, ++wu, ++cc) {
                    const int s = wu % GX_WSTAGES, buf = cc & 1;
                    mbar_wait(&w_full[s], (uint32_t)(wu / GX_WSTAGES) & 1u);
                    if (cc >= 2) mbar_wait(&acc_empty[buf], (uint32_t)((cc >> 1) - 1) & 1u);
                    tc_fence_after();
                    const uint32_t wbase = smem_u32(smem + L.w_off + s * L.w_bytes);
#pragma unroll
                    for (int t = 0; t < GX_MT; ++t) {
                        const uint32_t d = tmem_base + (uint32_t)buf * kAccCols + (uint32_t)(t * GX_NC);
                        for (int q = 0; q < qn; ++q) {
                            const uint32_t a_hi = tmem_base + kARing + (uint32_t)((q * GX_MT + t) * 2 * BK), a_lo = a_hi + BK;
                            const uint32_t whi = wbase + (uint32_t)(q * 2 * BK * GX_NC * 4), wlo = whi + (uint32_t)(BK * GX_NC * 4);
#pragma unroll
                            for (int k8 = 0; k8 < BK / 8; ++k8) {
                                const uint32_t bo = (uint32_t)(k8 * 2) * lbo_b;
                                const uint64_t dbh = umma_desc(whi + bo, lbo_b, 128), dbl = umma_desc(wlo + bo, lbo_b, 128);
                                umma_tf32_ts(d, a_hi + k8 * 8, dbh, idesc, (q | k8) != 0);
                                umma_tf32_ts(d, a_lo + k8 * 8, dbh, idesc, 1u);
                                umma_tf32_ts(d, a_hi + k8 * 8, dbl, idesc, 1u);
                            }
                        }
                    }
                    tc_commit(&w_empty[s]);
                    tc_commit(&acc_full[buf]);
                }
                tc_commit(a_free);  // every MMA that reads this item's A has completed once this fires
            }
        }
    } else if (warp < 2 + 4 * GX_MT) {
        // ===== converters: gY = gZ * lrelu'(Y + b), split, store to tensor memory (thread <-> row r of tile t) =====
        const int t = (warp - 2) >> 2;
        const int r = (warp & 3) * 32 + lane;
        int ru = 0, i_local = 0;
        for (int64_t ik = blockIdx.x; ik < items; ik += gridDim.x, ++i_local) {
            const int64_t item = item_at(ik);
            const int64_t row = (item * GX_MT + t) * BM + r;
            for (int q = 0; q < qn; ++q, ++ru) {
                mbar_wait(raw_full, (uint32_t)ru & 1u);
                const unsigned char *rg = smem + (t * 2 + 0) * kRawBytes + r * 128;
                const unsigned char *ry = smem + (t * 2 + 1) * kRawBytes + r * 128;
                float hi[32], lo[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float4 g = *reinterpret_cast<const float4 *>(rg + ((c ^ (r & 7)) << 4));
                    const bool real = q * BK + 4 * c < a.P;  // columns past P are TMA zero fill (P % 4 == 0)
                    if (act && real) {
                        const float4 y = *reinterpret_cast<const float4 *>(ry + ((c ^ (r & 7)) << 4));
                        const float4 b = __ldg(reinterpret_cast<const float4 *>(a.bias + q * BK) + c);
                        g.x = (y.x + b.x > 0.f) ? g.x : g.x * a.alpha;
                        g.y = (y.y + b.y > 0.f) ? g.y : g.y * a.alpha;
                        g.z = (y.z + b.z > 0.f) ? g.z : g.z * a.alpha;
                        g.w = (y.w + b.w > 0.f) ? g.w : g.w * a.alpha;
                    }
                    if (a.gY != nullptr && row < a.M && real) *(reinterpret_cast<float4 *>(a.gY + row * a.P + q * BK) + c) = g;
                    split_tf32(g.x, hi[4 * c + 0], lo[4 * c + 0]);
                    split_tf32(g.y, hi[4 * c + 1], lo[4 * c + 1]);
                    split_tf32(g.z, hi[4 * c + 2], lo[4 * c + 2]);
                    split_tf32(g.w, hi[4 * c + 3], lo[4 * c + 3]);
                }
                mbar_arrive(raw_empty);  // the slot's contents are in registers now
                if (q == 0 && i_local >= 1) mbar_wait(a_free, (uint32_t)(i_local - 1) & 1u);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kARing + (uint32_t)((q * GX_MT + t) * 2 * BK);
                tmem_st32(taddr, hi);
                tmem_st32(taddr + BK, lo);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(a_full);
        }
    } else {
        // ===== epilogue: accumulator chunk -> gX[row, 64 c .. 64 c + 63] =====
        // One warp per (lane quarter, tile).  tcgen05.ld hands every lane one ROW; the block is transposed, 32 columns
        // at a time, through a padded shared-memory staging area so that the global stores are coalesced.
        const int ew = warp - (2 + 4 * GX_MT);
        const int quarter = warp & 3, t = ew >> 2;
        float *stage = reinterpret_cast<float *>(smem + L.epi_off) + ew * 32 * kGxRowPad;
        const int sub = lane >> 3, l8 = lane & 7;
        int cc = 0;
        for (int64_t ik = blockIdx.x; ik < items; ik += gridDim.x) {
            const int64_t item = item_at(ik);
            const int64_t row0 = (item * GX_MT + t) * BM + quarter * 32;
            for (int c = 0; c < nchunks; ++c, ++cc) {
                const int buf = cc & 1;
                mbar_wait(&acc_full[buf], (uint32_t)(cc >> 1) & 1u);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)buf * kAccCols + (uint32_t)(t * GX_NC);
#pragma unroll 1
                for (int h = 0; h < GX_NC / kGxStageCols; ++h) {
#pragma unroll
                    for (int c0 = 0; c0 < kGxStageCols; c0 += 16) {
                        float v[16];
                        tmem_ld16(taddr + (uint32_t)(h * kGxStageCols + c0), v);
                        float4 *dst = reinterpret_cast<float4 *>(stage + lane * kGxRowPad + c0);
#pragma unroll
                        for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    }
                    if (h == GX_NC / kGxStageCols - 1) {
                        tc_fence_before();
                        mbar_arrive(&acc_empty[buf]);  // the accumulator has left tensor memory: the MMAs may reuse it
                    }
                    __syncwarp();
                    const int col = c * GX_NC + h * kGxStageCols + l8 * 4;
#pragma unroll 4
                    for (int rr = 0; rr < 32; rr += 4) {  // one store instruction = four complete 128-byte row segments
                        const int rl = rr + sub;
                        const int64_t row = row0 + rl;
                        float4 o = *reinterpret_cast<const float4 *>(stage + rl * kGxRowPad + l8 * 4);
                        if (row < a.M && col < a.K) {  // K % 4 == 0: whole float4s are in or out
                            float4 *g = reinterpret_cast<float4 *>(a.gX + row * a.K + col);
                            if (a.beta != 0.f) {
                                const float4 old = *g;
                                o.x = fmaf(a.beta, old.x, o.x);
                                o.y = fmaf(a.beta, old.y, o.y);
                                o.z = fmaf(a.beta, old.z, o.z);
                                o.w = fmaf(a.beta, old.w, o.w);
                            }
                            __stcs(g, o);
                        }
                    }
                    __syncwarp();  // staging area is reused by the next pass
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

// ===================================================================================================================
// grad-W (+ grad-bias) on the tensor cores:  gW[K, P] += X^T[K, M] gY[M, P],   gbias[P] += column sums of gY
//
// Replaces MatMul::backward's second product (MatMul.h:68-82) / kernel MatMul_backward_second (MatMul_gpu.h:94-111)
// and VectorAddTensor::backward's bias sum (VectorAddTensor.h:61-71).  The reduction runs over the rows m (millions):
// CTA (g, s) owns output rows k in [256 g, 256 g + 256) (two 128-lane accumulators) and the s-th slice of m, walks
// it 32 rows per stage, and adds its partial result to gW with atomics at the end (split-K).  The MMA needs both
// operands with the REDUCTION index contiguous, i.e. X and gY transposed: X^T falls out for free because the A
// operand lives in tensor memory -- thread k reads column k of the [32 x 256] TMA tile (conflict-free) and stores its
// 32 values along its own TMEM lane; the small gY^T operand is written as UMMA K-major panels to shared memory.
// ===================================================================================================================
constexpr int GW_KT = 2;         // 128-row output tiles (of K) per CTA
constexpr int GW_ROWS = 32;      // m rows per stage
constexpr int GW_STAGES = 3;
constexpr int kGwThreads = 32 * (2 + 4 * GW_KT + 2);

struct GwArgs {
    float *gW, *gbias;  // gbias may be null
    int64_t M, rows_per_split;
    int K, P, Pn;  // Pn = P rounded up to a multiple of 16 (the MMA's N); the padding columns of gY^T are zero
};

struct GwSmem {
    int P, x_bytes, g_bytes, b_bytes, stage_bytes, bar_off, total;
    __host__ __device__ GwSmem(int P_, int Pn) : P(P_) {
        x_bytes = GW_ROWS * GW_KT * BM * 4;  // [32 m][256 k]
        g_bytes = GW_ROWS * P * 4;           // [32 m][P]
        b_bytes = 2 * GW_ROWS * Pn * 4;      // hi | lo panels [8][Pn][4]
        stage_bytes = x_bytes + g_bytes + b_bytes;
        bar_off = GW_STAGES * stage_bytes;
        total = bar_off + 256;
    }
};

__global__ void __launch_bounds__(kGwThreads, 1) k_mix_gw_tc(const __grid_constant__ CUtensorMap tmapX, const __grid_constant__ CUtensorMap tmapG,
                                                             GwArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const GwSmem L(a.P, a.Pn);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bar_off);
    uint64_t *full = bars, *empty = full + GW_STAGES, *b_full = empty + GW_STAGES, *a_full = b_full + GW_STAGES;
    uint64_t *a_empty = a_full + kAStages, *acc_full = a_empty + kAStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x, sp = blockIdx.y;
    const int64_t m_begin = (int64_t)sp * a.rows_per_split;
    const int64_t m_end = m_begin + a.rows_per_split < a.M ? m_begin + a.rows_per_split : a.M;
    const int nst = m_end > m_begin ? (int)((m_end - m_begin + GW_ROWS - 1) / GW_ROWS) : 0;
    const int P = a.P, Pn = a.Pn;
    const uint32_t a_ring = (uint32_t)(GW_KT * Pn);  // TMEM: accumulators [GW_KT][P], then A ring [kAStages][GW_KT][hi 32 | lo 32]

    if (threadIdx.x == 0) {
        for (int s = 0; s < GW_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
            mbar_init(&b_full[s], Pn);
        }
        for (int i = 0; i < kAStages; ++i) {
            mbar_init(&a_full[i], 128 * GW_KT);
            mbar_init(&a_empty[i], 1);
        }
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512u);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const uint64_t pol_stream = l2_evict_first_policy();
            uint64_t pol_keep;
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
            for (int it = 0; it < nst; ++it) {
                const int s = it % GW_STAGES;
                if (it >= GW_STAGES) mbar_wait(&empty[s], (uint32_t)((it / GW_STAGES) - 1) & 1u);
                mbar_arrive_expect_tx(&full[s], (uint32_t)(L.x_bytes + L.g_bytes));
                const int row0 = (int)(m_begin + (int64_t)it * GW_ROWS);  // rows past M (or past this slice's end, see below) read as 0
                tma_load_2d(smem + s * L.stage_bytes, &tmapX, g * GW_KT * BM, row0, &full[s], pol_stream);
                tma_load_2d(smem + s * L.stage_bytes + L.x_bytes, &tmapG, 0, row0, &full[s], pol_keep);  // gY is re-read by every g
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(BM, Pn);
            const uint32_t lbo_b = (uint32_t)Pn * 16u;
            for (int it = 0; it < nst; ++it) {
                const int s = it % GW_STAGES, ar = it % kAStages;
                mbar_wait(&b_full[s], (uint32_t)(it / GW_STAGES) & 1u);
                mbar_wait(&a_full[ar], (uint32_t)(it / kAStages) & 1u);
                tc_fence_after();
                const uint32_t bhi = smem_u32(smem + s * L.stage_bytes + L.x_bytes + L.g_bytes), blo = bhi + (uint32_t)(L.b_bytes / 2);
#pragma unroll
                for (int t = 0; t < GW_KT; ++t) {
                    const uint32_t a_hi = tmem_base + a_ring + (uint32_t)((ar * GW_KT + t) * 2 * GW_ROWS), a_lo = a_hi + GW_ROWS;
                    const uint32_t d = tmem_base + (uint32_t)(t * Pn);
#pragma unroll
                    for (int k8 = 0; k8 < GW_ROWS / 8; ++k8) {
                        const uint32_t bo = (uint32_t)(k8 * 2) * lbo_b;
                        const uint64_t dbh = umma_desc(bhi + bo, lbo_b, 128), dbl = umma_desc(blo + bo, lbo_b, 128);
                        umma_tf32_ts(d, a_hi + k8 * 8, dbh, idesc, (it | k8) != 0);
                        umma_tf32_ts(d, a_lo + k8 * 8, dbh, idesc, 1u);
                        umma_tf32_ts(d, a_hi + k8 * 8, dbl, idesc, 1u);
                    }
                }
                tc_commit(&a_empty[ar]);
                tc_commit(&empty[s]);
            }
            tc_commit(acc_full);
        }
    } else if (warp < 2 + 4 * GW_KT) {
        // ===== A converters: thread <-> output row k = column k of the X tile = TMEM lane =====
        const int t = (warp - 2) >> 2, quarter = warp & 3;
        const int kk = t * BM + quarter * 32 + lane;  // column inside the [32 x 256] tile
        for (int it = 0; it < nst; ++it) {
            const int s = it % GW_STAGES, ar = it % kAStages;
            mbar_wait(&full[s], (uint32_t)(it / GW_STAGES) & 1u);
            const float *xs = reinterpret_cast<const float *>(smem + s * L.stage_bytes) + kk;
            const int64_t row0 = m_begin + (int64_t)it * GW_ROWS;
            float hi[32], lo[32];
#pragma unroll
            for (int m = 0; m < GW_ROWS; ++m) {
                const float x = (row0 + m < m_end) ? xs[m * (GW_KT * BM)] : 0.f;  // the slice ends where the next one begins
                split_tf32(x, hi[m], lo[m]);
            }
            if (it >= kAStages) mbar_wait(&a_empty[ar], (uint32_t)((it / kAStages) - 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a_ring + (uint32_t)((ar * GW_KT + t) * 2 * GW_ROWS);
            tmem_st32(taddr, hi);
            tmem_st32(taddr + GW_ROWS, lo);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&a_full[ar]);
        }
        // ===== epilogue (once): accumulator -> gW += (atomics; the other slices of m add into the same rows) =====
        if (nst > 0) {
            mbar_wait(acc_full, 0u);
            tc_fence_after();
            const int k = g * GW_KT * BM + kk;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * Pn);
            for (int c0 = 0; c0 < Pn; c0 += 16) {
                float v[16];
                tmem_ld16(taddr + (uint32_t)c0, v);
                if (k < a.K) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c0 + i < P) atomicAdd(a.gW + (int64_t)k * P + c0 + i, v[i]);
                }
            }
        }
    } else {
        // ===== B converters: thread <-> column p of gY; writes gY^T as UMMA K-major panels (hi | lo) =====
        const int pcol = (warp - (2 + 4 * GW_KT)) * 32 + lane;
        float bsum = 0.f;
        if (pcol < Pn) {
            const bool realp = pcol < P;  // padding columns contribute zeros
            for (int it = 0; it < nst; ++it) {
                const int s = it % GW_STAGES;
                mbar_wait(&full[s], (uint32_t)(it / GW_STAGES) & 1u);
                const float *gs = reinterpret_cast<const float *>(smem + s * L.stage_bytes + L.x_bytes) + pcol;
                unsigned char *bh = smem + s * L.stage_bytes + L.x_bytes + L.g_bytes + pcol * 16, *bl = bh + L.b_bytes / 2;
                const int64_t row0 = m_begin + (int64_t)it * GW_ROWS;
#pragma unroll
                for (int j = 0; j < GW_ROWS / 4; ++j) {
                    float4 h, l;
                    float x0 = (realp && row0 + 4 * j + 0 < m_end) ? gs[(4 * j + 0) * P] : 0.f;
                    float x1 = (realp && row0 + 4 * j + 1 < m_end) ? gs[(4 * j + 1) * P] : 0.f;
                    float x2 = (realp && row0 + 4 * j + 2 < m_end) ? gs[(4 * j + 2) * P] : 0.f;
                    float x3 = (realp && row0 + 4 * j + 3 < m_end) ? gs[(4 * j + 3) * P] : 0.f;
                    bsum += (x0 + x1) + (x2 + x3);
                    split_tf32(x0, h.x, l.x);
                    split_tf32(x1, h.y, l.y);
                    split_tf32(x2, h.z, l.z);
                    split_tf32(x3, h.w, l.w);
                    *reinterpret_cast<float4 *>(bh + j * (Pn * 16)) = h;
                    *reinterpret_cast<float4 *>(bl + j * (Pn * 16)) = l;
                }
                fence_proxy_async();
                mbar_arrive(&b_full[s]);
            }
            if (a.gbias != nullptr && g == 0 && nst > 0 && realp) atomicAdd(a.gbias + pcol, bsum);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        cudaGetLastError();
    }
    return fn;
}

int tiles_per_item(int Pn) { return (2 * 2 * Pn + kAStages * 2 * 2 * BK <= 512) ? 2 : 1; }  // TMEM: 512 columns

int stages_for(int P, int MT) {
    for (int s = kMaxSmemStages; s >= 2; --s)
        if (TcSmemLayout(P, MT, s).total <= 227 * 1024) return s;
    return 0;
}

}  // namespace

bool mix_tc_supported(const float *X, const float *Y, const float *Z, int64_t M, int K, int P) {
    auto al16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    return M > 0 && M < ((int64_t)1 << 31) - 256 && K > 0 && (K % 4) == 0 && P >= 4 && P <= 128 && (P % 4) == 0 && al16(X) &&
           al16(Y) && al16(Z) && stages_for((P + 15) & ~15, 1) > 0 && encode_fn() != nullptr;
}

size_t mix_tc_wprep_bytes(int K, int P) { return (size_t)((K + BK - 1) / BK) * 2 * BK * ((P + 15) & ~15) * sizeof(float); }

cudaError_t mix_tc_configure() {
    cudaError_t e = cudaFuncSetAttribute(k_mix_fwd_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_mix_fwd_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

namespace {

// Rows come in blocks of rows_per_inst per instance of which only the first n_i^2 are real (the level calls: the padding rows of X
// are zero and nobody reads the padding rows of Y / Z / gX).  A work item (rows_per_item consecutive rows) that holds no real row
// is left out of the list: list[0] = number of items kept, list[1..] their ids (any order: items are independent).
__global__ void k_mix_item_list(const int32_t *__restrict__ n_dev, int64_t rows_per_inst, int64_t M, int rows_per_item, int64_t nitems,
                                int *__restrict__ list) {
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= nitems) return;
    const int64_t r0 = item * rows_per_item, r1 = min(M, r0 + (int64_t)rows_per_item);
    bool real = false;
    for (int64_t i = r0 / rows_per_inst; i <= (r1 - 1) / rows_per_inst && !real; ++i) {
        const int64_t lo = i * rows_per_inst, hi = lo + (int64_t)n_dev[i] * n_dev[i];
        real = max(lo, r0) < min(hi, r1);
    }
    if (real) list[1 + atomicAdd(&list[0], 1)] = (int)item;
}

cudaError_t build_item_list(const int32_t *n_dev, int64_t rows_per_inst, int64_t M, int rows_per_item, int64_t nitems, int *list,
                            cudaStream_t st, LaunchLog *log) {
    cudaError_t e = cudaMemsetAsync(list, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    CCN_LAUNCH(log, K_MIX_PREP_W, st,
               k_mix_item_list<<<(unsigned)((nitems + 255) / 256), 256, 0, st>>>(n_dev, rows_per_inst, M, rows_per_item, nitems, list));
    return cudaGetLastError();
}

}  // namespace

size_t mix_item_list_bytes(int64_t M) { return (size_t)((M + BM - 1) / BM + 2) * sizeof(int); }

cudaError_t launch_mix_forward_tc(const float *X, const float *W, const float *bias, float *Y, float *Z, int64_t M, int K,
                                  int P, float alpha, float *wprep, int sm_count, int tiles_per_pass, cudaStream_t st,
                                  LaunchLog *log, const int32_t *rows_n, int64_t rows_per_inst, int *item_buf) {
    const int chunks = (K + BK - 1) / BK;
    const int Pn = (P + 15) & ~15;
    CCN_LAUNCH(log, K_MIX_PREP_W, st, k_mix_prep_w<<<(chunks * BK * Pn + 255) / 256, 256, 0, st>>>(W, wprep, K, P, Pn, chunks));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;

    CUtensorMap tmap;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode_fn()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(X), dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;

    const int64_t tiles = (M + BM - 1) / BM;
    int MT = tiles_per_pass > 0 ? tiles_per_pass : tiles_per_item(Pn);
    if (MT > tiles_per_item(Pn)) MT = tiles_per_item(Pn);
    if (tiles < 2 * (int64_t)sm_count) MT = 1;  // small problems: more, smaller work items
    TcArgs a;
    a.Wprep = wprep;
    a.bias = bias;
    a.Y = Y;
    a.Z = (bias && Z) ? Z : nullptr;
    a.M = M;
    a.K = K;
    a.P = P;
    a.Pn = Pn;
    a.stages = stages_for(Pn, MT);
    a.alpha = alpha;
    const int cols = 2 * MT * Pn + kAStages * MT * 2 * BK;
    a.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
    const TcSmemLayout L(Pn, MT, a.stages);
    const int64_t items = (tiles + MT - 1) / MT;
    const unsigned grid = (unsigned)(items < sm_count ? items : sm_count);
    a.items = nullptr;
    if (rows_n && item_buf) {
        e = build_item_list(rows_n, rows_per_inst, M, MT * BM, items, item_buf, st, log);
        if (e != cudaSuccess) return e;
        a.items = item_buf;
    }
    if (MT == 2)
        CCN_LAUNCH(log, K_MIX_FORWARD_TC, st, k_mix_fwd_tc<2><<<grid, tc_threads(2), L.total, st>>>(tmap, a));
    else
        CCN_LAUNCH(log, K_MIX_FORWARD_TC, st, k_mix_fwd_tc<1><<<grid, tc_threads(1), L.total, st>>>(tmap, a));
    return cudaGetLastError();
}

bool mix_gx_tc_supported(const float *gZ, const float *Y, const float *gX, const float *gYs, int64_t M, int K, int P) {
    auto al16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    return M > 0 && M < ((int64_t)1 << 31) - 256 && K > 0 && (K % 4) == 0 && P >= 4 && P <= 64 && (P % 4) == 0 && al16(gZ) && al16(Y) &&
           al16(gX) && al16(gYs) && encode_fn() != nullptr;
}

size_t mix_gx_tc_wprep_bytes(int K, int P) { return (size_t)((K + GX_NC - 1) / GX_NC) * ((P + BK - 1) / BK) * 2 * BK * GX_NC * sizeof(float); }

cudaError_t mix_gx_tc_configure() {
    return cudaFuncSetAttribute(k_mix_gx_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

cudaError_t launch_mix_grad_x_tc(const float *W, const float *bias, const float *Y, const float *gZ, float *gX, float *gY_out,
                                 int64_t M, int K, int P, float alpha, float beta_x, float *wtprep, int sm_count, cudaStream_t st,
                                 LaunchLog *log, const int32_t *rows_n, int64_t rows_per_inst, int *item_buf) {
    const int nchunks = (K + GX_NC - 1) / GX_NC;
    CCN_LAUNCH(log, K_MIX_PREP_W, st, k_mix_prep_wt<<<(nchunks * ((P + BK - 1) / BK) * BK * GX_NC + 255) / 256, 256, 0, st>>>(W, wtprep, K, P, nchunks));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    CUtensorMap tg, ty;
    const cuuint64_t dims[2] = {(cuuint64_t)P, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)P * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    const cuuint32_t estr[2] = {1, 1};
    for (int i = 0; i < 2; ++i) {
        const float *src = i == 0 ? gZ : (bias ? Y : gZ);
        const CUresult r = encode_fn()(i == 0 ? &tg : &ty, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(src), dims, strides, box,
                                       estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    }
    GxArgs a;
    a.Wtprep = wtprep;
    a.bias = bias;
    a.gX = gX;
    a.gY = gY_out;
    a.M = M;
    a.K = K;
    a.P = P;
    a.alpha = alpha;
    a.beta = beta_x;
    const GxSmem L(P);
    const int64_t items = ((M + BM - 1) / BM + GX_MT - 1) / GX_MT;
    const unsigned grid = (unsigned)(items < sm_count ? items : sm_count);
    a.items = nullptr;
    if (rows_n && item_buf) {
        if (gY_out) {  // grad-W reads every row of gY: the rows of the skipped items must read as zero
            e = cudaMemsetAsync(gY_out, 0, (size_t)M * P * sizeof(float), st);
            if (e != cudaSuccess) return e;
        }
        e = build_item_list(rows_n, rows_per_inst, M, GX_MT * BM, items, item_buf, st, log);
        if (e != cudaSuccess) return e;
        a.items = item_buf;
    }
    CCN_LAUNCH(log, K_MIX_GRAD_X_TC, st, k_mix_gx_tc<<<grid, kGxThreads, L.total, st>>>(tg, ty, a));
    return cudaGetLastError();
}

bool mix_gw_tc_supported(const float *X, const float *gY, const float *gW, int64_t M, int K, int P) {
    auto al16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    return M >= 4096 && M < ((int64_t)1 << 31) - 256 && K > 0 && (K % 4) == 0 && P >= 4 && P <= 64 && (P % 4) == 0 && al16(X) &&
           al16(gY) && al16(gW) && encode_fn() != nullptr;
}

cudaError_t mix_gw_tc_configure() {
    return cudaFuncSetAttribute(k_mix_gw_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

// gW += X^T gY, gbias += colsum(gY) (gbias may be null).  gY is the activation-corrected gradient [M, P].
cudaError_t launch_mix_grad_w_tc(const float *X, const float *gY, float *gW, float *gbias, int64_t M, int K, int P, int sm_count,
                                 cudaStream_t st, LaunchLog *log) {
    CUtensorMap tx, tg;
    const cuuint32_t estr[2] = {1, 1};
    {
        const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
        const cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(float)};
        const cuuint32_t box[2] = {(cuuint32_t)(GW_KT * BM), (cuuint32_t)GW_ROWS};
        if (encode_fn()(&tx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(X), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return cudaErrorInvalidValue;
    }
    {
        const cuuint64_t dims[2] = {(cuuint64_t)P, (cuuint64_t)M};
        const cuuint64_t strides[1] = {(cuuint64_t)P * sizeof(float)};
        const cuuint32_t box[2] = {(cuuint32_t)P, (cuuint32_t)GW_ROWS};
        if (encode_fn()(&tg, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(gY), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return cudaErrorInvalidValue;
    }
    const int groups = (K + GW_KT * BM - 1) / (GW_KT * BM);
    int splits = sm_count / groups;
    if (splits < 1) splits = 1;
    int64_t rows = (M + splits - 1) / splits;
    rows = ((rows + GW_ROWS - 1) / GW_ROWS) * GW_ROWS;
    splits = (int)((M + rows - 1) / rows);
    GwArgs a;
    a.gW = gW;
    a.gbias = gbias;
    a.M = M;
    a.rows_per_split = rows;
    a.K = K;
    a.P = P;
    a.Pn = (P + 15) & ~15;
    const GwSmem L(P, a.Pn);
    CCN_LAUNCH(log, K_MIX_GRAD_W_TC, st, (k_mix_gw_tc<<<dim3(groups, splits), kGwThreads, L.total, st>>>(tx, tg, a)));
    return cudaGetLastError();
}

}  // namespace ccn
