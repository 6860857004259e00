// ccn_common.cuh -- shared device/host definitions for the sm_100a CCN kernels.
//
// Vocabulary (the reference's): an *instance* is one RisiContraction_18 call = one (graph, vertex, level) with a
// receptive field of n vertices and C channels; T is the stacked neighbour tensor [n,n,n,C]; adj the reduced
// adjacency [n,n]; out the 18 contraction slabs [n,n,18C] (GraphFlow/RisiContraction_18.h:73-331).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

namespace ccn {

constexpr int kSlabs = 18;

// Kernel ids (ccn_kernel_name() in the C-ABI gives the strings) and the per-call launch log: counts launches and,
// when the context's kernel timing is on, brackets every launch with CUDA events on the launching stream.
enum KernelId {
    K_ADJ_PREPARE = 0,
    K_FWD_STREAM,
    K_FWD_FINISH,
    K_BWD_PLANES,
    K_BWD_STREAM,
    K_GEN_FWD_PLANES,
    K_GEN_FWD_SUMS,
    K_GEN_FWD_OUT,
    K_GEN_BWD_VECTORS,
    K_GEN_BWD_PLANES,
    K_GEN_BWD_SCATTER,
    K_MIX_FORWARD,
    K_MIX_GRAD_X,
    K_MIX_GRAD_W,
    K_MIX_GRAD_BIAS,
    K_FWD_FUSED,
    K_BWD_FUSED,
    K_MIX_PREP_W,
    K_MIX_FORWARD_TC,
    K_R50_ADJ,
    K_R50_FWD_PLANES,
    K_R50_FWD_VECTORS,
    K_R50_FWD_OUT,
    K_R50_BWD_VECTORS,
    K_R50_BWD_PLANES,
    K_R50_BWD_SCATTER,
    K_PROMOTE_FWD,
    K_PROMOTE_BWD,
    K_TENSOR_MUL,
    K_TRANSPOSE,
    K_MIX_GRAD_X_TC,
    K_MIX_GRAD_W_TC,
    K_OPTIMIZER,
    K_FWD_FUSED_GATHER,
    K_BWD_FUSED_SCATTER,
    K_READOUT,
    K_COUNT
};

struct LaunchRecord {
    int id;
    cudaEvent_t start, stop;
};

struct LaunchLog {
    int launches = 0;
    bool timing = false;
    std::vector<LaunchRecord> *records = nullptr;  // owned by the context
    std::vector<cudaEvent_t> *pool = nullptr;      // recycled events
    cudaEvent_t take() {
        cudaEvent_t e = nullptr;
        if (pool && !pool->empty()) {
            e = pool->back();
            pool->pop_back();
        } else {
            cudaEventCreate(&e);
        }
        return e;
    }
    void pre(int id, cudaStream_t st) {
        ++launches;
        if (!timing) return;
        LaunchRecord r{id, take(), take()};
        cudaEventRecord(r.start, st);
        records->push_back(r);
    }
    void post(cudaStream_t st) {
        if (timing) cudaEventRecord(records->back().stop, st);
    }
};

#define CCN_LAUNCH(log, id, st, ...) \
    do {                             \
        (log)->pre((id), (st));      \
        __VA_ARGS__;                 \
        (log)->post((st));           \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------
// Adjacency table: per instance, built once per call by k_adj_prepare and read by every other kernel.
//   A[n*n]      effective adjacency (positive part or raw), dense with row stride n
//   r[nm]       row sums  r[d] = sum_e A[d,e]
//   scal[4]     sA = sum r, tr = trace, nnz (as float), unused
//   rowptr[nm+1], rowidx[nm*nm], rowval[nm*nm]   CSR of A      : row d lists (e, A[d,e]) with A[d,e] != 0
//   colptr[nm+1], colidx[nm*nm], colval[nm*nm]   CSC of A      : column b lists (d, A[d,b]) with A[d,b] != 0
// All offsets are in 4-byte words from the start of the instance's table; nm = n_max of the call.
// ---------------------------------------------------------------------------------------------------------------
struct AdjTabLayout {
    int nm;
    __host__ __device__ int A() const { return 0; }
    __host__ __device__ int r() const { return nm * nm; }
    __host__ __device__ int scal() const { return r() + nm; }
    __host__ __device__ int rowptr() const { return scal() + 4; }
    __host__ __device__ int colptr() const { return rowptr() + nm + 1; }
    __host__ __device__ int rowidx() const { return colptr() + nm + 1; }
    __host__ __device__ int rowval() const { return rowidx() + nm * nm; }
    __host__ __device__ int colidx() const { return rowval() + nm * nm; }
    __host__ __device__ int colval() const { return colidx() + nm * nm; }
    __host__ __device__ int words() const { return ((colval() + nm * nm) + 3) & ~3; }  // keep 16-byte multiples
};

struct AdjView {
    const float *A, *r, *scal, *rowval, *colval;
    const int *rowptr, *colptr, *rowidx, *colidx;
};

__host__ __device__ inline AdjView adj_view(const float *tab, int nm) {
    AdjTabLayout L{nm};
    AdjView v;
    v.A = tab + L.A();
    v.r = tab + L.r();
    v.scal = tab + L.scal();
    v.rowptr = reinterpret_cast<const int *>(tab + L.rowptr());
    v.colptr = reinterpret_cast<const int *>(tab + L.colptr());
    v.rowidx = reinterpret_cast<const int *>(tab + L.rowidx());
    v.rowval = tab + L.rowval();
    v.colidx = reinterpret_cast<const int *>(tab + L.colidx());
    v.colval = tab + L.colval();
    return v;
}

// One batch (or chunk of a batch) of instances as the kernels see it.  Pointers are already offset to the
// first instance of the chunk.
struct Batch {
    const int32_t *n;  // per-instance n (device) or nullptr -> n_max
    int n_max;
    int C;
    int count;  // instances in this chunk
    __device__ int n_of(int inst) const { return n ? n[inst] : n_max; }
};

// ---------------------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP), sm_90+/sm_100a.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// A barrier that never completes is a bug (a faulted bulk copy, a wrong transaction count): after ~4 s of SM clocks the
// kernel traps, which surfaces as a CUDA error at the next API call instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000ll) asm volatile("trap;");
    }
}

// global -> shared bulk copy; bytes and both addresses must be multiples of 16.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// same, with an L2 cache policy (evict-first for data that is streamed exactly once)
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// shared -> global bulk copy (TMA store), tracked with bulk_group commit/wait.
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// streaming (evict-first) global store / load for data touched exactly once
__device__ __forceinline__ void st_stream(float *p, float v) { __stcs(p, v); }
__device__ __forceinline__ float ld_stream(const float *p) { return __ldcs(p); }

}  // namespace ccn
