// ccn_abi.cu -- the C-ABI of include/ccn_b200.h: context, workspace, batching/chunking, host-buffer pipelines.
// No kernel code lives here; see contract18_fused.cu, contract18_generic.cu, contract50.cu, mix_*.cu, aux_ops.cu.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/ccn_b200.h"
#include "contract18_kernels.cuh"
#include "mix_kernels.cuh"

using namespace ccn;

struct ccn_ctx {
    int device = 0;
    size_t ws_limit = (size_t)96 << 20;
    float *ws = nullptr;
    size_t ws_bytes = 0;
    int64_t launches = 0;
    int path = 0;  // CCN_PATH_AUTO / GENERIC / TILED
    int mix_path = 0;  // CCN_MIX_AUTO / SIMT / TENSOR
    int mix_tiles_per_pass = std::getenv("CCN_MIX_TILES") ? std::atoi(std::getenv("CCN_MIX_TILES")) : 0;  // A/B switch (0 = auto)
    int fused_variant = std::getenv("CCN_FUSED_VARIANT") ? std::atoi(std::getenv("CCN_FUSED_VARIANT")) : -1;  // A/B switch (-1 = default)
    float *wprep = nullptr;  // tensor-core mix: split + pre-arranged weights
    float *gybuf = nullptr;  // tensor-core mix backward: gY = gZ * lrelu'(Y + b), written by grad-X, read by grad-W
    size_t gybuf_bytes = 0;
    // set by the level calls around their mix: the mix rows are blocks of mix_rows_per_inst rows per instance with
    // mix_rows_n[i]^2 real rows each; the tensor-core kernels then skip the work items that hold no real row
    const int32_t *mix_rows_n = nullptr;
    int64_t mix_rows_per_inst = 0;
    int *mix_items = nullptr;
    size_t mix_items_bytes = 0;
    float *aux = nullptr;    // CustomMatMulTensor: transposed weights and their gradient
    size_t aux_bytes = 0;
    size_t wprep_bytes = 0;
    int sm_count = 148;
    int *ctl = nullptr;  // fused path control block (ticket + per-slot counters)
    size_t ctl_bytes = 0;
    unsigned long long *trace = nullptr;  // caller-owned device buffer, 8 words per fused-path tile
    int64_t trace_tiles = 0;
    std::string err;
    // Sticky device-side failure flag: a word of mapped pinned host memory that a fused-path tile sets when it gives up
    // waiting for its siblings.  Never cleared by a launch; every later entry point returns CCN_ERR_CUDA while it is set.
    int *fault_host = nullptr;  // host view
    int *fault_dev = nullptr;   // device view of the same word
    // One stream at a time: the context's scratch (ws, ctl, wprep, gybuf, aux) is shared by all calls, so a call on a
    // different stream than the previous one first waits (on the device) for the previous call's work.
    cudaStream_t last_stream = nullptr;
    bool last_stream_valid = false;
    cudaEvent_t scratch_done = nullptr;
    bool frozen = false;  // no buffer may grow (set while a CUDA graph that baked the pointers in is alive)
    // host-buffer pipeline (created lazily)
    static constexpr int kSlots = 3;
    cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[kSlots] = {}, ev_comp[kSlots] = {}, ev_out[kSlots] = {};
    float *stage = nullptr;  // device staging for all slots
    size_t stage_bytes = 0;
    // optional per-kernel timing (CUDA events on the launching stream)
    bool timing = false;
    std::vector<LaunchRecord> records;
    std::vector<cudaEvent_t> event_pool;
    double kernel_ms[K_COUNT] = {};
    int64_t kernel_launches[K_COUNT] = {};
};

namespace {

int fail(ccn_ctx *ctx, int status, const std::string &msg) {
    if (ctx) ctx->err = msg;
    return status;
}

int cuda_fail(ccn_ctx *ctx, cudaError_t e, const char *what) {
    return fail(ctx, e == cudaErrorMemoryAllocation ? CCN_ERR_OUT_OF_MEMORY : CCN_ERR_CUDA,
                std::string(what) + ": " + cudaGetErrorString(e));
}

#define CCN_CUDA(ctx, call)                                        \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Grows one of the context-owned device buffers.  Growth synchronises the device (earlier launches may still use the
// old block) and is refused while the context is frozen (ccn_ctx_set_frozen: a captured CUDA graph holds the pointers).
template <class T>
int grow_buffer(ccn_ctx *ctx, T **ptr, size_t *have, size_t need, const char *what) {
    if (*have >= need) return CCN_OK;
    if (ctx->frozen)
        return fail(ctx, CCN_ERR_UNSUPPORTED,
                    std::string("the context is frozen (a captured graph holds its buffers) and ") + what + " would have to grow");
    if (*ptr) {
        CCN_CUDA(ctx, cudaDeviceSynchronize());
        cudaFree(*ptr);
        *ptr = nullptr;
        *have = 0;
    }
    CCN_CUDA(ctx, cudaMalloc(ptr, need));
    *have = need;
    return CCN_OK;
}

int ensure_workspace(ccn_ctx *ctx, size_t bytes) { return grow_buffer(ctx, &ctx->ws, &ctx->ws_bytes, bytes, "the workspace"); }

// Entry / exit of every call that touches context scratch (see ccn_ctx::last_stream).  Skipped on a capturing stream:
// a capture owns its context and events recorded inside a capture cannot be waited on outside.
int scratch_enter(ccn_ctx *ctx, cudaStream_t st) {
    if (ctx->fault_host && *(volatile int *)ctx->fault_host != 0)
        return fail(ctx, CCN_ERR_CUDA,
                    "an earlier fused-path launch on this context timed out waiting for sibling tiles; its results are invalid "
                    "(ccn_ctx_fused_error_flag(ctx, &flag) reports and clears the condition)");
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) cudaGetLastError();
    if (cs != cudaStreamCaptureStatusNone) return CCN_OK;
    if (ctx->last_stream_valid && ctx->last_stream != st) CCN_CUDA(ctx, cudaStreamWaitEvent(st, ctx->scratch_done, 0));
    return CCN_OK;
}

int scratch_leave(ccn_ctx *ctx, cudaStream_t st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) cudaGetLastError();
    if (cs != cudaStreamCaptureStatusNone) return CCN_OK;
    CCN_CUDA(ctx, cudaEventRecord(ctx->scratch_done, st));
    ctx->last_stream = st;
    ctx->last_stream_valid = true;
    return CCN_OK;
}

// RAII form: `ScratchScope sc(ctx, st); if (sc.rc != CCN_OK) return sc.rc;` records the event on every exit path.
struct ScratchScope {
    ccn_ctx *ctx;
    cudaStream_t st;
    int rc;
    ScratchScope(ccn_ctx *c, cudaStream_t s) : ctx(c), st(s), rc(scratch_enter(c, s)) {}
    ~ScratchScope() {
        if (rc == CCN_OK) scratch_leave(ctx, st);
    }
};

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

LaunchLog make_log(ccn_ctx *ctx) {
    LaunchLog log;
    log.timing = ctx->timing;
    log.records = &ctx->records;
    log.pool = &ctx->event_pool;
    return log;
}

// Fold finished launch records into the per-kernel totals (synchronises on the recorded events).
void drain_records(ccn_ctx *ctx) {
    for (const LaunchRecord &r : ctx->records) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.stop) == cudaSuccess && cudaEventElapsedTime(&ms, r.start, r.stop) == cudaSuccess) {
            ctx->kernel_ms[r.id] += ms;
            ctx->kernel_launches[r.id] += 1;
        }
        ctx->event_pool.push_back(r.start);
        ctx->event_pool.push_back(r.stop);
    }
    ctx->records.clear();
    cudaGetLastError();
}

struct Plan {
    int adj_words;
    int64_t scratch_words;
    int64_t chunk;
};

// How many instances fit the workspace limit at once (at least 1; generic kernels index instances by blockIdx.y).
Plan make_plan(const ccn_ctx *ctx, int n_max, int C, int64_t batch, bool backward) {
    Plan p;
    p.adj_words = AdjTabLayout{n_max}.words();
    p.scratch_words = backward ? generic_bwd_scratch_words(n_max, C) : generic_fwd_scratch_words(n_max, C);
    const size_t per = (size_t)(p.adj_words + p.scratch_words) * 4;
    int64_t chunk = (int64_t)std::max<size_t>(1, ctx->ws_limit / per);
    chunk = std::min<int64_t>(chunk, batch);
    chunk = std::min<int64_t>(chunk, (int64_t)65535);
    p.chunk = chunk;
    return p;
}

// Fused path: scratch slots are recycled while they are still in L2.  Twice as many slots as instances that can be
// resident at once (2 CTAs per SM), so a tile practically never waits for its slot.
struct FusedPlan {
    int slots;
    int64_t scratch_words;
};

// geometry: 0 forward, 1 backward, 2 backward with fused promotion (see contract18_kernels.cuh)
int fused_prepare(ccn_ctx *ctx, int n_max, int C, int64_t batch, int geometry, FusedPlan *p) {
    const bool backward = geometry != 0;
    const int tiles = fused_tiles(n_max, C, geometry);
    const int resident = geometry == 1 ? fused_resident_ctas_bwd() : fused_resident_ctas_fwd();
    int64_t slots = 2 * ((resident * (int64_t)ctx->sm_count + tiles - 1) / tiles);
    slots = std::max<int64_t>(8, slots);
    slots = std::min<int64_t>(slots, batch);
    p->slots = (int)slots;
    p->scratch_words = geometry == 1 ? fused_bwd_scratch_words(n_max, C)
                                     : geometry == 2 ? fused_bwd_scatter_scratch_words(n_max, C) : fused_fwd_scratch_words(n_max, C);
    (void)backward;
    int rc = ensure_workspace(ctx, (size_t)slots * p->scratch_words * 4);
    if (rc != CCN_OK) return rc;
    return grow_buffer(ctx, &ctx->ctl, &ctx->ctl_bytes, (size_t)fused_ctl_words(p->slots) * sizeof(int), "the fused-path control block");
}

int check_common(ccn_ctx *ctx, const void *adj, int n_max, int C, int64_t batch, int adj_mode) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    if (!adj) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "adj_dev is NULL");
    if (n_max <= 0 || C <= 0 || batch < 0) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "n_max, C must be > 0 and batch >= 0");
    if (adj_mode != CCN_ADJ_POSITIVE_PART && adj_mode != CCN_ADJ_RAW)
        return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "adj_mode must be CCN_ADJ_POSITIVE_PART or CCN_ADJ_RAW");
    if ((int64_t)n_max * n_max > (int64_t)1 << 24) return fail(ctx, CCN_ERR_UNSUPPORTED, "n_max too large");
    if ((int64_t)n_max * n_max * n_max * C >= (int64_t)1 << 31)
        return fail(ctx, CCN_ERR_UNSUPPORTED, "one instance (n_max^3 * C elements) must stay below 2^31 elements");
    return CCN_OK;
}

}  // namespace

extern "C" {

int ccn_abi_version(void) { return CCN_B200_ABI_VERSION; }

const char *ccn_status_string(int status) {
    switch (status) {
        case CCN_OK: return "ok";
        case CCN_ERR_INVALID_ARGUMENT: return "invalid argument";
        case CCN_ERR_CUDA: return "CUDA error";
        case CCN_ERR_OUT_OF_MEMORY: return "out of memory";
        case CCN_ERR_NO_DEVICE: return "no usable sm_100 device";
        case CCN_ERR_UNSUPPORTED: return "unsupported shape";
    }
    return "unknown status";
}

const char *ccn_last_error(const ccn_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int ccn_ctx_create(ccn_ctx **out, int device) {
    if (!out) return CCN_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return CCN_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return CCN_ERR_NO_DEVICE;
    if (prop.major != 10) return CCN_ERR_NO_DEVICE;  // the kernels are built for sm_100a only
    ccn_ctx *ctx = new (std::nothrow) ccn_ctx();
    if (!ctx) return CCN_ERR_OUT_OF_MEMORY;
    ctx->device = device;
    DeviceGuard g(device);
    ctx->sm_count = prop.multiProcessorCount;
    cudaError_t e = fused_path_configure_fwd();
    if (e == cudaSuccess) e = fused_path_configure_bwd();
    if (e == cudaSuccess) e = mix_configure();
    if (e == cudaSuccess) e = mix_tc_configure();
    if (e == cudaSuccess) e = r50_configure();
    if (e == cudaSuccess) e = mix_gx_tc_configure();
    if (e == cudaSuccess) e = mix_gw_tc_configure();
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->scratch_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void **>(&ctx->fault_host), 64, cudaHostAllocMapped);
    if (e == cudaSuccess) {
        *ctx->fault_host = 0;
        e = cudaHostGetDevicePointer(reinterpret_cast<void **>(&ctx->fault_dev), ctx->fault_host, 0);
    }
    if (e != cudaSuccess) {
        if (ctx->scratch_done) cudaEventDestroy(ctx->scratch_done);
        if (ctx->fault_host) cudaFreeHost(ctx->fault_host);
        delete ctx;
        cudaGetLastError();
        return CCN_ERR_CUDA;
    }
    *out = ctx;
    return CCN_OK;
}

int ccn_ctx_destroy(ccn_ctx *ctx) {
    if (!ctx) return CCN_OK;
    DeviceGuard g(ctx->device);
    cudaDeviceSynchronize();
    drain_records(ctx);
    for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->ctl) cudaFree(ctx->ctl);
    if (ctx->wprep) cudaFree(ctx->wprep);
    if (ctx->aux) cudaFree(ctx->aux);
    if (ctx->gybuf) cudaFree(ctx->gybuf);
    if (ctx->mix_items) cudaFree(ctx->mix_items);
    if (ctx->stage) cudaFree(ctx->stage);
    if (ctx->scratch_done) cudaEventDestroy(ctx->scratch_done);
    if (ctx->fault_host) cudaFreeHost(ctx->fault_host);
    for (int i = 0; i < ccn_ctx::kSlots; ++i) {
        if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
        if (ctx->ev_comp[i]) cudaEventDestroy(ctx->ev_comp[i]);
        if (ctx->ev_out[i]) cudaEventDestroy(ctx->ev_out[i]);
    }
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_comp) cudaStreamDestroy(ctx->s_comp);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    delete ctx;
    return CCN_OK;
}

int ccn_ctx_set_workspace_limit(ccn_ctx *ctx, size_t bytes) {
    if (!ctx || bytes == 0) return CCN_ERR_INVALID_ARGUMENT;
    ctx->ws_limit = bytes;
    return CCN_OK;
}

int64_t ccn_ctx_kernel_launches(const ccn_ctx *ctx) { return ctx ? ctx->launches : 0; }

int ccn_num_kernels(void) { return K_COUNT; }

const char *ccn_kernel_name(int kernel_id) {
    static const char *names[K_COUNT] = {"adj_prepare",     "fwd_stream",      "fwd_finish",      "bwd_planes",
                                         "bwd_stream",      "gen_fwd_planes",  "gen_fwd_sums",    "gen_fwd_out",
                                         "gen_bwd_vectors", "gen_bwd_planes",  "gen_bwd_scatter", "mix_forward",
                                         "mix_grad_x",      "mix_grad_w",      "mix_grad_bias",   "fwd_fused",
                                         "bwd_fused",       "mix_prep_w",      "mix_forward_tc",
                                         "r50_adj",         "r50_fwd_planes",  "r50_fwd_vectors", "r50_fwd_out",
                                         "r50_bwd_vectors", "r50_bwd_planes",  "r50_bwd_scatter", "promote_fwd",
                                         "promote_bwd",     "tensor_mul",      "transpose",
                                         "mix_grad_x_tc",   "mix_grad_w_tc",   "optimizer",
                                         "fwd_fused_gather", "bwd_fused_scatter", "readout"};
    return (kernel_id >= 0 && kernel_id < K_COUNT) ? names[kernel_id] : "?";
}

int ccn_ctx_set_kernel_timing(ccn_ctx *ctx, int enable) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    DeviceGuard g(ctx->device);
    drain_records(ctx);
    for (int i = 0; i < K_COUNT; ++i) {
        ctx->kernel_ms[i] = 0.0;
        ctx->kernel_launches[i] = 0;
    }
    ctx->timing = enable != 0;
    return CCN_OK;
}

int ccn_ctx_get_kernel_timing(ccn_ctx *ctx, int kernel_id, double *total_ms, int64_t *launches) {
    if (!ctx || kernel_id < 0 || kernel_id >= K_COUNT) return CCN_ERR_INVALID_ARGUMENT;
    DeviceGuard g(ctx->device);
    drain_records(ctx);
    if (total_ms) *total_ms = ctx->kernel_ms[kernel_id];
    if (launches) *launches = ctx->kernel_launches[kernel_id];
    return CCN_OK;
}

int ccn_ctx_set_kernel_path(ccn_ctx *ctx, int path) {
    if (!ctx || path < CCN_PATH_AUTO || path > CCN_PATH_GENERIC) return CCN_ERR_INVALID_ARGUMENT;
    ctx->path = path;
    return CCN_OK;
}

int ccn_ctx_set_mix_path(ccn_ctx *ctx, int path) {
    if (!ctx || path < CCN_MIX_AUTO || path > CCN_MIX_TENSOR) return CCN_ERR_INVALID_ARGUMENT;
    ctx->mix_path = path;
    return CCN_OK;
}

int ccn_ctx_set_frozen(ccn_ctx *ctx, int frozen) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    ctx->frozen = frozen != 0;
    return CCN_OK;
}

int ccn_ctx_set_phase_trace(ccn_ctx *ctx, void *trace_dev, int64_t tiles) {
    if (!ctx || tiles < 0) return CCN_ERR_INVALID_ARGUMENT;
    ctx->trace = static_cast<unsigned long long *>(trace_dev);
    ctx->trace_tiles = trace_dev ? tiles : 0;
    return CCN_OK;
}

int ccn_ctx_fused_error_flag(ccn_ctx *ctx, int *flag) {
    if (!ctx || !flag) return CCN_ERR_INVALID_ARGUMENT;
    *flag = 0;
    DeviceGuard g(ctx->device);
    CCN_CUDA(ctx, cudaDeviceSynchronize());
    *flag = *(volatile int *)ctx->fault_host;
    *(volatile int *)ctx->fault_host = 0;  // reported: the caller decides what to recompute
    return CCN_OK;
}

}  // extern "C"

namespace {

// Can the promotion be read inside the fused kernels?  (shape supported, 16-byte pieces: C % 4 == 0, aligned base; the
// offsets are multiples of C by contract, include/ccn_b200.h ccn_promote_forward)
bool gather_fusable(const ccn_ctx *ctx, const void *f_dev, int n_max, int C, int64_t batch) {
    return ctx->path == CCN_PATH_AUTO && fused_path_supported(n_max, C) && aligned16(f_dev) && (C % 4) == 0 &&
           batch * fused_tiles(n_max, C, 1) < (int64_t)1 << 31;
}

// ccn_contract18_forward with an optional fused promotion (G != nullptr: the input is gathered from f_{l-1}, and the
// caller has checked gather_fusable).
// keep / out_scale: the slab mask and scale of RisiContraction_18_dropout, applied inside the fused kernels; *masked_by_kernel
// reports whether they were (the generic kernels ignore them and the caller runs the mask passes instead).
int contract18_forward_impl(ccn_ctx *ctx, const float *T_dev, const float *const *slabs_dev, const GatherRef *G,
                            const float *adj_dev, float *out_dev, const int32_t *n_dev, int n_max, int C, int64_t batch,
                            int64_t stride_T, int64_t stride_adj, int64_t stride_out, int adj_mode, void *stream,
                            uint32_t keep = 0x3ffffu, float out_scale = 1.f, bool *masked_by_kernel = nullptr) {
    if (masked_by_kernel) *masked_by_kernel = false;
    int rc = check_common(ctx, adj_dev, n_max, C, batch, adj_mode);
    if (rc != CCN_OK) return rc;
    if (!G && (T_dev == nullptr) == (slabs_dev == nullptr))
        return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "exactly one of T_dev / slabs_dev must be given");
    if (!out_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "out_dev is NULL");
    if (batch == 0) return CCN_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ScratchScope scope(ctx, st);
    if (scope.rc != CCN_OK) return scope.rc;
    bool fused = ctx->path == CCN_PATH_AUTO && fused_path_supported(n_max, C);
    // bulk copies need 16-byte alignment (a slab-pointer table cannot be checked here: its entries must be 16-byte aligned)
    if (!G && T_dev && (!aligned16(T_dev) || (stride_T & 3) != 0)) fused = false;
    if (G && !fused) return fail(ctx, CCN_ERR_UNSUPPORTED, "fused promotion needs a shape of the fused kernels");
    if (fused && batch * fused_tiles(n_max, C, 0) < (int64_t)1 << 31) {
        FusedPlan fp;
        rc = fused_prepare(ctx, n_max, C, batch, 0, &fp);
        if (rc != CCN_OK) return rc;
        Fused18Fwd a;
        a.G = G ? *G : GatherRef{nullptr, nullptr, nullptr, nullptr};
        a.T.base = const_cast<float *>(T_dev);
        a.T.slabs = const_cast<float *const *>(slabs_dev);
        a.T.stride = stride_T;
        a.out = out_dev;
        a.stride_out = stride_out;
        a.adj = adj_dev;
        a.stride_adj = stride_adj;
        a.positive_part = adj_mode == CCN_ADJ_POSITIVE_PART;
        a.b = Batch{n_dev, n_max, C, (int)batch};
        a.scratch = ctx->ws;
        a.scratch_words = fp.scratch_words;
        a.ctl = ctx->ctl;
        a.slots = fp.slots;
        a.trace = (ctx->trace && batch * fused_tiles(n_max, C, 0) <= ctx->trace_tiles) ? ctx->trace : nullptr;
        a.fault = ctx->fault_dev;
        a.variant = ctx->fused_variant;
        a.keep = keep & 0x3ffffu;
        a.out_scale = out_scale;
        if (masked_by_kernel) *masked_by_kernel = true;
        LaunchLog flog = make_log(ctx);
        CCN_CUDA(ctx, launch_fused_forward(a, st, &flog));
        ctx->launches += flog.launches;
        return CCN_OK;
    }
    const Plan plan = make_plan(ctx, n_max, C, batch, false);
    rc = ensure_workspace(ctx, (size_t)plan.chunk * (plan.adj_words + plan.scratch_words) * 4);
    if (rc != CCN_OK) return rc;
    float *adjtab = ctx->ws;
    float *scratch = ctx->ws + (size_t)plan.chunk * plan.adj_words;

    LaunchLog log = make_log(ctx);
    for (int64_t i0 = 0; i0 < batch; i0 += plan.chunk) {
        const int cnt = (int)std::min<int64_t>(plan.chunk, batch - i0);
        Batch b{n_dev ? n_dev + i0 : nullptr, n_max, C, cnt};
        CCN_CUDA(ctx, launch_adj_prepare(adj_dev + i0 * stride_adj, stride_adj, b, adj_mode, adjtab, st, &log));
        Contract18Fwd a;
        a.T.base = const_cast<float *>(T_dev ? T_dev + i0 * stride_T : nullptr);
        a.T.slabs = const_cast<float *const *>(slabs_dev ? slabs_dev + i0 * n_max : nullptr);
        a.T.stride = stride_T;
        a.out = out_dev + i0 * stride_out;
        a.stride_out = stride_out;
        a.b = b;
        a.adjtab = adjtab;
        a.adjtab_words = plan.adj_words;
        a.scratch = scratch;
        a.scratch_words = plan.scratch_words;
        CCN_CUDA(ctx, launch_generic_forward(a, st, &log));
    }
    ctx->launches += log.launches;
    return CCN_OK;
}

int contract18_backward_impl(ccn_ctx *ctx, const float *gout_dev, const float *adj_dev, float *gT_dev, float *const *gslabs_dev,
                             const GatherRef *G, const int32_t *n_dev, int n_max, int C, int64_t batch, int64_t stride_gout,
                             int64_t stride_adj, int64_t stride_gT, int adj_mode, float beta, void *stream,
                             uint32_t keep = 0x3ffffu, bool *masked_by_kernel = nullptr) {
    if (masked_by_kernel) *masked_by_kernel = false;
    int rc = check_common(ctx, adj_dev, n_max, C, batch, adj_mode);
    if (rc != CCN_OK) return rc;
    if (!G && (gT_dev == nullptr) == (gslabs_dev == nullptr))
        return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "exactly one of gT_dev / gslabs_dev must be given");
    if (!gout_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "gout_dev is NULL");
    if (batch == 0) return CCN_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ScratchScope scope(ctx, st);
    if (scope.rc != CCN_OK) return scope.rc;
    bool fused = ctx->path == CCN_PATH_AUTO && fused_path_supported(n_max, C);
    // the fused kernel reads gout with 16-byte cp.async (a gslabs table cannot be checked here: 16-byte aligned entries)
    if (!aligned16(gout_dev) || (stride_gout & 3) != 0) fused = false;
    if (!G && gT_dev && (!aligned16(gT_dev) || (stride_gT & 3) != 0)) fused = false;
    if (G && !fused) return fail(ctx, CCN_ERR_UNSUPPORTED, "fused promotion backward needs a shape of the fused kernels and a 16-byte aligned gout");
    const int geometry = G ? 2 : 1;
    if (fused && batch * fused_tiles(n_max, C, geometry) < (int64_t)1 << 31) {
        FusedPlan fp;
        rc = fused_prepare(ctx, n_max, C, batch, geometry, &fp);
        if (rc != CCN_OK) return rc;
        Fused18Bwd a;
        a.G = G ? *G : GatherRef{nullptr, nullptr, nullptr, nullptr};
        a.gout = gout_dev;
        a.stride_gout = stride_gout;
        a.gT.base = gT_dev;
        a.gT.slabs = gslabs_dev;
        a.gT.stride = stride_gT;
        a.adj = adj_dev;
        a.stride_adj = stride_adj;
        a.positive_part = adj_mode == CCN_ADJ_POSITIVE_PART;
        a.b = Batch{n_dev, n_max, C, (int)batch};
        a.scratch = ctx->ws;
        a.scratch_words = fp.scratch_words;
        a.ctl = ctx->ctl;
        a.slots = fp.slots;
        a.beta = beta;
        a.trace = (ctx->trace && batch * fused_tiles(n_max, C, geometry) <= ctx->trace_tiles) ? ctx->trace : nullptr;
        a.fault = ctx->fault_dev;
        a.variant = ctx->fused_variant;
        a.keep = keep & 0x3ffffu;
        if (masked_by_kernel) *masked_by_kernel = true;
        LaunchLog flog = make_log(ctx);
        CCN_CUDA(ctx, G ? launch_fused_backward_scatter(a, st, &flog) : launch_fused_backward(a, st, &flog));
        ctx->launches += flog.launches;
        return CCN_OK;
    }
    const Plan plan = make_plan(ctx, n_max, C, batch, true);
    rc = ensure_workspace(ctx, (size_t)plan.chunk * (plan.adj_words + plan.scratch_words) * 4);
    if (rc != CCN_OK) return rc;
    float *adjtab = ctx->ws;
    float *scratch = ctx->ws + (size_t)plan.chunk * plan.adj_words;

    LaunchLog log = make_log(ctx);
    for (int64_t i0 = 0; i0 < batch; i0 += plan.chunk) {
        const int cnt = (int)std::min<int64_t>(plan.chunk, batch - i0);
        Batch b{n_dev ? n_dev + i0 : nullptr, n_max, C, cnt};
        CCN_CUDA(ctx, launch_adj_prepare(adj_dev + i0 * stride_adj, stride_adj, b, adj_mode, adjtab, st, &log));
        Contract18Bwd a;
        a.gout = gout_dev + i0 * stride_gout;
        a.stride_gout = stride_gout;
        a.gT.base = gT_dev ? gT_dev + i0 * stride_gT : nullptr;
        a.gT.slabs = gslabs_dev ? gslabs_dev + i0 * n_max : nullptr;
        a.gT.stride = stride_gT;
        a.b = b;
        a.adjtab = adjtab;
        a.adjtab_words = plan.adj_words;
        a.scratch = scratch;
        a.scratch_words = plan.scratch_words;
        a.beta = beta;
        CCN_CUDA(ctx, launch_generic_backward(a, st, &log));
    }
    ctx->launches += log.launches;
    return CCN_OK;
}

}  // namespace

extern "C" {

int ccn_contract18_forward(ccn_ctx *ctx, const float *T_dev, const float *const *slabs_dev, const float *adj_dev,
                           float *out_dev, const int32_t *n_dev, int n_max, int C, int64_t batch, int64_t stride_T,
                           int64_t stride_adj, int64_t stride_out, int adj_mode, void *stream) {
    return contract18_forward_impl(ctx, T_dev, slabs_dev, nullptr, adj_dev, out_dev, n_dev, n_max, C, batch, stride_T, stride_adj,
                                   stride_out, adj_mode, stream);
}

int ccn_contract18_backward(ccn_ctx *ctx, const float *gout_dev, const float *adj_dev, float *gT_dev,
                            float *const *gslabs_dev, const int32_t *n_dev, int n_max, int C, int64_t batch,
                            int64_t stride_gout, int64_t stride_adj, int64_t stride_gT, int adj_mode, float beta,
                            void *stream) {
    return contract18_backward_impl(ctx, gout_dev, adj_dev, gT_dev, gslabs_dev, nullptr, n_dev, n_max, C, batch, stride_gout,
                                    stride_adj, stride_gT, adj_mode, beta, stream);
}

// ---------------------------------------------------------------------------------------------------------------
// Host-buffer pipelines.  Chunk k uses slot k % kSlots of a device staging area; three streams (upload, compute,
// download) are chained with events so chunk k+1 uploads while chunk k computes and chunk k-1 downloads.
// ---------------------------------------------------------------------------------------------------------------
namespace {

int ensure_pipeline(ccn_ctx *ctx, size_t stage_bytes) {
    if (!ctx->s_in) {
        CCN_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
        CCN_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_comp, cudaStreamNonBlocking));
        CCN_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < ccn_ctx::kSlots; ++i) {
            CCN_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
            CCN_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_comp[i], cudaEventDisableTiming));
            CCN_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming));
        }
    }
    return grow_buffer(ctx, &ctx->stage, &ctx->stage_bytes, stage_bytes, "the device staging ring");
}

// what: bit 0 = forward, bit 1 = backward
int host_pipeline(ccn_ctx *ctx, int what, const float *T_host, const float *adj_host, const float *gout_host,
                  float *out_host, float *gT_host, int n, int C, int64_t batch, int adj_mode, float beta) {
    int rc = check_common(ctx, adj_host, n, C, batch, adj_mode);
    if (rc != CCN_OK) return rc;
    const bool fwd = what & 1, bwd = what & 2;
    if (fwd && (!T_host || !out_host)) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "T_host / out_host is NULL");
    if (bwd && (!gout_host || !gT_host)) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "gout_host / gT_host is NULL");
    if (batch == 0) return CCN_OK;
    DeviceGuard g(ctx->device);

    const int64_t szT = (int64_t)n * n * n * C, szA = (int64_t)n * n, szO = (int64_t)n * n * kSlabs * C;
    // per-instance device staging (floats), 64-float aligned sections
    auto up = [](int64_t v) { return (v + 63) & ~(int64_t)63; };
    const int64_t per = (fwd ? szT + szO : 0) + (bwd ? szO + szT : 0) + szA;
    // chunk size: about 256 MiB of staging per slot (measured best of 32..256 MiB on PCIe gen5; CCN_STAGE_MIB overrides), at least one instance
    const char *env_mib = std::getenv("CCN_STAGE_MIB");
    const int64_t slot_mib = env_mib ? std::max(1, std::atoi(env_mib)) : 256;
    int64_t chunk = std::max<int64_t>(1, (slot_mib << 20) / (per * 4));
    chunk = std::min(chunk, batch);
    const int64_t offT = 0, offO = offT + (fwd ? up(chunk * szT) : 0), offG = offO + (fwd ? up(chunk * szO) : 0),
                  offGT = offG + (bwd ? up(chunk * szO) : 0), offA = offGT + (bwd ? up(chunk * szT) : 0),
                  slot_words = offA + up(chunk * szA);
    rc = ensure_pipeline(ctx, (size_t)slot_words * 4 * ccn_ctx::kSlots);
    if (rc != CCN_OK) return rc;

    int64_t k = 0;
    for (int64_t i0 = 0; i0 < batch; i0 += chunk, ++k) {
        const int slot = (int)(k % ccn_ctx::kSlots);
        const int64_t cnt = std::min(chunk, batch - i0);
        float *base = ctx->stage + (size_t)slot * slot_words;
        float *dT = base + offT, *dO = base + offO, *dG = base + offG, *dGT = base + offGT, *dA = base + offA;
        // upload (the slot's previous results must have left the device first)
        if (k >= ccn_ctx::kSlots) CCN_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_out[slot], 0));
        CCN_CUDA(ctx, cudaMemcpyAsync(dA, adj_host + i0 * szA, (size_t)cnt * szA * 4, cudaMemcpyHostToDevice, ctx->s_in));
        if (fwd)
            CCN_CUDA(ctx, cudaMemcpyAsync(dT, T_host + i0 * szT, (size_t)cnt * szT * 4, cudaMemcpyHostToDevice, ctx->s_in));
        if (bwd) {
            CCN_CUDA(ctx, cudaMemcpyAsync(dG, gout_host + i0 * szO, (size_t)cnt * szO * 4, cudaMemcpyHostToDevice, ctx->s_in));
            if (beta != 0.f)
                CCN_CUDA(ctx, cudaMemcpyAsync(dGT, gT_host + i0 * szT, (size_t)cnt * szT * 4, cudaMemcpyHostToDevice, ctx->s_in));
        }
        CCN_CUDA(ctx, cudaEventRecord(ctx->ev_in[slot], ctx->s_in));
        // compute
        CCN_CUDA(ctx, cudaStreamWaitEvent(ctx->s_comp, ctx->ev_in[slot], 0));
        if (fwd) {
            rc = ccn_contract18_forward(ctx, dT, nullptr, dA, dO, nullptr, n, C, cnt, szT, szA, szO, adj_mode, ctx->s_comp);
            if (rc != CCN_OK) return rc;
        }
        if (bwd) {
            rc = ccn_contract18_backward(ctx, dG, dA, dGT, nullptr, nullptr, n, C, cnt, szO, szA, szT, adj_mode, beta,
                                         ctx->s_comp);
            if (rc != CCN_OK) return rc;
        }
        CCN_CUDA(ctx, cudaEventRecord(ctx->ev_comp[slot], ctx->s_comp));
        // download
        CCN_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_comp[slot], 0));
        if (fwd)
            CCN_CUDA(ctx, cudaMemcpyAsync(out_host + i0 * szO, dO, (size_t)cnt * szO * 4, cudaMemcpyDeviceToHost, ctx->s_out));
        if (bwd)
            CCN_CUDA(ctx, cudaMemcpyAsync(gT_host + i0 * szT, dGT, (size_t)cnt * szT * 4, cudaMemcpyDeviceToHost, ctx->s_out));
        CCN_CUDA(ctx, cudaEventRecord(ctx->ev_out[slot], ctx->s_out));
    }
    CCN_CUDA(ctx, cudaStreamSynchronize(ctx->s_out));
    CCN_CUDA(ctx, cudaStreamSynchronize(ctx->s_comp));
    CCN_CUDA(ctx, cudaStreamSynchronize(ctx->s_in));
    return CCN_OK;
}

}  // namespace

int ccn_contract18_forward_host(ccn_ctx *ctx, const float *T_host, const float *adj_host, float *out_host, int n,
                                int C, int64_t batch, int adj_mode) {
    return host_pipeline(ctx, 1, T_host, adj_host, nullptr, out_host, nullptr, n, C, batch, adj_mode, 0.f);
}

int ccn_contract18_backward_host(ccn_ctx *ctx, const float *gout_host, const float *adj_host, float *gT_host, int n,
                                 int C, int64_t batch, int adj_mode, float beta) {
    return host_pipeline(ctx, 2, nullptr, adj_host, gout_host, nullptr, gT_host, n, C, batch, adj_mode, beta);
}

int ccn_contract18_forward_backward_host(ccn_ctx *ctx, const float *T_host, const float *adj_host,
                                         const float *gout_host, float *out_host, float *gT_host, int n, int C,
                                         int64_t batch, int adj_mode) {
    return host_pipeline(ctx, 3, T_host, adj_host, gout_host, out_host, gT_host, n, C, batch, adj_mode, 0.f);
}

// ---------------------------------------------------------------------------------------------------------------
// RisiContraction_50
// ---------------------------------------------------------------------------------------------------------------
namespace {

int contract50_run(ccn_ctx *ctx, int variant, uint64_t keep_mask, bool backward, const float *in_dev, float *T_dev,
                   float *const *slabs_dev, const float *adj_dev, const int32_t *n_dev, int n_max, int C, int64_t batch,
                   int64_t stride_T, int64_t stride_adj, int64_t stride_out, int adj_mode, float beta, float out_scale,
                   void *stream) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    if (variant == 4 && out_scale != 1.f)
        return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "out_scale needs a variant with an adjacency operand");
    R50Plan plan;
    std::memset(&plan, 0, sizeof(plan));
    if (r50_make_plan(variant, keep_mask, &plan) != 0)
        return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "variant must be 4, 10, 18 or 50");
    // RisiContraction_4 has no adjacency operand; any non-null pointer passes the common checks and is never read
    int rc = check_common(ctx, (variant == 4 && !adj_dev) ? (const void *)ctx : (const void *)adj_dev, n_max, C, batch, adj_mode);
    if (rc != CCN_OK) return rc;
    if (variant == 4) adj_dev = nullptr;
    if ((T_dev == nullptr) == (slabs_dev == nullptr))
        return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "exactly one of the stacked tensor / slab table must be given");
    if (!in_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "out_dev / gout_dev is NULL");
    if (batch == 0) return CCN_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ScratchScope scope(ctx, st);
    if (scope.rc != CCN_OK) return scope.rc;
    const int adj_words = r50_adj_words(n_max);
    const int64_t scratch_words = r50_scratch_words(n_max, C);
    const size_t per = (size_t)(adj_words + scratch_words) * 4;
    // The 15 planes of an instance are ~18 MB at N = 48, C = 128: far beyond L2 residency anyway, so the chunk is sized
    // for occupancy (>= 64 instances per launch when they fit in 2 GiB), not by the L2-oriented workspace limit.
    const size_t limit = std::max<size_t>(ctx->ws_limit, (size_t)2 << 30);
    int64_t chunk = (int64_t)std::max<size_t>(1, limit / per);
    chunk = std::min<int64_t>(std::min<int64_t>(chunk, batch), 65535);
    if (const char *ev = getenv("CCN_R50_CHUNK")) chunk = std::max<int64_t>(1, std::min<int64_t>(chunk, atoll(ev)));  // tuning knob
    rc = ensure_workspace(ctx, (size_t)chunk * per);
    if (rc != CCN_OK) return rc;
    float *adjtab = ctx->ws, *scratch = ctx->ws + (size_t)chunk * adj_words;
    LaunchLog log = make_log(ctx);
    for (int64_t i0 = 0; i0 < batch; i0 += chunk) {
        const int cnt = (int)std::min<int64_t>(chunk, batch - i0);
        Batch b{n_dev ? n_dev + i0 : nullptr, n_max, C, cnt};
        TensorRef T;
        T.base = T_dev ? T_dev + i0 * stride_T : nullptr;
        T.slabs = slabs_dev ? slabs_dev + i0 * n_max : nullptr;
        T.stride = stride_T;
        CCN_CUDA(ctx, launch_r50(backward, plan, T, const_cast<float *>(in_dev) + i0 * stride_out, stride_out,
                                 adj_dev ? adj_dev + i0 * stride_adj : nullptr, stride_adj, b, adj_mode, out_scale, adjtab, scratch,
                                 beta, st, &log));
    }
    ctx->launches += log.launches;
    return CCN_OK;
}

}  // namespace

int ccn_contract50_forward(ccn_ctx *ctx, const float *T_dev, const float *const *slabs_dev, const float *adj_dev,
                           float *out_dev, const int32_t *n_dev, int n_max, int C, int64_t batch, int64_t stride_T,
                           int64_t stride_adj, int64_t stride_out, int adj_mode, void *stream) {
    return contract50_run(ctx, 50, ~0ull, false, out_dev, const_cast<float *>(T_dev), const_cast<float *const *>(slabs_dev), adj_dev,
                          n_dev, n_max, C, batch, stride_T, stride_adj, stride_out, adj_mode, 0.f, 1.f, stream);
}

int ccn_contract50_backward(ccn_ctx *ctx, const float *gout_dev, const float *adj_dev, float *gT_dev,
                            float *const *gslabs_dev, const int32_t *n_dev, int n_max, int C, int64_t batch,
                            int64_t stride_gout, int64_t stride_adj, int64_t stride_gT, int adj_mode, float beta,
                            void *stream) {
    return contract50_run(ctx, 50, ~0ull, true, gout_dev, gT_dev, gslabs_dev, adj_dev, n_dev, n_max, C, batch, stride_gT, stride_adj,
                          stride_gout, adj_mode, beta, 1.f, stream);
}

int ccn_contract_family_forward(ccn_ctx *ctx, int variant, uint64_t keep_mask, const float *T_dev, const float *const *slabs_dev,
                                const float *adj_dev, float *out_dev, const int32_t *n_dev, int n_max, int C, int64_t batch,
                                int64_t stride_T, int64_t stride_adj, int64_t stride_out, int adj_mode, float out_scale,
                                void *stream) {
    if (variant == 18) {
        // the 18-way kernels: the fused ones apply the slab mask and the scale as they store (no extra pass); after the generic
        // ones the dropped slabs are zeroed / the kept ones scaled in place
        const uint32_t keep = (uint32_t)(keep_mask & 0x3ffffu);
        bool masked = false;
        int rc = contract18_forward_impl(ctx, T_dev, slabs_dev, nullptr, adj_dev, out_dev, n_dev, n_max, C, batch, stride_T, stride_adj,
                                         stride_out, adj_mode, stream, keep, out_scale, &masked);
        if (rc != CCN_OK || batch == 0 || masked || (keep == 0x3ffffu && out_scale == 1.f)) return rc;
        DeviceGuard g(ctx->device);
        LaunchLog log = make_log(ctx);
        for (int64_t i0 = 0; i0 < batch; i0 += 65535) {
            const int cnt = (int)std::min<int64_t>(65535, batch - i0);
            Batch b{n_dev ? n_dev + i0 : nullptr, n_max, C, cnt};
            CCN_CUDA(ctx, launch_slab_mask_inplace(out_dev + i0 * stride_out, stride_out, b, keep, out_scale, 18,
                                                   static_cast<cudaStream_t>(stream), &log));
        }
        ctx->launches += log.launches;
        return CCN_OK;
    }
    return contract50_run(ctx, variant, keep_mask, false, out_dev, const_cast<float *>(T_dev),
                          const_cast<float *const *>(slabs_dev), adj_dev, n_dev, n_max, C, batch, stride_T, stride_adj, stride_out,
                          adj_mode, 0.f, out_scale, stream);
}

int ccn_contract_family_backward(ccn_ctx *ctx, int variant, uint64_t keep_mask, const float *gout_dev, const float *adj_dev,
                                 float *gT_dev, float *const *gslabs_dev, const int32_t *n_dev, int n_max, int C, int64_t batch,
                                 int64_t stride_gout, int64_t stride_adj, int64_t stride_gT, int adj_mode, float beta, void *stream) {
    if (variant == 18) {
        const uint32_t keep = (uint32_t)(keep_mask & 0x3ffffu);
        if (keep == 0x3ffffu)
            return ccn_contract18_backward(ctx, gout_dev, adj_dev, gT_dev, gslabs_dev, n_dev, n_max, C, batch, stride_gout, stride_adj,
                                           stride_gT, adj_mode, beta, stream);
        // dropped slabs must not reach gT: the fused kernels simply never read them; the generic ones read a copy of gout with
        // those slabs zeroed, chunk by chunk
        int rc = check_common(ctx, adj_dev, n_max, C, batch, adj_mode);
        if (rc != CCN_OK) return rc;
        if (!gout_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "gout_dev is NULL");
        if (batch == 0) return CCN_OK;
        if (ctx->path == CCN_PATH_AUTO && fused_path_supported(n_max, C) && aligned16(gout_dev) && (stride_gout & 3) == 0 &&
            (!gT_dev || (aligned16(gT_dev) && (stride_gT & 3) == 0)) && batch * fused_tiles(n_max, C, 1) < (int64_t)1 << 31) {
            bool masked = false;
            rc = contract18_backward_impl(ctx, gout_dev, adj_dev, gT_dev, gslabs_dev, nullptr, n_dev, n_max, C, batch, stride_gout,
                                          stride_adj, stride_gT, adj_mode, beta, stream, keep, &masked);
            if (rc != CCN_OK || masked) return rc;
            return fail(ctx, CCN_ERR_CUDA, "internal: the fused backward was expected to apply the slab mask");
        }
        DeviceGuard g(ctx->device);
        ScratchScope scope(ctx, static_cast<cudaStream_t>(stream));
        if (scope.rc != CCN_OK) return scope.rc;
        const int64_t per = (int64_t)18 * n_max * n_max * C;
        const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(batch, 65535), ((int64_t)256 << 20) / (per * 4)));
        const size_t need = (size_t)chunk * per * 4;
        rc = grow_buffer(ctx, &ctx->gybuf, &ctx->gybuf_bytes, need, "the gradient staging buffer");
        if (rc != CCN_OK) return rc;
        for (int64_t i0 = 0; i0 < batch; i0 += chunk) {
            const int cnt = (int)std::min<int64_t>(chunk, batch - i0);
            Batch b{n_dev ? n_dev + i0 : nullptr, n_max, C, cnt};
            LaunchLog log = make_log(ctx);
            CCN_CUDA(ctx, launch_slab_mask_copy(gout_dev + i0 * stride_gout, stride_gout, ctx->gybuf, per, b, keep, 18,
                                                static_cast<cudaStream_t>(stream), &log));
            ctx->launches += log.launches;
            rc = ccn_contract18_backward(ctx, ctx->gybuf, adj_dev + i0 * stride_adj, gT_dev ? gT_dev + i0 * stride_gT : nullptr,
                                         gslabs_dev ? gslabs_dev + i0 * n_max : nullptr, n_dev ? n_dev + i0 : nullptr, n_max, C, cnt,
                                         per, stride_adj, stride_gT, adj_mode, beta, stream);
            if (rc != CCN_OK) return rc;
        }
        return CCN_OK;
    }
    return contract50_run(ctx, variant, keep_mask, true, gout_dev, gT_dev, gslabs_dev, adj_dev, n_dev, n_max, C, batch, stride_gT,
                          stride_adj, stride_gout, adj_mode, beta, 1.f, stream);
}

// ---------------------------------------------------------------------------------------------------------------
// Feature mix
// ---------------------------------------------------------------------------------------------------------------
int ccn_mix_forward(ccn_ctx *ctx, const float *X_dev, const float *W_dev, const float *bias_dev, float *Y_dev,
                    float *Z_dev, int64_t M, int K, int P, float lrelu_alpha, void *stream) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    if (!X_dev || !W_dev || (!Y_dev && !Z_dev)) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "X/W NULL or no output given");
    if (Z_dev && !bias_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "Z_dev needs bias_dev");
    if (M < 0 || K <= 0 || P <= 0) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "bad GEMM shape");
    if (M == 0) return CCN_OK;
    DeviceGuard g(ctx->device);
    ScratchScope scope(ctx, static_cast<cudaStream_t>(stream));
    if (scope.rc != CCN_OK) return scope.rc;
    LaunchLog log = make_log(ctx);
    const bool tc_ok = mix_tc_supported(X_dev, Y_dev, Z_dev, M, K, P);
    if (ctx->mix_path == CCN_MIX_TENSOR && !tc_ok)
        return fail(ctx, CCN_ERR_UNSUPPORTED, "tensor-core mix needs K % 4 == 0, P % 4 == 0, 4 <= P <= 128, 16-byte aligned buffers");
    if (tc_ok && ctx->mix_path != CCN_MIX_SIMT) {
        int rcw = grow_buffer(ctx, &ctx->wprep, &ctx->wprep_bytes, mix_tc_wprep_bytes(K, P), "the prepared-weights buffer");
        if (rcw != CCN_OK) return rcw;
        if (ctx->mix_rows_n) {
            rcw = grow_buffer(ctx, &ctx->mix_items, &ctx->mix_items_bytes, mix_item_list_bytes(M), "the mix work-item list");
            if (rcw != CCN_OK) return rcw;
        }
        CCN_CUDA(ctx, launch_mix_forward_tc(X_dev, W_dev, bias_dev, Y_dev, Z_dev, M, K, P, lrelu_alpha, ctx->wprep,
                                            ctx->sm_count, ctx->mix_tiles_per_pass, static_cast<cudaStream_t>(stream), &log,
                                            ctx->mix_rows_n, ctx->mix_rows_per_inst, ctx->mix_rows_n ? ctx->mix_items : nullptr));
        ctx->launches += log.launches;
        return CCN_OK;
    }
    CCN_CUDA(ctx, launch_mix_forward(X_dev, W_dev, bias_dev, Y_dev, Z_dev, M, K, P, lrelu_alpha,
                                     static_cast<cudaStream_t>(stream), &log));
    ctx->launches += log.launches;
    return CCN_OK;
}

int ccn_mix_backward(ccn_ctx *ctx, const float *X_dev, const float *W_dev, const float *bias_dev, const float *Y_dev,
                     const float *gZ_dev, float *gX_dev, float *gW_dev, float *gbias_dev, int64_t M, int K, int P,
                     float lrelu_alpha, float beta_x, void *stream) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    if (!gZ_dev || !W_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "gZ_dev / W_dev is NULL");
    if (gW_dev && !X_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "gW_dev needs X_dev");
    if (bias_dev && !Y_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "bias_dev needs Y_dev (pre-activation)");
    if (gbias_dev && !bias_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "gbias_dev needs bias_dev");
    if (M < 0 || K <= 0 || P <= 0) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "bad GEMM shape");
    if (M == 0) return CCN_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ScratchScope scope(ctx, st);
    if (scope.rc != CCN_OK) return scope.rc;
    LaunchLog log = make_log(ctx);
    int simt_parts = 7;
    const bool gx_tc = gX_dev && ctx->mix_path != CCN_MIX_SIMT && mix_gx_tc_supported(gZ_dev, Y_dev, gX_dev, nullptr, M, K, P);
    if (ctx->mix_path == CCN_MIX_TENSOR && gX_dev && !gx_tc)
        return fail(ctx, CCN_ERR_UNSUPPORTED, "tensor-core grad-X needs K % 4 == 0, P % 4 == 0, 4 <= P <= 64, 16-byte aligned buffers");
    if (gx_tc) {
        int rcw = grow_buffer(ctx, &ctx->wprep, &ctx->wprep_bytes, mix_gx_tc_wprep_bytes(K, P), "the prepared-weights buffer");
        if (rcw != CCN_OK) return rcw;
        // grad-W on the tensor cores needs gY as an array: grad-X writes it on the way (no extra pass over gZ / Y)
        const float *gy_for_w = nullptr;
        float *gy_out = nullptr;
        if (gW_dev && X_dev && mix_gw_tc_supported(X_dev, gZ_dev, gW_dev, M, K, P)) {
            if (!bias_dev) {
                gy_for_w = gZ_dev;
            } else {
                int rcg = grow_buffer(ctx, &ctx->gybuf, &ctx->gybuf_bytes, (size_t)M * P * sizeof(float), "the gY buffer");
                if (rcg != CCN_OK) return rcg;
                gy_out = ctx->gybuf;
                gy_for_w = ctx->gybuf;
            }
        }
        if (ctx->mix_rows_n) {
            rcw = grow_buffer(ctx, &ctx->mix_items, &ctx->mix_items_bytes, mix_item_list_bytes(M), "the mix work-item list");
            if (rcw != CCN_OK) return rcw;
        }
        CCN_CUDA(ctx, launch_mix_grad_x_tc(W_dev, bias_dev, Y_dev, gZ_dev, gX_dev, gy_out, M, K, P, lrelu_alpha, beta_x, ctx->wprep,
                                           ctx->sm_count, st, &log, ctx->mix_rows_n, ctx->mix_rows_per_inst,
                                           ctx->mix_rows_n ? ctx->mix_items : nullptr));
        simt_parts &= ~1;
        if (gy_for_w) {
            CCN_CUDA(ctx, launch_mix_grad_w_tc(X_dev, gy_for_w, gW_dev, gbias_dev, M, K, P, ctx->sm_count, st, &log));
            simt_parts &= ~(2 | 4);
        }
    }
    CCN_CUDA(ctx, launch_mix_backward_parts(simt_parts, X_dev, W_dev, bias_dev, Y_dev, gZ_dev, gX_dev, gW_dev, gbias_dev, M, K, P,
                                            lrelu_alpha, beta_x, st, &log));
    ctx->launches += log.launches;
    return CCN_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Promotion (MatTensorMul + TensorMatMul + StackTensor3D as a gather), TensorMul, CustomMatMulTensor
// ---------------------------------------------------------------------------------------------------------------
namespace {

int promote_run(ccn_ctx *ctx, bool backward, float *f_dev, const int64_t *f_off_dev, const int32_t *m_dev, const int32_t *pos_dev,
                float *T_dev, const int32_t *n_dev, int n_max, int C, int64_t batch, int64_t stride_T, void *stream) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    if (!f_dev || !f_off_dev || !m_dev || !pos_dev || !T_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "NULL argument");
    if (n_max <= 0 || C <= 0 || batch < 0 || n_max > 65535) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "bad sizes");
    if (batch == 0) return CCN_OK;
    DeviceGuard g(ctx->device);
    LaunchLog log = make_log(ctx);
    for (int64_t i0 = 0; i0 < batch; i0 += 65535) {
        const int cnt = (int)std::min<int64_t>(65535, batch - i0);
        PromoteArgs a;
        a.f = f_dev;
        a.f_off = f_off_dev + i0 * n_max;
        a.m = m_dev + i0 * n_max;
        a.pos = pos_dev + i0 * n_max * n_max;
        a.T = T_dev + i0 * stride_T;
        a.stride_T = stride_T;
        a.n = n_dev ? n_dev + i0 : nullptr;
        a.n_max = n_max;
        a.C = C;
        a.v4_offsets_ok = (C % 4) == 0;  // offsets are sums of whole [m, m, C] tensors: multiples of C
        CCN_CUDA(ctx, launch_promote(backward, a, cnt, static_cast<cudaStream_t>(stream), &log));
    }
    ctx->launches += log.launches;
    return CCN_OK;
}

int ensure_aux(ccn_ctx *ctx, size_t bytes) { return grow_buffer(ctx, &ctx->aux, &ctx->aux_bytes, bytes, "the transposed-weights buffer"); }

}  // namespace

int ccn_promote_forward(ccn_ctx *ctx, const float *f_dev, const int64_t *f_off_dev, const int32_t *m_dev,
                        const int32_t *pos_dev, float *T_dev, const int32_t *n_dev, int n_max, int C, int64_t batch,
                        int64_t stride_T, void *stream) {
    return promote_run(ctx, false, const_cast<float *>(f_dev), f_off_dev, m_dev, pos_dev, T_dev, n_dev, n_max, C, batch, stride_T,
                       stream);
}

int ccn_promote_backward(ccn_ctx *ctx, const float *gT_dev, const int64_t *f_off_dev, const int32_t *m_dev,
                         const int32_t *pos_dev, float *gf_dev, const int32_t *n_dev, int n_max, int C, int64_t batch,
                         int64_t stride_T, void *stream) {
    return promote_run(ctx, true, gf_dev, f_off_dev, m_dev, pos_dev, const_cast<float *>(gT_dev), n_dev, n_max, C, batch, stride_T,
                       stream);
}

// ---- chained entry points (the same kernels as the single entry points, in order) ----------------------------------
int ccn_gather_contract18_forward(ccn_ctx *ctx, const float *f_dev, const int64_t *f_off_dev, const int32_t *m_dev,
                                  const int32_t *pos_dev, const float *adj_dev, float *T_scratch_dev, float *out_dev,
                                  const int32_t *n_dev, int n_max, int C, int64_t batch, int64_t stride_adj,
                                  int64_t stride_out, int adj_mode, void *stream) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    if (!f_dev || !f_off_dev || !m_dev || !pos_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "NULL argument");
    if (gather_fusable(ctx, f_dev, n_max, C, batch)) {  // one kernel: the promotion is read inside the contraction's stream
        const GatherRef G{const_cast<float *>(f_dev), f_off_dev, m_dev, pos_dev};
        return contract18_forward_impl(ctx, nullptr, nullptr, &G, adj_dev, out_dev, n_dev, n_max, C, batch, 0, stride_adj, stride_out,
                                       adj_mode, stream);
    }
    if (!T_scratch_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "this shape needs T_scratch_dev (the promotion is not fused)");
    const int64_t stride_T = (int64_t)n_max * n_max * n_max * C;
    int rc = ccn_promote_forward(ctx, f_dev, f_off_dev, m_dev, pos_dev, T_scratch_dev, n_dev, n_max, C, batch, stride_T, stream);
    if (rc != CCN_OK) return rc;
    return ccn_contract18_forward(ctx, T_scratch_dev, nullptr, adj_dev, out_dev, n_dev, n_max, C, batch, stride_T, stride_adj,
                                  stride_out, adj_mode, stream);
}

int ccn_gather_contract18_backward(ccn_ctx *ctx, const float *gout_dev, const float *adj_dev, const int64_t *f_off_dev,
                                   const int32_t *m_dev, const int32_t *pos_dev, float *gT_scratch_dev, float *gf_dev,
                                   const int32_t *n_dev, int n_max, int C, int64_t batch, int64_t stride_gout,
                                   int64_t stride_adj, int adj_mode, void *stream) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    if (!gf_dev || !f_off_dev || !m_dev || !pos_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "NULL argument");
    if (gather_fusable(ctx, gf_dev, n_max, C, batch) && aligned16(gout_dev) && (stride_gout & 3) == 0) {
        const GatherRef G{gf_dev, f_off_dev, m_dev, pos_dev};  // gT is never written: its rows are added straight into gf
        return contract18_backward_impl(ctx, gout_dev, adj_dev, nullptr, nullptr, &G, n_dev, n_max, C, batch, stride_gout, stride_adj,
                                        0, adj_mode, 0.f, stream);
    }
    if (!gT_scratch_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "this shape needs gT_scratch_dev (the promotion is not fused)");
    const int64_t stride_T = (int64_t)n_max * n_max * n_max * C;
    int rc = ccn_contract18_backward(ctx, gout_dev, adj_dev, gT_scratch_dev, nullptr, n_dev, n_max, C, batch, stride_gout, stride_adj,
                                     stride_T, adj_mode, 0.f, stream);
    if (rc != CCN_OK) return rc;
    return ccn_promote_backward(ctx, gT_scratch_dev, f_off_dev, m_dev, pos_dev, gf_dev, n_dev, n_max, C, batch, stride_T, stream);
}

int ccn_level_forward(ccn_ctx *ctx, const float *T_dev, const float *const *slabs_dev, const float *adj_dev, const float *K_dev,
                      const float *bias_dev, float *X_dev, float *Y_dev, float *Z_dev, const int32_t *n_dev, int n_max, int C_in,
                      int C_out, int64_t batch, int64_t stride_T, int64_t stride_adj, int adj_mode, float lrelu_alpha, void *stream) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    if (!X_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "X_dev is NULL");
    const int64_t stride_X = (int64_t)18 * n_max * n_max * C_in;
    int rc = ccn_contract18_forward(ctx, T_dev, slabs_dev, adj_dev, X_dev, n_dev, n_max, C_in, batch, stride_T, stride_adj, stride_X,
                                    adj_mode, stream);
    if (rc != CCN_OK) return rc;
    return ccn_mix_forward(ctx, X_dev, K_dev, bias_dev, Y_dev, Z_dev, batch * n_max * n_max, 18 * C_in, C_out, lrelu_alpha, stream);
}

int ccn_level_backward(ccn_ctx *ctx, const float *gZ_dev, const float *X_dev, const float *Y_dev, const float *K_dev,
                       const float *bias_dev, const float *adj_dev, float *gX_scratch_dev, float *gT_dev, float *const *gslabs_dev,
                       float *gK_dev, float *gbias_dev, const int32_t *n_dev, int n_max, int C_in, int C_out, int64_t batch,
                       int64_t stride_adj, int64_t stride_gT, int adj_mode, float lrelu_alpha, float beta, void *stream) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    if (!gX_scratch_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "gX_scratch_dev is NULL");
    const int64_t stride_X = (int64_t)18 * n_max * n_max * C_in;
    int rc = ccn_mix_backward(ctx, X_dev, K_dev, bias_dev, Y_dev, gZ_dev, gX_scratch_dev, gK_dev, gbias_dev, batch * n_max * n_max,
                              18 * C_in, C_out, lrelu_alpha, 0.f, stream);
    if (rc != CCN_OK) return rc;
    return ccn_contract18_backward(ctx, gX_scratch_dev, adj_dev, gT_dev, gslabs_dev, n_dev, n_max, C_in, batch, stride_X, stride_adj,
                                   stride_gT, adj_mode, beta, stream);
}

// ---- one whole CCN level from the level l-1 tensors: promotion fused into the contraction, then the feature mix ------
int ccn_gather_level_forward(ccn_ctx *ctx, const float *f_dev, const int64_t *f_off_dev, const int32_t *m_dev, const int32_t *pos_dev,
                             const float *adj_dev, const float *K_dev, const float *bias_dev, float *T_scratch_dev, float *X_dev,
                             float *Y_dev, float *Z_dev, const int32_t *n_dev, int n_max, int C_in, int C_out, int64_t batch,
                             int64_t stride_adj, int adj_mode, float lrelu_alpha, void *stream) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    if (!X_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "X_dev is NULL");
    const int64_t stride_X = (int64_t)18 * n_max * n_max * C_in;
    int rc = ccn_gather_contract18_forward(ctx, f_dev, f_off_dev, m_dev, pos_dev, adj_dev, T_scratch_dev, X_dev, n_dev, n_max, C_in,
                                           batch, stride_adj, stride_X, adj_mode, stream);
    if (rc != CCN_OK) return rc;
    // ragged batches: the mix skips the work items that lie entirely in the padding rows [n_i^2, n_max^2) of the instances
    ctx->mix_rows_n = n_dev;
    ctx->mix_rows_per_inst = (int64_t)n_max * n_max;
    rc = ccn_mix_forward(ctx, X_dev, K_dev, bias_dev, Y_dev, Z_dev, batch * n_max * n_max, 18 * C_in, C_out, lrelu_alpha, stream);
    ctx->mix_rows_n = nullptr;
    return rc;
}

int ccn_gather_level_backward(ccn_ctx *ctx, const float *gZ_dev, const float *X_dev, const float *Y_dev, const float *K_dev,
                              const float *bias_dev, const float *adj_dev, const int64_t *f_off_dev, const int32_t *m_dev,
                              const int32_t *pos_dev, float *gX_scratch_dev, float *gT_scratch_dev, float *gf_dev, float *gK_dev,
                              float *gbias_dev, const int32_t *n_dev, int n_max, int C_in, int C_out, int64_t batch,
                              int64_t stride_adj, int adj_mode, float lrelu_alpha, void *stream) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    if (!gX_scratch_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "gX_scratch_dev is NULL");
    const int64_t stride_X = (int64_t)18 * n_max * n_max * C_in;
    ctx->mix_rows_n = n_dev;
    ctx->mix_rows_per_inst = (int64_t)n_max * n_max;
    int rc = ccn_mix_backward(ctx, X_dev, K_dev, bias_dev, Y_dev, gZ_dev, gX_scratch_dev, gK_dev, gbias_dev, batch * n_max * n_max,
                              18 * C_in, C_out, lrelu_alpha, 0.f, stream);
    ctx->mix_rows_n = nullptr;
    if (rc != CCN_OK) return rc;
    return ccn_gather_contract18_backward(ctx, gX_scratch_dev, adj_dev, f_off_dev, m_dev, pos_dev, gT_scratch_dev, gf_dev, n_dev, n_max,
                                          C_in, batch, stride_X, stride_adj, adj_mode, stream);
}

// Host-buffer form of the calls above for a STACK of L levels that stays on the device between the levels (what a model with
// host-resident inputs calls once per batch): groups of instances (graphs) are uploaded, computed and downloaded on three
// streams through the device staging ring.  See include/ccn_b200.h.
}  // extern "C"

namespace {

// The read-out head of the models on the device (ccn_readout_*): when W_host is given, the last level's output goes through
// ShrinkTensor -> LeakyReLU -> SumVectors (per group = graph) -> InnerProduct(W) -> SquaredLoss(target) and the gradient of the
// last level's output is produced on the device, so neither Z nor gZ crosses PCIe.
struct ReadoutHost {
    const float *W_host, *target_host;  // [C_out], [groups]
    float *predict_host, *loss_host;    // [groups]
    float *gW_host;                     // [C_out]
};

int levels_host_impl(ccn_ctx *ctx, int levels, const float *f_host, const int64_t *f_group_ptr, const int64_t *inst_group_ptr,
                     int64_t groups, const int64_t *const *f_off_host, const int32_t *const *m_host, const int32_t *const *pos_host,
                     const float *const *adj_host, const float *const *K_host, const float *const *bias_host, const float *gZ_host,
                     float *Z_host, float *gf_host, float *const *gK_host, float *const *gbias_host, int n, int C_in, int C_out,
                     int adj_mode, float lrelu_alpha, const ReadoutHost *ro) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    if (levels < 1 || levels > 16) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "levels must be in 1..16");
    if (!f_host || !f_group_ptr || !inst_group_ptr || !f_off_host || !m_host || !pos_host || !adj_host || !K_host || !bias_host ||
        !gf_host || !gK_host || !gbias_host)
        return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "NULL argument");
    if (ro ? (!ro->W_host || !ro->target_host || !ro->predict_host || !ro->loss_host || !ro->gW_host) : (!gZ_host || !Z_host))
        return fail(ctx, CCN_ERR_INVALID_ARGUMENT, ro ? "NULL read-out argument" : "gZ_host / Z_host is NULL");
    for (int l = 0; l < levels; ++l)
        if (!f_off_host[l] || !m_host[l] || !pos_host[l] || !adj_host[l] || !K_host[l] || !bias_host[l] || !gK_host[l] || !gbias_host[l])
            return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "NULL per-level argument");
    if (groups < 0 || n <= 0 || C_in <= 0 || C_out <= 0) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "bad sizes");
    if (levels > 1 && C_in != C_out) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "a stack of levels needs C_in == C_out");
    if (!fused_path_supported(n, C_in) || (C_in % 4) != 0)
        return fail(ctx, CCN_ERR_UNSUPPORTED, "the host-buffer level call needs a shape of the fused kernels (n <= 32, C_in in {8,16,32,64,128})");
    if (groups == 0) return CCN_OK;
    DeviceGuard g(ctx->device);
    const int L = levels;
    const int64_t nn = (int64_t)n * n, Kd = (int64_t)18 * C_in, sX = nn * Kd, sY = nn * C_out;
    for (int64_t q = 0; q < groups; ++q)
        if (inst_group_ptr[q + 1] < inst_group_ptr[q] || f_group_ptr[q + 1] < f_group_ptr[q] || (f_group_ptr[q] & 3))
            return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "group pointers must be non-decreasing and the f boundaries multiples of 4 elements");
    // chunk = consecutive groups of about 256 instances (CCN_LEVEL_CHUNK overrides; measured 64 / 128 / 256: 78.5 / 83.3 / 87.3 k
    // contractions/s at 4 levels x 2048 instances): several waves of tiles per launch
    const char *env = std::getenv("CCN_LEVEL_CHUNK");
    const int64_t want = env ? std::max(1, std::atoi(env)) : 256;
    std::vector<int64_t> cut(1, 0);
    for (int64_t q = 1; q <= groups; ++q)
        if (q == groups || inst_group_ptr[q + 1] - inst_group_ptr[cut.back()] > want) cut.push_back(q);
    int64_t max_inst = 0, max_f = 0;
    for (size_t c = 0; c + 1 < cut.size(); ++c) {
        max_inst = std::max(max_inst, inst_group_ptr[cut[c + 1]] - inst_group_ptr[cut[c]]);
        max_f = std::max(max_f, f_group_ptr[cut[c + 1]] - f_group_ptr[cut[c]]);
    }
    auto up = [](int64_t v) { return (v + 63) & ~(int64_t)63; };
    // per slot (floats): f, gf, Z_L, gZ_L, then per level adj and the integer tables (f_off as 2 words each)
    const int64_t tabA = up(max_inst * nn), tabOff = up(2 * max_inst * n), tabM = up(max_inst * n), tabPos = up(max_inst * nn);
    const int64_t per_level_tab = tabA + tabOff + tabM + tabPos;
    int64_t max_groups = 0;
    for (size_t c = 0; c + 1 < cut.size(); ++c) max_groups = std::max<int64_t>(max_groups, cut[c + 1] - cut[c]);
    // read-out per slot: group pointers (int64, chunk-relative), instance -> group (int32), targets, predictions, losses
    const int64_t roPtr = up(2 * (max_groups + 1)), roIG = up(max_inst), roG = up(max_groups);
    const int64_t ro_words = ro ? roPtr + roIG + 3 * roG : 0;
    const int64_t oF = 0, oGF = oF + up(max_f), oZ = oGF + up(max_f), oGZ = oZ + up(max_inst * sY), oTab = oGZ + up(max_inst * sY),
                  oRo = oTab + L * per_level_tab, slot_words = oRo + ro_words;
    // shared by the slots (compute is serial): X_l and Y_l of every level (kept for the backward), the activations between the
    // levels, two gradient buffers between the levels, gX, then the parameters and their gradients
    const int64_t oX = slot_words * ccn_ctx::kSlots, oY = oX + L * up(max_inst * sX), oMid = oY + L * up(max_inst * sY),
                  oGMid = oMid + (L - 1) * up(max_inst * sY), oGX = oGMid + 2 * up(max_inst * sY), oK = oGX + up(max_inst * sX),
                  oB = oK + L * up(Kd * C_out), oW = oB + L * up(C_out), oShr = oW + up(C_out), oGFe = oShr + up(max_inst * C_out),
                  oGK = oGFe + up(max_groups * C_out), oGB = oGK + L * up(Kd * C_out), oGW = oGB + L * up(C_out),
                  total_words = oGW + up(C_out);
    int rc = ensure_pipeline(ctx, (size_t)total_words * 4);
    if (rc != CCN_OK) return rc;
    float *base = ctx->stage;
    auto dXl = [&](int l) { return base + oX + l * up(max_inst * sX); };
    auto dYl = [&](int l) { return base + oY + l * up(max_inst * sY); };
    auto dMid = [&](int l) { return base + oMid + l * up(max_inst * sY); };  // output of level l (0-based), l < L-1
    auto dGMid = [&](int k) { return base + oGMid + k * up(max_inst * sY); };
    auto dKl = [&](int l) { return base + oK + l * up(Kd * C_out); };
    auto dBl = [&](int l) { return base + oB + l * up(C_out); };
    auto dGKl = [&](int l) { return base + oGK + l * up(Kd * C_out); };
    auto dGBl = [&](int l) { return base + oGB + l * up(C_out); };
    float *dGX = base + oGX, *dW = base + oW, *dShr = base + oShr, *dGFe = base + oGFe, *dGW = base + oGW;
    if (ro) CCN_CUDA(ctx, cudaMemcpyAsync(dW, ro->W_host, (size_t)C_out * 4, cudaMemcpyHostToDevice, ctx->s_comp));
    // chunk-relative group pointers and instance -> group maps of every chunk (kept alive until the final synchronisation)
    std::vector<std::vector<int64_t>> h_ptr(ro ? cut.size() : 0);
    std::vector<std::vector<int32_t>> h_ig(ro ? cut.size() : 0);
    for (int l = 0; l < L; ++l) {
        CCN_CUDA(ctx, cudaMemcpyAsync(dKl(l), K_host[l], (size_t)Kd * C_out * 4, cudaMemcpyHostToDevice, ctx->s_comp));
        CCN_CUDA(ctx, cudaMemcpyAsync(dBl(l), bias_host[l], (size_t)C_out * 4, cudaMemcpyHostToDevice, ctx->s_comp));
    }
    CCN_CUDA(ctx, cudaMemsetAsync(base + oGK, 0, (size_t)(total_words - oGK) * 4, ctx->s_comp));
    for (size_t c = 0; c + 1 < cut.size(); ++c) {
        const int slot = (int)(c % ccn_ctx::kSlots);
        const int64_t i0 = inst_group_ptr[cut[c]], cnt = inst_group_ptr[cut[c + 1]] - i0;
        const int64_t f0 = f_group_ptr[cut[c]], fcnt = f_group_ptr[cut[c + 1]] - f0;
        if (cnt == 0) continue;
        float *sb = base + (size_t)slot * slot_words;
        float *dF = sb + oF, *dGF = sb + oGF, *dZ = sb + oZ, *dGZ = sb + oGZ;
        auto tA = [&](int l) { return sb + oTab + l * per_level_tab; };
        auto tOff = [&](int l) { return reinterpret_cast<int64_t *>(sb + oTab + l * per_level_tab + tabA); };
        auto tM = [&](int l) { return reinterpret_cast<int32_t *>(sb + oTab + l * per_level_tab + tabA + tabOff); };
        auto tPos = [&](int l) { return reinterpret_cast<int32_t *>(sb + oTab + l * per_level_tab + tabA + tabOff + tabM); };
        if (c >= (size_t)ccn_ctx::kSlots) CCN_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_out[slot], 0));
        CCN_CUDA(ctx, cudaMemcpyAsync(dF, f_host + f0, (size_t)fcnt * 4, cudaMemcpyHostToDevice, ctx->s_in));
        const int64_t ng = cut[c + 1] - cut[c];
        int64_t *dPtr = reinterpret_cast<int64_t *>(sb + oRo);
        int32_t *dIG = reinterpret_cast<int32_t *>(sb + oRo + roPtr);
        float *dTgt = sb + oRo + roPtr + roIG, *dPred = dTgt + roG, *dLoss = dPred + roG;
        if (ro) {
            h_ptr[c].resize(ng + 1);
            h_ig[c].resize(cnt);
            for (int64_t q = 0; q <= ng; ++q) h_ptr[c][q] = inst_group_ptr[cut[c] + q] - i0;
            for (int64_t q = 0; q < ng; ++q)
                for (int64_t i = h_ptr[c][q]; i < h_ptr[c][q + 1]; ++i) h_ig[c][i] = (int32_t)q;
            CCN_CUDA(ctx, cudaMemcpyAsync(dPtr, h_ptr[c].data(), (size_t)(ng + 1) * 8, cudaMemcpyHostToDevice, ctx->s_in));
            CCN_CUDA(ctx, cudaMemcpyAsync(dIG, h_ig[c].data(), (size_t)cnt * 4, cudaMemcpyHostToDevice, ctx->s_in));
            CCN_CUDA(ctx, cudaMemcpyAsync(dTgt, ro->target_host + cut[c], (size_t)ng * 4, cudaMemcpyHostToDevice, ctx->s_in));
        } else {
            CCN_CUDA(ctx, cudaMemcpyAsync(dGZ, gZ_host + i0 * sY, (size_t)cnt * sY * 4, cudaMemcpyHostToDevice, ctx->s_in));
        }
        for (int l = 0; l < L; ++l) {
            CCN_CUDA(ctx, cudaMemcpyAsync(tA(l), adj_host[l] + i0 * nn, (size_t)cnt * nn * 4, cudaMemcpyHostToDevice, ctx->s_in));
            CCN_CUDA(ctx, cudaMemcpyAsync(tOff(l), f_off_host[l] + i0 * n, (size_t)cnt * n * 8, cudaMemcpyHostToDevice, ctx->s_in));
            CCN_CUDA(ctx, cudaMemcpyAsync(tM(l), m_host[l] + i0 * n, (size_t)cnt * n * 4, cudaMemcpyHostToDevice, ctx->s_in));
            CCN_CUDA(ctx, cudaMemcpyAsync(tPos(l), pos_host[l] + i0 * nn, (size_t)cnt * nn * 4, cudaMemcpyHostToDevice, ctx->s_in));
        }
        CCN_CUDA(ctx, cudaEventRecord(ctx->ev_in[slot], ctx->s_in));
        CCN_CUDA(ctx, cudaStreamWaitEvent(ctx->s_comp, ctx->ev_in[slot], 0));
        // The offsets are absolute: level 1's into f_host, the later levels' into the previous level's output array
        // [batch, n^2, C_out].  Shift the device base pointers instead of rewriting the tables.
        for (int l = 0; l < L; ++l) {
            const float *src = l == 0 ? dF - f0 : dMid(l - 1) - i0 * sY;
            float *dst = l == L - 1 ? dZ : dMid(l);
            rc = ccn_gather_level_forward(ctx, src, tOff(l), tM(l), tPos(l), tA(l), dKl(l), dBl(l), nullptr, dXl(l), dYl(l), dst, nullptr, n,
                                          C_in, C_out, cnt, nn, adj_mode, lrelu_alpha, ctx->s_comp);
            if (rc != CCN_OK) return rc;
        }
        if (ro) {  // read-out head + loss on the last level's output; its gradient is produced into dGZ on the device
            rc = ccn_readout_forward(ctx, dZ, sY, nullptr, n, C_out, cnt, dPtr, ng, dW, dTgt, lrelu_alpha, dShr, dGFe, dPred, dLoss,
                                     ctx->s_comp);
            if (rc != CCN_OK) return rc;
            rc = ccn_readout_backward(ctx, dShr, dGFe, dPred, dTgt, dW, dIG, nullptr, n, C_out, cnt, ng, lrelu_alpha, dGZ, sY, dGW,
                                      ctx->s_comp);
            if (rc != CCN_OK) return rc;
        }
        for (int l = L - 1; l >= 0; --l) {
            const float *gout = l == L - 1 ? dGZ : dGMid(l & 1);
            float *gin = l == 0 ? dGF : dGMid((l - 1) & 1);
            const int64_t gin_words = l == 0 ? fcnt : cnt * sY;
            CCN_CUDA(ctx, cudaMemsetAsync(gin, 0, (size_t)gin_words * 4, ctx->s_comp));
            rc = ccn_gather_level_backward(ctx, gout, dXl(l), dYl(l), dKl(l), dBl(l), tA(l), tOff(l), tM(l), tPos(l), dGX, nullptr,
                                           l == 0 ? gin - f0 : gin - i0 * sY, dGKl(l), dGBl(l), nullptr, n, C_in, C_out, cnt, nn, adj_mode,
                                           lrelu_alpha, ctx->s_comp);
            if (rc != CCN_OK) return rc;
        }
        CCN_CUDA(ctx, cudaEventRecord(ctx->ev_comp[slot], ctx->s_comp));
        CCN_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_comp[slot], 0));
        if (ro) {
            CCN_CUDA(ctx, cudaMemcpyAsync(ro->predict_host + cut[c], dPred, (size_t)ng * 4, cudaMemcpyDeviceToHost, ctx->s_out));
            CCN_CUDA(ctx, cudaMemcpyAsync(ro->loss_host + cut[c], dLoss, (size_t)ng * 4, cudaMemcpyDeviceToHost, ctx->s_out));
        } else {
            CCN_CUDA(ctx, cudaMemcpyAsync(Z_host + i0 * sY, dZ, (size_t)cnt * sY * 4, cudaMemcpyDeviceToHost, ctx->s_out));
        }
        CCN_CUDA(ctx, cudaMemcpyAsync(gf_host + f0, dGF, (size_t)fcnt * 4, cudaMemcpyDeviceToHost, ctx->s_out));
        CCN_CUDA(ctx, cudaEventRecord(ctx->ev_out[slot], ctx->s_out));
    }
    for (int l = 0; l < L; ++l) {
        CCN_CUDA(ctx, cudaMemcpyAsync(gK_host[l], dGKl(l), (size_t)Kd * C_out * 4, cudaMemcpyDeviceToHost, ctx->s_comp));
        CCN_CUDA(ctx, cudaMemcpyAsync(gbias_host[l], dGBl(l), (size_t)C_out * 4, cudaMemcpyDeviceToHost, ctx->s_comp));
    }
    if (ro) CCN_CUDA(ctx, cudaMemcpyAsync(ro->gW_host, dGW, (size_t)C_out * 4, cudaMemcpyDeviceToHost, ctx->s_comp));
    CCN_CUDA(ctx, cudaStreamSynchronize(ctx->s_out));
    CCN_CUDA(ctx, cudaStreamSynchronize(ctx->s_comp));
    CCN_CUDA(ctx, cudaStreamSynchronize(ctx->s_in));
    return CCN_OK;
}

}  // namespace

extern "C" {

int ccn_gather_levels_forward_backward_host(ccn_ctx *ctx, int levels, const float *f_host, const int64_t *f_group_ptr,
                                            const int64_t *inst_group_ptr, int64_t groups, const int64_t *const *f_off_host,
                                            const int32_t *const *m_host, const int32_t *const *pos_host, const float *const *adj_host,
                                            const float *const *K_host, const float *const *bias_host, const float *gZ_host,
                                            float *Z_host, float *gf_host, float *const *gK_host, float *const *gbias_host, int n, int C_in,
                                            int C_out, int adj_mode, float lrelu_alpha) {
    return levels_host_impl(ctx, levels, f_host, f_group_ptr, inst_group_ptr, groups, f_off_host, m_host, pos_host, adj_host, K_host,
                            bias_host, gZ_host, Z_host, gf_host, gK_host, gbias_host, n, C_in, C_out, adj_mode, lrelu_alpha, nullptr);
}

int ccn_gather_levels_readout_forward_backward_host(ccn_ctx *ctx, int levels, const float *f_host, const int64_t *f_group_ptr,
                                                    const int64_t *inst_group_ptr, int64_t groups, const int64_t *const *f_off_host,
                                                    const int32_t *const *m_host, const int32_t *const *pos_host,
                                                    const float *const *adj_host, const float *const *K_host,
                                                    const float *const *bias_host, const float *W_host, const float *target_host,
                                                    float *predict_host, float *loss_host, float *gf_host, float *const *gK_host,
                                                    float *const *gbias_host, float *gW_host, int n, int C, int adj_mode,
                                                    float lrelu_alpha) {
    const ReadoutHost ro{W_host, target_host, predict_host, loss_host, gW_host};
    return levels_host_impl(ctx, levels, f_host, f_group_ptr, inst_group_ptr, groups, f_off_host, m_host, pos_host, adj_host, K_host,
                            bias_host, nullptr, nullptr, gf_host, gK_host, gbias_host, n, C, C, adj_mode, lrelu_alpha, &ro);
}

int ccn_gather_level_forward_backward_host(ccn_ctx *ctx, const float *f_host, const int64_t *f_group_ptr, const int64_t *inst_group_ptr,
                                           int64_t groups, const int64_t *f_off_host, const int32_t *m_host, const int32_t *pos_host,
                                           const float *adj_host, const float *K_host, const float *bias_host, const float *gZ_host,
                                           float *Z_host, float *gf_host, float *gK_host, float *gbias_host, int n, int C_in, int C_out,
                                           int adj_mode, float lrelu_alpha) {
    return ccn_gather_levels_forward_backward_host(ctx, 1, f_host, f_group_ptr, inst_group_ptr, groups, &f_off_host, &m_host, &pos_host,
                                                   &adj_host, &K_host, &bias_host, gZ_host, Z_host, gf_host, &gK_host, &gbias_host, n, C_in,
                                                   C_out, adj_mode, lrelu_alpha);
}

// ---- read-out head + loss ---------------------------------------------------------------------------------------------
int ccn_readout_forward(ccn_ctx *ctx, const float *Z_dev, int64_t stride_Z, const int32_t *n_dev, int n_max, int C, int64_t batch,
                        const int64_t *inst_graph_ptr_dev, int64_t graphs, const float *W_dev, const float *target_dev,
                        float lrelu_alpha, float *shrinked_dev, float *graph_feature_dev, float *predict_dev, float *loss_dev,
                        void *stream) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    if (!Z_dev || !inst_graph_ptr_dev || !W_dev || !shrinked_dev || !graph_feature_dev || !predict_dev || (loss_dev && !target_dev))
        return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "NULL argument");
    if (n_max <= 0 || C <= 0 || batch < 0 || graphs < 0) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "bad sizes");
    if (batch == 0 || graphs == 0) return CCN_OK;
    DeviceGuard g(ctx->device);
    LaunchLog log = make_log(ctx);
    CCN_CUDA(ctx, launch_readout_forward(Z_dev, stride_Z, n_dev, n_max, C, batch, inst_graph_ptr_dev, graphs, W_dev, target_dev, lrelu_alpha,
                                         shrinked_dev, graph_feature_dev, predict_dev, loss_dev, static_cast<cudaStream_t>(stream), &log));
    ctx->launches += log.launches;
    return CCN_OK;
}

int ccn_readout_backward(ccn_ctx *ctx, const float *shrinked_dev, const float *graph_feature_dev, const float *predict_dev,
                         const float *target_dev, const float *W_dev, const int32_t *inst_graph_dev, const int32_t *n_dev, int n_max,
                         int C, int64_t batch, int64_t graphs, float lrelu_alpha, float *gZ_dev, int64_t stride_gZ, float *gW_dev,
                         void *stream) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    if (!shrinked_dev || !graph_feature_dev || !predict_dev || !target_dev || !W_dev || !inst_graph_dev || !gZ_dev)
        return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "NULL argument");
    if (n_max <= 0 || C <= 0 || batch < 0 || graphs < 0) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "bad sizes");
    if (batch == 0 || graphs == 0) return CCN_OK;
    DeviceGuard g(ctx->device);
    LaunchLog log = make_log(ctx);
    CCN_CUDA(ctx, launch_readout_backward(shrinked_dev, graph_feature_dev, predict_dev, target_dev, W_dev, inst_graph_dev, n_dev, n_max, C,
                                          batch, graphs, lrelu_alpha, gZ_dev, stride_gZ, gW_dev, static_cast<cudaStream_t>(stream), &log));
    ctx->launches += log.launches;
    return CCN_OK;
}

int ccn_level_features_forward(ccn_ctx *ctx, const float *Z_dev, int64_t stride_Z, const int32_t *n_dev, int n_max, int C, int64_t batch,
                               const int64_t *inst_graph_ptr_dev, int64_t graphs, float lrelu_alpha, float *shrinked_dev,
                               float *feature_dev, int64_t ld_feature, void *stream) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    if (!Z_dev || !inst_graph_ptr_dev || !shrinked_dev || !feature_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "NULL argument");
    if (n_max <= 0 || C <= 0 || batch < 0 || graphs < 0 || ld_feature < C) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "bad sizes");
    if (batch == 0 || graphs == 0) return CCN_OK;
    DeviceGuard g(ctx->device);
    LaunchLog log = make_log(ctx);
    CCN_CUDA(ctx, launch_level_features_forward(Z_dev, stride_Z, n_dev, n_max, C, batch, inst_graph_ptr_dev, graphs, lrelu_alpha, shrinked_dev,
                                                feature_dev, ld_feature, static_cast<cudaStream_t>(stream), &log));
    ctx->launches += log.launches;
    return CCN_OK;
}

int ccn_level_features_backward(ccn_ctx *ctx, const float *shrinked_dev, const float *dfeature_dev, int64_t ld_feature,
                                const int32_t *inst_graph_dev, const int32_t *n_dev, int n_max, int C, int64_t batch, float lrelu_alpha,
                                float *gZ_dev, int64_t stride_gZ, void *stream) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    if (!shrinked_dev || !dfeature_dev || !inst_graph_dev || !gZ_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "NULL argument");
    if (n_max <= 0 || C <= 0 || batch < 0 || ld_feature < C) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "bad sizes");
    if (batch == 0) return CCN_OK;
    DeviceGuard g(ctx->device);
    LaunchLog log = make_log(ctx);
    CCN_CUDA(ctx, launch_level_features_backward(shrinked_dev, dfeature_dev, ld_feature, inst_graph_dev, n_dev, n_max, C, batch, lrelu_alpha,
                                                 gZ_dev, stride_gZ, static_cast<cudaStream_t>(stream), &log));
    ctx->launches += log.launches;
    return CCN_OK;
}

// ---- pinning the caller's own host arrays (the reference allocates value[] / gradient[] with plain new[], Vector.h:24-25) ----
int ccn_host_register(ccn_ctx *ctx, void *ptr_host, size_t bytes) {
    if (!ctx || !ptr_host || bytes == 0) return CCN_ERR_INVALID_ARGUMENT;
    DeviceGuard g(ctx->device);
    CCN_CUDA(ctx, cudaHostRegister(ptr_host, bytes, cudaHostRegisterPortable));
    return CCN_OK;
}

int ccn_host_unregister(ccn_ctx *ctx, void *ptr_host) {
    if (!ctx || !ptr_host) return CCN_ERR_INVALID_ARGUMENT;
    DeviceGuard g(ctx->device);
    CCN_CUDA(ctx, cudaHostUnregister(ptr_host));
    return CCN_OK;
}

// ---- gradient all-reduce: NCCL bound lazily (no link dependency) ----------------------------------------------------
namespace {
typedef int (*nccl_allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef const char *(*nccl_errstr_fn)(int);
nccl_allreduce_fn g_nccl_allreduce = nullptr;
nccl_errstr_fn g_nccl_errstr = nullptr;
std::once_flag g_nccl_once;

void bind_nccl_once() {
    void *sym = dlsym(RTLD_DEFAULT, "ncclAllReduce");
    void *h = nullptr;
    if (!sym) {
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (h) sym = dlsym(h, "ncclAllReduce");
    }
    g_nccl_allreduce = reinterpret_cast<nccl_allreduce_fn>(sym);
    void *es = dlsym(h ? h : RTLD_DEFAULT, "ncclGetErrorString");
    g_nccl_errstr = reinterpret_cast<nccl_errstr_fn>(es);
}
}  // namespace

int ccn_allreduce_grads(ccn_ctx *ctx, void *nccl_comm, float *buf_dev, int64_t count, void *stream) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    if (!nccl_comm || !buf_dev || count < 0) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "NULL communicator / buffer or negative count");
    if (count == 0) return CCN_OK;
    std::call_once(g_nccl_once, bind_nccl_once);  // the two pointers are written once, before any reader passes this line
    if (!g_nccl_allreduce) return fail(ctx, CCN_ERR_UNSUPPORTED, "no NCCL in this process and libnccl.so.2 not found");
    DeviceGuard g(ctx->device);
    // ncclFloat32 = 7, ncclSum = 0 (nccl.h: ncclDataType_t, ncclRedOp_t -- stable across NCCL 2.x)
    const int rc = g_nccl_allreduce(buf_dev, buf_dev, (size_t)count, 7, 0, nccl_comm, static_cast<cudaStream_t>(stream));
    if (rc != 0) return fail(ctx, CCN_ERR_CUDA, std::string("ncclAllReduce: ") + (g_nccl_errstr ? g_nccl_errstr(rc) : "error"));
    return CCN_OK;
}

int ccn_adam_step(ccn_ctx *ctx, float *params_dev, const float *grads_dev, float *m_dev, float *v_dev, int64_t count, double alpha,
                  double beta1, double beta2, double epsilon, int n_batch, int64_t updates_before, int per_element_bias,
                  void *stream) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    if (!params_dev || !grads_dev || !m_dev || !v_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "NULL argument");
    if (count < 0 || n_batch <= 0 || updates_before < 0 || !(beta1 > 0 && beta1 < 1) || !(beta2 > 0 && beta2 < 1))
        return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "bad optimizer arguments");
    if (count == 0) return CCN_OK;
    DeviceGuard g(ctx->device);
    LaunchLog log = make_log(ctx);
    CCN_CUDA(ctx, launch_adam_step(params_dev, grads_dev, m_dev, v_dev, count, alpha, beta1, beta2, epsilon, 1.0 / n_batch,
                                   updates_before, per_element_bias != 0, static_cast<cudaStream_t>(stream), &log));
    ctx->launches += log.launches;
    return CCN_OK;
}

int ccn_momentum_step(ccn_ctx *ctx, float *params_dev, const float *grads_dev, float *moments_dev, int64_t count,
                      double learning_rate, double gamma, int n_batch, void *stream) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    if (!params_dev || !grads_dev || !moments_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "NULL argument");
    if (count < 0 || n_batch <= 0) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "bad optimizer arguments");
    if (count == 0) return CCN_OK;
    DeviceGuard g(ctx->device);
    LaunchLog log = make_log(ctx);
    CCN_CUDA(ctx, launch_momentum_step(params_dev, grads_dev, moments_dev, count, learning_rate, gamma, 1.0 / n_batch,
                                       static_cast<cudaStream_t>(stream), &log));
    ctx->launches += log.launches;
    return CCN_OK;
}

int ccn_tensor_mul_forward(ccn_ctx *ctx, const float *A_dev, const float *B_dev, float *out_dev, int R, int K, int Cc, int D,
                           int64_t batch, void *stream) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    if (!A_dev || !B_dev || !out_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "NULL argument");
    if (R <= 0 || K <= 0 || Cc <= 0 || D <= 0 || batch < 0 || batch > 65535) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "bad sizes");
    if (batch == 0) return CCN_OK;
    DeviceGuard g(ctx->device);
    LaunchLog log = make_log(ctx);
    CCN_CUDA(ctx, launch_tensor_mul_forward(A_dev, B_dev, out_dev, R, K, Cc, D, (int)batch, static_cast<cudaStream_t>(stream), &log));
    ctx->launches += log.launches;
    return CCN_OK;
}

int ccn_tensor_mul_backward(ccn_ctx *ctx, const float *A_dev, const float *B_dev, const float *gout_dev, float *gA_dev,
                            float *gB_dev, int R, int K, int Cc, int D, int64_t batch, float beta, void *stream) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    if (!A_dev || !B_dev || !gout_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "NULL argument");
    if (R <= 0 || K <= 0 || Cc <= 0 || D <= 0 || batch < 0 || batch > 65535) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "bad sizes");
    if (batch == 0) return CCN_OK;
    DeviceGuard g(ctx->device);
    LaunchLog log = make_log(ctx);
    CCN_CUDA(ctx, launch_tensor_mul_backward(A_dev, B_dev, gout_dev, gA_dev, gB_dev, R, K, Cc, D, (int)batch, beta,
                                             static_cast<cudaStream_t>(stream), &log));
    ctx->launches += log.launches;
    return CCN_OK;
}

int ccn_custom_matmul_tensor_forward(ccn_ctx *ctx, const float *Kt_dev, const float *X_dev, float *Y_dev, int64_t M, int V, int P,
                                     void *stream) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    if (!Kt_dev || !X_dev || !Y_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "NULL argument");
    if (M < 0 || V <= 0 || P <= 0) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "bad GEMM shape");
    if (M == 0) return CCN_OK;
    DeviceGuard g(ctx->device);
    ScratchScope scope(ctx, static_cast<cudaStream_t>(stream));
    if (scope.rc != CCN_OK) return scope.rc;
    int rc = ensure_aux(ctx, (size_t)2 * V * P * sizeof(float));
    if (rc != CCN_OK) return rc;
    LaunchLog log = make_log(ctx);
    CCN_CUDA(ctx, launch_transpose_add(Kt_dev, ctx->aux, P, V, 0.f, static_cast<cudaStream_t>(stream), &log));  // W[V,P] = Kt^T
    ctx->launches += log.launches;
    return ccn_mix_forward(ctx, X_dev, ctx->aux, nullptr, Y_dev, nullptr, M, V, P, 0.f, stream);
}

int ccn_custom_matmul_tensor_backward(ccn_ctx *ctx, const float *Kt_dev, const float *X_dev, const float *gY_dev, float *gKt_dev,
                                      float *gX_dev, int64_t M, int V, int P, float beta_x, void *stream) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    if (!Kt_dev || !X_dev || !gY_dev) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "NULL argument");
    if (M < 0 || V <= 0 || P <= 0) return fail(ctx, CCN_ERR_INVALID_ARGUMENT, "bad GEMM shape");
    if (M == 0) return CCN_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ScratchScope scope(ctx, st);
    if (scope.rc != CCN_OK) return scope.rc;
    int rc = ensure_aux(ctx, (size_t)2 * V * P * sizeof(float));
    if (rc != CCN_OK) return rc;
    float *W = ctx->aux, *gW = ctx->aux + (size_t)V * P;
    LaunchLog log = make_log(ctx);
    CCN_CUDA(ctx, launch_transpose_add(Kt_dev, W, P, V, 0.f, st, &log));
    CCN_CUDA(ctx, cudaMemsetAsync(gW, 0, (size_t)V * P * sizeof(float), st));
    ctx->launches += log.launches;
    rc = ccn_mix_backward(ctx, X_dev, W, nullptr, nullptr, gY_dev, gX_dev, gKt_dev ? gW : nullptr, nullptr, M, V, P, 0.f, beta_x, stream);
    if (rc != CCN_OK) return rc;
    if (gKt_dev) {
        LaunchLog log2 = make_log(ctx);
        CCN_CUDA(ctx, launch_transpose_add(gW, gKt_dev, V, P, 1.f, st, &log2));  // gKt[P,V] += gW^T
        ctx->launches += log2.launches;
    }
    return CCN_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------------------------
int ccn_device_alloc(ccn_ctx *ctx, void **ptr_dev, size_t bytes) {
    if (!ctx || !ptr_dev) return CCN_ERR_INVALID_ARGUMENT;
    DeviceGuard g(ctx->device);
    CCN_CUDA(ctx, cudaMalloc(ptr_dev, bytes ? bytes : 16));
    return CCN_OK;
}

int ccn_device_free(ccn_ctx *ctx, void *ptr_dev) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    DeviceGuard g(ctx->device);
    CCN_CUDA(ctx, cudaFree(ptr_dev));
    return CCN_OK;
}

int ccn_h2d(ccn_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes, void *stream) {
    if (!ctx || (bytes && (!dst_dev || !src_host))) return CCN_ERR_INVALID_ARGUMENT;
    DeviceGuard g(ctx->device);
    CCN_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
    return CCN_OK;
}

int ccn_d2h(ccn_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes, void *stream) {
    if (!ctx || (bytes && (!dst_host || !src_dev))) return CCN_ERR_INVALID_ARGUMENT;
    DeviceGuard g(ctx->device);
    CCN_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    return CCN_OK;
}

int ccn_memset_zero(ccn_ctx *ctx, void *dst_dev, size_t bytes, void *stream) {
    if (!ctx || (bytes && !dst_dev)) return CCN_ERR_INVALID_ARGUMENT;
    DeviceGuard g(ctx->device);
    CCN_CUDA(ctx, cudaMemsetAsync(dst_dev, 0, bytes, static_cast<cudaStream_t>(stream)));
    return CCN_OK;
}

int ccn_stream_synchronize(ccn_ctx *ctx, void *stream) {
    if (!ctx) return CCN_ERR_INVALID_ARGUMENT;
    DeviceGuard g(ctx->device);
    CCN_CUDA(ctx, cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    if (*(volatile int *)ctx->fault_host != 0)
        return fail(ctx, CCN_ERR_CUDA, "a fused-path tile timed out waiting for its siblings; the results of that launch are invalid");
    return CCN_OK;
}

int ccn_stream_create(ccn_ctx *ctx, void **stream) {
    if (!ctx || !stream) return CCN_ERR_INVALID_ARGUMENT;
    DeviceGuard g(ctx->device);
    cudaStream_t st = nullptr;
    CCN_CUDA(ctx, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    *stream = st;
    return CCN_OK;
}

int ccn_stream_destroy(ccn_ctx *ctx, void *stream) {
    if (!ctx || !stream) return CCN_ERR_INVALID_ARGUMENT;
    DeviceGuard g(ctx->device);
    CCN_CUDA(ctx, cudaStreamDestroy(static_cast<cudaStream_t>(stream)));
    return CCN_OK;
}

}  // extern "C"
