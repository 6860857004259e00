/*
 * ccn_b200.h -- C-ABI of the B200-native (sm_100a) second-order CCN message-passing hot path.
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes, no CUDA or torch types in any signature
 * (a stream is passed as `void*` holding a cudaStream_t; NULL = the legacy default stream).  Every entry
 * point names the reference interface (HyTruongSon/GraphFlow, file:line) that it replaces.  The
 * reference-side bindings (header-compatible C++ classes, ctypes) are shown in INTEGRATION.md and shipped in
 * include/graphflow_b200/ and graphflow_b200/.
 *
 * Conventions
 *   - All tensors are fp32, row-major, channel innermost -- the reference's own layouts:
 *       T    [n, n, n, C]   ((a*n + b)*n + c)*C + f     Tensor4D::index            (Tensor4D.h:29-31)
 *       adj  [n, n]         d*n + e                      Matrix::index              (Matrix.h:34-36)
 *       out  [n, n, 18*C]   (x*n + y)*18C + k*C + f      Tensor3D::index, slab k    (Tensor3D.h:37-39,
 *                                                                                    RisiContraction_18.h:102-318)
 *   - A *batch* is `batch` independent instances (one per (graph, vertex, level)).  Instance i lives at
 *     base + i*stride (strides in ELEMENTS) and is stored densely with its own n_i = n[i] <= n_max
 *     (n == NULL means every instance has n_i = n_max).  Ragged batches are first class.
 *   - `*_dev` pointers are device pointers of the ctx's device; `*_host` pointers are host pointers
 *     (pageable or pinned).  Nothing here allocates on behalf of the caller except the ctx workspace.
 *   - Every function returns a ccn_status (0 = CCN_OK, negative = error) and never throws or aborts.
 *     ccn_last_error() gives a human-readable message for the last failure on that ctx.
 *   - Re-entrant across contexts; one context must be used by one host thread at a time (the reference's own
 *     rule for op instances, SMP_beta.h:722-729).  No global mutable state.
 *   - A context's scratch is shared by all of its calls, so its calls execute in issue order even when they name
 *     different streams: a call on another stream than the previous one first waits (on the device, no host
 *     synchronisation) for the previous call.  For concurrent streams use one context per stream, like the
 *     reference's one-op-instance-per-thread replicas (SMP_beta_gpu_multistreams.h:701-718).
 *   - Slab-pointer tables (slabs_dev / gslabs_dev) must hold 16-byte aligned pointers; base pointers and strides
 *     that are not 16-byte aligned are accepted and take the shape-generic kernels.
 *   - adj_mode: CCN_ADJ_POSITIVE_PART reproduces RisiContraction_18 (entries <= 0 are skipped,
 *     RisiContraction_18.h:90,345); CCN_ADJ_RAW reproduces RisiContraction_18_thread / _50 (raw product,
 *     RisiContraction_18_thread.h:70-72).  Identical for the 0/1(+I) adjacency every model builds
 *     (SMP_beta.h:505-526).
 */
#ifndef CCN_B200_H_INCLUDED
#define CCN_B200_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CCN_API __attribute__((visibility("default")))
#else
#define CCN_API
#endif

#define CCN_B200_ABI_VERSION 2
#define CCN_NUM_CONTRACTIONS 18 /* RisiContraction_18_gpu::nContractions, RisiContraction_18_gpu.h:1749 */

typedef struct ccn_ctx ccn_ctx;

typedef enum {
    CCN_OK = 0,
    CCN_ERR_INVALID_ARGUMENT = -1, /* NULL pointer, non-positive size, misaligned base, bad mode */
    CCN_ERR_CUDA = -2,             /* a CUDA runtime call or kernel launch failed                 */
    CCN_ERR_OUT_OF_MEMORY = -3,    /* workspace / staging allocation failed                       */
    CCN_ERR_NO_DEVICE = -4,        /* no usable sm_100 device                                     */
    CCN_ERR_UNSUPPORTED = -5       /* shape outside what this build supports                      */
} ccn_status;

enum { CCN_ADJ_POSITIVE_PART = 0, CCN_ADJ_RAW = 1 };

/* ---- context -------------------------------------------------------------------------------------------------
 * Replaces the per-object cudaMalloc/cudaFree in the reference op constructors/destructors
 * (RisiContraction_18_gpu.h:849-918, 1797-1803; MatMul_gpu.h:115-165): one context owns the scratch
 * workspace, the adjacency tables and the device staging ring of the host-buffer entry points, and is reused
 * across calls. */
CCN_API int ccn_ctx_create(ccn_ctx **ctx, int device);
CCN_API int ccn_ctx_destroy(ccn_ctx *ctx);
CCN_API const char *ccn_last_error(const ccn_ctx *ctx);
CCN_API const char *ccn_status_string(int status);
CCN_API int ccn_abi_version(void);
/* Upper bound (bytes) for the scratch the context may hold at once; batches are processed in chunks that
 * fit.  Default 96 MiB, chosen to stay inside the 126 MB L2 together with the streamed data's footprint. */
CCN_API int ccn_ctx_set_workspace_limit(ccn_ctx *ctx, size_t bytes);
/* Number of kernels launched by this context since creation (bench.py's gpu_launches evidence). */
CCN_API int64_t ccn_ctx_kernel_launches(const ccn_ctx *ctx);
/* Per-kernel device timing: when enabled, every kernel launch of this context is bracketed by CUDA events on the
 * launching stream.  set(…) also clears the totals; get(…) synchronises on the recorded events and returns the
 * summed duration and launch count of one kernel id in [0, ccn_num_kernels()).  Used by bench.py's roofline. */
CCN_API int ccn_num_kernels(void);
CCN_API const char *ccn_kernel_name(int kernel_id);
CCN_API int ccn_ctx_set_kernel_timing(ccn_ctx *ctx, int enable);
CCN_API int ccn_ctx_get_kernel_timing(ccn_ctx *ctx, int kernel_id, double *total_ms, int64_t *launches);
/* Selects the contraction implementation.  CCN_PATH_AUTO: the fused single-kernel path when the shape allows
 * (n_max <= 32, C in {8, 16, 32, 64, 128}), else the generic kernels.  CCN_PATH_GENERIC: the shape-agnostic kernels. */
enum { CCN_PATH_AUTO = 0, CCN_PATH_GENERIC = 1 };
CCN_API int ccn_ctx_set_kernel_path(ccn_ctx *ctx, int path);
/* Selects the feature-mix forward implementation.  CCN_MIX_AUTO: tcgen05 tensor cores with split-precision (3xTF32,
 * fp32-accurate) operands when the shape allows (K % 4 == 0, P % 4 == 0, P <= 128 forward / P <= 64 backward), else the fp32 SIMT
 * kernel.  CCN_MIX_SIMT: always the SIMT kernel.  CCN_MIX_TENSOR: tensor cores or CCN_ERR_UNSUPPORTED. */
enum { CCN_MIX_AUTO = 0, CCN_MIX_SIMT = 1, CCN_MIX_TENSOR = 2 };
CCN_API int ccn_ctx_set_mix_path(ccn_ctx *ctx, int path);
/* A fused-path tile that gives up waiting for its sibling tiles (a bug or a wedged device, never a normal wait) sets a
 * STICKY flag on the context: every later entry point that launches work, and ccn_stream_synchronize, returns
 * CCN_ERR_CUDA while it is set.  This call synchronises the device, reports the flag (0 = never happened) and clears it. */
CCN_API int ccn_ctx_fused_error_flag(ccn_ctx *ctx, int *flag);
/* While frozen != 0 no context-owned buffer may grow: a call that would have to reallocate scratch returns
 * CCN_ERR_UNSUPPORTED instead.  Set it after the warm-up calls and before capturing the context's launches in a CUDA
 * graph (the graph bakes the scratch pointers in); clear it when the graph is destroyed. */
CCN_API int ccn_ctx_set_frozen(ccn_ctx *ctx, int frozen);
/* Profiling aid: when trace_dev != NULL, thread 0 of every fused-path tile (work item w = instance * tiles + tile,
 * in ticket order) stores up to 8 %globaltimer marks at trace_dev[8*w .. 8*w+7] (uint64 nanoseconds) for calls with
 * at most `tiles` work items.  NULL switches it off.  The buffer is owned by the caller. */
CCN_API int ccn_ctx_set_phase_trace(ccn_ctx *ctx, void *trace_dev, int64_t tiles);

/* ---- StackTensor3D + RisiContraction_18, forward ---------------------------------------------------------------
 * Replaces StackTensor3D::forward (StackTensor3D.h:54-72) followed by RisiContraction_18::forward
 * (RisiContraction_18.h:73-331) / RisiContraction_18_gpu::forward_GPU + kernel
 * RisiContraction_18_forward_job (RisiContraction_18_gpu.h:49-379, 1509-1568), for a whole batch.
 *
 * Input is EITHER the stacked tensor `T_dev` (instance i at T_dev + i*stride_T)   [RisiContraction_18_gpu API]
 *          OR a table of slab pointers `slabs_dev` (device array of batch*n_max device pointers; entry
 *          i*n_max + a points at vertex a's [n_i, n_i, C] tensor) which fuses the stack into the read
 *          [RisiContraction_18::add_tensor API, RisiContraction_18.h:49-55].  Exactly one must be non-NULL.
 * out_dev is fully overwritten (the reference zeroes then accumulates, :76-78). */
CCN_API int ccn_contract18_forward(ccn_ctx *ctx, const float *T_dev, const float *const *slabs_dev, const float *adj_dev,
                           float *out_dev, const int32_t *n_dev, int n_max, int C, int64_t batch, int64_t stride_T,
                           int64_t stride_adj, int64_t stride_out, int adj_mode, void *stream);

/* ---- RisiContraction_18 + StackTensor3D, backward --------------------------------------------------------------
 * Replaces RisiContraction_18::backward (RisiContraction_18.h:333-560) / RisiContraction_18_gpu::backward_GPU +
 * kernel RisiContraction_18_backward_job (RisiContraction_18_gpu.h:541-685, 1639-1689) followed by
 * StackTensor3D::backward (StackTensor3D.h:74-90).
 *
 * gT = beta*gT + contraction^T(gout).  beta = 1 is the reference's `+=` into tensors[a]->gradient; beta = 0
 * writes a fresh gradient without reading gT (what a graph executor needs after Vector::forward zeroed it,
 * Vector.h:28-32).  Destination is `gT_dev` (stacked) or `gslabs_dev` (pointer table as above). */
CCN_API int ccn_contract18_backward(ccn_ctx *ctx, const float *gout_dev, const float *adj_dev, float *gT_dev,
                            float *const *gslabs_dev, const int32_t *n_dev, int n_max, int C, int64_t batch,
                            int64_t stride_gout, int64_t stride_adj, int64_t stride_gT, int adj_mode, float beta,
                            void *stream);

/* ---- StackTensor3D + RisiContraction_50 --------------------------------------------------------------------------
 * Replaces RisiContraction_50::forward / backward (RisiContraction_50.h:73-441, 443-802): all 50 ways of keeping two
 * of the five indices of T[a,b,c,f] * adj[d,e]; out is [n, n, 50*C], slab k-1 at depth (k-1)*C + f (:96).  Same
 * argument meaning as the 18-way entry points.  The reference multiplies by the raw adjacency entry here
 * (value_at, :63-65), so pass CCN_ADJ_RAW for reference semantics. */
#define CCN_NUM_CONTRACTIONS_50 50
CCN_API int ccn_contract50_forward(ccn_ctx *ctx, const float *T_dev, const float *const *slabs_dev, const float *adj_dev,
                           float *out_dev, const int32_t *n_dev, int n_max, int C, int64_t batch, int64_t stride_T,
                           int64_t stride_adj, int64_t stride_out, int adj_mode, void *stream);
CCN_API int ccn_contract50_backward(ccn_ctx *ctx, const float *gout_dev, const float *adj_dev, float *gT_dev,
                            float *const *gslabs_dev, const int32_t *n_dev, int n_max, int C, int64_t batch,
                            int64_t stride_gout, int64_t stride_adj, int64_t stride_gT, int adj_mode, float beta,
                            void *stream);

/* ---- the other members of the contraction family -------------------------------------------------------------------
 * One entry point pair for the operators that differ from RisiContraction_50 only in WHICH index patterns they emit:
 *   variant 50  RisiContraction_50                                   out [n, n, 50 C]
 *   variant 10  RisiContraction_10::forward/backward (RisiContraction_10.h:72-150,152-230): cases 1..10 of the 50
 *               (three summed indices), raw adjacency                out [n, n, 10 C]
 *   variant 4   RisiContraction_4::forward/backward (RisiContraction_4.h:68-123,125-180): sum_c T[a,b,c] -> (a,b),
 *               sum_a -> (b,c), T[a,a,c] -> (a,c), T[a,b,b] -> (a,b); no adjacency (adj_dev may be NULL)
 *                                                                    out [n, n, 4 C]
 *   variant 18  the 18 slabs of RisiContraction_18 (run by the 18-way kernels, fused where supported, plus a slab-mask
 *               pass); with keep_mask this is
 *               RisiContraction_18_dropout::forward/backward (RisiContraction_18_dropout.h:104-478,480-797): bit k of
 *               keep_mask = use[k] of the reference (:113-131, chosen by the caller); a dropped slab is written as
 *               zeros in the forward and ignored in the backward; pass CCN_ADJ_POSITIVE_PART (`adj_value > 0`, :148)
 * Bits of keep_mask above the variant's slab count are ignored; ~0 keeps everything.  out_scale multiplies the forward
 * output: 1, or nKept/18 for the dropout operator's test mode (:467-472; all slabs kept, no backward in the
 * reference).  All other arguments as in ccn_contract50_forward / _backward. */
CCN_API int ccn_contract_family_forward(ccn_ctx *ctx, int variant, uint64_t keep_mask, const float *T_dev,
                                const float *const *slabs_dev, const float *adj_dev, float *out_dev, const int32_t *n_dev,
                                int n_max, int C, int64_t batch, int64_t stride_T, int64_t stride_adj, int64_t stride_out,
                                int adj_mode, float out_scale, void *stream);
CCN_API int ccn_contract_family_backward(ccn_ctx *ctx, int variant, uint64_t keep_mask, const float *gout_dev,
                                 const float *adj_dev, float *gT_dev, float *const *gslabs_dev, const int32_t *n_dev,
                                 int n_max, int C, int64_t batch, int64_t stride_gout, int64_t stride_adj,
                                 int64_t stride_gT, int adj_mode, float beta, void *stream);

/* ---- host-buffer (end-to-end) variants -------------------------------------------------------------------------
 * Same operators with HOST arrays, as the reference op classes present them (value[]/gradient[] live on the host,
 * Vector.h:22-26; the reference does H2D -> kernel -> D2H per call, RisiContraction_18_gpu.h:1523-1540).  The
 * batch is cut into chunks that are uploaded, computed and downloaded on three streams through a DEVICE staging
 * ring (three slots of about 256 MiB), so PCIe transfers in both directions overlap the kernels.  The host arrays are
 * used where they lie: pinned arrays are copied at PCIe speed; pageable arrays (the reference's plain `new[]`,
 * Vector.h:24-25) go through the driver's bounce buffers unless ccn_host_register pinned them first.  Uniform n (= n_max) per call; instances are contiguous
 * (stride = dense instance size).  Synchronous: results are in the host arrays on return. */
CCN_API int ccn_contract18_forward_host(ccn_ctx *ctx, const float *T_host, const float *adj_host, float *out_host, int n,
                                int C, int64_t batch, int adj_mode);
CCN_API int ccn_contract18_backward_host(ccn_ctx *ctx, const float *gout_host, const float *adj_host, float *gT_host, int n,
                                 int C, int64_t batch, int adj_mode, float beta);
/* forward + backward of the same batch in one pass over the staging ring (bench.py's e2e metric):
 * uploads T, adj, gout; downloads out and gT (beta = 0). */
CCN_API int ccn_contract18_forward_backward_host(ccn_ctx *ctx, const float *T_host, const float *adj_host,
                                         const float *gout_host, float *out_host, float *gT_host, int n, int C,
                                         int64_t batch, int adj_mode);

/* ---- feature mix ------------------------------------------------------------------------------------------------
 * Replaces Reshape2D (Reshape2D.h:42-71) + MatMul::forward (MatMul.h:48-66) / MatMul_gpu::forward_GPU + kernel
 * Matrix_Multiplication_GPU (MatMul_gpu.h:28-65, 273-313) [+ Reshape3D + VectorAddTensor::forward
 * (VectorAddTensor.h:46-59) + LeakyReLU3D::forward (LeakyReLU3D.h:60-72) when bias_dev != NULL]:
 *   Y[M, P] = X[M, K] * W[K, P]           (for the CCN level: M = sum n_i^2, K = 18*C, P = C_out)
 *   Z[M, P] = lrelu(Y + bias, alpha)       only if bias_dev and Z_dev are given (alpha = 0.01 in every model)
 * Y_dev may be NULL when only Z is wanted.  fp32 in / fp32 out; products accumulate in fp32. */
CCN_API int ccn_mix_forward(ccn_ctx *ctx, const float *X_dev, const float *W_dev, const float *bias_dev, float *Y_dev,
                    float *Z_dev, int64_t M, int K, int P, float lrelu_alpha, void *stream);

/* Replaces LeakyReLU3D::backward (LeakyReLU3D.h:74-82) + VectorAddTensor::backward (VectorAddTensor.h:61-71) +
 * MatMul::backward (MatMul.h:68-82) / kernels MatMul_backward_first/second (MatMul_gpu.h:71-111, 412-466):
 *   gY = gZ * (Y + bias > 0 ? 1 : alpha)           (if bias_dev != NULL, else gY = gZ_dev)
 *   gX = beta_x*gX + gY * W^T ;  gW += X^T * gY ;  gbias += column sums of gY
 * gW and gbias always accumulate (they are parameter gradients summed over a batch, SMP_beta.h:677-687).
 * Any of gX_dev / gW_dev / gbias_dev may be NULL to skip that product. */
CCN_API int ccn_mix_backward(ccn_ctx *ctx, const float *X_dev, const float *W_dev, const float *bias_dev, const float *Y_dev,
                     const float *gZ_dev, float *gX_dev, float *gW_dev, float *gbias_dev, int64_t M, int K, int P,
                     float lrelu_alpha, float beta_x, void *stream);

/* ---- promotion (+ stack) ----------------------------------------------------------------------------------------------
 * Replaces MatTensorMul::forward (MatTensorMul.h:47-65) + TensorMatMul::forward (TensorMatMul.h:46-64) as wired at
 * SMP_beta.h:588-594 with the 0/1 selection matrices of init_permutation_matrix (SMP_beta.h:446-459), followed by
 * StackTensor3D::forward (StackTensor3D.h:54-72): slab a of instance i of the stacked T is
 *     T[i][a][r][c][:] = f[f_off[i*n_max + a]][pos[r], pos[c], :]   (zero when pos[r] < 0 or pos[c] < 0)
 * where the source tensor f_{l-1}[w] is [m, m, C] with m = m_dev[i*n_max + a] and pos = pos_dev + (i*n_max + a)*n_max
 * gives, for every member of phi_l(v), its position inside phi_{l-1}(w) or -1.  The backward
 * (MatTensorMul.h:67-85, TensorMatMul.h:66-84, StackTensor3D.h:74-90) adds gT back into gf (atomically: one f_{l-1}[w] is
 * promoted into many stacks); no gradient flows to the selection matrices' entries that are structurally zero.
 * Every f_off entry must be a multiple of C (true whenever the level l-1 tensors, m*m*C elements each, are packed back
 * to back): with C % 4 == 0 the copies then run as 16-byte vector accesses / vector atomics. */
CCN_API int ccn_promote_forward(ccn_ctx *ctx, const float *f_dev, const int64_t *f_off_dev, const int32_t *m_dev,
                        const int32_t *pos_dev, float *T_dev, const int32_t *n_dev, int n_max, int C, int64_t batch,
                        int64_t stride_T, void *stream);
CCN_API int ccn_promote_backward(ccn_ctx *ctx, const float *gT_dev, const int64_t *f_off_dev, const int32_t *m_dev,
                         const int32_t *pos_dev, float *gf_dev, const int32_t *n_dev, int n_max, int C, int64_t batch,
                         int64_t stride_T, void *stream);

/* ---- chained entry points: one call per stage of a CCN level ---------------------------------------------------------
 * The per-vertex block of SMP_beta::complete_computation_graph (SMP_beta.h:588-616) for a whole batch of vertices:
 *   ccn_gather_contract18_*   MatTensorMul + TensorMatMul (promotion) + StackTensor3D + RisiContraction_18: f_{l-1} -> out.
 *                             For the shapes of the fused kernels (n_max <= 32, C in {8,16,32,64,128}, f_dev 16-byte aligned)
 *                             this is ONE kernel per direction: the forward reads slab a of the stack straight out of
 *                             f_{l-1} through the promotion table inside its streaming loop, the backward adds the rows of
 *                             gT into gf with reductions as it produces them -- the stacked T / gT is never materialised
 *                             and T_scratch_dev / gT_scratch_dev may be NULL.  Other shapes run promotion + contraction in
 *                             sequence through the caller-provided scratch ([batch, n_max^3 C]).  The backward ADDS into gf
 *                             (one f_{l-1}[w] feeds many stacks): zero it first for a fresh gradient.
 *   ccn_level_*               RisiContraction_18 + Reshape2D + MatMul(K) + Reshape3D + VectorAddTensor(b) + LeakyReLU3D:
 *                             T -> X [batch, n_max^2, 18 C_in] (kept for the backward) -> Y = X K -> Z = lrelu(Y + b),
 *                             Y, Z [batch * n_max^2, C_out]; the backward gives gT (beta as in ccn_contract18_backward)
 *                             and accumulates gK, gbias; gX_scratch_dev [batch, n_max^2, 18 C_in] is scratch.
 *   ccn_gather_level_*        both of the above chained: f_{l-1} -> X -> Y -> Z and gZ -> gX -> gf, gK, gbias: the whole
 *                             per-vertex block of SMP_beta.h:588-616 for a batch of vertices, reading only the level l-1
 *                             tensors and the index tables.
 *   ccn_gather_level_forward_backward_host
 *                             the same with HOST arrays, for callers whose activations live on the host (the reference's
 *                             value[] / gradient[]): only f_{l-1} (n^2 C per vertex, not the n^3 C stack), gZ and the tables
 *                             are uploaded, only Z and gf downloaded.  Instances come in `groups` (graphs): group q owns
 *                             instances [inst_group_ptr[q], inst_group_ptr[q+1]) and its instances reference only
 *                             f_host[f_group_ptr[q] .. f_group_ptr[q+1]) (f_off_host entries are absolute element offsets
 *                             into f_host; boundaries multiples of 4).  Consecutive groups are cut into chunks that are
 *                             uploaded, computed and downloaded on three streams through the device staging ring.  Uniform
 *                             n per call.  Z_host [batch n^2, C_out] and gf_host (same layout as f_host) are overwritten;
 *                             gK_host [18 C_in, C_out] and gbias_host [C_out] receive the sums over the batch.
 * Ragged batches (n_dev != NULL): an
 * instance's contraction output is compact ([n, n, 18 C_in] at the start of its n_max^2 rows), so zero X_dev once before
 * the first call, ignore the rows of Y / Z past n^2 of each instance, and pass zeros in those rows of gZ_dev. */
CCN_API int ccn_gather_contract18_forward(ccn_ctx *ctx, const float *f_dev, const int64_t *f_off_dev, const int32_t *m_dev,
                                  const int32_t *pos_dev, const float *adj_dev, float *T_scratch_dev, float *out_dev,
                                  const int32_t *n_dev, int n_max, int C, int64_t batch, int64_t stride_adj,
                                  int64_t stride_out, int adj_mode, void *stream);
CCN_API int ccn_gather_contract18_backward(ccn_ctx *ctx, const float *gout_dev, const float *adj_dev, const int64_t *f_off_dev,
                                   const int32_t *m_dev, const int32_t *pos_dev, float *gT_scratch_dev, float *gf_dev,
                                   const int32_t *n_dev, int n_max, int C, int64_t batch, int64_t stride_gout,
                                   int64_t stride_adj, int adj_mode, void *stream);
CCN_API int ccn_level_forward(ccn_ctx *ctx, const float *T_dev, const float *const *slabs_dev, const float *adj_dev,
                      const float *K_dev, const float *bias_dev, float *X_dev, float *Y_dev, float *Z_dev,
                      const int32_t *n_dev, int n_max, int C_in, int C_out, int64_t batch, int64_t stride_T,
                      int64_t stride_adj, int adj_mode, float lrelu_alpha, void *stream);
CCN_API int ccn_level_backward(ccn_ctx *ctx, const float *gZ_dev, const float *X_dev, const float *Y_dev, const float *K_dev,
                       const float *bias_dev, const float *adj_dev, float *gX_scratch_dev, float *gT_dev,
                       float *const *gslabs_dev, float *gK_dev, float *gbias_dev, const int32_t *n_dev, int n_max, int C_in,
                       int C_out, int64_t batch, int64_t stride_adj, int64_t stride_gT, int adj_mode, float lrelu_alpha,
                       float beta, void *stream);

CCN_API int ccn_gather_level_forward(ccn_ctx *ctx, const float *f_dev, const int64_t *f_off_dev, const int32_t *m_dev,
                             const int32_t *pos_dev, const float *adj_dev, const float *K_dev, const float *bias_dev,
                             float *T_scratch_dev, float *X_dev, float *Y_dev, float *Z_dev, const int32_t *n_dev, int n_max,
                             int C_in, int C_out, int64_t batch, int64_t stride_adj, int adj_mode, float lrelu_alpha, void *stream);
CCN_API int ccn_gather_level_backward(ccn_ctx *ctx, const float *gZ_dev, const float *X_dev, const float *Y_dev, const float *K_dev,
                              const float *bias_dev, const float *adj_dev, const int64_t *f_off_dev, const int32_t *m_dev,
                              const int32_t *pos_dev, float *gX_scratch_dev, float *gT_scratch_dev, float *gf_dev, float *gK_dev,
                              float *gbias_dev, const int32_t *n_dev, int n_max, int C_in, int C_out, int64_t batch,
                              int64_t stride_adj, int adj_mode, float lrelu_alpha, void *stream);
/* The same for a STACK of `levels` levels that stays on the device between the levels (C_in == C_out): per-level arrays of
 * tables and parameters (arrays of `levels` pointers).  Level 1's f_off entries are absolute element offsets into f_host; level
 * l > 1's are absolute offsets into level l-1's output array [batch, n^2, C_out] (instance i's tensor at i * n^2 * C_out), as a
 * model's next level addresses them.  Only f_host / gZ_host (gradient w.r.t. the LAST level's output) go up and Z_host (the last
 * level's output) / gf_host come back; every X, Y and intermediate activation lives and dies on the device. */
CCN_API int ccn_gather_levels_forward_backward_host(ccn_ctx *ctx, int levels, const float *f_host, const int64_t *f_group_ptr,
                                            const int64_t *inst_group_ptr, int64_t groups, const int64_t *const *f_off_host,
                                            const int32_t *const *m_host, const int32_t *const *pos_host,
                                            const float *const *adj_host, const float *const *K_host,
                                            const float *const *bias_host, const float *gZ_host, float *Z_host, float *gf_host,
                                            float *const *gK_host, float *const *gbias_host, int n, int C_in, int C_out,
                                            int adj_mode, float lrelu_alpha);
/* The same stack followed by the models' read-out head and loss ON THE DEVICE (ccn_readout_*: ShrinkTensor -> LeakyReLU ->
 * SumVectors per group (= graph) -> InnerProduct(W) -> SquaredLoss(target), SMP_beta.h:620-639): the last level's output and its
 * gradient never cross PCIe either.  Up: f_host, the tables, the parameters, target_host [groups].  Down: predict_host and
 * loss_host [groups], gf_host, and the parameter gradients gK_host[l], gbias_host[l], gW_host [C] summed over the batch -- a
 * training step of the model minus level 0.  C_in == C_out == C. */
CCN_API int ccn_gather_levels_readout_forward_backward_host(ccn_ctx *ctx, int levels, const float *f_host, const int64_t *f_group_ptr,
                                                    const int64_t *inst_group_ptr, int64_t groups,
                                                    const int64_t *const *f_off_host, const int32_t *const *m_host,
                                                    const int32_t *const *pos_host, const float *const *adj_host,
                                                    const float *const *K_host, const float *const *bias_host,
                                                    const float *W_host, const float *target_host, float *predict_host,
                                                    float *loss_host, float *gf_host, float *const *gK_host,
                                                    float *const *gbias_host, float *gW_host, int n, int C, int adj_mode,
                                                    float lrelu_alpha);
CCN_API int ccn_gather_level_forward_backward_host(ccn_ctx *ctx, const float *f_host, const int64_t *f_group_ptr,
                                           const int64_t *inst_group_ptr, int64_t groups, const int64_t *f_off_host,
                                           const int32_t *m_host, const int32_t *pos_host, const float *adj_host,
                                           const float *K_host, const float *bias_host, const float *gZ_host, float *Z_host,
                                           float *gf_host, float *gK_host, float *gbias_host, int n, int C_in, int C_out,
                                           int adj_mode, float lrelu_alpha);

/* ---- read-out head and loss on the device -----------------------------------------------------------------------------
 * Replaces, for a batch of graphs, the tail of SMP_beta::complete_computation_graph (SMP_beta.h:620-639): ShrinkTensor
 * (ShrinkTensor.h:37-51: per-channel sum over the n x n cells of f_L[v]) -> LeakyReLU (alpha) -> SumVectors over the graph's
 * vertices -> InnerProduct with W -> SquaredLoss = 0.5 (predict - target)^2 (SquaredLoss.h:46-54), and their backward passes.
 * Z_dev: the last level's activations, instance i at Z_dev + i*stride_Z as n_i^2 compact rows of C floats; instances of graph
 * g are [inst_graph_ptr[g], inst_graph_ptr[g+1]); inst_graph_dev[i] = graph of instance i.  Outputs: shrinked [batch, C]
 * (kept for the backward), graph_feature [graphs, C], predict [graphs], loss [graphs] (loss_dev / target_dev may be NULL for
 * inference).  Backward: gZ_dev (same layout as Z, all n_max^2 rows of every instance written, zeros in the padding) and
 * gW_dev [C] += (predict - target) graph_feature (may be NULL). */
CCN_API int ccn_readout_forward(ccn_ctx *ctx, const float *Z_dev, int64_t stride_Z, const int32_t *n_dev, int n_max, int C,
                        int64_t batch, const int64_t *inst_graph_ptr_dev, int64_t graphs, const float *W_dev,
                        const float *target_dev, float lrelu_alpha, float *shrinked_dev, float *graph_feature_dev,
                        float *predict_dev, float *loss_dev, void *stream);
CCN_API int ccn_readout_backward(ccn_ctx *ctx, const float *shrinked_dev, const float *graph_feature_dev, const float *predict_dev,
                         const float *target_dev, const float *W_dev, const int32_t *inst_graph_dev, const int32_t *n_dev,
                         int n_max, int C, int64_t batch, int64_t graphs, float lrelu_alpha, float *gZ_dev, int64_t stride_gZ,
                         float *gW_dev, void *stream);

/* Level features of the multi-level read-outs (SMP_omega_physics.h:560-583, SMP_omega_pairgraphs.h:637-655): ShrinkTensor ->
 * LeakyReLU -> SumVectors of ONE level for a batch of graphs, written into columns [0, C) of the row-major [graphs, ld_feature]
 * matrix at feature_dev (offset the pointer to the level's columns of the concatenated graph feature).  shrinked_dev [batch, C]
 * is kept for the backward.  Instances (vertices) of a graph are consecutive: inst_graph_ptr_dev [graphs + 1]. */
CCN_API int ccn_level_features_forward(ccn_ctx *ctx, const float *Z_dev, int64_t stride_Z, const int32_t *n_dev, int n_max, int C,
                                       int64_t batch, const int64_t *inst_graph_ptr_dev, int64_t graphs, float lrelu_alpha,
                                       float *shrinked_dev, float *feature_dev, int64_t ld_feature, void *stream);
/* The transpose: gZ[i][r][c] = dfeature[graph(i)][c] * lrelu'(shrinked[i][c]) for the n_i^2 real rows r of instance i, zero in
 * the padding rows up to n_max^2 (WRITTEN, not added: call it before the next level's backward adds into the same array). */
CCN_API int ccn_level_features_backward(ccn_ctx *ctx, const float *shrinked_dev, const float *dfeature_dev, int64_t ld_feature,
                                        const int32_t *inst_graph_dev, const int32_t *n_dev, int n_max, int C, int64_t batch,
                                        float lrelu_alpha, float *gZ_dev, int64_t stride_gZ, void *stream);

/* ---- TensorMul ------------------------------------------------------------------------------------------------------
 * Replaces TensorMul::forward / backward (TensorMul.h:48-86): out[i,j,d] = sum_k A[i,k,d] B[k,j,d] per channel d, for
 * `batch` dense instances (A [R,K,D], B [K,Cc,D], out [R,Cc,D]).  backward: gA = beta gA + g . B^T, gB = beta gB + A^T . g
 * (beta = 1 is the reference's +=); either gradient pointer may be NULL. */
CCN_API int ccn_tensor_mul_forward(ccn_ctx *ctx, const float *A_dev, const float *B_dev, float *out_dev, int R, int K, int Cc, int D,
                           int64_t batch, void *stream);
CCN_API int ccn_tensor_mul_backward(ccn_ctx *ctx, const float *A_dev, const float *B_dev, const float *gout_dev, float *gA_dev,
                            float *gB_dev, int R, int K, int Cc, int D, int64_t batch, float beta, void *stream);

/* ---- CustomMatMulTensor ---------------------------------------------------------------------------------------------
 * Replaces CustomMatMulTensor::forward / backward (CustomMatMulTensor.h:47-85; SMP_2D_ver8.h:526-527): the feature mix
 * with the weights stored transposed, Y[r,k] = sum_v Kt[k,v] X[r,v] (rows r = (i,j) flattened; Kt is [P, V]).
 * backward: gX = beta_x gX + gY Kt ; gKt += gY^T X.  Runs on the same (tensor-core) kernels as ccn_mix_*. */
CCN_API int ccn_custom_matmul_tensor_forward(ccn_ctx *ctx, const float *Kt_dev, const float *X_dev, float *Y_dev, int64_t M, int V,
                                     int P, void *stream);
CCN_API int ccn_custom_matmul_tensor_backward(ccn_ctx *ctx, const float *Kt_dev, const float *X_dev, const float *gY_dev,
                                      float *gKt_dev, float *gX_dev, int64_t M, int V, int P, float beta_x, void *stream);

/* ---- parameter update on the device ----------------------------------------------------------------------------------
 * The reference's optimizers on the flat parameter vector (registration order H, K_l, b_l ..., W; SMP_beta.h:274-280).
 * ccn_adam_step = Adam::Learn (Adam.h:76-137): g = grad / n_batch; m, v updated; params -= alpha m^ / (sqrt(v^) + eps).
 *   per_element_bias != 0 reproduces the `Learn(alpha, nBatch)` overload, whose bias-correction powers advance once per
 *   ELEMENT (:123,127): element i uses beta^(updates_before + i + 1); the caller adds `count` to updates_before after
 *   every call.  per_element_bias == 0 is `Learn(alpha)` (:82-83): beta^(updates_before + 1), updates_before = number
 *   of earlier calls (pass n_batch = 1).
 * ccn_momentum_step = Momentum::Learn (Momentum.h:51-67): moments = gamma moments + lr grad / n_batch; params -= moments;
 *   gamma = 0 gives SGD::Learn (SGD.h:36-50). */
CCN_API int ccn_adam_step(ccn_ctx *ctx, float *params_dev, const float *grads_dev, float *m_dev, float *v_dev, int64_t count,
                  double alpha, double beta1, double beta2, double epsilon, int n_batch, int64_t updates_before,
                  int per_element_bias, void *stream);
CCN_API int ccn_momentum_step(ccn_ctx *ctx, float *params_dev, const float *grads_dev, float *moments_dev, int64_t count,
                      double learning_rate, double gamma, int n_batch, void *stream);

/* ---- gradient all-reduce -------------------------------------------------------------------------------------------
 * The sum of the parameter gradients over the data-parallel replicas (add_gradient over the threads' instances,
 * SMP_beta.h:677-687, 731-733) as ONE in-place ncclAllReduce(sum, float) over a flat buffer on `stream`.  `nccl_comm` is the
 * caller's ncclComm_t.  NCCL is bound at call time (the symbols already in the process -- e.g. the NCCL a framework
 * loaded -- else libnccl.so.2), so the library itself has no NCCL link dependency; CCN_ERR_UNSUPPORTED when no NCCL is
 * found.  (The Python binding uses torch.distributed instead: graphflow_b200/shard.py.) */
CCN_API int ccn_allreduce_grads(ccn_ctx *ctx, void *nccl_comm, float *buf_dev, int64_t count, void *stream);

/* ---- host-side graph preprocessing -> index tables ---------------------------------------------------------------------
 * What SMP_beta::complete_computation_graph computes from a DenseGraph before it wires operators (SMP_beta.h:531-552:
 * floyd_warshall :343-365, weisfeiler_lehman :367-389, rank_vertices :403-419, the receptive fields :461-489 and reduced
 * adjacency :505-526; kind CCN_GRAPH_OMEGA: the insertion-ordered, max_field-limited fields of
 * SMP_omega_physics.h:367-418 with the raw vertex features), as flat tables for ccn_promote_* / ccn_contract18_*.  Pure
 * host code (no device needed).  adj: [V, V] (> 0 = edge), feat: [V, F].  Levels are 1..n_levels; level 0 fields are {v}.
 *   features   [V, width], width = F (n_depth + 1) for CCN_GRAPH_BETA (the level-0 input features), F for CCN_GRAPH_OMEGA
 *   field      phi_level(v): returns its size n and the member list
 *   vertex     for phi_level(v): n, the reduced adjacency [n, n] (1 on the diagonal), and for every slab a the source
 *              vertex src[a] = phi_level(v)[a], m[a] = |phi_{level-1}(src[a])| and pos[a][i] = position of phi_level(v)[i]
 *              inside phi_{level-1}(src[a]) or -1 (the promotion gather table)
 * Pointers stay valid until ccn_graph_tables_destroy. */
#define CCN_GRAPH_BETA 0
#define CCN_GRAPH_OMEGA 1    /* SMP_omega_physics: raw features, insertion-ordered fields limited to max_field */
#define CCN_GRAPH_OMEGA_WL 2 /* SMP_omega (SMP_omega.h:476-531): SMP_beta's WL features and ranking; a field larger than max_field is
                                cut to its nearest members (by distance, ties by rank, whole outermost shells dropped) and then
                                ordered by rank */
typedef struct ccn_graph_tables ccn_graph_tables;
CCN_API int ccn_graph_tables_create(const int32_t *adj, const double *feat, int V, int F, int n_levels, int n_depth, int kind,
                            int max_field, ccn_graph_tables **out);
CCN_API void ccn_graph_tables_destroy(ccn_graph_tables *g);
CCN_API int ccn_graph_tables_feature_width(const ccn_graph_tables *g);
CCN_API const double *ccn_graph_tables_features(const ccn_graph_tables *g);
CCN_API const int32_t *ccn_graph_tables_rank(const ccn_graph_tables *g);
CCN_API int ccn_graph_tables_field(const ccn_graph_tables *g, int level, int v, const int32_t **members);
CCN_API int ccn_graph_tables_vertex(const ccn_graph_tables *g, int level, int v, const float **adj_red, const int32_t **src,
                            const int32_t **m, const int32_t **pos);

/* ---- small helpers for host-side callers (the C++ facade's lazily synchronised mirrors) ------------------------- */
CCN_API int ccn_device_alloc(ccn_ctx *ctx, void **ptr_dev, size_t bytes);
CCN_API int ccn_device_free(ccn_ctx *ctx, void *ptr_dev);
CCN_API int ccn_h2d(ccn_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes, void *stream);
CCN_API int ccn_d2h(ccn_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes, void *stream);
CCN_API int ccn_memset_zero(ccn_ctx *ctx, void *dst_dev, size_t bytes, void *stream);
CCN_API int ccn_stream_synchronize(ccn_ctx *ctx, void *stream);
/* A non-blocking stream on the context's device for callers without the CUDA runtime headers (one stream per host
 * thread is the reference's multi-stream replica scheme, GraphFlow_gpu/SMP_beta_gpu_multistreams.h:701-718). */
CCN_API int ccn_stream_create(ccn_ctx *ctx, void **stream);
/* Page-locks / releases a host array the caller already owns (cudaHostRegister), so that the *_host entry points and
 * ccn_h2d / ccn_d2h copy it at full PCIe speed and asynchronously; the reference's operators allocate value[] / gradient[]
 * with plain new[] (Vector.h:24-25), which is pageable.  Register once per array, not per call. */
CCN_API int ccn_host_register(ccn_ctx *ctx, void *ptr_host, size_t bytes);
CCN_API int ccn_host_unregister(ccn_ctx *ctx, void *ptr_host);
CCN_API int ccn_stream_destroy(ccn_ctx *ctx, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CCN_B200_H_INCLUDED */
