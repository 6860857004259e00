// contract18_kernels.cuh -- internal launch interface between the C-ABI (ccn_abi.cu) and the kernel files.
#pragma once
#include "ccn_common.cuh"

namespace ccn {

// Where the neighbour tensors of a chunk live: either one stacked tensor per instance (base + inst*stride) or a
// table of per-vertex slab pointers (StackTensor3D fused away): entry inst*n_max + a -> [n, n, C].
struct TensorRef {
    float *base;
    float *const *slabs;
    int64_t stride;
};

struct Contract18Fwd {
    TensorRef T;  // read only
    float *out;
    int64_t stride_out;
    Batch b;
    const float *adjtab;  // per instance, stride adjtab_words
    int adjtab_words;
    float *scratch;  // per instance, stride scratch_words
    int64_t scratch_words;
};

struct Contract18Bwd {
    const float *gout;
    int64_t stride_gout;
    TensorRef gT;  // written (beta = 0) or accumulated (beta = 1)
    Batch b;
    const float *adjtab;
    int adjtab_words;
    float *scratch;
    int64_t scratch_words;
    float beta;
};

// adjacency tables (all paths)
cudaError_t launch_adj_prepare(const float *adj, int64_t stride_adj, Batch b, int adj_mode, float *adjtab,
                               cudaStream_t st, LaunchLog *log);

// generic path: any n, any C
int64_t generic_fwd_scratch_words(int n_max, int C);
int64_t generic_bwd_scratch_words(int n_max, int C);
cudaError_t launch_generic_forward(const Contract18Fwd &a, cudaStream_t st, LaunchLog *log);
cudaError_t launch_generic_backward(const Contract18Bwd &a, cudaStream_t st, LaunchLog *log);

// Promotion fused into the contraction (fusion step 2 of SURVEY 7.2): instead of a materialised stacked tensor the fused
// kernels read slab a of instance i straight out of the level l-1 buffer,
//     T[i][a][r][c][:] = f[f_off[i*n_max + a]][pos[r], pos[c], :]   (zero when pos[r] < 0 or pos[c] < 0),
// with pos = pos + (i*n_max + a)*n_max and the source tensor [m, m, C], m = m[i*n_max + a] (the arguments of
// ccn_promote_forward); the backward adds gT into gf at the same addresses with atomic reductions instead of writing it.
struct GatherRef {
    float *f;              // forward: f_{l-1} (read); backward: its gradient (red.add).  nullptr = no gather
    const int64_t *f_off;  // [batch * n_max]
    const int32_t *m;      // [batch * n_max]
    const int32_t *pos;    // [batch * n_max * n_max]
};

// fused path: n_max <= 32, C in {32, 64, 128}; one kernel per direction, the tiles of an instance cooperate through
// an L2-resident scratch slot (see contract18_fused.cu).  `ctl` is the control block (ticket + per-slot counters).
struct Fused18Fwd {
    TensorRef T;   // read only
    GatherRef G;   // when G.f != nullptr the input is gathered from f_{l-1} and T is ignored
    float *out;
    int64_t stride_out;
    const float *adj;  // raw adjacency, instance stride stride_adj
    int64_t stride_adj;
    int positive_part;
    Batch b;
    float *scratch;  // per slot, stride scratch_words
    int64_t scratch_words;
    int *ctl;
    int slots;
    unsigned long long *trace;  // optional: 8 globaltimer marks per tile (debug), else nullptr
    int *fault;                 // sticky failure word (device view of mapped host memory, owned by the context)
    int variant;                // A/B switches (bit mask), -1 = the defaults of contract18_fused_impl.cuh
    uint32_t keep;              // bit k set: slab k is kept (RisiContraction_18_dropout's use[k]); 0x3ffff = the plain operator
    float out_scale;            // multiplies the kept slabs (1, or nKept/18 in the dropout operator's test mode)
};

struct Fused18Bwd {
    const float *gout;
    int64_t stride_gout;
    TensorRef gT;  // written (beta = 0) or accumulated (beta != 0)
    GatherRef G;   // when G.f != nullptr the gradient is scattered (added) into gf_{l-1} = G.f and gT is ignored
    const float *adj;
    int64_t stride_adj;
    int positive_part;
    Batch b;
    float *scratch;
    int64_t scratch_words;
    int *ctl;
    int slots;
    float beta;
    unsigned long long *trace;
    int *fault;
    int variant;
    uint32_t keep;  // bit k clear: slab k of gout is ignored (never read)
};

bool fused_path_supported(int n_max, int C);
// tiles per instance and resident CTAs per SM differ by direction (256-thread tiles forward, 128-thread tiles backward)
int fused_tiles_fwd(int n_max, int C);
int fused_tiles_bwd(int n_max, int C);
int fused_resident_ctas_fwd();
int fused_resident_ctas_bwd();
// geometry: 0 = forward, 1 = backward, 2 = backward with fused promotion (runs on the forward's 256-thread tiles)
inline int fused_tiles(int n_max, int C, int geometry) { return geometry == 1 ? fused_tiles_bwd(n_max, C) : fused_tiles_fwd(n_max, C); }
int fused_ctl_words(int slots);
int64_t fused_fwd_scratch_words(int n_max, int C);
int64_t fused_bwd_scratch_words(int n_max, int C);
int64_t fused_bwd_scatter_scratch_words(int n_max, int C);
cudaError_t fused_path_configure_fwd();
cudaError_t fused_path_configure_bwd();
cudaError_t launch_fused_forward(const Fused18Fwd &a, cudaStream_t st, LaunchLog *log);
cudaError_t launch_fused_backward(const Fused18Bwd &a, cudaStream_t st, LaunchLog *log);          // a.G.f == nullptr
cudaError_t launch_fused_backward_scatter(const Fused18Bwd &a, cudaStream_t st, LaunchLog *log);  // a.G.f != nullptr

// aux_ops.cu: promotion gather / scatter-add, TensorMul, small transposes.
struct PromoteArgs {
    float *f;              // forward: source f_{l-1} buffer (read); backward: its gradient (atomic +=)
    const int64_t *f_off;  // [batch * n_max] element offset of slab (inst, a)'s source tensor inside f
    const int32_t *m;      // [batch * n_max] side of that source tensor ([m, m, C])
    const int32_t *pos;    // [batch * n_max * n_max] row i of slab -> row of the source, or -1
    float *T;              // forward: stacked output; backward: its gradient (read)
    int64_t stride_T;
    const int32_t *n;      // per instance (or nullptr -> n_max)
    int n_max, C;
    bool v4_offsets_ok;    // every f_off is a multiple of 4 elements (the caller guarantees it): 16-byte vector path allowed
};
cudaError_t launch_promote(bool backward, const PromoteArgs &a, int batch, cudaStream_t st, LaunchLog *log);
cudaError_t launch_tensor_mul_forward(const float *A, const float *B, float *out, int R, int K, int Cc, int D, int batch,
                                      cudaStream_t st, LaunchLog *log);
cudaError_t launch_tensor_mul_backward(const float *A, const float *B, const float *g, float *gA, float *gB, int R, int K, int Cc,
                                       int D, int batch, float beta, cudaStream_t st, LaunchLog *log);
cudaError_t launch_transpose_add(const float *src, float *dst, int rows, int cols, float beta, cudaStream_t st, LaunchLog *log);
// slab dropout around the 18-way kernels (K_TRANSPOSE in the launch log: small elementwise passes)
cudaError_t launch_slab_mask_inplace(float *out, int64_t stride, Batch b, uint32_t keep, float scale, int S, cudaStream_t st,
                                     LaunchLog *log);
cudaError_t launch_slab_mask_copy(const float *g, int64_t stride_g, float *dst, int64_t stride_d, Batch b, uint32_t keep, int S,
                                  cudaStream_t st, LaunchLog *log);
// the reference's parameter updates (Adam.h:76-137, Momentum.h:51-67) on the flat parameter vector
cudaError_t launch_adam_step(float *p, const float *g, float *m, float *v, int64_t n, double alpha, double beta1, double beta2,
                             double eps, double inv_batch, int64_t updates_before, bool per_element, cudaStream_t st, LaunchLog *log);
cudaError_t launch_momentum_step(float *p, const float *g, float *mom, int64_t n, double lr, double gamma, double inv_batch,
                                 cudaStream_t st, LaunchLog *log);

// readout.cu: ShrinkTensor -> LeakyReLU -> SumVectors -> InnerProduct -> SquaredLoss for a batch of graphs, and back
cudaError_t launch_readout_forward(const float *Z, int64_t stride, const int32_t *n_dev, int n_max, int C, int64_t batch,
                                   const int64_t *inst_ptr, int64_t graphs, const float *W, const float *target, float alpha,
                                   float *shrinked, float *graph_feature, float *predict, float *loss, cudaStream_t st, LaunchLog *log);
cudaError_t launch_readout_backward(const float *shrinked, const float *graph_feature, const float *predict, const float *target,
                                    const float *W, const int32_t *inst_graph, const int32_t *n_dev, int n_max, int C, int64_t batch,
                                    int64_t graphs, float alpha, float *gZ, int64_t stride, float *gW, cudaStream_t st, LaunchLog *log);
cudaError_t launch_level_features_forward(const float *Z, int64_t stride, const int32_t *n_dev, int n_max, int C, int64_t batch,
                                          const int64_t *inst_ptr, int64_t graphs, float alpha, float *shrinked, float *feature, int64_t ld,
                                          cudaStream_t st, LaunchLog *log);
cudaError_t launch_level_features_backward(const float *shrinked, const float *dfeature, int64_t ld, const int32_t *inst_graph,
                                           const int32_t *n_dev, int n_max, int C, int64_t batch, float alpha, float *gZ, int64_t stride,
                                           cudaStream_t st, LaunchLog *log);

// RisiContraction_50 (contract50.cu): generic kernels, any n and C.  T is the input (forward) or the gT destination
// (backward); `out` is out (forward) or gout (backward).
cudaError_t r50_configure();
int r50_adj_words(int n_max);
int64_t r50_scratch_words(int n_max, int C);
// plan of a contraction family member over the 15 planes / 6 vectors / 5 scalars of contract50.cu
constexpr int kR50MaxCases = 50;
struct R50Case {
    unsigned char form, id, aux, flags;
};
struct R50Plan {
    R50Case c[kR50MaxCases];
    int ncases;
};
// variant: 50 (RisiContraction_50), 10 (RisiContraction_10 = its cases 1..10), 18 (the 18-way subset, for the slab-dropout
// operator), 4 (RisiContraction_4); bit k of keep_mask clear = slab k dropped.  Returns non-zero for an unknown variant.
int r50_make_plan(int variant, uint64_t keep_mask, R50Plan *plan);
cudaError_t launch_r50(bool backward, const R50Plan &plan, TensorRef T, float *out, int64_t stride_out, const float *adj, int64_t stride_adj,
                       Batch b, int adj_mode, float adj_scale, float *adjtab, float *scratch, float beta, cudaStream_t st,
                       LaunchLog *log);

}  // namespace ccn
