// readout.cu -- the read-out head and the loss of the second-order CCN models on the device,
// sm_100a.  Replaces, for a whole batch of graphs, the tail of SMP_beta::complete_computation_graph (SMP_beta.h:620-639):
//
//   shrinked[v]     = ShrinkTensor(f_L[v])            sum over the n_v x n_v cells, per channel   (ShrinkTensor.h:37-51)
//   vertex_feature  = LeakyReLU(shrinked[v])          alpha = 0.01                                 (LeakyReLU.h:31,50-60)
//   graph_feature   = SumVectors over the vertices of the graph                                    (SumVectors.h)
//   predict         = InnerProduct(graph_feature, W)                                               (InnerProduct.h:40-47)
//   loss            = SquaredLoss(predict, target) = 0.5 (predict - target)^2                      (SquaredLoss.h:46-54)
//
// and their backward passes (SquaredLoss.h:56-62, InnerProduct.h:49-56, SumVectors / LeakyReLU / ShrinkTensor.h:53-63): the
// gradient of f_L[v] is the same C-vector in every cell.  All of it is O(sum n_v^2 C) streaming work: one pass over the last
// level's activations forward, one broadcast write backward.
#include "contract18_kernels.cuh"

namespace ccn {

namespace {

constexpr int kThreads = 256;

// grid = instances.  Z of instance i: n_i^2 rows of C floats at Z + i * stride (compact).  Thread (c, g) sums the rows
// g, g + G, ... of channel c (coalesced over c); the G partial sums meet in shared memory.
__global__ void __launch_bounds__(kThreads) k_readout_shrink(const float *__restrict__ Z, int64_t stride, const int32_t *__restrict__ n_dev,
                                                             int n_max, int C, float *__restrict__ shrinked) {
    extern __shared__ float part[];  // [G][C]
    const int inst = blockIdx.x;
    const int n = n_dev ? n_dev[inst] : n_max;
    const int rows = n * n;
    const float *z = Z + inst * stride;
    const int G = kThreads / C > 0 ? kThreads / C : 1;
    for (int c = threadIdx.x % C; c < C; c += kThreads) {  // C > kThreads: each thread takes several channels, G = 1
        const int g = threadIdx.x / C;
        float acc = 0.f;
        if (g < G)
            for (int r = g; r < rows; r += G) acc += z[(int64_t)r * C + c];
        if (g < G) part[g * C + c] = acc;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kThreads) {
        float s = 0.f;
        for (int g = 0; g < G; ++g) s += part[g * C + c];
        shrinked[(int64_t)inst * C + c] = s;
    }
}

// grid = graphs, block = 128.  graph_feature[g][c] = sum over the graph's instances of lrelu(shrinked), then predict and loss.
__global__ void __launch_bounds__(128) k_readout_graph(const float *__restrict__ shrinked, const int64_t *__restrict__ inst_ptr, int C,
                                                       const float *__restrict__ W, const float *__restrict__ target, float alpha,
                                                       float *__restrict__ graph_feature, int64_t ld, float *__restrict__ predict,
                                                       float *__restrict__ loss) {
    __shared__ float red[128];
    const int g = blockIdx.x;
    const int64_t i0 = inst_ptr[g], i1 = inst_ptr[g + 1];
    float dot = 0.f;
    for (int c = threadIdx.x; c < C; c += 128) {
        float s = 0.f;
        for (int64_t i = i0; i < i1; ++i) {
            const float x = shrinked[i * C + c];
            s += x > 0.f ? x : alpha * x;
        }
        graph_feature[(int64_t)g * ld + c] = s;
        if (W) dot = fmaf(s, W[c], dot);
    }
    if (!W) return;  // level features only (block-uniform)
    red[threadIdx.x] = dot;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float p = red[0];
        predict[g] = p;
        if (loss) {
            const float d = p - target[g];
            loss[g] = 0.5f * d * d;
        }
    }
}

// grid = graphs: gW += (predict - target) * graph_feature   (InnerProduct.h:52-54 with SquaredLoss.h:59)
__global__ void __launch_bounds__(128) k_readout_bwd_w(const float *__restrict__ graph_feature, const float *__restrict__ predict,
                                                       const float *__restrict__ target, int C, float *__restrict__ gW) {
    const int g = blockIdx.x;
    const float d = predict[g] - target[g];
    for (int c = threadIdx.x; c < C; c += 128) atomicAdd(gW + c, d * graph_feature[(int64_t)g * C + c]);
}

// grid = (instances, row chunks): gZ[i][r][c] = (predict - target)[graph(i)] * W[c] * lrelu'(shrinked[i][c]) for r < n_i^2,
// zero in the padding rows (the next level's backward reads whole n_max^2-row instances).
__global__ void __launch_bounds__(kThreads) k_readout_bwd_bcast(const float *__restrict__ shrinked, const int32_t *__restrict__ inst_graph,
                                                                const float *__restrict__ predict, const float *__restrict__ target,
                                                                const float *__restrict__ W, const float *__restrict__ dfeat, int64_t ld,
                                                                const int32_t *__restrict__ n_dev, int n_max,
                                                                int C, float alpha, float *__restrict__ gZ, int64_t stride) {
    const int inst = blockIdx.x;
    const int n = n_dev ? n_dev[inst] : n_max;
    const int g = inst_graph[inst];
    const float d = dfeat ? 0.f : predict[g] - target[g];
    const int64_t real = (int64_t)n * n * C, total = (int64_t)n_max * n_max * C;
    float *out = gZ + inst * stride;
    for (int64_t idx = (int64_t)blockIdx.y * kThreads + threadIdx.x; idx < total; idx += (int64_t)gridDim.y * kThreads) {
        const int c = (int)(idx % C);
        float v = 0.f;
        if (idx < real) {
            const float x = shrinked[(int64_t)inst * C + c];
            v = (dfeat ? dfeat[(int64_t)g * ld + c] : d * W[c]) * (x > 0.f ? 1.f : alpha);  // dfeat: the gradient of the level feature
        }
        out[idx] = v;
    }
}

}  // namespace

cudaError_t launch_readout_forward(const float *Z, int64_t stride, const int32_t *n_dev, int n_max, int C, int64_t batch,
                                   const int64_t *inst_ptr, int64_t graphs, const float *W, const float *target, float alpha,
                                   float *shrinked, float *graph_feature, float *predict, float *loss, cudaStream_t st, LaunchLog *log) {
    const int G = kThreads / C > 0 ? kThreads / C : 1;
    CCN_LAUNCH(log, K_READOUT, st,
               (k_readout_shrink<<<(unsigned)batch, kThreads, (size_t)G * C * sizeof(float), st>>>(Z, stride, n_dev, n_max, C, shrinked)));
    CCN_LAUNCH(log, K_READOUT, st,
               (k_readout_graph<<<(unsigned)graphs, 128, 0, st>>>(shrinked, inst_ptr, C, W, target, alpha, graph_feature, C, predict, loss)));
    return cudaGetLastError();
}

cudaError_t launch_readout_backward(const float *shrinked, const float *graph_feature, const float *predict, const float *target,
                                    const float *W, const int32_t *inst_graph, const int32_t *n_dev, int n_max, int C, int64_t batch,
                                    int64_t graphs, float alpha, float *gZ, int64_t stride, float *gW, cudaStream_t st, LaunchLog *log) {
    if (gW) CCN_LAUNCH(log, K_READOUT, st, (k_readout_bwd_w<<<(unsigned)graphs, 128, 0, st>>>(graph_feature, predict, target, C, gW)));
    const int64_t total = (int64_t)n_max * n_max * C;
    const unsigned chunks = (unsigned)std::min<int64_t>(32, (total + kThreads * 8 - 1) / (kThreads * 8));
    dim3 grid((unsigned)batch, chunks > 0 ? chunks : 1);
    CCN_LAUNCH(log, K_READOUT, st,
               (k_readout_bwd_bcast<<<grid, kThreads, 0, st>>>(shrinked, inst_graph, predict, target, W, nullptr, 0, n_dev, n_max, C, alpha, gZ, stride)));
    return cudaGetLastError();
}

// Level features of the multi-level read-outs (SMP_omega_physics.h:560-583, SMP_omega_pairgraphs.h:637-655): ShrinkTensor ->
// LeakyReLU -> SumVectors of ONE level, written into columns [0, C) of a row-major [graphs, ld] matrix (the caller offsets the
// pointer to the level's columns of the concatenated graph feature), and the transpose from the gradient of that matrix.
cudaError_t launch_level_features_forward(const float *Z, int64_t stride, const int32_t *n_dev, int n_max, int C, int64_t batch,
                                          const int64_t *inst_ptr, int64_t graphs, float alpha, float *shrinked, float *feature, int64_t ld,
                                          cudaStream_t st, LaunchLog *log) {
    const int G = kThreads / C > 0 ? kThreads / C : 1;
    CCN_LAUNCH(log, K_READOUT, st,
               (k_readout_shrink<<<(unsigned)batch, kThreads, (size_t)G * C * sizeof(float), st>>>(Z, stride, n_dev, n_max, C, shrinked)));
    CCN_LAUNCH(log, K_READOUT, st,
               (k_readout_graph<<<(unsigned)graphs, 128, 0, st>>>(shrinked, inst_ptr, C, nullptr, nullptr, alpha, feature, ld, nullptr, nullptr)));
    return cudaGetLastError();
}

cudaError_t launch_level_features_backward(const float *shrinked, const float *dfeature, int64_t ld, const int32_t *inst_graph,
                                           const int32_t *n_dev, int n_max, int C, int64_t batch, float alpha, float *gZ, int64_t stride,
                                           cudaStream_t st, LaunchLog *log) {
    const int64_t total = (int64_t)n_max * n_max * C;
    const unsigned chunks = (unsigned)std::min<int64_t>(32, (total + kThreads * 8 - 1) / (kThreads * 8));
    dim3 grid((unsigned)batch, chunks > 0 ? chunks : 1);
    CCN_LAUNCH(log, K_READOUT, st,
               (k_readout_bwd_bcast<<<grid, kThreads, 0, st>>>(shrinked, inst_graph, nullptr, nullptr, nullptr, dfeature, ld, n_dev, n_max, C,
                                                               alpha, gZ, stride)));
    return cudaGetLastError();
}

}  // namespace ccn
