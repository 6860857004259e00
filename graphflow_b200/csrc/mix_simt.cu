// mix_simt.cu -- feature-mix GEMMs on the fp32 CUDA cores (exact-fp32 path).
//
// Replaces Reshape2D + MatMul::forward/backward (GraphFlow/MatMul.h:48-82), the reference kernels
// Matrix_Multiplication_GPU / MatMul_backward_first / MatMul_backward_second (GraphFlow_gpu/MatMul_gpu.h:28-111)
// and, fused into the epilogue / prologue, VectorAddTensor (VectorAddTensor.h:46-71) and LeakyReLU3D
// (LeakyReLU3D.h:60-82).
//
// One register-tiled kernel (128 x 64 x 16 tiles, 8 x 4 outputs per thread) instantiated for the three products:
//   forward   Y = X W              A(i,j) = X[i,j]        B(j,n) = W[j,n]      i<M, j<K, n<P
//   grad-X    gX = gY W^T          A(i,j) = gY[i,j]       B(j,n) = W[n,j]      i<M, j<P, n<K
//   grad-W    gW += X^T gY         A(i,j) = X[j,i]        B(j,n) = gY[j,n]     i<K, j<M (split over grid.z), n<P
// with gY = gZ * lrelu'(Y + bias) evaluated on the fly when it is loaded.
#include "mix_kernels.cuh"

namespace ccn {

namespace {

constexpr int BM = 128, BN = 64, BK = 16, kThreads = 256;

struct GradY {  // gY(m, p) = gZ[m,p] * (Y[m,p] + bias[p] > 0 ? 1 : alpha)
    const float *gZ, *Y, *bias;
    float alpha;
    int P;
    __device__ __forceinline__ float operator()(int64_t m, int p) const {
        const float g = gZ[m * P + p];
        if (!bias) return g;
        return (Y[m * P + p] + bias[p] > 0.f) ? g : g * alpha;
    }
};

enum Mode { kForward = 0, kGradX = 1, kGradW = 2 };

struct MixArgs {
    const float *X, *W, *bias;
    float *Y, *Z;       // forward outputs
    float *gX, *gW;     // backward outputs
    GradY gy;
    int64_t M;
    int K, P;
    float alpha, beta_x;
    int64_t split;  // grad-W: rows of M per grid.z slice
};

template <int MODE>
__global__ void __launch_bounds__(kThreads) k_mix_gemm(MixArgs a) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    // problem dims in GEMM terms
    const int64_t Mi = (MODE == kGradW) ? a.K : a.M;                                   // rows of C
    const int Nn = (MODE == kGradX) ? a.K : a.P;                                       // cols of C
    const int64_t Jtot = (MODE == kForward) ? a.K : (MODE == kGradX) ? a.P : a.M;      // reduction length
    const int64_t i0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    int64_t j_begin = 0, j_end = Jtot;
    if (MODE == kGradW) {
        j_begin = (int64_t)blockIdx.z * a.split;
        j_end = min(Jtot, j_begin + a.split);
    }
    float acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

    for (int64_t j0 = j_begin; j0 < j_end; j0 += BK) {
        // A tile: BM x BK
        for (int t = tid; t < BM * BK; t += kThreads) {
            int ii, jj;
            if (MODE == kGradW) {  // X[j, i]: i fastest in memory
                ii = t % BM;
                jj = t / BM;
            } else {  // row-major [i, j]: j fastest
                jj = t % BK;
                ii = t / BK;
            }
            const int64_t i = i0 + ii, j = j0 + jj;
            float v = 0.f;
            if (i < Mi && j < j_end) {
                if (MODE == kForward) v = a.X[i * a.K + j];
                else if (MODE == kGradX) v = a.gy(i, (int)j);
                else v = a.X[j * a.K + i];
            }
            As[jj][ii] = v;
        }
        // B tile: BK x BN
        for (int t = tid; t < BK * BN; t += kThreads) {
            int jj, nn;
            if (MODE == kGradX) {  // W[n, j]: j fastest
                jj = t % BK;
                nn = t / BK;
            } else {
                nn = t % BN;
                jj = t / BN;
            }
            const int64_t j = j0 + jj;
            const int n = n0 + nn;
            float v = 0.f;
            if (j < j_end && n < Nn) {
                if (MODE == kForward) v = a.W[j * a.P + n];
                else if (MODE == kGradX) v = a.W[(int64_t)n * a.P + j];
                else v = a.gy(j, n);
            }
            Bs[jj][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int jj = 0; jj < BK; ++jj) {
            float av[8], bv[4];
#pragma unroll
            for (int r = 0; r < 8; ++r) av[r] = As[jj][ty * 8 + r];
#pragma unroll
            for (int c = 0; c < 4; ++c) bv[c] = Bs[jj][tx * 4 + c];
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int64_t i = i0 + ty * 8 + r;
        if (i >= Mi) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int n = n0 + tx * 4 + c;
            if (n >= Nn) continue;
            const float v = acc[r][c];
            if (MODE == kForward) {
                if (a.Y) a.Y[i * a.P + n] = v;
                if (a.Z) {
                    const float s = v + a.bias[n];
                    a.Z[i * a.P + n] = s > 0.f ? s : a.alpha * s;
                }
            } else if (MODE == kGradX) {
                float *dst = a.gX + i * a.K + n;
                *dst = (a.beta_x != 0.f) ? fmaf(a.beta_x, *dst, v) : v;
            } else {
                atomicAdd(a.gW + i * a.P + n, v);
            }
        }
    }
}

// gbias[p] += sum_m gY(m, p): each CTA reduces a slice of rows, one atomic per column per CTA.
__global__ void __launch_bounds__(kThreads) k_mix_grad_bias(GradY gy, float *gbias, int64_t M, int P, int64_t rows_per_cta) {
    const int64_t m0 = (int64_t)blockIdx.x * rows_per_cta, m1 = min(M, m0 + rows_per_cta);
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        float s = 0.f;
        for (int64_t m = m0; m < m1; ++m) s += gy(m, p);
        atomicAdd(gbias + p, s);
    }
}

}  // namespace

cudaError_t mix_configure() { return cudaSuccess; }

cudaError_t launch_mix_forward(const float *X, const float *W, const float *bias, float *Y, float *Z, int64_t M, int K,
                               int P, float alpha, cudaStream_t st, LaunchLog *log) {
    MixArgs a{};
    a.X = X; a.W = W; a.bias = bias; a.Y = Y; a.Z = Z; a.M = M; a.K = K; a.P = P; a.alpha = alpha;
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((P + BN - 1) / BN));
    CCN_LAUNCH(log, K_MIX_FORWARD, st, k_mix_gemm<kForward><<<grid, kThreads, 0, st>>>(a));
    return cudaGetLastError();
}

cudaError_t launch_mix_backward(const float *X, const float *W, const float *bias, const float *Y, const float *gZ,
                                float *gX, float *gW, float *gbias, int64_t M, int K, int P, float alpha, float beta_x,
                                cudaStream_t st, LaunchLog *log) {
    return launch_mix_backward_parts(7, X, W, bias, Y, gZ, gX, gW, gbias, M, K, P, alpha, beta_x, st, log);
}

cudaError_t launch_mix_backward_parts(int what, const float *X, const float *W, const float *bias, const float *Y, const float *gZ,
                                      float *gX, float *gW, float *gbias, int64_t M, int K, int P, float alpha, float beta_x,
                                      cudaStream_t st, LaunchLog *log) {
    if (!(what & 1)) gX = nullptr;
    if (!(what & 2)) gW = nullptr;
    if (!(what & 4)) gbias = nullptr;
    MixArgs a{};
    a.X = X; a.W = W; a.bias = bias; a.gX = gX; a.gW = gW; a.M = M; a.K = K; a.P = P; a.alpha = alpha; a.beta_x = beta_x;
    a.gy = GradY{gZ, Y, bias, alpha, P};
    if (gX) {
        dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((K + BN - 1) / BN));
        CCN_LAUNCH(log, K_MIX_GRAD_X, st, k_mix_gemm<kGradX><<<grid, kThreads, 0, st>>>(a));
    }
    if (gW) {
        // split the reduction over M so that about 4 waves of CTAs are in flight
        const int64_t tiles = (int64_t)((K + BM - 1) / BM) * ((P + BN - 1) / BN);
        int64_t slices = (4 * 148 + tiles - 1) / tiles;
        int64_t split = (M + slices - 1) / slices;
        split = ((split + BK - 1) / BK) * BK;
        if (split < BK) split = BK;
        slices = (M + split - 1) / split;
        if (slices > 65535) {
            split = ((M + 65534) / 65535 + BK - 1) / BK * BK;
            slices = (M + split - 1) / split;
        }
        a.split = split;
        dim3 grid((unsigned)((K + BM - 1) / BM), (unsigned)((P + BN - 1) / BN), (unsigned)slices);
        CCN_LAUNCH(log, K_MIX_GRAD_W, st, k_mix_gemm<kGradW><<<grid, kThreads, 0, st>>>(a));
    }
    if (gbias) {
        const int64_t rows_per_cta = 512;
        CCN_LAUNCH(log, K_MIX_GRAD_BIAS, st,
                   k_mix_grad_bias<<<(unsigned)((M + rows_per_cta - 1) / rows_per_cta), kThreads, 0, st>>>(
                       a.gy, gbias, M, P, rows_per_cta));
    }
    return cudaGetLastError();
}

}  // namespace ccn
