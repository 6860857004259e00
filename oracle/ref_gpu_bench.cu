// ref_gpu_bench.cu -- TEST / MEASUREMENT INFRASTRUCTURE, not product code.
//
// Times the reference's OWN CUDA implementation of the hot path on the same B200 (SURVEY.md section 8d(iii),
// BASELINE.md section 4.4): the UNMODIFIED GraphFlow_gpu_32bit/RisiContraction_18_gpu.h (kernels
// RisiContraction_18_forward_job :29, RisiContraction_18_backward_job :521; forward_GPU / backward_GPU with their
// pageable-memory cudaMemcpy calls) and GraphFlow_gpu_32bit/MatMul_gpu.h (Matrix_Multiplication_GPU :28,
// MatMul_backward_first :71, MatMul_backward_second :94), included from the reference tree where it lies and
// recompiled with `nvcc -arch=sm_100a` by oracle/Makefile (target refgpu -> oracle/_ref/ref_gpu_bench).  Nothing is copied.
//
//   ref_gpu_bench [N] [C] [reps]     prints one JSON object:
//     contract18: op_ms_fwd / op_ms_bwd      forward_GPU() / backward_GPU() as a model calls them (H2D + kernel + D2H,
//                                            host arrays are the reference's pageable new[] buffers)
//                 kernel_ms_fwd / _bwd       the kernels alone (CUDA events, data resident), same grid as forward_GPU
//     matmul:     the same four figures for MatMul_gpu at [N*N, 18C] x [18C, C]
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <chrono>

#include "Matrix.h"
#include "Tensor3D.h"
#include "Tensor4D.h"
#include "RisiContraction_18_gpu.h"
#include "MatMul_gpu.h"

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static float frand() { return (float)rand() / (float)RAND_MAX * 2.f - 1.f; }

int main(int argc, char **argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 32;
    const int C = argc > 2 ? atoi(argv[2]) : 64;
    const int reps = argc > 3 ? atoi(argv[3]) : 5;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        printf("{\"unavailable\": \"no CUDA device\"}\n");
        return 0;
    }
    srand(123456789);

    // ---- RisiContraction_18_gpu: inputs as in tests/test_RisiContraction_18_gpu.cu (a Tensor4D and a Matrix) ----
    Tensor4D *T = new Tensor4D(N, N, N, C);
    Matrix *adj = new Matrix(N, N);
    for (int i = 0; i < T->size; ++i) T->value[i] = frand();
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) adj->value[adj->index(i, j)] = 0.f;
    for (int i = 0; i < N; ++i) {  // a path plus the diagonal: ~3 non-zeros per row, like a molecular graph
        adj->value[adj->index(i, i)] = 1.f;
        if (i + 1 < N) adj->value[adj->index(i, i + 1)] = adj->value[adj->index(i + 1, i)] = 1.f;
    }
    RisiContraction_18_gpu *op = new RisiContraction_18_gpu(T, adj);
    for (int i = 0; i < op->size; ++i) op->gradient[i] = frand();
    for (int i = 0; i < T->size; ++i) T->gradient[i] = 0.f;

    op->forward_GPU();
    op->backward_GPU();
    cudaDeviceSynchronize();
    double t_f = 0, t_b = 0;
    for (int r = 0; r < reps; ++r) {
        double t0 = now_ms();
        op->forward_GPU();
        cudaDeviceSynchronize();
        double t1 = now_ms();
        for (int i = 0; i < op->size; ++i) op->gradient[i] = 1.f;  // forward_GPU zeroes the gradient; not timed
        double t2 = now_ms();
        op->backward_GPU();
        cudaDeviceSynchronize();
        double t3 = now_ms();
        t_f += t1 - t0;
        t_b += t3 - t2;
    }
    const double op_f = t_f / reps, op_b = t_b / reps;

    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int nThreads = RisiContraction_18_gpu::nThreads;
    dim3 gridF(op->rounded_division(op->size, nThreads)), gridB(op->rounded_division(T->size, nThreads)), block(nThreads);
    float ms = 0.f, k_f = 0.f, k_b = 0.f;
    for (int r = 0; r < reps + 1; ++r) {
        cudaMemset(op->device_value, 0, op->this_size);
        cudaEventRecord(e0);
        RisiContraction_18_forward_job<<<gridF, block>>>(op->device_tensor_value, op->device_adj_value, op->device_value, N, C);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (r) k_f += ms;
        cudaEventRecord(e0);
        RisiContraction_18_backward_job<<<gridB, block>>>(op->device_tensor_gradient, op->device_adj_value, op->device_gradient, N, C);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (r) k_b += ms;
    }
    k_f /= reps;
    k_b /= reps;
    const cudaError_t err1 = cudaGetLastError();

    // ---- MatMul_gpu at the feature-mix shape [N*N, 18C] x [18C, C] (SMP_beta.h:604-605) ----
    const int M = N * N, K = 18 * C;
    Matrix *X = new Matrix(M, K), *W = new Matrix(K, C);
    for (int i = 0; i < X->size; ++i) X->value[i] = frand();
    for (int i = 0; i < W->size; ++i) W->value[i] = frand() * 0.1f;
    MatMul_gpu *mm = new MatMul_gpu(X, W);
    mm->forward_GPU();
    for (int i = 0; i < mm->size; ++i) mm->gradient[i] = 1.f;
    mm->backward_GPU();
    cudaDeviceSynchronize();
    double m_f = 0, m_b = 0;
    for (int r = 0; r < reps; ++r) {
        double t0 = now_ms();
        mm->forward_GPU();
        cudaDeviceSynchronize();
        double t1 = now_ms();
        for (int i = 0; i < mm->size; ++i) mm->gradient[i] = 1.f;
        double t2 = now_ms();
        mm->backward_GPU();
        cudaDeviceSynchronize();
        double t3 = now_ms();
        m_f += t1 - t0;
        m_b += t3 - t2;
    }
    m_f /= reps;
    m_b /= reps;
    float mk_f = 0.f, mk_b = 0.f;
    {
        float *dA, *dB, *dC, *dgA, *dgB, *dgC;
        cudaMalloc(&dA, sizeof(float) * M * K);
        cudaMalloc(&dB, sizeof(float) * K * C);
        cudaMalloc(&dC, sizeof(float) * M * C);
        cudaMalloc(&dgA, sizeof(float) * M * K);
        cudaMalloc(&dgB, sizeof(float) * K * C);
        cudaMalloc(&dgC, sizeof(float) * M * C);
        cudaMemcpy(dA, X->value, sizeof(float) * M * K, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, W->value, sizeof(float) * K * C, cudaMemcpyHostToDevice);
        cudaMemset(dgA, 0, sizeof(float) * M * K);
        cudaMemset(dgB, 0, sizeof(float) * K * C);
        cudaMemset(dgC, 0, sizeof(float) * M * C);
        const int BS = MATMUL_GPU_BLOCK_SIZE;
        dim3 blk(BS, BS), gC((M + BS - 1) / BS, (C + BS - 1) / BS), gA((M + BS - 1) / BS, (K + BS - 1) / BS),
            gB((K + BS - 1) / BS, (C + BS - 1) / BS);
        for (int r = 0; r < reps + 1; ++r) {
            cudaEventRecord(e0);
            Matrix_Multiplication_GPU<<<gC, blk>>>(dA, dB, dC, M, K, K, C);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            if (r) mk_f += ms;
            cudaEventRecord(e0);
            MatMul_backward_first<<<gA, blk>>>(dgA, dB, dgC, M, K, K, C);
            MatMul_backward_second<<<gB, blk>>>(dA, dgB, dgC, M, K, K, C);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            if (r) mk_b += ms;
        }
        mk_f /= reps;
        mk_b /= reps;
    }
    const cudaError_t err2 = cudaGetLastError();

    printf("{\"N\": %d, \"C\": %d, \"reps\": %d, \"source\": \"unmodified GraphFlow_gpu_32bit, nvcc -arch=sm_100a\", "
           "\"contract18\": {\"op_ms_fwd\": %.4f, \"op_ms_bwd\": %.4f, \"kernel_ms_fwd\": %.4f, \"kernel_ms_bwd\": %.4f, "
           "\"per_s_with_copies\": %.2f, \"per_s_kernels_only\": %.2f, \"cuda_error\": \"%s\"}, "
           "\"matmul\": {\"M\": %d, \"K\": %d, \"P\": %d, \"op_ms_fwd\": %.4f, \"op_ms_bwd\": %.4f, \"kernel_ms_fwd\": %.4f, "
           "\"kernel_ms_bwd\": %.4f, \"cuda_error\": \"%s\"}}\n",
           N, C, reps, op_f, op_b, k_f, k_b, 1e3 / (op_f + op_b), 1e3 / (k_f + k_b), cudaGetErrorString(err1), M, K, C, m_f, m_b,
           mk_f, mk_b, cudaGetErrorString(err2));
    return 0;
}
