// mix_kernels.cuh -- internal launch interface of the feature-mix GEMM kernels (see include/ccn_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ccn_common.cuh"

namespace ccn {

cudaError_t mix_configure();

// Y[M,P] = X[M,K] W[K,P];  Z = lrelu(Y + bias) when bias && Z.  Y may be null.
cudaError_t launch_mix_forward(const float *X, const float *W, const float *bias, float *Y, float *Z, int64_t M, int K,
                               int P, float alpha, cudaStream_t st, LaunchLog *log);

// gY = gZ * lrelu'(Y + bias) (or gZ when bias == null); gX = beta_x gX + gY W^T; gW += X^T gY; gbias += colsum(gY).
cudaError_t launch_mix_backward(const float *X, const float *W, const float *bias, const float *Y, const float *gZ,
                                float *gX, float *gW, float *gbias, int64_t M, int K, int P, float alpha, float beta_x,
                                cudaStream_t st, LaunchLog *log);

// tensor-core (tcgen05, 3xTF32) forward: K % 4 == 0, P % 4 == 0, 4 <= P <= 128 (N is padded to a multiple of 16 inside
// the MMA), 16-byte aligned X / Y / Z.
// `wprep` is a device buffer of mix_tc_wprep_bytes(K, P) bytes owned by the context.
bool mix_tc_supported(const float *X, const float *Y, const float *Z, int64_t M, int K, int P);
size_t mix_tc_wprep_bytes(int K, int P);
cudaError_t mix_tc_configure();
cudaError_t launch_mix_forward_tc(const float *X, const float *W, const float *bias, float *Y, float *Z, int64_t M, int K,
                                  int P, float alpha, float *wprep, int sm_count, int tiles_per_pass, cudaStream_t st,
                                  LaunchLog *log, const int32_t *rows_n = nullptr, int64_t rows_per_inst = 0, int *item_buf = nullptr);
// rows_n / rows_per_inst / item_buf (mix_item_list_bytes(M) bytes): the rows are blocks of rows_per_inst per instance with
// rows_n[i]^2 real rows each; work items without a real row are skipped (their output rows are NOT written)
size_t mix_item_list_bytes(int64_t M);

// tensor-core grad-X (P % 4 == 0, P <= 64); optionally also writes gY = gZ * lrelu'(Y + bias) to gY_out [M, P].
bool mix_gx_tc_supported(const float *gZ, const float *Y, const float *gX, const float *gYs, int64_t M, int K, int P);
size_t mix_gx_tc_wprep_bytes(int K, int P);
cudaError_t mix_gx_tc_configure();
cudaError_t launch_mix_grad_x_tc(const float *W, const float *bias, const float *Y, const float *gZ, float *gX, float *gY_out,
                                 int64_t M, int K, int P, float alpha, float beta_x, float *wtprep, int sm_count, cudaStream_t st,
                                 LaunchLog *log, const int32_t *rows_n = nullptr, int64_t rows_per_inst = 0, int *item_buf = nullptr);
// tensor-core grad-W (+ grad-bias) from the activation-corrected gradient gY [M, P] (P % 4 == 0, P <= 64, M >= 4096).
bool mix_gw_tc_supported(const float *X, const float *gY, const float *gW, int64_t M, int K, int P);
cudaError_t mix_gw_tc_configure();
cudaError_t launch_mix_grad_w_tc(const float *X, const float *gY, float *gW, float *gbias, int64_t M, int K, int P, int sm_count,
                                 cudaStream_t st, LaunchLog *log);
// the SIMT backward, one product at a time (what == 1: grad-X, 2: grad-W, 4: grad-bias; or-able)
cudaError_t launch_mix_backward_parts(int what, const float *X, const float *W, const float *bias, const float *Y, const float *gZ,
                                      float *gX, float *gW, float *gbias, int64_t M, int K, int P, float alpha, float beta_x,
                                      cudaStream_t st, LaunchLog *log);

}  // namespace ccn
