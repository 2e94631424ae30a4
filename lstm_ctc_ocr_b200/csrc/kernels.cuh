// Launchers for the SIMT / HBM-bound kernels (kernels.cu).
#pragma once
#include "common.cuh"

struct SumsqSegs {
  int n;
  long long off[8];
  long long cnt[8];
};

int launch_conv1_pool(const float* data, const float* w, const float* b, __nv_bfloat16* out, uint8_t* argmax, int N, int W,
                      int num_sms, cudaStream_t st);
int launch_bn_finalize(const double* stats, double count, const float* gamma, const float* beta, float eps, float* scale,
                       float* shift, float* save_mean, float* save_invstd, int C, cudaStream_t st);
int launch_bn_apply_relu(const __nv_bfloat16* in, __nv_bfloat16* out, const float* scale, const float* shift, size_t rows,
                         int C, cudaStream_t st);
int launch_bn_apply_relu_pool12(const __nv_bfloat16* in, __nv_bfloat16* out, const float* scale, const float* shift,
                                size_t out_positions, int C, cudaStream_t st);
int launch_transpose_cast(const float* src, int R, int Cc, int ld_src, __nv_bfloat16* dst, int ld_dst, int perm_mode,
                          cudaStream_t st);
int launch_lstm_bias_prep(const float* b_fw, const float* b_bw, float* xbias, int upc, cudaStream_t st);
int launch_sumsq(const float* params, const SumsqSegs& segs, double* out, cudaStream_t st);
int launch_total_loss(const float* costs, int N, const double* sumsq, float wd, float* loss, cudaStream_t st);
int launch_bf16_to_f32(const __nv_bfloat16* in, float* out, size_t n, cudaStream_t st);
