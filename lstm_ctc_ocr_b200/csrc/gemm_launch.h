// Launch helpers for the tcgen05 GEMM kernels (shared by model.cu and backward.cu).
#pragma once
#include "gemm.cuh"
#include "gemm_tn.cuh"

template <int BN, int AM, int EPI, int ST>
static int launch_gemm(const CUtensorMap& a, const CUtensorMap& b, const gemm::Params& p, int num_sms, cudaStream_t st) {
  auto kern = gemm::gemm_kernel<BN, AM, EPI, ST>;
  constexpr int smem = gemm::Smem<BN, ST>::BYTES;
  static bool attr = false;
  if (!attr) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < num_sms ? tiles : num_sms;
  kern<<<grid, gemm::NUM_THREADS, smem, st>>>(a, b, p);
  CUDA_TRY(cudaGetLastError());
  return CRNN_OK;
}

template <int BN, int AM, int ST>
static int launch_gemm_tn(const CUtensorMap& a, const CUtensorMap& b, gemm_tn::Params p, int num_sms, cudaStream_t st) {
  auto kern = gemm_tn::gemm_tn_kernel<BN, AM, ST>;
  constexpr int smem = gemm_tn::Smem<BN, ST>::BYTES;
  static bool attr = false;
  if (!attr) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  const int tiles = p.num_taps * p.num_m_tiles * p.num_n_tiles;
  if (p.k_splits <= 0) {
    int s = (3 * num_sms + tiles - 1) / tiles;
    if (s > p.k_blocks_total) s = p.k_blocks_total;
    if (s < 1) s = 1;
    p.k_splits = s;
  }
  const int items = tiles * p.k_splits;
  const int grid = items < num_sms ? items : num_sms;
  kern<<<grid, gemm_tn::NUM_THREADS, smem, st>>>(a, b, p);
  CUDA_TRY(cudaGetLastError());
  return CRNN_OK;
}
