// Launch helpers for the tcgen05 GEMM kernels (shared by model.cu and backward.cu).
#pragma once
#include "gemm.cuh"
#include "gemm2.cuh"
#include "gemm_tn.cuh"

template <int BN, int AM, int EPI, int ST, int KIND = 0>
static int launch_gemm(const CUtensorMap& a, const CUtensorMap& b, const gemm::Params& p, int num_sms, cudaStream_t st) {
  auto kern = gemm::gemm_kernel<BN, AM, EPI, ST, KIND>;
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

// 2-CTA pairs: B map must have box rows = 128 (each CTA stages half of the 256-row N tile)
template <int AM, int EPI, int ST, int BN = 256>
static int launch_gemm2(const CUtensorMap& a, const CUtensorMap& b_half, const gemm::Params& p, int num_sms, cudaStream_t st) {
  auto kern = gemm::gemm2_kernel<BN, AM, EPI, ST>;
  constexpr int smem = gemm::Smem2<BN, ST>::BYTES;
  static bool attr = false;
  if (!attr) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  const int pairs = ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
  int clusters = num_sms / 2;
  if (pairs < clusters) clusters = pairs;
  kern<<<2 * clusters, gemm::NUM_THREADS, smem, st>>>(a, b_half, p);
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

// 2-CTA weight-gradient GEMM: p.num_m_tiles must count 256-row PAIRS and p.num_n_tiles 256-column tiles.
template <int AM, int ST>
static int launch_gemm_tn2(const CUtensorMap& a, const CUtensorMap& b, gemm_tn::Params p, int num_sms, cudaStream_t st) {
  auto kern = gemm_tn::gemm_tn2_kernel<AM, ST>;
  constexpr int smem = gemm_tn::Smem2<ST>::BYTES;
  static bool attr = false;
  if (!attr) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  const int tiles = p.num_taps * p.num_m_tiles * p.num_n_tiles;
  const int max_clusters = num_sms / 2;
  if (p.k_splits <= 0) {
    int s = (3 * max_clusters + tiles - 1) / tiles;
    if (s > p.k_blocks_total) s = p.k_blocks_total;
    if (s < 1) s = 1;
    p.k_splits = s;
  }
  const int items = tiles * p.k_splits;
  const int clusters = items < max_clusters ? items : max_clusters;
  kern<<<2 * clusters, gemm_tn::NUM_THREADS, smem, st>>>(a, b, p);
  CUDA_TRY(cudaGetLastError());
  return CRNN_OK;
}
