// 2-CTA (cta_group::2) variant of the persistent tcgen05 GEMM: a cluster of two CTAs computes a 256 x 256 output tile.
//
// Why: the 1-CTA kernel is bound by the L2 -> shared-memory feed (measured 14.2 TB/s chip-wide = 48.8 B/clk/SM for
// BLOCK_N = 64/128/256 alike, tools/gemm_probe.py), not by the tensor pipe.  With cta_group::2 each CTA stages its own 128
// rows of A but only HALF of the B tile (128 of the 256 N rows); the paired MMA reads both halves from the two CTAs' shared
// memory.  Bytes per MMA cycle drop from 48 KB to 32 KB per K-block.
//
// Protocol (leader = cluster rank 0):
//   producers (one lane in each CTA)  wait own empty[s]  ->  TMA A (own 128 rows) + B (own 128 N-rows) into own smem,
//                                     completion bytes credited to the LEADER's full[s] (leader arms expect_tx for both)
//   MMA lane (leader only)            wait full[s] -> 4 x tcgen05.mma.cta_group::2 (M = 256) -> commit multicast to empty[s] of
//                                     both CTAs; after the last K-block commit multicast to tmem_full[acc] of both CTAs
//   epilogue warps (both CTAs)        wait own tmem_full[acc] -> run_epilogue on own TMEM (own 128 rows) -> arrive on the
//                                     LEADER's tmem_empty[acc] (count 8 = 4 warps x 2 CTAs)
#pragma once
#include "gemm.cuh"

namespace gemm {

template <int BLOCK_N, int STAGES>
struct Smem2 {
  static constexpr int A_STAGE = BLOCK_M * 128;       // 16 KB
  static constexpr int B_STAGE = (BLOCK_N / 2) * 128; // this CTA's half of the BLOCK_N rows of B
  static constexpr int BAR_OFFSET = STAGES * (A_STAGE + B_STAGE);
  static constexpr int BYTES = BAR_OFFSET + 256 + 1024;
};

template <int BLOCK_N, int AMODE, int EPI, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
  static_assert(BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N");
  constexpr int A_STAGE = Smem2<BLOCK_N, STAGES>::A_STAGE, B_STAGE = Smem2<BLOCK_N, STAGES>::B_STAGE;
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(2 * BLOCK_M, BLOCK_N);     // M = 256 across the CTA pair

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Smem2<BLOCK_N, STAGES>::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int num_pairs = ((p.num_m_tiles + 1) >> 1) * p.num_n_tiles;     // 256-row tile pairs x N tiles
  const int num_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tmem_full[s], 1); ptx::mbar_init(&tmem_empty[s], 2 * NUM_EPI_WARPS); }
    ptx::fence_barrier_init();
  }
  if (warp_idx == 1) { ptx::tmem_alloc_2cta(tmem_ptr, TMEM_COLS); ptx::tmem_relinquish_2cta(); }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();            // barrier inits + TMEM allocation visible in both CTAs before any cross-CTA signal
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp_idx == 0) {
    // ===================== TMA producer (both CTAs): one lane per TMA box, see gemm.cuh =====================
    const int nA = (AMODE == A_CONV3 && !p.merged) ? 4 : 1;
    if (lane <= nA) {
      int stage = 0; uint32_t phase = 0;
      for (int pr = cluster_id; pr < num_pairs; pr += num_clusters) {
        const int m_blk = p.m_tile0 + (pr / p.num_n_tiles) * 2 + (int)rank, n_blk = pr % p.num_n_tiles;
        const int b_row = n_blk * BLOCK_N + (int)rank * (BLOCK_N / 2);
        int cn = 0, ch0 = 0;
        if (AMODE == A_CONV3 && lane < nA) {
          const int g = m_blk * 4 + lane;
          cn = g / p.sb_per_img;
          ch0 = (g - cn * p.sb_per_img) * p.bh;
        }
        int tap = 0, cb = 0;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          if (p.debug_skip_tma) {
            if (leader && lane == 0) ptx::mbar_arrive(&full_bar[stage]);
          } else {
            const uint32_t lead_full = ptx::mapa(ptx::smem_u32(&full_bar[stage]), 0);
            if (leader && lane == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * (A_STAGE + B_STAGE));
            if (lane < nA) {
              uint8_t* a_dst = smem_a + stage * A_STAGE;
              if (AMODE == A_PLAIN) {
                const int rs = kb / p.kb_per_shift, kc = kb - rs * p.kb_per_shift;
                ptx::tma_load_2d_2cta(&tmA, lead_full, a_dst, kc * BLOCK_K, m_blk * BLOCK_M + rs * p.row_shift_mul);
              } else {
                const int r = tap / 3, sx = tap - 3 * r;
                ptx::tma_load_4d_2cta(&tmA, lead_full, a_dst + lane * 4096, cb * BLOCK_K, sx - 1, ch0 + r - 1, cn);
              }
            } else {
              ptx::tma_load_2d_2cta(&tmB, lead_full, smem_b + stage * B_STAGE, kb * BLOCK_K, b_row);
            }
          }
          if (++cb == p.cin_blocks) { cb = 0; ++tap; }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      int stage = 0; uint32_t phase = 0; int it = 0;
      for (int pr = cluster_id; pr < num_pairs; pr += num_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint64_t a_desc = ptx::make_desc_k_sw128(ptx::smem_u32(smem_a + stage * A_STAGE));
          const uint64_t b_desc = ptx::make_desc_k_sw128(ptx::smem_u32(smem_b + stage * B_STAGE));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) ptx::mma_f16_ss_2cta(d_tmem, a_desc + 2 * k, b_desc + 2 * k, IDESC, (kb | k) != 0);
          ptx::tc_commit_2cta_mc(&empty_bar[stage], 3);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::tc_commit_2cta_mc(&tmem_full[acc], 3);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps (both CTAs, own 128 rows) =====================
    const int q = warp_idx & 3;
    int it = 0;
    for (int pr = cluster_id; pr < num_pairs; pr += num_clusters, ++it) {
      const int m_blk = p.m_tile0 + (pr / p.num_n_tiles) * 2 + (int)rank, n_blk = pr % p.num_n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N;
      const int chalf = (warp_idx - 2) >> 2;
      if (m_blk < p.m_tile0 + p.num_m_tiles)
        run_epilogue<BLOCK_N, EPI>(p, tbase, m_blk, n_blk, q, lane, chalf * (BLOCK_N / 2), (chalf + 1) * (BLOCK_N / 2));
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&tmem_empty[acc]), 0));
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();            // neither CTA may exit (or free TMEM) while its peer can still touch its smem / barriers
  if (warp_idx == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc_2cta(tmem_base, TMEM_COLS); }
}

}  // namespace gemm
