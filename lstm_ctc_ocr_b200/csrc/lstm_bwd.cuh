// Persistent backward recurrence (BPTT) of the BiLSTM for sm_100a; mirror image of csrc/lstm.cuh.
//
// Restates what tf.gradients produces for tf.contrib.rnn.LSTMCell under bidirectional_dynamic_rnn
// (lib/networks/network.py:104-107, lib/lstm/train.py:82).  Per step s = T-1 .. 0 (step space: frame t = s for the
// forward direction, len-1-s for the backward direction; inactive when s >= len):
//   dh = d_out[t] + dz_{s+1} W_h^T          dc = dc_{s+1->s} + dh * o * (1 - tanh(c_s)^2)
//   do = dh * tanh(c_s) * o(1-o)   di = dc * j * i(1-i)   dj = dc * i * (1-j^2)   df = dc * c_{s-1} * f(1-f)
//   dc_{s->s-1} = dc * f
// Cluster of 8 CTAs per (direction, 128-sample tile); CTA `rank` owns 32 hidden units: it keeps dc for them in
// registers, holds W_h[units, all 1024 gate columns] (64 KB bf16, K-major over gates) resident in shared memory, and per
// step computes dh_rec[128 x 32] = dz_{s+1}[128 x 1024] * W_h^T on tensor cores (64 x tcgen05.mma 128x32x16), streaming
// dz_{s+1} (written to global/L2 by the whole cluster one step earlier) through a 6-stage TMA ring.  The tile is identical
// for the 8 CTAs of a cluster, so each K-block is read from L2 ONCE and TMA-multicast into all 8 shared memories.
// Outputs: dz for every (sample, frame) in FRAME order (`dz_all`, consumed by the dW_x / dW_h / dx GEMMs).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "lstm.cuh"

namespace lstm_bwd {

constexpr int NUM_THREADS = 192;
constexpr int BLOCK_M = 128;
constexpr int CS = 8;
constexpr int UPC = 32;
constexpr int STAGES = 6;
constexpr int B_BYTES = 16 * UPC * 128;          // 16 K-blocks x [32 rows x 128 B] = 64 KB
constexpr int A_STAGE = BLOCK_M * 128;           // 16 KB
constexpr int BAR_OFFSET = B_BYTES + STAGES * A_STAGE;
constexpr int SMEM_BYTES = BAR_OFFSET + 256 + 1024;

struct Params {
  const __nv_bfloat16* gates;     // [2][Nimg][T][4][256]  (saved by the forward kernel)
  const float* csave;             // [2][Nimg][T][256]
  const __nv_bfloat16* d_out;     // [Nimg*H, 512] gradient w.r.t. the LSTM output (frame order)
  __nv_bfloat16* dz_state;        // [2 bufs][2 dirs][Npad][1024] step-order exchange buffer
  __nv_bfloat16* dz_all;          // [Nimg*H, 2048] frame order, permuted gate columns, [fw | bw]
  const int* seq_len;
  int Nimg, Npad, H, T, tiles_per_dir;
};

__device__ __forceinline__ uint4 pack8(const float* v) {
  return make_uint4(ptx::pack_bf16x2(v[0], v[1]), ptx::pack_bf16x2(v[2], v[3]), ptx::pack_bf16x2(v[4], v[5]),
                    ptx::pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ void unpack8(const uint4 q, float* v) {
  v[0] = ptx::bf16_lo(q.x); v[1] = ptx::bf16_hi(q.x); v[2] = ptx::bf16_lo(q.y); v[3] = ptx::bf16_hi(q.y);
  v[4] = ptx::bf16_lo(q.z); v[5] = ptx::bf16_hi(q.z); v[6] = ptx::bf16_lo(q.w); v[7] = ptx::bf16_hi(q.w);
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
lstm_bwd_kernel(const __grid_constant__ CUtensorMap tmDz, const __grid_constant__ CUtensorMap tmW, const Params p) {
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(BLOCK_M, UPC);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem;
  uint8_t* smem_a = smem + B_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + BAR_OFFSET);
  uint64_t* a_empty = a_full + STAGES;
  uint64_t* b_full = a_empty + STAGES;
  uint64_t* acc_full = b_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)lstm::cluster_ctarank();
  const int unit = blockIdx.x / CS;
  const int dir = unit / p.tiles_per_dir;
  const int tile = unit - dir * p.tiles_per_dir;

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmDz);
    ptx::prefetch_tmap(&tmW);
    for (int i = 0; i < STAGES; ++i) { ptx::mbar_init(&a_full[i], 1); ptx::mbar_init(&a_empty[i], CS); }   // slot free = all 8 CTAs consumed it
    ptx::mbar_init(b_full, 1);
    ptx::mbar_init(acc_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp_idx == 1) { ptx::tmem_alloc(tmem_ptr, 32); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp_idx == 0 && lane == 0) {      // resident W_h rows [dir*256 + rank*32, +32) x 1024 gate columns
    ptx::mbar_arrive_expect_tx(b_full, B_BYTES);
    for (int kb = 0; kb < 16; ++kb) ptx::tma_load_2d(&tmW, b_full, smem_b + kb * UPC * 128, kb * 64, dir * 256 + rank * UPC);
  }

  const int q = warp_idx & 3;
  const int row = q * 32 + lane;
  const int n = tile * BLOCK_M + row;
  const bool is_epi = warp_idx >= 2;
  const bool okn = is_epi && (n < p.Nimg);
  const int len = okn ? min(max(__ldg(p.seq_len + n), 0), p.T) : 0;
  float dcr[UPC];
#pragma unroll
  for (int i = 0; i < UPC; ++i) dcr[i] = 0.f;

  if (warp_idx == 1 && lane == 0) ptx::mbar_wait(b_full, 0);

  uint32_t prod_parity = 0;  // producer lane l: parity of ring slot l (flips on every use of that slot)
  uint32_t mma_parity = 0;   // MMA lane: bit s = parity of ring slot s
  int nmma = 0;              // number of accumulations completed (acc_full parity)

  for (int s = p.T - 1; s >= 0; --s) {
    const bool has_rec = (s < p.T - 1);
    if (warp_idx == 0) {
      // K-block kb of every step lives in ring slot kb % STAGES.  Lanes 0..STAGES-1 own one slot each and issue their
      // K-blocks in lock-step (one SIMD cp.async.bulk.tensor per round instead of 16 serial single-thread issues).
      if (lane < STAGES && has_rec) {
        lstm::fence_proxy_async_all();
        const int zrow = ((((s + 1) & 1) * 2 + dir) * p.Npad) + tile * BLOCK_M;
        for (int kb = lane; kb < 16; kb += STAGES) {
          // all 8 CTAs of the cluster need the SAME dz tile: CTA (kb % 8) loads K-block kb once and multicasts it; every
          // CTA arms its own barrier.  a_empty[slot] counts the MMA commits of all 8 CTAs (multicast commit below).
          ptx::mbar_wait(&a_empty[lane], prod_parity ^ 1);
          ptx::mbar_arrive_expect_tx(&a_full[lane], A_STAGE);
          if ((kb & (CS - 1)) == rank)
            ptx::tma_load_2d_mc(&tmDz, &a_full[lane], smem_a + lane * A_STAGE, kb * 64, zrow, (uint16_t)0xFF);
          prod_parity ^= 1;
        }
      }
      __syncwarp();
    } else if (warp_idx == 1) {
      if (lane == 0 && has_rec) {
        for (int kb = 0; kb < 16; ++kb) {
          const int slot = kb % STAGES;
          ptx::mbar_wait(&a_full[slot], (mma_parity >> slot) & 1u);
          mma_parity ^= (1u << slot);
          ptx::tc_fence_after();
          const uint64_t a_desc = ptx::make_desc_k_sw128(ptx::smem_u32(smem_a + slot * A_STAGE));
          const uint64_t b_desc = ptx::make_desc_k_sw128(ptx::smem_u32(smem_b + kb * UPC * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(tmem_base, a_desc + 2 * k, b_desc + 2 * k, IDESC, (kb | k) != 0);
          ptx::tc_commit_mc(&a_empty[slot], (uint16_t)0xFF);
        }
        ptx::tc_commit(acc_full);
      }
      __syncwarp();
    } else {
      const bool active = s < len;
      const int t = active ? (dir ? (len - 1 - s) : s) : s;
      const size_t srow = ((size_t)dir * p.Nimg + n) * p.T + s;
      if (has_rec) {
        ptx::mbar_wait(acc_full, nmma & 1);
        ptx::tc_fence_after();
      }
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      __nv_bfloat16* zs = p.dz_state + ((size_t)(((s & 1) * 2 + dir) * p.Npad) + n) * 1024 + rank * 128;
      __nv_bfloat16* za = p.dz_all + ((size_t)n * p.H + t) * 2048 + dir * 1024 + rank * 128;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int u0 = hh * 16;
        uint32_t acc[16];
        if (has_rec) {
          ptx::tmem_ld_32x32b_x16(tbase + u0, acc);
          ptx::tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[i] = 0u;
        }
        float dzi[16], dzj[16], dzf[16], dzo[16];
        if (active) {
          const __nv_bfloat16* gs = p.gates + srow * 1024 + rank * UPC + u0;
          const float* cs = p.csave + srow * 256 + rank * UPC + u0;
          const __nv_bfloat16* dout = p.d_out + ((size_t)n * p.H + t) * 512 + dir * 256 + rank * UPC + u0;
          float gi[16], gj[16], gf[16], go[16], cc[16], cp[16], dh[16];
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            unpack8(__ldg(reinterpret_cast<const uint4*>(gs + 0 * 256) + v), gi + 8 * v);
            unpack8(__ldg(reinterpret_cast<const uint4*>(gs + 1 * 256) + v), gj + 8 * v);
            unpack8(__ldg(reinterpret_cast<const uint4*>(gs + 2 * 256) + v), gf + 8 * v);
            unpack8(__ldg(reinterpret_cast<const uint4*>(gs + 3 * 256) + v), go + 8 * v);
            unpack8(__ldg(reinterpret_cast<const uint4*>(dout) + v), dh + 8 * v);
          }
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(cs) + v);
            cc[4 * v] = a.x; cc[4 * v + 1] = a.y; cc[4 * v + 2] = a.z; cc[4 * v + 3] = a.w;
            const float4 b = (s > 0) ? __ldg(reinterpret_cast<const float4*>(cs - 256) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
            cp[4 * v] = b.x; cp[4 * v + 1] = b.y; cp[4 * v + 2] = b.z; cp[4 * v + 3] = b.w;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float dht = dh[i] + __uint_as_float(acc[i]);
            const float tc = ptx::fast_tanh(cc[i]);
            const float dc = dcr[u0 + i] + dht * go[i] * (1.f - tc * tc);
            dzo[i] = dht * tc * go[i] * (1.f - go[i]);
            dzi[i] = dc * gj[i] * gi[i] * (1.f - gi[i]);
            dzj[i] = dc * gi[i] * (1.f - gj[i] * gj[i]);
            dzf[i] = dc * cp[i] * gf[i] * (1.f - gf[i]);
            dcr[u0 + i] = dc * gf[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) { dzi[i] = 0.f; dzj[i] = 0.f; dzf[i] = 0.f; dzo[i] = 0.f; }
        }
        if (okn) {
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            const uint4 qi = pack8(dzi + 8 * v), qj = pack8(dzj + 8 * v), qf = pack8(dzf + 8 * v), qo = pack8(dzo + 8 * v);
            *(reinterpret_cast<uint4*>(zs + 0 * 32 + u0) + v) = qi;
            *(reinterpret_cast<uint4*>(zs + 1 * 32 + u0) + v) = qj;
            *(reinterpret_cast<uint4*>(zs + 2 * 32 + u0) + v) = qf;
            *(reinterpret_cast<uint4*>(zs + 3 * 32 + u0) + v) = qo;
            *(reinterpret_cast<uint4*>(za + 0 * 32 + u0) + v) = qi;
            *(reinterpret_cast<uint4*>(za + 1 * 32 + u0) + v) = qj;
            *(reinterpret_cast<uint4*>(za + 2 * 32 + u0) + v) = qf;
            *(reinterpret_cast<uint4*>(za + 3 * 32 + u0) + v) = qo;
          }
        }
      }
      if (has_rec) ++nmma;
      lstm::fence_proxy_async_all();
      ptx::tc_fence_before();
    }
    lstm::cluster_arrive_release();
    lstm::cluster_wait_acquire();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 32); }
}

}  // namespace lstm_bwd
