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
// (Two alternatives were built and measured on the B200 and were NOT faster -- the step is bound by its serial latency chain,
// not by MMA count or exchange volume: splitting the product along K with the dz slice written straight into shared memory and
// the 8 partial products reduced through L2 (1.98 ms) or through DSMEM inboxes (1.81 ms) vs this version (1.68 ms).)
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
  const __nv_bfloat16* gates;     // saved by the forward kernel, coalesced per batch tile (common.cuh: lstm_gate_off)
  const float* csave;             // (common.cuh: lstm_c_off)
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
      const size_t dts = (size_t)unit * p.T + s;                    // coalesced saved-state layout, common.cuh
      // The saved forward state of step s-1 (gates, c, c_prev, d_out: ~0.7 KB per thread, long evicted from L2) is pulled
      // into L2 one step ahead so the dependent loads of the next iteration do not pay HBM latency on the serial chain.
      if (s >= 1 && (s - 1) < len) {
        const int tp = dir ? (len - s) : (s - 1);
#pragma unroll
        for (int g = 0; g < 4; ++g) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.gates + lstm_gate_off(dts - 1, g, rank * UPC, row)));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.csave + lstm_c_off(dts - 1, rank * UPC, row)));
        if (s >= 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.csave + lstm_c_off(dts - 2, rank * UPC, row)));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.d_out + ((size_t)n * p.H + tp) * 512 + dir * 256 + rank * UPC));
      }
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
          const __nv_bfloat16* gs = p.gates + lstm_gate_off(dts, 0, rank * UPC + u0, row);
          const float* cs = p.csave + lstm_c_off(dts, rank * UPC + u0, row);
          const __nv_bfloat16* dout = p.d_out + ((size_t)n * p.H + t) * 512 + dir * 256 + rank * UPC + u0;
          float gi[16], gj[16], gf[16], go[16], cc[16], cp[16], dh[16];
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            unpack8(__ldg(reinterpret_cast<const uint4*>(gs + 0 * LSTM_GATE_STRIDE + v * LSTM_GCHUNK_STRIDE)), gi + 8 * v);
            unpack8(__ldg(reinterpret_cast<const uint4*>(gs + 1 * LSTM_GATE_STRIDE + v * LSTM_GCHUNK_STRIDE)), gj + 8 * v);
            unpack8(__ldg(reinterpret_cast<const uint4*>(gs + 2 * LSTM_GATE_STRIDE + v * LSTM_GCHUNK_STRIDE)), gf + 8 * v);
            unpack8(__ldg(reinterpret_cast<const uint4*>(gs + 3 * LSTM_GATE_STRIDE + v * LSTM_GCHUNK_STRIDE)), go + 8 * v);
            unpack8(__ldg(reinterpret_cast<const uint4*>(dout) + v), dh + 8 * v);
          }
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(cs + v * LSTM_CCHUNK_STRIDE));
            cc[4 * v] = a.x; cc[4 * v + 1] = a.y; cc[4 * v + 2] = a.z; cc[4 * v + 3] = a.w;
            const float4 b = (s > 0) ? __ldg(reinterpret_cast<const float4*>(cs - LSTM_CSTEP_STRIDE + v * LSTM_CCHUNK_STRIDE)) : make_float4(0.f, 0.f, 0.f, 0.f);
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


// ---------------------------------------------------------------------------------------------------------------------------
// v2 ("ks"): the same recurrence with the product split along K and NOTHING but generic-proxy traffic between CTAs.
//
// v1 above moves the whole dz_{s+1} tile (128 x 1024, 256 KB) into every CTA each step (multicast ring with cluster-wide slot
// hand-shakes), issues 64 tcgen05.mma of N = 32 (an MMA instruction costs >= ~128 cycles whatever its N) and pays two
// fence.proxy.async and a cluster barrier per step: 26 us per step.  Here CTA `rank` multiplies ONLY ITS OWN dz slice
// (128 x 128 gate columns, written by its own epilogue straight into shared memory as the no-swizzle A operand) with the
// resident W_h[all 256 units, its 128 gate columns]: 8 tcgen05.mma of 128 x 256 x 16 give its partial dh for ALL units.  The
// partials are exchanged all-to-all through L2 as bf16 (8 KB per (source, destination) pair): plain st.global, one
// release.cluster arrive on every peer's mbarrier, acquire.cluster wait, ld.global.cg of the 8 partial rows, f32 sum --
// no async proxy, no proxy fences, no cluster barrier on the critical path.
//   X[buf = s & 1][unit][dst][src][128 rows][32 units] bf16 is the exchange buffer (double buffered: a source can only reach
//   step s-2 after every peer finished reading step s, because its own step s-1 needs all peers' step s-1 partials).
namespace ks {
constexpr int NUM_THREADS = 320;                 // warp 0 setup, warp 1 MMA, warps 2..9 epilogue
constexpr int EPI_THREADS = 256;
constexpr int B_BYTES = 2 * 256 * 128;           // 2 K-blocks x [256 unit rows x 128 B] (SW128) = 64 KB
constexpr int A_BYTES = 16 * BLOCK_M * 16;       // [16 K-chunks][128 rows][16 B] = 32 KB, no swizzle
constexpr int BAR_OFFSET = B_BYTES + A_BYTES;
constexpr int SMEM_BYTES = BAR_OFFSET + 128 + 1024;
constexpr int PAIR_BYTES = BLOCK_M * UPC * 2;    // one (source, destination) block of partial sums: 8 KB
}  // namespace ks

__global__ void __launch_bounds__(ks::NUM_THREADS, 1)
lstm_bwd_ks_kernel(const __grid_constant__ CUtensorMap tmW, const Params p, uint8_t* __restrict__ xbuf) {
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(BLOCK_M, 256);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_b = smem;
  uint8_t* smem_a = smem + ks::B_BYTES;
  uint64_t* b_full = reinterpret_cast<uint64_t*>(smem + ks::BAR_OFFSET);
  uint64_t* acc_full = b_full + 1;
  uint64_t* a_ready = acc_full + 1;
  uint64_t* part_ready = a_ready + 1;            // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(part_ready + 2);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)lstm::cluster_ctarank();
  const int unit = blockIdx.x / CS;
  const int dir = unit / p.tiles_per_dir;
  const int tile = unit - dir * p.tiles_per_dir;
  const int num_units = 2 * p.tiles_per_dir;

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmW);
    ptx::mbar_init(b_full, 1);
    ptx::mbar_init(acc_full, 1);
    ptx::mbar_init(a_ready, ks::EPI_THREADS);
    ptx::mbar_init(&part_ready[0], CS);
    ptx::mbar_init(&part_ready[1], CS);
    ptx::fence_barrier_init();
    // resident W_h[dir: all 256 units][gate columns rank*128 .. +128)
    ptx::mbar_arrive_expect_tx(b_full, ks::B_BYTES);
    for (int kb = 0; kb < 2; ++kb) ptx::tma_load_2d(&tmW, b_full, smem_b + kb * 256 * 128, rank * 128 + kb * 64, dir * 256);
  }
  if (warp_idx == 1) { ptx::tmem_alloc(tmem_ptr, 256); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  lstm::cluster_arrive_release();                  // peers arrive on this CTA's part_ready barriers
  lstm::cluster_wait_acquire();

  if (warp_idx == 1) {
    // ===================== MMA issuer: partial dh[128 x 256] = dz_{s+1}[128 x own 128 gate columns] * W_h^T =====================
    if (lane == 0) {
      ptx::mbar_wait(b_full, 0);
      for (int s = p.T - 2; s >= 0; --s) {
        const uint32_t ph = (uint32_t)(p.T - 2 - s) & 1u;
        ptx::mbar_wait(a_ready, ph);
        ptx::tc_fence_after();
        const uint32_t a_base = ptx::smem_u32(smem_a);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t a_desc = ptx::make_desc_k_nosw(a_base + k * 4096, 2048, 128);
          const uint64_t b_desc = ptx::make_desc_k_sw128(ptx::smem_u32(smem_b + (k >> 2) * 256 * 128)) + 2 * (k & 3);
          ptx::mma_f16_ss(tmem_base, a_desc, b_desc, IDESC, k != 0);
        }
        ptx::tc_commit(acc_full);
      }
    }
    __syncwarp();
  } else if (warp_idx >= 2) {
    // ===================== epilogue: thread = (sample row, half) =====================
    const int q = warp_idx & 3;
    const int hh = (warp_idx - 2) >> 2;            // exchange: destination CTAs hh*4 .. hh*4+3; cell: units hh*16 .. +16 of this CTA
    const int u0 = hh * 16;
    const int row = q * 32 + lane;
    const int n = tile * BLOCK_M + row;
    const bool okn = n < p.Nimg;
    const int len = okn ? min(max(__ldg(p.seq_len + n), 0), p.T) : 0;
    const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + hh * 128;
    float dcr[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) dcr[i] = 0.f;
    const size_t buf_stride = (size_t)num_units * CS * CS * ks::PAIR_BYTES;
    uint8_t* x_unit = xbuf + (size_t)unit * CS * CS * ks::PAIR_BYTES;

    for (int s = p.T - 1; s >= 0; --s) {
      const bool has_rec = (s < p.T - 1);
      const bool active = s < len;
      const int t = active ? (dir ? (len - 1 - s) : s) : s;
      const size_t dts = (size_t)unit * p.T + s;                    // coalesced saved-state layout, common.cuh
      // pull the next step's saved state into L2 one step ahead (it was evicted long ago)
      if (s >= 1 && (s - 1) < len) {
        const int tp = dir ? (len - s) : (s - 1);
#pragma unroll
        for (int g = 0; g < 4; ++g) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.gates + lstm_gate_off(dts - 1, g, rank * UPC + u0, row)));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.csave + lstm_c_off(dts - 1, rank * UPC + u0, row)));
        if (s >= 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.csave + lstm_c_off(dts - 2, rank * UPC + u0, row)));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.d_out + ((size_t)n * p.H + tp) * 512 + dir * 256 + rank * UPC + u0));
      }

      float rec[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) rec[i] = 0.f;
      if (has_rec) {
        const int b = s & 1;
        uint8_t* xb = x_unit + (size_t)b * buf_stride;
        ptx::mbar_wait(acc_full, (uint32_t)(p.T - 2 - s) & 1u);
        ptx::tc_fence_after();
        // ---- this CTA's partial sums for destinations hh*4 .. hh*4+3 -> X[dst][src = rank][row]
#pragma unroll 1
        for (int jj = 0; jj < 4; ++jj) {
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(tbase + jj * 32, v);
          ptx::tmem_ld_wait();
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) w[i] = ptx::pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
          uint8_t* dst = xb + ((size_t)((hh * 4 + jj) * CS + rank) * BLOCK_M + row) * 64;
          ptx::st_global_v8(dst, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
          ptx::st_global_v8(dst + 32, w[8], w[9], w[10], w[11], w[12], w[13], w[14], w[15]);
        }
        ptx::tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(ks::EPI_THREADS) : "memory");
        // one release.cluster arrive per peer (cumulative over the barrier above: covers every thread's stores)
        if (warp_idx == 2 && lane < CS) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&part_ready[b]), (uint32_t)lane));
      }
      // saved forward state of this step: issued before the exchange wait so that its latency hides behind it
      uint4 qg[4][2], qd[2];
      float4 qc[4], qp[4];
      if (active) {
        const __nv_bfloat16* gs = p.gates + lstm_gate_off(dts, 0, rank * UPC + u0, row);
        const float* cs = p.csave + lstm_c_off(dts, rank * UPC + u0, row);
        const __nv_bfloat16* dout = p.d_out + ((size_t)n * p.H + t) * 512 + dir * 256 + rank * UPC + u0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          qg[g][0] = __ldg(reinterpret_cast<const uint4*>(gs + g * LSTM_GATE_STRIDE));
          qg[g][1] = __ldg(reinterpret_cast<const uint4*>(gs + g * LSTM_GATE_STRIDE + LSTM_GCHUNK_STRIDE));
        }
        qd[0] = __ldg(reinterpret_cast<const uint4*>(dout));
        qd[1] = __ldg(reinterpret_cast<const uint4*>(dout) + 1);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          qc[v] = __ldg(reinterpret_cast<const float4*>(cs + v * LSTM_CCHUNK_STRIDE));
          qp[v] = (s > 0) ? __ldg(reinterpret_cast<const float4*>(cs - LSTM_CSTEP_STRIDE + v * LSTM_CCHUNK_STRIDE)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if (has_rec) {
        const int b = s & 1;
        uint8_t* xb = x_unit + (size_t)b * buf_stride;
        // ---- all 8 partial rows for this thread's 16 units
        ptx::mbar_wait_cluster(&part_ready[b], (uint32_t)((p.T - 2 - s) >> 1) & 1u);
        const uint8_t* src = xb + ((size_t)(rank * CS) * BLOCK_M + row) * 64 + hh * 32;
#pragma unroll
        for (int r = 0; r < CS; ++r) {
          const uint4 a = __ldcg(reinterpret_cast<const uint4*>(src + (size_t)r * BLOCK_M * 64));
          const uint4 c = __ldcg(reinterpret_cast<const uint4*>(src + (size_t)r * BLOCK_M * 64) + 1);
          float f[16];
          unpack8(a, f);
          unpack8(c, f + 8);
#pragma unroll
          for (int i = 0; i < 16; ++i) rec[i] += f[i];
        }
      }

      float dzi[16], dzj[16], dzf[16], dzo[16];
      if (active) {
        float gi[16], gj[16], gf[16], go[16], dh[16];
        unpack8(qg[0][0], gi); unpack8(qg[0][1], gi + 8);
        unpack8(qg[1][0], gj); unpack8(qg[1][1], gj + 8);
        unpack8(qg[2][0], gf); unpack8(qg[2][1], gf + 8);
        unpack8(qg[3][0], go); unpack8(qg[3][1], go + 8);
        unpack8(qd[0], dh); unpack8(qd[1], dh + 8);
        const float* cc = reinterpret_cast<const float*>(qc);
        const float* cp = reinterpret_cast<const float*>(qp);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float dht = dh[i] + rec[i];
          const float tc = ptx::fast_tanh(cc[i]);
          const float dc = dcr[i] + dht * go[i] * (1.f - tc * tc);
          dzo[i] = dht * tc * go[i] * (1.f - go[i]);
          dzi[i] = dc * gj[i] * gi[i] * (1.f - gi[i]);
          dzj[i] = dc * gi[i] * (1.f - gj[i] * gj[i]);
          dzf[i] = dc * cp[i] * gf[i] * (1.f - gf[i]);
          dcr[i] = dc * gf[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) { dzi[i] = 0.f; dzj[i] = 0.f; dzf[i] = 0.f; dzo[i] = 0.f; }
      }
      const uint4 zi0 = pack8(dzi), zi1 = pack8(dzi + 8), zj0 = pack8(dzj), zj1 = pack8(dzj + 8);
      const uint4 zf0 = pack8(dzf), zf1 = pack8(dzf + 8), zo0 = pack8(dzo), zo1 = pack8(dzo + 8);
      if (s > 0) {
        // A operand of the next step's product: K index = gate*32 + unit -> chunks gate*4 + hh*2 + {0, 1}
        uint8_t* a = smem_a + (size_t)(hh * 2) * 2048 + row * 16;
        *reinterpret_cast<uint4*>(a + 0 * 8192) = zi0; *reinterpret_cast<uint4*>(a + 0 * 8192 + 2048) = zi1;
        *reinterpret_cast<uint4*>(a + 1 * 8192) = zj0; *reinterpret_cast<uint4*>(a + 1 * 8192 + 2048) = zj1;
        *reinterpret_cast<uint4*>(a + 2 * 8192) = zf0; *reinterpret_cast<uint4*>(a + 2 * 8192 + 2048) = zf1;
        *reinterpret_cast<uint4*>(a + 3 * 8192) = zo0; *reinterpret_cast<uint4*>(a + 3 * 8192 + 2048) = zo1;
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(a_ready);
      }
      if (okn) {
        __nv_bfloat16* za = p.dz_all + ((size_t)n * p.H + t) * 2048 + dir * 1024 + rank * 128 + u0;
        ptx::st_global_v8(za + 0 * 32, zi0.x, zi0.y, zi0.z, zi0.w, zi1.x, zi1.y, zi1.z, zi1.w);
        ptx::st_global_v8(za + 1 * 32, zj0.x, zj0.y, zj0.z, zj0.w, zj1.x, zj1.y, zj1.z, zj1.w);
        ptx::st_global_v8(za + 2 * 32, zf0.x, zf0.y, zf0.z, zf0.w, zf1.x, zf1.y, zf1.z, zf1.w);
        ptx::st_global_v8(za + 3 * 32, zo0.x, zo0.y, zo0.z, zo0.w, zo1.x, zo1.y, zo1.z, zo1.w);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  lstm::cluster_arrive_release();                  // no CTA leaves while a peer may still arrive on its barriers
  lstm::cluster_wait_acquire();
  if (warp_idx == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 256); }
}

}  // namespace lstm_bwd
