// Persistent backward recurrence (BPTT) of the BiLSTM for sm_100a; mirror image of csrc/lstm.cuh.
//
// Restates what tf.gradients produces for tf.contrib.rnn.LSTMCell under bidirectional_dynamic_rnn
// (lib/networks/network.py:104-107, lib/lstm/train.py:82).  Per step s = T-1 .. 0 (step space: frame t = s for the
// forward direction, len-1-s for the backward direction; inactive when s >= len):
//   dh = d_out[t] + dz_{s+1} W_h^T          dc = dc_{s+1->s} + dh * o * (1 - tanh(c_s)^2)
//   do = dh * tanh(c_s) * o(1-o)   di = dc * j * i(1-i)   dj = dc * i * (1-j^2)   df = dc * c_{s-1} * f(1-f)
//   dc_{s->s-1} = dc * f
// Cluster of 8 CTAs per (direction, 128-sample tile); CTA `rank` owns 32 hidden units (dc in registers) and their 128 gate
// columns.  The recurrent product is split along K (the gate axis): each CTA multiplies ITS OWN dz slice [128 x 128]
// (written by its epilogue straight into shared memory in the swizzled K-major layout -- no global round trip for the A
// operand) with the resident W_h[all 256 units, its 128 gate columns] (64 KB bf16): 8 x tcgen05.mma 128x256x16 per step.
// (A first version multiplied the full dz [128 x 1024] by W_h[32 units] -- 64 MMAs of N = 32 per step; a tcgen05.mma costs
// ~128 cycles for any N <= 256, so that spent 4 us per step in the tensor pipe alone.)  The 8 partial products are reduced
// across the cluster through DISTRIBUTED SHARED MEMORY: every CTA sends the [128 x 32] slice that belongs to owner CTA i
// (bf16) straight into CTA i's double-buffered "inbox" with st.shared::cluster, barrier.cluster (release/acquire), then each
// thread sums the 8 inbox entries of its own 32 units from local shared memory.  (A version that went through L2 with f32
// partials spent ~40 % of the step draining 1 KB of stores per thread at the barrier and ~15 % in the 8 dependent reads.)
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
constexpr int B_BYTES = 2 * 256 * 128;           // 2 K-blocks x [256 unit rows x 128 B] = 64 KB
constexpr int A_BYTES = 2 * BLOCK_M * 128;       // 2 K-blocks x [128 rows x 128 B] = 32 KB
constexpr int INBOX_BYTES = 2 * CS * BLOCK_M * UPC * 2;   // [2 bufs][8 src CTAs][128 rows][32 units] bf16 = 128 KB
constexpr int BAR_OFFSET = B_BYTES + A_BYTES + INBOX_BYTES;
constexpr int SMEM_BYTES = BAR_OFFSET + 64 + 1024;

struct Params {
  const __nv_bfloat16* gates;     // [2][Nimg][T][4][256]  (saved by the forward kernel)
  const float* csave;             // [2][Nimg][T][256]
  const __nv_bfloat16* d_out;     // [Nimg*H, 512] gradient w.r.t. the LSTM output (frame order)
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
lstm_bwd_kernel(const __grid_constant__ CUtensorMap tmW, const Params p) {
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(BLOCK_M, 256);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem;
  uint8_t* smem_a = smem + B_BYTES;
  uint8_t* inbox = smem + B_BYTES + A_BYTES;
  uint64_t* b_full = reinterpret_cast<uint64_t*>(smem + BAR_OFFSET);
  uint64_t* a_ready = b_full + 1;
  uint64_t* acc_full = a_ready + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)lstm::cluster_ctarank();
  const int unit = blockIdx.x / CS;
  const int dir = unit / p.tiles_per_dir;
  const int tile = unit - dir * p.tiles_per_dir;

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmW);
    ptx::mbar_init(b_full, 1);
    ptx::mbar_init(a_ready, 4);          // one arrive per epilogue warp
    ptx::mbar_init(acc_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp_idx == 1) { ptx::tmem_alloc(tmem_ptr, 256); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp_idx == 0 && lane == 0) {      // resident W_h[dir, all 256 units, gate columns rank*128 .. +128), K-major over gates
    ptx::mbar_arrive_expect_tx(b_full, B_BYTES);
    for (int kb = 0; kb < 2; ++kb) ptx::tma_load_2d(&tmW, b_full, smem_b + kb * 256 * 128, rank * 128 + kb * 64, dir * 256);
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

  for (int s = p.T - 1; s >= 0; --s) {
    const uint32_t par = (uint32_t)(p.T - 1 - s) & 1u;                         // parity of this step's a_ready / acc_full use
    if (warp_idx == 1) {
      if (lane == 0 && s > 0) {                                                // step 0's partial product is never consumed
        ptx::mbar_wait(a_ready, par);
        ptx::tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t a_desc = ptx::make_desc_k_sw128(ptx::smem_u32(smem_a + kb * BLOCK_M * 128));
          const uint64_t b_desc = ptx::make_desc_k_sw128(ptx::smem_u32(smem_b + kb * 256 * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(tmem_base, a_desc + 2 * k, b_desc + 2 * k, IDESC, (kb | k) != 0);
        }
        ptx::tc_commit(acc_full);
      }
      __syncwarp();
    } else if (warp_idx >= 2) {
      const bool active = s < len;
      const int t = active ? (dir ? (len - 1 - s) : s) : s;
      const size_t srow = ((size_t)dir * p.Nimg + n) * p.T + s;
      // The saved forward state of step s-1 (gates, c, c_prev, d_out: ~0.7 KB per thread, long evicted from L2) is pulled
      // into L2 one step ahead so the dependent loads of the next iteration do not pay HBM latency on the serial chain.
      if (s >= 1 && (s - 1) < len) {
        const size_t prow = srow - 1;
        const int tp = dir ? (len - s) : (s - 1);
#pragma unroll
        for (int g = 0; g < 4; ++g) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.gates + prow * 1024 + g * 256 + rank * UPC));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.csave + prow * 256 + rank * UPC));
        if (s >= 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.csave + (prow - 1) * 256 + rank * UPC));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.d_out + ((size_t)n * p.H + tp) * 512 + dir * 256 + rank * UPC));
      }
      // ---- recurrent gradient for this thread's 32 units = sum of the 8 CTAs' partial products of step s+1
      float dh_rec[UPC];
#pragma unroll
      for (int i = 0; i < UPC; ++i) dh_rec[i] = 0.f;
      if (s < p.T - 1) {
        const uint8_t* src = inbox + (size_t)((s + 1) & 1) * (CS * BLOCK_M * UPC * 2) + (size_t)row * (UPC * 2);
#pragma unroll
        for (int jj = 0; jj < CS; ++jj) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float f[8];
            unpack8(*reinterpret_cast<const uint4*>(src + (size_t)jj * (BLOCK_M * UPC * 2) + 16 * i), f);
#pragma unroll
            for (int e = 0; e < 8; ++e) dh_rec[8 * i + e] += f[e];
          }
        }
      }
      __nv_bfloat16* za = p.dz_all + ((size_t)n * p.H + t) * 2048 + dir * 1024 + rank * 128;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int u0 = hh * 16;
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
            const float dht = dh[i] + dh_rec[u0 + i];
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
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const uint4 qg[4] = {pack8(dzi + 8 * v), pack8(dzj + 8 * v), pack8(dzf + 8 * v), pack8(dzo + 8 * v)};
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            // A operand (own dz slice), K index = g*32 + u: K-block g>>1, 16-byte chunk ((g&1)*32 + u0)/8 + v, 128B swizzle
            const int chunk = ((g & 1) * 32 + u0) / 8 + v;
            *reinterpret_cast<uint4*>(smem_a + (g >> 1) * (BLOCK_M * 128) + row * 128 + ((chunk ^ (row & 7)) << 4)) = qg[g];
            if (okn) *(reinterpret_cast<uint4*>(za + g * 32 + u0) + v) = qg[g];
          }
        }
      }
      if (s > 0) {
        ptx::fence_proxy_async_smem();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(a_ready);
        // ---- this CTA's partial product P[128 x 256]: columns 32*i .. 32*i+31 belong to owner CTA i -> its inbox (DSMEM)
        ptx::mbar_wait(acc_full, par);
        ptx::tc_fence_after();
        const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const uint32_t my_slot = ptx::smem_u32(inbox) + (uint32_t)((s & 1) * (CS * BLOCK_M * UPC * 2) + (rank * BLOCK_M + row) * (UPC * 2));
#pragma unroll 1
        for (int owner = 0; owner < CS; ++owner) {
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(tbase + owner * 32, v);
          ptx::tmem_ld_wait();
          const uint32_t dst = ptx::mapa(my_slot, (uint32_t)owner);
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 2 * i),
                         "r"(ptx::pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1]))),
                         "r"(ptx::pack_bf16x2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]))),
                         "r"(ptx::pack_bf16x2(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]))),
                         "r"(ptx::pack_bf16x2(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])))
                         : "memory");
          }
        }
        ptx::tc_fence_before();
      }
    }
    lstm::cluster_arrive_release();
    lstm::cluster_wait_acquire();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 256); }
}

}  // namespace lstm_bwd
