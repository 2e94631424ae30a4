// Persistent BiLSTM recurrence for sm_100a: one thread-block CLUSTER per (direction, 128-sample batch tile).
//
// Replaces the tf.while_loop of tf.contrib.rnn.LSTMCell(256) under bidirectional_dynamic_rnn
// (lib/networks/network.py:104-107): per step  z = x_t W_x + b (precomputed, `xproj`) + h_{t-1} W_h ;
// i,j,f,o = split(z); c = sigma(f+1) c + sigma(i) tanh(j); h = sigma(o) tanh(c); zero output past sequence_length.
//
// Two generations live in this file:
//   lstm_mc_kernel<CS, MODE, EW>   (further down; the default, MODE 2 = "ms") -- h exchanged as 8 KB slices that land directly in
//                                  every CTA's no-swizzle A operand, signalled through mbarriers; no cluster barrier per step
//   lstm_persistent_kernel<CS>     (first, below; CRNN_LSTM_IMPL=persistent) -- h through global memory, a 64 KB TMA fetch per CTA,
//                                  fence.proxy.async and one barrier.cluster per step; its measured step timeline is the
//                                  reason the second generation exists (see the comment above lstm_mc_kernel)
//
// First generation:
// Cluster of CS CTAs; CTA `rank` owns UPC = 256/CS hidden units (4*UPC gate columns, ordered [i|j|f|o]):
//   * its W_h slice [4*UPC x 256] bf16 stays resident in shared memory for all T steps (loaded once by TMA)
//   * per step: TMA loads h_{t-1} [128 x 256] (written to global/L2 by the whole cluster in the previous step),
//     one elected thread issues 16 tcgen05.mma (128 x 4*UPC x 16) into TMEM, 4 epilogue warps (thread = sample row)
//     add the precomputed input projection (prefetched to registers before the MMA wait), run the cell with the
//     f32 cell state held in REGISTERS for the whole sequence, write h (bf16) for the next step and the output row
//   * one barrier.cluster per step orders the h exchange (generic-proxy global stores -> fence.proxy.async ->
//     release/acquire cluster barrier -> TMA loads of the next step)
// Clusters are independent (no grid-wide sync), so partial residency cannot deadlock.
// The backward direction reads xproj rows that the projection GEMM already stored reversed-by-length, so both
// directions index step s uniformly; outputs are written back at t = len-1-s.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace lstm {

constexpr int NUM_THREADS = 192;
constexpr int BLOCK_M = 128;

struct Params {
  const __nv_bfloat16* xproj;   // [Nimg*H, 2048]: [fw 1024 | bw 1024 (rows reversed by length)], permuted columns
  __nv_bfloat16* h_state;       // [2 bufs][2 dirs][Npad][256]
  __nv_bfloat16* lstm_out;      // [Nimg*H, 512]
  const int* seq_len;           // [Nimg]
  int Nimg, Npad, H, T, tiles_per_dir;
  // training only (nullptr for inference): activations the backward recurrence needs
  __nv_bfloat16* gates;         // post-activation gates i,j,f,o, coalesced per batch tile: layout in common.cuh (lstm_gate_off)
  float* csave;                 // cell state after the step (lstm_c_off)
  int swap_ls;                  // debug (CRNN_LSTM_SWAPLS=1): exchange the LBO/SBO fields of the no-swizzle A descriptor
  long long* trace;             // debug (CRNN_LSTM_TRACE=1): clock64 stamps of CTAs 0 and 5, steps 8..11, 16 events each
};

// debug timeline: one stamp per (selected CTA, step, event); all stamps of a CTA come from the same SM clock
#define LSTM_TRACE(ev)                                                                                  \
  do {                                                                                                  \
    if (p.trace != nullptr && lane == 0 && s >= 8 && s < 12 && (blockIdx.x == 0 || blockIdx.x == 5))    \
      p.trace[(((blockIdx.x ? 1 : 0) * 4 + (s - 8)) * 16) + (ev)] = clock64();                          \
  } while (0)

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <int CS>
struct Cfg {
  static constexpr int UPC = 256 / CS;          // hidden units per CTA
  static constexpr int NCOLS = 4 * UPC;         // gate columns per CTA
  static constexpr int B_BYTES = 4 * NCOLS * 128;   // 4 K-blocks x [NCOLS rows x 128 B]
  static constexpr int A_BYTES = 4 * BLOCK_M * 128; // 4 K-blocks x [128 rows x 128 B]
  static constexpr int BAR_OFFSET = A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFFSET + 128 + 1024;
};

template <int CS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
lstm_persistent_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmW, const Params p) {
  static_assert(CS == 8, "register budget of the epilogue is sized for 32 units per CTA");
  using C = Cfg<CS>;
  constexpr int UPC = C::UPC, NCOLS = C::NCOLS;
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(BLOCK_M, NCOLS);
  constexpr int HALF = UPC / 2;                 // units processed per epilogue pass (register budget)

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::A_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + C::BAR_OFFSET);   // [4] one per K-block
  uint64_t* b_full = a_full + 4;
  uint64_t* acc_full = b_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int unit = blockIdx.x / CS;                   // (dir, batch tile)
  const int dir = unit / p.tiles_per_dir;
  const int tile = unit - dir * p.tiles_per_dir;

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmH);
    ptx::prefetch_tmap(&tmW);
    for (int i = 0; i < 4; ++i) ptx::mbar_init(&a_full[i], 1);
    ptx::mbar_init(b_full, 1);
    ptx::mbar_init(acc_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp_idx == 1) {
    ptx::tmem_alloc(tmem_ptr, NCOLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // resident recurrent weights: rows [dir*1024 + rank*NCOLS, +NCOLS) of Bh [2048, 256]
  if (warp_idx == 0 && lane == 0) {
    ptx::mbar_arrive_expect_tx(b_full, C::B_BYTES);
    for (int kb = 0; kb < 4; ++kb)
      ptx::tma_load_2d(&tmW, b_full, smem_b + kb * NCOLS * 128, kb * 64, dir * 1024 + rank * NCOLS);
  }

  // ---- per-thread epilogue state (warps 2..5): one sample row, UPC units, cell state in registers
  const int q = warp_idx & 3;
  const int row = q * 32 + lane;
  const int n = tile * BLOCK_M + row;
  const bool is_epi = warp_idx >= 2;
  const bool okn = is_epi && (n < p.Nimg);
  const int len = okn ? min(max(__ldg(p.seq_len + n), 0), p.T) : 0;
  float cst[UPC];
#pragma unroll
  for (int i = 0; i < UPC; ++i) cst[i] = 0.f;

  if (warp_idx == 1 && lane == 0) ptx::mbar_wait(b_full, 0);      // W_h slice resident before the first MMA / exit

  for (int s = 0; s < p.T; ++s) {
    if (warp_idx == 0) {
      LSTM_TRACE(0);
      if (lane < 4 && s > 0) {        // one lane per K-block: the four TMA issues overlap instead of serialising
        fence_proxy_async_all();
        const int hrow = ((s & 1) * 2 + dir) * p.Npad + tile * BLOCK_M;
        ptx::mbar_arrive_expect_tx(&a_full[lane], BLOCK_M * 128);
        ptx::tma_load_2d(&tmH, &a_full[lane], smem_a + lane * BLOCK_M * 128, lane * 64, hrow);
      }
      LSTM_TRACE(1);
      __syncwarp();
    } else if (warp_idx == 1) {
      if (lane == 0 && s > 0) {
        const uint32_t ph = (s - 1) & 1;
        for (int kb = 0; kb < 4; ++kb) {
          ptx::mbar_wait(&a_full[kb], ph);
          if (kb == 0) LSTM_TRACE(2);
          if (kb == 3) LSTM_TRACE(3);
          ptx::tc_fence_after();
          const uint64_t a_desc = ptx::make_desc_k_sw128(ptx::smem_u32(smem_a + kb * BLOCK_M * 128));
          const uint64_t b_desc = ptx::make_desc_k_sw128(ptx::smem_u32(smem_b + kb * NCOLS * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(tmem_base, a_desc + 2 * k, b_desc + 2 * k, IDESC, (kb | k) != 0);
        }
        ptx::tc_commit(acc_full);
        LSTM_TRACE(4);
      }
      __syncwarp();
    } else {
      const bool active = s < len;
      const int t = active ? (dir ? (len - 1 - s) : s) : s;
      // prefetch this step's input projection (row n, step s; bw rows were stored reversed by the projection GEMM)
      uint4 xp[UPC / 2];                                   // 4 gates x UPC bf16 = UPC/2 x 16 B
      if (active) {
        const uint4* src = reinterpret_cast<const uint4*>(p.xproj + ((size_t)n * p.H + s) * 2048 + dir * 1024 + rank * NCOLS);
#pragma unroll
        for (int i = 0; i < UPC / 2; ++i) xp[i] = __ldg(src + i);
      }
      if (warp_idx == 2) LSTM_TRACE(5);
      if (s > 0) {
        ptx::mbar_wait(acc_full, (s - 1) & 1);
        ptx::tc_fence_after();
      }
      if (warp_idx == 2) LSTM_TRACE(6);
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      __nv_bfloat16* hn = p.h_state + ((size_t)(((s + 1) & 1) * 2 + dir) * p.Npad + n) * 256 + rank * UPC;
      __nv_bfloat16* lo = p.lstm_out + ((size_t)n * p.H + t) * 512 + dir * 256 + rank * UPC;
      const uint32_t* xw = reinterpret_cast<const uint32_t*>(xp);    // gate g, unit u -> bf16 index g*UPC + u
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int u0 = hh * HALF;
        uint32_t gi[HALF], gj[HALF], gf[HALF], go[HALF];
        if (s > 0) {
          ptx::tmem_ld_32x32b_x16(tbase + 0 * UPC + u0, gi);
          ptx::tmem_ld_32x32b_x16(tbase + 1 * UPC + u0, gj);
          ptx::tmem_ld_32x32b_x16(tbase + 2 * UPC + u0, gf);
          ptx::tmem_ld_32x32b_x16(tbase + 3 * UPC + u0, go);
          ptx::tmem_ld_wait();
          if (warp_idx == 2 && hh == 0) LSTM_TRACE(7);
        } else {
#pragma unroll
          for (int i = 0; i < HALF; ++i) { gi[i] = 0u; gj[i] = 0u; gf[i] = 0u; go[i] = 0u; }
        }
        uint32_t hp[HALF / 2];
        if (active) {
          float hv[HALF];
#pragma unroll
          for (int i = 0; i < HALF; ++i) {
            const int u = u0 + i;
            const uint32_t wi = xw[(0 * UPC + u) >> 1], wj = xw[(1 * UPC + u) >> 1], wf = xw[(2 * UPC + u) >> 1], wo = xw[(3 * UPC + u) >> 1];
            const float zi = __uint_as_float(gi[i]) + ((u & 1) ? ptx::bf16_hi(wi) : ptx::bf16_lo(wi));
            const float zj = __uint_as_float(gj[i]) + ((u & 1) ? ptx::bf16_hi(wj) : ptx::bf16_lo(wj));
            const float zf = __uint_as_float(gf[i]) + ((u & 1) ? ptx::bf16_hi(wf) : ptx::bf16_lo(wf));
            const float zo = __uint_as_float(go[i]) + ((u & 1) ? ptx::bf16_hi(wo) : ptx::bf16_lo(wo));
            const float ai = ptx::fast_sigmoid(zi), aj = ptx::fast_tanh(zj), af = ptx::fast_sigmoid(zf), ao = ptx::fast_sigmoid(zo);
            const float c = af * cst[u] + ai * aj;                 // forget_bias (+1.0) is folded into the projected bias
            cst[u] = c;
            hv[i] = ao * ptx::fast_tanh(c);
            if (p.gates != nullptr) {                              // reuse the (now dead) accumulator registers as staging
              gi[i] = __float_as_uint(ai); gj[i] = __float_as_uint(aj); gf[i] = __float_as_uint(af); go[i] = __float_as_uint(ao);
            }
          }
          if (p.gates != nullptr) {
            const size_t dts = (size_t)unit * p.T + s;                  // coalesced saved-state layout, common.cuh
            __nv_bfloat16* gs = p.gates + lstm_gate_off(dts, 0, rank * UPC + u0, row);
            float* cs = p.csave + lstm_c_off(dts, rank * UPC + u0, row);
#pragma unroll
            for (int i = 0; i < HALF; i += 8) {
              *reinterpret_cast<uint4*>(gs + 0 * LSTM_GATE_STRIDE + (i >> 3) * LSTM_GCHUNK_STRIDE) = make_uint4(ptx::pack_bf16x2(__uint_as_float(gi[i]), __uint_as_float(gi[i + 1])), ptx::pack_bf16x2(__uint_as_float(gi[i + 2]), __uint_as_float(gi[i + 3])), ptx::pack_bf16x2(__uint_as_float(gi[i + 4]), __uint_as_float(gi[i + 5])), ptx::pack_bf16x2(__uint_as_float(gi[i + 6]), __uint_as_float(gi[i + 7])));
              *reinterpret_cast<uint4*>(gs + 1 * LSTM_GATE_STRIDE + (i >> 3) * LSTM_GCHUNK_STRIDE) = make_uint4(ptx::pack_bf16x2(__uint_as_float(gj[i]), __uint_as_float(gj[i + 1])), ptx::pack_bf16x2(__uint_as_float(gj[i + 2]), __uint_as_float(gj[i + 3])), ptx::pack_bf16x2(__uint_as_float(gj[i + 4]), __uint_as_float(gj[i + 5])), ptx::pack_bf16x2(__uint_as_float(gj[i + 6]), __uint_as_float(gj[i + 7])));
              *reinterpret_cast<uint4*>(gs + 2 * LSTM_GATE_STRIDE + (i >> 3) * LSTM_GCHUNK_STRIDE) = make_uint4(ptx::pack_bf16x2(__uint_as_float(gf[i]), __uint_as_float(gf[i + 1])), ptx::pack_bf16x2(__uint_as_float(gf[i + 2]), __uint_as_float(gf[i + 3])), ptx::pack_bf16x2(__uint_as_float(gf[i + 4]), __uint_as_float(gf[i + 5])), ptx::pack_bf16x2(__uint_as_float(gf[i + 6]), __uint_as_float(gf[i + 7])));
              *reinterpret_cast<uint4*>(gs + 3 * LSTM_GATE_STRIDE + (i >> 3) * LSTM_GCHUNK_STRIDE) = make_uint4(ptx::pack_bf16x2(__uint_as_float(go[i]), __uint_as_float(go[i + 1])), ptx::pack_bf16x2(__uint_as_float(go[i + 2]), __uint_as_float(go[i + 3])), ptx::pack_bf16x2(__uint_as_float(go[i + 4]), __uint_as_float(go[i + 5])), ptx::pack_bf16x2(__uint_as_float(go[i + 6]), __uint_as_float(go[i + 7])));
            }
#pragma unroll
            for (int i = 0; i < HALF; i += 4)
              *reinterpret_cast<float4*>(cs + (i >> 2) * LSTM_CCHUNK_STRIDE) = make_float4(cst[u0 + i], cst[u0 + i + 1], cst[u0 + i + 2], cst[u0 + i + 3]);
          }
#pragma unroll
          for (int i = 0; i < HALF / 2; ++i) hp[i] = ptx::pack_bf16x2(hv[2 * i], hv[2 * i + 1]);
        } else {
#pragma unroll
          for (int i = 0; i < HALF / 2; ++i) hp[i] = 0u;     // zero output past sequence_length
        }
        if (okn) {
#pragma unroll
          for (int i = 0; i < HALF / 2; i += 4) {
            const uint4 v = make_uint4(hp[i], hp[i + 1], hp[i + 2], hp[i + 3]);
            *reinterpret_cast<uint4*>(hn + u0 + 2 * i) = v;
            *reinterpret_cast<uint4*>(lo + u0 + 2 * i) = v;
          }
        }
      }
      // make this thread's h stores visible to the async proxy (TMA loads of the next step) of the whole cluster
      if (warp_idx == 2) LSTM_TRACE(8);
      fence_proxy_async_all();
      ptx::tc_fence_before();
      if (warp_idx == 2) LSTM_TRACE(9);
    }
    cluster_arrive_release();
    if (warp_idx == 2) LSTM_TRACE(10);
    cluster_wait_acquire();
    if (warp_idx == 2) LSTM_TRACE(11);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, NCOLS);
  }
}


// ---------------------------------------------------------------------------------------------------------------------------
// v2: the same recurrence WITHOUT a cluster barrier (and without fence.proxy.async + a 64 KB TMA fetch per CTA) on the
// per-step critical path.  Measured timeline of the v1 kernel above (clock64, B200, one step = 8866 cycles): producer-side
// fence.proxy.async 2117, 64 KB TMA 1441, MMA tail 469, cell epilogue 2360, epilogue-side fence.proxy.async 1304,
// barrier.cluster arrive(release)+wait 1012.
//
//   * the A operand (h_{t-1}, [128 rows x 256]) lives in shared memory WITHOUT swizzle as [32 K-chunks][128 rows][16 B]
//     (8-row x 16-B core matrices, LBO = 2048, SBO = 128), so the 32 units one CTA produces are ONE contiguous 8 KB slice
//   * 8 epilogue warps (2 per TMEM lane quadrant, 16 units each) run the cell and write their h_t slice
//       DS = true : straight into this CTA's own copy of the next A buffer (st.shared), then 7 bulk copies
//                   shared::cta -> shared::cluster push the slice into the peers' A buffers and credit THEIR mbarriers
//       DS = false: into a private 8 KB global buffer, then one bulk copy global -> shared::cluster with cluster MULTICAST
//                   lands it in all 8 CTAs (L2 is read once per slice instead of 8 times)
//   * the MMA thread of each CTA waits on its own mbarrier (8 slices) and goes; nobody waits for a cluster barrier
//   * A is double buffered: a slice of h_t can only be sent after its sender saw h_{t-1} from every CTA, i.e. after every CTA's
//     MMA of step t-1 (the last reader of that buffer) has completed -- causality replaces a "buffer free" handshake
//   * the accumulator is single buffered for the same reason (MMA t+1 needs this CTA's own h_t, sent after its TMEM reads)
// Global exchange buffer (DS = false): p.h_state viewed as [2 bufs][2*tiles_per_dir units][8 ranks][8 KB].
// EW epilogue warps (8 or 16): EW/4 per TMEM lane quadrant, each owning 32/(EW/4) of the CTA's units.  The cell is a chain of
// MUFU latencies (ncu: XU pipe 22 %, issue slots 22 % busy), but 16 warps x 8 units measured SLOWER than 8 x 16 on the B200
// (0.422 vs 0.402 ms: the longer 512-thread barrier and 4 tcgen05.ld per 8 units eat the gain); model.cu instantiates EW = 8.
template <int EW>
struct McThreads {
  static constexpr int EPI = EW * 32;
  static constexpr int ALL = 64 + EPI;       // warp 0 setup, warp 1 MMA, warps 2.. epilogue
};

template <int CS>
struct CfgMc {
  static constexpr int UPC = 256 / CS;
  static constexpr int NCOLS = 4 * UPC;
  static constexpr int B_BYTES = 4 * NCOLS * 128;         // resident W_h slice (SW128 K-major, 4 K-blocks of 64)
  static constexpr int A_BYTES = 32 * BLOCK_M * 16;       // [32 K-chunks][128 rows][16 B] = 64 KB per buffer
  static constexpr int SLICE_BYTES = A_BYTES / CS;        // 8 KB: the 4 K-chunks one CTA produces
  static constexpr int BAR_OFFSET = 2 * A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFFSET + 128 + 1024;
};

__device__ __forceinline__ void bulk_copy_s2s_cluster(uint32_t dst_cluster_addr, const void* src_smem, uint32_t bytes,
                                                      uint32_t bar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster_addr),
               "r"(ptx::smem_u32(src_smem)), "r"(bytes), "r"(bar_cluster_addr)
               : "memory");
}

// MODE: 0 = global slice + multicast ("mc"), 1 = smem -> peer smem pushes ("ds"), 2 = smem slice -> bulk store to global ->
// multicast to the 7 peers ("ms": no generic-proxy global stores, hence no full fence.proxy.async on the critical path),
// 3 = generic proxy only ("gx"): st.global slice -> release.cluster arrive on every peer's mbarrier -> acquire.cluster wait ->
// the 8 epilogue warps copy the 64 KB h tile L2 -> smem with ld.global.cg / st.shared (the exchange the K-split BPTT uses)
template <int CS, int MODE, int EW>
__global__ void __launch_bounds__(McThreads<EW>::ALL, 1)
lstm_mc_kernel(const __grid_constant__ CUtensorMap tmW, const Params p) {
  static_assert(CS == 8, "8 CTAs x 32 units");
  using C = CfgMc<CS>;
  constexpr int UPC = C::UPC, NCOLS = C::NCOLS;
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(BLOCK_M, NCOLS);
  constexpr int HALF = UPC / (EW / 4);                      // units per epilogue warp (16 or 8)
  constexpr int MC_EPI_THREADS = McThreads<EW>::EPI;
  static_assert(EW == 8 || EW == 16, "epilogue warps");
  constexpr bool DS = (MODE == 1), MS = (MODE == 2), GX = (MODE == 3), LOCAL = DS || MS;   // LOCAL: own slice written in place
  constexpr uint32_t FILL_TX = LOCAL ? (CS - 1) * C::SLICE_BYTES : C::A_BYTES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (STS)
  uint8_t* smem_a = smem;                                   // [2][A_BYTES]
  uint8_t* smem_b = smem + 2 * C::A_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + C::BAR_OFFSET);   // [2] one per A buffer
  uint64_t* b_full = a_full + 2;
  uint64_t* acc_full = b_full + 1;
  uint64_t* part_ready = acc_full + 1;                      // [2] GX: the 8 CTAs' slices of one h buffer are in L2
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(part_ready + 2);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int unit = blockIdx.x / CS;                   // (dir, batch tile)
  const int dir = unit / p.tiles_per_dir;
  const int tile = unit - dir * p.tiles_per_dir;

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmW);
    ptx::mbar_init(&part_ready[0], CS);
    ptx::mbar_init(&part_ready[1], CS);
    // DS: one arrival arms the byte count (MMA thread), the other says "the local slice is in place" (epilogue)
    ptx::mbar_init(&a_full[0], GX ? MC_EPI_THREADS : LOCAL ? 2 : 1);      // GX: every epilogue thread copied its part of the tile
    ptx::mbar_init(&a_full[1], GX ? MC_EPI_THREADS : LOCAL ? 2 : 1);
    ptx::mbar_init(b_full, 1);
    ptx::mbar_init(acc_full, 1);
    ptx::fence_barrier_init();
    // first fills: buffer 1 receives h_0 (consumed at step 1), buffer 0 receives h_1 (consumed at step 2)
    if (!GX && p.T > 1) ptx::mbar_arrive_expect_tx(&a_full[1], FILL_TX);
    if (!GX && p.T > 2) ptx::mbar_arrive_expect_tx(&a_full[0], FILL_TX);
    ptx::mbar_arrive_expect_tx(b_full, C::B_BYTES);
    for (int kb = 0; kb < 4; ++kb)
      ptx::tma_load_2d(&tmW, b_full, smem_b + kb * NCOLS * 128, kb * 64, dir * 1024 + rank * NCOLS);
  }
  if (warp_idx == 1) {
    ptx::tmem_alloc(tmem_ptr, NCOLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // peers write into this CTA's smem and credit its mbarriers: everything above must be in place cluster-wide first
  cluster_arrive_release();
  cluster_wait_acquire();

  if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      ptx::mbar_wait(b_full, 0);
      for (int s = 1; s < p.T; ++s) {
        const int b = s & 1;
        ptx::mbar_wait(&a_full[b], ((s - 1) >> 1) & 1);
        LSTM_TRACE(3);
        ptx::tc_fence_after();
        // next fill of this buffer is h_{s+1}, consumed at step s+2; its senders are all behind this wait (see header)
        if (!GX && s + 2 < p.T) ptx::mbar_arrive_expect_tx(&a_full[b], FILL_TX);
        const uint32_t a_base = ptx::smem_u32(smem_a + b * C::A_BYTES);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const uint64_t a_desc = p.swap_ls ? ptx::make_desc_k_nosw(a_base + k * 4096, 128, 2048) : ptx::make_desc_k_nosw(a_base + k * 4096, 2048, 128);
          const uint64_t b_desc = ptx::make_desc_k_sw128(ptx::smem_u32(smem_b + (k >> 2) * NCOLS * 128)) + 2 * (k & 3);
          ptx::mma_f16_ss(tmem_base, a_desc, b_desc, IDESC, k != 0);
        }
        ptx::tc_commit(acc_full);
        LSTM_TRACE(4);
      }
    }
    __syncwarp();
  } else if (warp_idx >= 2) {
    // ===================== epilogue: thread = (sample row, 16 of the CTA's 32 units); cell state in registers =====================
    const int q = warp_idx & 3;
    const int hh = (warp_idx - 2) >> 2;               // which HALF-unit group of the CTA's units
    const int u0 = hh * HALF;
    const int row = q * 32 + lane;
    const int n = tile * BLOCK_M + row;
    const bool okn = n < p.Nimg;
    const int len = okn ? min(max(__ldg(p.seq_len + n), 0), p.T) : 0;
    float cst[HALF];
#pragma unroll
    for (int i = 0; i < HALF; ++i) cst[i] = 0.f;
    const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // this thread's HALF/8 x 16 B of the h slice: K-chunks u0/8 .. at chunk*2048 + row*16 inside the CTA's 8 KB slice
    const uint32_t slice_off = rank * C::SLICE_BYTES + (u0 / 8) * 2048 + row * 16;
    uint8_t* hx0 = reinterpret_cast<uint8_t*>(p.h_state) + ((size_t)unit * CS + rank) * C::SLICE_BYTES;
    const size_t hx_buf_stride = (size_t)2 * p.tiles_per_dir * CS * C::SLICE_BYTES;

    for (int s = 0; s < p.T; ++s) {
      const bool active = s < len;
      const int t = active ? (dir ? (len - 1 - s) : s) : s;
      // this step's input projection (row n, step s; bw rows were stored reversed by the projection GEMM): 4 gates x 16 units
      uint4 xp[4][HALF / 8];
      if (active) {
        const __nv_bfloat16* src = p.xproj + ((size_t)n * p.H + s) * 2048 + dir * 1024 + rank * NCOLS + u0;
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int v = 0; v < HALF / 8; ++v) xp[g][v] = __ldg(reinterpret_cast<const uint4*>(src + g * UPC) + v);
      }
      if (warp_idx == 2) LSTM_TRACE(5);
      uint32_t gi[HALF], gj[HALF], gf[HALF], go[HALF];
      if (s > 0) {
        ptx::mbar_wait(acc_full, (s - 1) & 1);
        ptx::tc_fence_after();
        if (warp_idx == 2) LSTM_TRACE(6);
        if constexpr (HALF == 16) {
          ptx::tmem_ld_32x32b_x16(tbase + 0 * UPC + u0, gi);
          ptx::tmem_ld_32x32b_x16(tbase + 1 * UPC + u0, gj);
          ptx::tmem_ld_32x32b_x16(tbase + 2 * UPC + u0, gf);
          ptx::tmem_ld_32x32b_x16(tbase + 3 * UPC + u0, go);
        } else {
          ptx::tmem_ld_32x32b_x8(tbase + 0 * UPC + u0, gi);
          ptx::tmem_ld_32x32b_x8(tbase + 1 * UPC + u0, gj);
          ptx::tmem_ld_32x32b_x8(tbase + 2 * UPC + u0, gf);
          ptx::tmem_ld_32x32b_x8(tbase + 3 * UPC + u0, go);
        }
        ptx::tmem_ld_wait();
        if (warp_idx == 2) LSTM_TRACE(7);
      } else {
#pragma unroll
        for (int i = 0; i < HALF; ++i) { gi[i] = 0u; gj[i] = 0u; gf[i] = 0u; go[i] = 0u; }
      }
      uint32_t hp[HALF / 2];
      if (active) {
        const uint32_t* xi = reinterpret_cast<const uint32_t*>(xp[0]);
        const uint32_t* xj = reinterpret_cast<const uint32_t*>(xp[1]);
        const uint32_t* xf = reinterpret_cast<const uint32_t*>(xp[2]);
        const uint32_t* xo = reinterpret_cast<const uint32_t*>(xp[3]);
        float hv[HALF];
#pragma unroll
        for (int i = 0; i < HALF; ++i) {
          const float zi = __uint_as_float(gi[i]) + ((i & 1) ? ptx::bf16_hi(xi[i >> 1]) : ptx::bf16_lo(xi[i >> 1]));
          const float zj = __uint_as_float(gj[i]) + ((i & 1) ? ptx::bf16_hi(xj[i >> 1]) : ptx::bf16_lo(xj[i >> 1]));
          const float zf = __uint_as_float(gf[i]) + ((i & 1) ? ptx::bf16_hi(xf[i >> 1]) : ptx::bf16_lo(xf[i >> 1]));
          const float zo = __uint_as_float(go[i]) + ((i & 1) ? ptx::bf16_hi(xo[i >> 1]) : ptx::bf16_lo(xo[i >> 1]));
          const float ai = ptx::fast_sigmoid(zi), aj = ptx::fast_tanh(zj), af = ptx::fast_sigmoid(zf), ao = ptx::fast_sigmoid(zo);
          const float c = af * cst[i] + ai * aj;                 // forget_bias (+1.0) is folded into the projected bias
          cst[i] = c;
          hv[i] = ao * ptx::fast_tanh(c);
          if (p.gates != nullptr) {                              // reuse the (now dead) accumulator registers as staging
            gi[i] = __float_as_uint(ai); gj[i] = __float_as_uint(aj); gf[i] = __float_as_uint(af); go[i] = __float_as_uint(ao);
          }
        }
#pragma unroll
        for (int i = 0; i < HALF / 2; ++i) hp[i] = ptx::pack_bf16x2(hv[2 * i], hv[2 * i + 1]);
      } else {
#pragma unroll
        for (int i = 0; i < HALF / 2; ++i) hp[i] = 0u;     // zero output past sequence_length (and for padding rows)
      }
      // ---- h_t slice first (it is on the critical path), bookkeeping stores afterwards
      if (s + 1 < p.T) {
        const int nb = (s + 1) & 1;
        if (LOCAL) {
          uint8_t* dst = smem_a + nb * C::A_BYTES + slice_off;
#pragma unroll
          for (int v = 0; v < HALF / 8; ++v)
            *reinterpret_cast<uint4*>(dst + v * 2048) = make_uint4(hp[4 * v], hp[4 * v + 1], hp[4 * v + 2], hp[4 * v + 3]);
          ptx::fence_proxy_async_smem();                    // generic-proxy smem writes -> async proxy (bulk copies, tcgen05.mma)
        } else {
          uint8_t* dst = hx0 + (size_t)nb * hx_buf_stride + (u0 / 8) * 2048 + row * 16;
#pragma unroll
          for (int v = 0; v < HALF / 8; ++v)
            *reinterpret_cast<uint4*>(dst + v * 2048) = make_uint4(hp[4 * v], hp[4 * v + 1], hp[4 * v + 2], hp[4 * v + 3]);
          if (!GX) fence_proxy_async_all();                 // generic-proxy global writes -> async proxy (bulk copy)
        }
        if (warp_idx == 2) LSTM_TRACE(8);
        ptx::tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(MC_EPI_THREADS) : "memory");
        if (warp_idx == 2) {
          uint8_t* my_slice = smem_a + nb * C::A_BYTES + rank * C::SLICE_BYTES;
          if (DS) {
            if (lane < CS) {
              if (lane == rank) {
                ptx::mbar_arrive(&a_full[nb]);
              } else {
                const uint32_t dst = ptx::mapa(ptx::smem_u32(my_slice), (uint32_t)lane);
                const uint32_t bar = ptx::mapa(ptx::smem_u32(&a_full[nb]), (uint32_t)lane);
                bulk_copy_s2s_cluster(dst, my_slice, C::SLICE_BYTES, bar);
              }
            }
          } else if (MS) {
            if (lane == 0) {
              uint8_t* g = hx0 + (size_t)nb * hx_buf_stride;
              ptx::bulk_store_1d(g, my_slice, C::SLICE_BYTES);                // async proxy: smem -> global (L2)
              ptx::bulk_commit();
              asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");       // writes performed, not just the smem reads
              ptx::bulk_load_1d_mc(my_slice, g, C::SLICE_BYTES, &a_full[nb], (uint16_t)(((1u << CS) - 1) & ~(1u << rank)));
              ptx::mbar_arrive(&a_full[nb]);                                  // the local slice is already in place
            }
          } else if (GX) {
            // cumulative over the barrier above: the release covers every epilogue thread's slice stores
            if (lane < CS) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&part_ready[nb]), (uint32_t)lane));
          } else if (lane == 0) {
            ptx::bulk_load_1d_mc(my_slice, hx0 + (size_t)nb * hx_buf_stride, C::SLICE_BYTES, &a_full[nb], (uint16_t)((1u << CS) - 1));
          }
          __syncwarp();
          LSTM_TRACE(9);
        }
        if (GX) {
          // h_t of the whole tile = the cluster's 8 contiguous slices = exactly the A operand image: 64 KB, 16 B per thread per pass
          ptx::mbar_wait_cluster(&part_ready[nb], (uint32_t)(s >> 1) & 1u);
          const uint8_t* g = reinterpret_cast<const uint8_t*>(p.h_state) + (size_t)nb * hx_buf_stride + (size_t)unit * CS * C::SLICE_BYTES;
          uint8_t* d = smem_a + nb * C::A_BYTES;
          const int et = threadIdx.x - 64;                       // 0 .. MC_EPI_THREADS-1
          uint4 v[C::A_BYTES / 16 / MC_EPI_THREADS];
#pragma unroll
          for (int k = 0; k < C::A_BYTES / 16 / MC_EPI_THREADS; ++k) v[k] = __ldcg(reinterpret_cast<const uint4*>(g) + k * MC_EPI_THREADS + et);
#pragma unroll
          for (int k = 0; k < C::A_BYTES / 16 / MC_EPI_THREADS; ++k) *(reinterpret_cast<uint4*>(d) + k * MC_EPI_THREADS + et) = v[k];
          ptx::fence_proxy_async_smem();
          ptx::mbar_arrive(&a_full[nb]);
        }
      }
      if (okn) {
        __nv_bfloat16* lo = p.lstm_out + ((size_t)n * p.H + t) * 512 + dir * 256 + rank * UPC + u0;
        if constexpr (HALF == 16) ptx::st_global_v8(lo, hp[0], hp[1], hp[2], hp[3], hp[4], hp[5], hp[6], hp[7]);
        else *reinterpret_cast<uint4*>(lo) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
      }
      if (active && p.gates != nullptr) {
        const size_t dts = (size_t)unit * p.T + s;                  // coalesced saved-state layout, common.cuh
        __nv_bfloat16* gs = p.gates + lstm_gate_off(dts, 0, rank * UPC + u0, row);
        float* cs = p.csave + lstm_c_off(dts, rank * UPC + u0, row);
#pragma unroll
        for (int i = 0; i < HALF; i += 8) {
          *reinterpret_cast<uint4*>(gs + 0 * LSTM_GATE_STRIDE + (i >> 3) * LSTM_GCHUNK_STRIDE) = make_uint4(ptx::pack_bf16x2(__uint_as_float(gi[i]), __uint_as_float(gi[i + 1])), ptx::pack_bf16x2(__uint_as_float(gi[i + 2]), __uint_as_float(gi[i + 3])), ptx::pack_bf16x2(__uint_as_float(gi[i + 4]), __uint_as_float(gi[i + 5])), ptx::pack_bf16x2(__uint_as_float(gi[i + 6]), __uint_as_float(gi[i + 7])));
          *reinterpret_cast<uint4*>(gs + 1 * LSTM_GATE_STRIDE + (i >> 3) * LSTM_GCHUNK_STRIDE) = make_uint4(ptx::pack_bf16x2(__uint_as_float(gj[i]), __uint_as_float(gj[i + 1])), ptx::pack_bf16x2(__uint_as_float(gj[i + 2]), __uint_as_float(gj[i + 3])), ptx::pack_bf16x2(__uint_as_float(gj[i + 4]), __uint_as_float(gj[i + 5])), ptx::pack_bf16x2(__uint_as_float(gj[i + 6]), __uint_as_float(gj[i + 7])));
          *reinterpret_cast<uint4*>(gs + 2 * LSTM_GATE_STRIDE + (i >> 3) * LSTM_GCHUNK_STRIDE) = make_uint4(ptx::pack_bf16x2(__uint_as_float(gf[i]), __uint_as_float(gf[i + 1])), ptx::pack_bf16x2(__uint_as_float(gf[i + 2]), __uint_as_float(gf[i + 3])), ptx::pack_bf16x2(__uint_as_float(gf[i + 4]), __uint_as_float(gf[i + 5])), ptx::pack_bf16x2(__uint_as_float(gf[i + 6]), __uint_as_float(gf[i + 7])));
          *reinterpret_cast<uint4*>(gs + 3 * LSTM_GATE_STRIDE + (i >> 3) * LSTM_GCHUNK_STRIDE) = make_uint4(ptx::pack_bf16x2(__uint_as_float(go[i]), __uint_as_float(go[i + 1])), ptx::pack_bf16x2(__uint_as_float(go[i + 2]), __uint_as_float(go[i + 3])), ptx::pack_bf16x2(__uint_as_float(go[i + 4]), __uint_as_float(go[i + 5])), ptx::pack_bf16x2(__uint_as_float(go[i + 6]), __uint_as_float(go[i + 7])));
        }
#pragma unroll
        for (int i = 0; i < HALF; i += 4) *reinterpret_cast<float4*>(cs + (i >> 2) * LSTM_CCHUNK_STRIDE) = make_float4(cst[i], cst[i + 1], cst[i + 2], cst[i + 3]);
      }
      if (warp_idx == 2) LSTM_TRACE(10);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  cluster_arrive_release();          // no CTA leaves (and frees its smem / mbarriers) while a peer may still write into it
  cluster_wait_acquire();
  if (warp_idx == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, NCOLS);
  }
}

}  // namespace lstm
