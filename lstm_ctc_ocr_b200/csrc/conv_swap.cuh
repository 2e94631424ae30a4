// conv2 + bias + ReLU + 2x2 max-pool (lib/networks/LSTM_train.py:26-27) with the GEMM operands SWAPPED:
//
//   D^T[128 out-channels x 256 positions, f32 in TMEM] = W[128 x K] (bf16, K-major) * X[256 positions x K]^T (bf16, K-major)
//
// Why: with Cout = 128 the position-major kernel (gemm.cuh, M = 128 positions, N = 128 channels) moves 8 KB of operands
// through shared memory per 128x128x16 MMA and cannot keep the tensor pipe fed (measured 874 TFLOP/s, pipe 43.6 %).  Putting
// the 128 channels on the M side lets N be 256 POSITIONS: 12 KB per 128x256x16 MMA -- the same smem bytes per flop as the
// Cout = 256 layers that run at 1450-1700 TFLOP/s.
//
// A tile = 16 H-rows x Wd(16) of one image = 256 positions; per K-block (= one 3x3 tap, Cin = 64) the producer issues two
// 4-D TMA boxes of 128 positions at (r-1, s-1)-shifted coordinates (OOB zero fill = SAME padding, also past the image end)
// and one 2-D box of the weights.  Accumulator lane = output channel, column = position, so the 2x2 pool is register-local
// in the epilogue thread (columns hl*16+w: window = {2pw, 2pw+1} x {row, row+1}); a warp stores 32 consecutive channels.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace convsw {

constexpr int STAGES = 4;
constexpr int W_BYTES = 128 * 128;          // weights: 128 channels x 64 K (128 B rows, SW128)
constexpr int X_BYTES = 256 * 128;          // activations: 256 positions x 64 K
constexpr int STAGE_BYTES = W_BYTES + X_BYTES;
constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
constexpr int SMEM_BYTES = BAR_OFFSET + 256 + 1024;
constexpr int NUM_THREADS = 320;            // warp 0 producer, warp 1 MMA, warps 2..9 epilogue
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_TAPS = 9;

struct Params {
  int Nimg, H;            // this launch covers images img0 .. img0+Nimg-1 of [*, H, 16, 64] NHWC (H = image width / 2) -> [*, H/2, 8, 128]
  int img0;
  int tiles_per_img;      // ceil(H / 16)
  const float* bias;      // [128]
  __nv_bfloat16* out;
  uint8_t* argmax;        // TRAIN: window index (dy*2+dx) of the max, same shape as out
};

template <bool TRAIN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv2_swap_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const Params p) {
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(128, 256);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.Nimg * p.tiles_per_img;

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmX);
    ptx::prefetch_tmap(&tmW);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      ptx::mbar_init(&tmem_empty[s], NUM_EPI_WARPS);
    }
    ptx::fence_barrier_init();
  }
  if (warp_idx == 1) {
    ptx::tmem_alloc(tmem_ptr, 512);           // two 256-column accumulators
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp_idx == 0) {
    // ===================== TMA producer: lanes 0/1 = the two activation boxes, lane 2 = the weight box =====================
    if (lane < 3) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int nl = tile / p.tiles_per_img, n = p.img0 + nl;
        const int h0 = (tile - nl * p.tiles_per_img) * 16;
        for (int tap = 0; tap < NUM_TAPS; ++tap) {
          const int r = tap / 3, sx = tap - 3 * r;
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * STAGE_BYTES;
          if (lane == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          if (lane < 2) ptx::tma_load_4d(&tmX, &full_bar[stage], st + W_BYTES + lane * (X_BYTES / 2), 0, sx - 1, h0 + lane * 8 + r - 1, n);
          else ptx::tma_load_2d(&tmW, &full_bar[stage], st, tap * 64, 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        ptx::mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int tap = 0; tap < NUM_TAPS; ++tap) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint64_t a_desc = ptx::make_desc_k_sw128(ptx::smem_u32(smem + stage * STAGE_BYTES));
          const uint64_t b_desc = ptx::make_desc_k_sw128(ptx::smem_u32(smem + stage * STAGE_BYTES + W_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, IDESC, (tap | k) != 0);
          ptx::tc_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::tc_commit(&tmem_full[acc]);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue: lane quadrant q = 32 channels, column half ch = 8 of the tile's 16 H-rows =====================
    const int q = warp_idx & 3;
    const int ch = (warp_idx - 2) >> 2;
    const int c = q * 32 + lane;                       // output channel of this thread
    const float bias = __ldg(p.bias + c);
    const int Hp = p.H >> 1;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int nl = tile / p.tiles_per_img, n = p.img0 + nl;
      const int h0 = (tile - nl * p.tiles_per_img) * 16;
      const int acc = it & 1;
      ptx::mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256 + ch * 128;
#pragma unroll 1
      for (int pr = 0; pr < 4; ++pr) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(tbase + pr * 32, v);   // columns: [row h (16 w) | row h+1 (16 w)]
        ptx::tmem_ld_wait();
        const int h = h0 + ch * 8 + 2 * pr;
        if (h < p.H) {                                  // H is even: both rows of the window are inside or outside together
          const size_t off = (((size_t)n * Hp + (h >> 1)) * 8) * 128 + c;
#pragma unroll
          for (int pw = 0; pw < 8; ++pw) {
            const float x00 = __uint_as_float(v[2 * pw]), x01 = __uint_as_float(v[2 * pw + 1]);
            const float x10 = __uint_as_float(v[16 + 2 * pw]), x11 = __uint_as_float(v[16 + 2 * pw + 1]);
            if (!TRAIN) {
              const float mx = fmaxf(fmaxf(x00, x01), fmaxf(x10, x11));
              p.out[off + (size_t)pw * 128] = __float2bfloat16_rn(fmaxf(mx + bias, 0.f));
            } else {
              // key = (bf16 bits of relu(x + b) << 2) | (3 - window index): the largest value wins, ties go to the FIRST
              // window position in (dy, dx) row-major order (same rule as gemm.cuh EPI_RELU_POOL22_T)
              const uint32_t p0 = ptx::pack_bf16x2(fmaxf(x00 + bias, 0.f), fmaxf(x01 + bias, 0.f));
              const uint32_t p1 = ptx::pack_bf16x2(fmaxf(x10 + bias, 0.f), fmaxf(x11 + bias, 0.f));
              const uint32_t k0 = ((p0 & 0xFFFFu) << 2) | 3u, k1 = ((p0 >> 16) << 2) | 2u;
              const uint32_t k2 = ((p1 & 0xFFFFu) << 2) | 1u, k3 = ((p1 >> 16) << 2) | 0u;
              const uint32_t k = max(max(k0, k1), max(k2, k3));
              reinterpret_cast<unsigned short*>(p.out)[off + (size_t)pw * 128] = (unsigned short)(k >> 2);
              p.argmax[off + (size_t)pw * 128] = (uint8_t)(3u - (k & 3u));
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}


// ---------------------------------------------------------------------------------------------------------------------------
// DATA gradients of the narrow layers (conv2: 64 input channels, conv3_1: 128) with the same operand swap (r2).  conv2:  d_a1[p, ci] = sum over taps, co of d_pre2[p + tap', co] * W'[ci][(tap', co)]
// is a 3x3 SAME convolution of the [N, H, 16, 128] gradient with the flipped / transposed kernel (Bd_c2, backward_kernels.cu):
// only 64 output channels.  Position-major (gemm.cuh, BLOCK_N = 64) it ran the MMA at N = 64, a quarter of the 128x256x16 rate
// (0.60 ms for 309 GFLOP).  Here the 64 channels sit on the M side (rows 64..127 of the weight box are out of bounds of the
// tensor map, i.e. zero-filled by TMA: half of the M = 128 MMA is padding) and N is 256 positions: half the padded work at the
// full rate.  K-blocks = 9 taps x 2 blocks of 64 gradient channels.  Epilogue: lane = input channel (quadrants 0 and 1 only),
// column = position; plain bf16 store (the ReLU / pool1 backward is folded into conv1's weight-gradient kernel).
// ---------------------------------------------------------------------------------------------------------------------------
// Templated on the geometry so that conv3_1's data gradient (Cin = 128: N = 128 position-major, half the MMA rate) takes the same
// route: WD = positions per H row (16 / 8), CB = 64-channel blocks of the incoming gradient (2 / 4), MVALID = output channels (64 / 128).
struct DgradParams {
  int Nimg, H;            // gradient [Nimg, H, WD, 64*CB] -> [Nimg, H, WD, MVALID]
  int tiles_per_img;      // ceil(H / (256 / WD))
  __nv_bfloat16* out;
};

template <int WD, int CB, int MVALID>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_dgrad_swap_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const DgradParams p) {
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(128, 256);
  constexpr int NUM_KB = 9 * CB;             // 9 taps x CB channel blocks
  constexpr int RT = 256 / WD;               // H rows per tile (two TMA boxes of RT/2 rows)
  constexpr int RC = 32 / WD;                // H rows per 32-column accumulator chunk
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.Nimg * p.tiles_per_img;

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmX);
    ptx::prefetch_tmap(&tmW);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      ptx::mbar_init(&tmem_empty[s], NUM_EPI_WARPS);
    }
    ptx::fence_barrier_init();
  }
  if (warp_idx == 1) {
    ptx::tmem_alloc(tmem_ptr, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp_idx == 0) {
    if (lane < 3) {                          // lanes 0/1: the two 128-position gradient boxes, lane 2: the weight box
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n = tile / p.tiles_per_img;
        const int h0 = (tile - n * p.tiles_per_img) * RT;
        for (int kb = 0; kb < NUM_KB; ++kb) {
          const int tap = kb / CB, cb = kb - tap * CB;
          const int r = tap / 3, sx = tap - 3 * r;
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * STAGE_BYTES;
          if (lane == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          if (lane < 2) ptx::tma_load_4d(&tmX, &full_bar[stage], st + W_BYTES + lane * (X_BYTES / 2), cb * 64, sx - 1, h0 + lane * (RT / 2) + r - 1, n);
          else ptx::tma_load_2d(&tmW, &full_bar[stage], st, kb * 64, 0);       // MVALID = 64: rows 64..127 out of bounds -> zero fill
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp_idx == 1) {
    if (lane == 0) {
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        ptx::mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kb = 0; kb < NUM_KB; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint64_t a_desc = ptx::make_desc_k_sw128(ptx::smem_u32(smem + stage * STAGE_BYTES));
          const uint64_t b_desc = ptx::make_desc_k_sw128(ptx::smem_u32(smem + stage * STAGE_BYTES + W_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, IDESC, (kb | k) != 0);
          ptx::tc_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::tc_commit(&tmem_full[acc]);
      }
    }
    __syncwarp();
  } else {
    const int q = warp_idx & 3;
    const int ch = (warp_idx - 2) >> 2;      // column half: RT/2 of the tile's RT H-rows
    const int c = q * 32 + lane;              // input channel of this thread (valid for c < MVALID)
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int n = tile / p.tiles_per_img;
      const int h0 = (tile - n * p.tiles_per_img) * RT;
      const int acc = it & 1;
      ptx::mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      ptx::tc_fence_after();
      if (q * 32 < MVALID) {
        const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256 + ch * 128;
#pragma unroll 1
        for (int pr = 0; pr < 4; ++pr) {
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(tbase + pr * 32, v);     // columns: RC consecutive H rows of WD positions each
          ptx::tmem_ld_wait();
#pragma unroll
          for (int hr = 0; hr < RC; ++hr) {
            const int h = h0 + ch * (RT / 2) + pr * RC + hr;
            if (h < p.H) {
              __nv_bfloat16* o = p.out + (((size_t)n * p.H + h) * WD) * MVALID + c;
#pragma unroll
              for (int w = 0; w < WD; ++w) o[(size_t)w * MVALID] = __float2bfloat16_rn(__uint_as_float(v[hr * WD + w]));
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace convsw

template <bool TRAIN>
static int launch_conv2_swap(const CUtensorMap& x, const CUtensorMap& w, const convsw::Params& p, int num_sms, cudaStream_t st) {
  auto kern = convsw::conv2_swap_kernel<TRAIN>;
  static bool attr = false;
  if (!attr) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, convsw::SMEM_BYTES));
    attr = true;
  }
  const int tiles = p.Nimg * p.tiles_per_img;
  const int grid = tiles < num_sms ? tiles : num_sms;
  kern<<<grid, convsw::NUM_THREADS, convsw::SMEM_BYTES, st>>>(x, w, p);
  CUDA_TRY(cudaGetLastError());
  return CRNN_OK;
}

template <int WD, int CB, int MVALID>
static int launch_conv_dgrad_swap(const CUtensorMap& x, const CUtensorMap& w, const convsw::DgradParams& p, int num_sms, cudaStream_t st) {
  auto kern = convsw::conv_dgrad_swap_kernel<WD, CB, MVALID>;
  static bool attr = false;
  if (!attr) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, convsw::SMEM_BYTES));
    attr = true;
  }
  const int tiles = p.Nimg * p.tiles_per_img;
  const int grid = tiles < num_sms ? tiles : num_sms;
  kern<<<grid, convsw::NUM_THREADS, convsw::SMEM_BYTES, st>>>(x, w, p);
  CUDA_TRY(cudaGetLastError());
  return CRNN_OK;
}
