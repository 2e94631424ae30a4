// Persistent warp-specialised tcgen05 GEMM for sm_100a with fused epilogues.
//
//   D[128 x BLOCK_N tile, f32 in TMEM] = A[rows x K] (bf16, K-major, TMA) * B[Nc x K]^T (bf16, K-major, TMA)
//
// Roles (320 threads): warp 0 = TMA producer (one lane per box), warp 1 = tcgen05.mma issuer (one lane,
// also owns the TMEM allocation), warps 2..9 = epilogue (TMEM lane quadrant = warp_idx % 4; two warps per quadrant,
// each draining half of the tile's columns -- the plain-store epilogues were the bottleneck of the short-K GEMMs).
// Pipelines: STAGES-deep smem ring (full/empty mbarriers, TMA <-> MMA) and a 2-deep TMEM accumulator
// ring (tmem_full/tmem_empty, MMA <-> epilogue) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// A-operand modes
//   A_PLAIN  rows are consecutive rows of a 2-D [rows, K] tensor map; conv5 (2x2 VALID over [N,H,2,512]) reads
//            row m for the first half of K and row m+1 for the second half (kb_per_shift).
//   A_CONV3  implicit GEMM for a 3x3 SAME convolution over an NHWC activation [N, H, Wd, C]: a tile is
//            4 sub-boxes of 32 output positions (bh = 32/Wd rows of H x full Wd); K-block kb = tap (r,s) x
//            64-channel block; the producer issues one 4-D TMA box per sub-box at coordinates shifted by
//            (r-1, s-1) -- out-of-bounds elements are zero-filled by TMA, which *is* the SAME padding.
//
// Epilogues: see enum Epi.  Every epilogue thread owns one accumulator row (TMEM lane).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace gemm {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                      // bf16 elements per K-block = one 128 B swizzle row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * 128;     // 16 KB
constexpr int NUM_THREADS = 320;                 // warp 0 producer, warp 1 MMA, warps 2..9 epilogue
constexpr int NUM_EPI_WARPS = 8;                 // two warps per TMEM lane quadrant, each draining half of the tile's columns
constexpr int EPI_WARP0 = 2;

enum AMode { A_PLAIN = 0, A_CONV3 = 1 };
enum Epi {
  EPI_F32 = 0,          // D f32 row-major [M, Nc] (tests)
  EPI_BIAS_BF16 = 1,    // + bias -> bf16 row-major [M, ldo]            (conv5, LSTM input projection)
  EPI_RELU = 2,         // conv: + bias, ReLU -> bf16 NHWC              (conv3_1)
  EPI_RELU_POOL22 = 3,  // conv: + bias, ReLU, 2x2/2 max-pool           (conv2 + pool2), needs Wd=16
  EPI_RELU_POOL12 = 4,  // conv: + bias, ReLU, max over Wd pairs        (conv3_2 + pool), needs Wd=8
  EPI_STATS = 5,        // conv: + bias -> bf16 pre-BN, per-channel sum / sum^2 (f64 atomics) (conv4_x)
  EPI_LSTM = 6,         // recurrent step: gates = acc + xproj; LSTM cell; writes h, c, output
  EPI_LOGITS = 7,       // + bias -> f32 time-major [T, N, 64]
  EPI_XPROJ = 8,        // + bias -> bf16 [N*H, 2048]; columns >= 1024 (backward direction) stored reversed-by-length
  EPI_CONV_STORE = 9,   // conv: plain bf16 NHWC store (data-gradient convolutions)
  EPI_RELU_POOL22_T = 10,  // training variants of the pooled epilogues: also emit the arg-max window index (uint8)
  EPI_RELU_POOL12_T = 11,
  EPI_CONV_STORE_MASK = 13,  // EPI_CONV_STORE with the ReLU backward of the PRODUCING layer folded in: zero where p.mask (its bf16 output, same NHWC layout) is 0
  EPI_CONV_STORE_BNRED = 14, // EPI_CONV_STORE + pass 1 of the BatchNorm/ReLU backward of the PRODUCING layer: per-channel f64 sums of the
                             // ReLU-masked gradient and of gradient * xhat (p.mask = its pre-BN bf16 output, p.bnp = scale|shift|mean|invstd)
  EPI_CONV_F32 = 12     // conv: raw f32 accumulators, NHWC store (f32-class path, forward_x3.cu: bias/BN/ReLU/pool + hi/lo split follow)
};

struct Params {
  int num_m_tiles, num_n_tiles, num_k_blocks;
  int m_tile0;           // first M tile of this launch (batch-chunked launches of the conv front end); tiles m_tile0 .. m_tile0+num_m_tiles-1
  int kb_per_shift;      // A_PLAIN: K-block kb reads A columns (kb % kb_per_shift)*64 of row (m + kb / kb_per_shift);
                         // == num_k_blocks for an ordinary GEMM; conv5 (2x2 VALID) uses 16 -> rows t and t+1
  int merged;            // conv: the tile's 4 sub-boxes are contiguous H rows of one image -> one 128-position TMA box
  int debug_skip_tma;    // probe only: producer arrives without loading (measures the MMA/epilogue ceiling)
  int row_shift_mul;     // +1 (conv5 forward: rows m, m+1) or -1 (conv5 data gradient: rows m, m-1)
  // split-bf16 ("3xbf16", f32-class) operands: the activation tensor stores [hi | lo] halves and the K loop visits
  // [hi | lo | hi] against weights [wh | wh | wl]; a virtual K-block index >= the fold wraps back onto the hi half.  0 = off.
  int cin_phys;          // A_CONV3: physical 64-channel blocks per tap (= 2*Cin/64 when cin_blocks = 3*Cin/64)
  int kb_phys;           // A_PLAIN: physical K-blocks per row shift (kb_per_shift counts the virtual ones)
  int M;                 // valid rows (plain modes)
  int Nc;                // total output columns
  // conv geometry (A_CONV3 and conv epilogues)
  int cin_blocks;        // Cin / 64
  int sb_per_img;        // ceil(H / bh)
  int bh, Wd, H, Nimg;
  // epilogue
  const float* bias;     // [Nc]
  void* out;             // primary output
  int ldo;               // row stride of `out` in elements (plain modes)
  double* stats;         // [2][Nc] (EPI_STATS)
  const __nv_bfloat16* mask;   // EPI_CONV_STORE_MASK: post-ReLU activation of the layer whose pre-activation gradient is being written
  const float* bnp;      // EPI_CONV_STORE_BNRED: [4][Nc] scale, shift, mean, invstd of the producing layer's BatchNorm
  uint8_t* argmax;       // pooled-epilogue training variants: window index of the max, same shape as `out`
  // EPI_LSTM
  const __nv_bfloat16* xproj;   // [Nimg*H, 2048] gate pre-activations (x part + bias), permuted columns
  float* c_state;               // [2][Npad][256]
  __nv_bfloat16* h_next;        // [2][Npad][256]
  __nv_bfloat16* lstm_out;      // [Nimg*H, 512]
  const int* seq_len;           // [Nimg]
  int step, Npad, m_tiles_per_dir, T;
};

__device__ __forceinline__ float warp_colsum32(const float (&v)[32], int lane) {
  // Sum over the 32 lanes of a warp for each of 32 per-lane registers; lane l returns column l's total.
  float r16[16], r8[8], r4[4], r2[2];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float send = (lane & 16) ? v[i] : v[i + 16];
    float keep = (lane & 16) ? v[i + 16] : v[i];
    r16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float send = (lane & 8) ? r16[i] : r16[i + 8];
    float keep = (lane & 8) ? r16[i + 8] : r16[i];
    r8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float send = (lane & 4) ? r8[i] : r8[i + 4];
    float keep = (lane & 4) ? r8[i + 4] : r8[i];
    r4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float send = (lane & 2) ? r4[i] : r4[i + 2];
    float keep = (lane & 2) ? r4[i + 2] : r4[i];
    r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  float send = (lane & 1) ? r2[0] : r2[1];
  float keep = (lane & 1) ? r2[1] : r2[0];
  return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

template <int BLOCK_N, int STAGES>
struct Smem {
  static constexpr int B_STAGE_BYTES = BLOCK_N * 128;
  static constexpr int BAR_OFFSET = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);
  static constexpr int BYTES = BAR_OFFSET + 256 + 1024;   // barriers + alignment slack
};

// Per-tile epilogue shared by the 1-CTA and 2-CTA kernels: thread (q, lane) owns accumulator row q*32+lane of the 128-row
// tile held in its CTA's TMEM at `tbase` (lane quadrant and accumulator stage already applied).
template <int BLOCK_N, int EPI>
__device__ __forceinline__ void run_epilogue(const Params& p, const uint32_t tbase, const int m_blk, const int n_blk, const int q,
                                             const int lane, const int c_lo = 0, const int c_hi = BLOCK_N) {
  const int row = q * 32 + lane;
  const int col0 = n_blk * BLOCK_N;

  // ---- conv row geometry (one sub-box of 32 positions per warp)
  int n_img = 0, h = 0, w = 0;
  bool valid = true;
  if (EPI == EPI_RELU || EPI == EPI_RELU_POOL22 || EPI == EPI_RELU_POOL12 || EPI == EPI_STATS || EPI == EPI_CONV_STORE ||
      EPI == EPI_RELU_POOL22_T || EPI == EPI_RELU_POOL12_T || EPI == EPI_CONV_F32 || EPI == EPI_CONV_STORE_MASK ||
      EPI == EPI_CONV_STORE_BNRED) {
    const int g = m_blk * 4 + q;
    n_img = g / p.sb_per_img;
    const int hb = g - n_img * p.sb_per_img;
    const int hl = lane / p.Wd;
    w = lane - hl * p.Wd;
    h = hb * p.bh + hl;
    valid = (n_img < p.Nimg) && (h < p.H);
  }

  if (EPI == EPI_F32) {
    const int grow = m_blk * BLOCK_M + row;
    float* out = reinterpret_cast<float*>(p.out) + (size_t)grow * p.Nc + col0;
#pragma unroll 1
    for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(tbase + c0, v);
      ptx::tmem_ld_wait();
      if (grow < p.M) {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<uint4*>(out + c0 + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    }
  } else if (EPI == EPI_BIAS_BF16 || EPI == EPI_XPROJ) {
    const int grow = m_blk * BLOCK_M + row;
    int drow = grow;
    if (EPI == EPI_XPROJ && col0 >= 1024 && grow < p.M) {
      // tf.reverse_sequence(len) on the backward direction's input, done once at write time
      const int n = grow / p.H, t = grow - n * p.H;
      const int len = min(max(__ldg(p.seq_len + n), 0), p.T);
      if (t < len) drow = n * p.H + (len - 1 - t);
    }
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)drow * p.ldo + col0;
#pragma unroll 1
    for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(tbase + c0, v);
      ptx::tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 b = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + col0 + c0 + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
        pk[i / 2] = ptx::pack_bf16x2(__uint_as_float(v[i]) + b.x, __uint_as_float(v[i + 1]) + b.y);
        pk[i / 2 + 1] = ptx::pack_bf16x2(__uint_as_float(v[i + 2]) + b.z, __uint_as_float(v[i + 3]) + b.w);
      }
      if (grow < p.M) {
#pragma unroll
        for (int i = 0; i < 16; i += 8)
          ptx::st_global_v8(out + c0 + 2 * i, pk[i], pk[i + 1], pk[i + 2], pk[i + 3], pk[i + 4], pk[i + 5], pk[i + 6], pk[i + 7]);
      }
    }
  } else if (EPI == EPI_LOGITS) {
    const int grow = m_blk * BLOCK_M + row;
    const int n = grow / p.H, t = grow - n * p.H;
    const bool ok = (grow < p.M) && (t < p.T);
    float* out = reinterpret_cast<float*>(p.out) + ((size_t)t * p.Nimg + n) * p.Nc + col0;
#pragma unroll 1
    for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(tbase + c0, v);
      ptx::tmem_ld_wait();
      if (ok) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + c0 + i));
          *reinterpret_cast<float4*>(out + c0 + i) =
              make_float4(__uint_as_float(v[i]) + b.x, __uint_as_float(v[i + 1]) + b.y,
                          __uint_as_float(v[i + 2]) + b.z, __uint_as_float(v[i + 3]) + b.w);
        }
      }
    }
  } else if (EPI == EPI_RELU || EPI == EPI_RELU_POOL22 || EPI == EPI_RELU_POOL12) {
    __nv_bfloat16* outb = reinterpret_cast<__nv_bfloat16*>(p.out);
    size_t off;
    if (EPI == EPI_RELU) off = (((size_t)n_img * p.H + h) * p.Wd + w) * p.Nc;
    else if (EPI == EPI_RELU_POOL22) off = (((size_t)n_img * (p.H >> 1) + (h >> 1)) * (p.Wd >> 1) + (w >> 1)) * p.Nc;
    else off = (((size_t)n_img * p.H + h) * (p.Wd >> 1) + (w >> 1)) * p.Nc;
    __nv_bfloat16* out = outb + off + col0;
#pragma unroll 1
    for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(tbase + c0, v);
      ptx::tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + c0 + i));
        pk[i / 2] = ptx::pack_bf16x2(fmaxf(__uint_as_float(v[i]) + b.x, 0.f), fmaxf(__uint_as_float(v[i + 1]) + b.y, 0.f));
        pk[i / 2 + 1] = ptx::pack_bf16x2(fmaxf(__uint_as_float(v[i + 2]) + b.z, 0.f), fmaxf(__uint_as_float(v[i + 3]) + b.w, 0.f));
      }
      if (EPI == EPI_RELU) {
        if (valid) {
#pragma unroll
          for (int i = 0; i < 16; i += 8)
            ptx::st_global_v8(out + c0 + 2 * i, pk[i], pk[i + 1], pk[i + 2], pk[i + 3], pk[i + 4], pk[i + 5], pk[i + 6], pk[i + 7]);
        }
      } else if (EPI == EPI_RELU_POOL22) {
        // lane = hl*16 + w : partners lane^1 (w pair) and lane^16 (h pair); rounding to bf16 is monotonic,
        // so max after packing == packing after max
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          pk[i] = ptx::hmax2_bf16(pk[i], __shfl_xor_sync(0xffffffffu, pk[i], 1));
          pk[i] = ptx::hmax2_bf16(pk[i], __shfl_xor_sync(0xffffffffu, pk[i], 16));
        }
        const int sub = (lane & 1) | ((lane >> 3) & 2);     // which quarter of the 32 columns this lane stores
        uint4 o;
        o.x = sub == 0 ? pk[0] : sub == 1 ? pk[4] : sub == 2 ? pk[8] : pk[12];
        o.y = sub == 0 ? pk[1] : sub == 1 ? pk[5] : sub == 2 ? pk[9] : pk[13];
        o.z = sub == 0 ? pk[2] : sub == 1 ? pk[6] : sub == 2 ? pk[10] : pk[14];
        o.w = sub == 0 ? pk[3] : sub == 1 ? pk[7] : sub == 2 ? pk[11] : pk[15];
        if (valid) *reinterpret_cast<uint4*>(out + c0 + 8 * sub) = o;
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = ptx::hmax2_bf16(pk[i], __shfl_xor_sync(0xffffffffu, pk[i], 1));
        const int sub = lane & 1;
        uint4 o0, o1;
        o0.x = sub ? pk[8] : pk[0];  o0.y = sub ? pk[9] : pk[1];  o0.z = sub ? pk[10] : pk[2]; o0.w = sub ? pk[11] : pk[3];
        o1.x = sub ? pk[12] : pk[4]; o1.y = sub ? pk[13] : pk[5]; o1.z = sub ? pk[14] : pk[6]; o1.w = sub ? pk[15] : pk[7];
        if (valid) {
          *reinterpret_cast<uint4*>(out + c0 + 16 * sub) = o0;
          *reinterpret_cast<uint4*>(out + c0 + 16 * sub + 8) = o1;
        }
      }
    }
  } else if (EPI == EPI_CONV_STORE) {
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out) + (((size_t)n_img * p.H + h) * p.Wd + w) * p.Nc + col0;
#pragma unroll 1
    for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(tbase + c0, v);
      ptx::tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int i = 0; i < 32; i += 16)
          ptx::st_global_v8(out + c0 + i,
                            ptx::pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])),
                            ptx::pack_bf16x2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])),
                            ptx::pack_bf16x2(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])),
                            ptx::pack_bf16x2(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])),
                            ptx::pack_bf16x2(__uint_as_float(v[i + 8]), __uint_as_float(v[i + 9])),
                            ptx::pack_bf16x2(__uint_as_float(v[i + 10]), __uint_as_float(v[i + 11])),
                            ptx::pack_bf16x2(__uint_as_float(v[i + 12]), __uint_as_float(v[i + 13])),
                            ptx::pack_bf16x2(__uint_as_float(v[i + 14]), __uint_as_float(v[i + 15])));
      }
    }
  } else if (EPI == EPI_CONV_STORE_MASK) {
    const size_t off = (((size_t)n_img * p.H + h) * p.Wd + w) * p.Nc + col0;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out) + off;
    const __nv_bfloat16* msk = p.mask + off;
#pragma unroll 1
    for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
      uint4 mk[4];
      if (valid) {
#pragma unroll
        for (int i = 0; i < 4; ++i) mk[i] = __ldg(reinterpret_cast<const uint4*>(msk + c0) + i);     // issued before the TMEM wait
      }
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(tbase + c0, v);
      ptx::tmem_ld_wait();
      if (valid) {
        const uint32_t* mw = reinterpret_cast<const uint32_t*>(mk);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          // the activation is post-ReLU (>= 0): "> 0" is "bf16 bits, sign aside, non-zero"; rounding then masking == masking then rounding
          const uint32_t keep = ((mw[i] & 0x7FFFu) ? 0xFFFFu : 0u) | ((mw[i] & 0x7FFF0000u) ? 0xFFFF0000u : 0u);
          pk[i] = ptx::pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])) & keep;
        }
#pragma unroll
        for (int i = 0; i < 16; i += 8)
          ptx::st_global_v8(out + c0 + 2 * i, pk[i], pk[i + 1], pk[i + 2], pk[i + 3], pk[i + 4], pk[i + 5], pk[i + 6], pk[i + 7]);
      }
    }
  } else if (EPI == EPI_CONV_STORE_BNRED) {
    // same sums as bn_bwd_reduce_kernel<false> (backward_kernels.cu), taken on the bf16-rounded gradient the apply pass will read
    const size_t off = (((size_t)n_img * p.H + h) * p.Wd + w) * p.Nc + col0;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out) + off;
    const __nv_bfloat16* xp = p.mask + off;
#pragma unroll 1
    for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
      uint4 xq[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xq[i] = valid ? __ldg(reinterpret_cast<const uint4*>(xp + c0) + i) : make_uint4(0u, 0u, 0u, 0u);
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(tbase + c0, v);
      ptx::tmem_ld_wait();
      const uint32_t* xw = reinterpret_cast<const uint32_t*>(xq);
      uint32_t pk[16];
      float f[32], f2[32];
      const float* bc = p.bnp + col0 + c0;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(bc + i));
        const float4 sh = __ldg(reinterpret_cast<const float4*>(bc + p.Nc + i));
        const float4 mu = __ldg(reinterpret_cast<const float4*>(bc + 2 * p.Nc + i));
        const float4 is = __ldg(reinterpret_cast<const float4*>(bc + 3 * p.Nc + i));
        pk[i / 2] = ptx::pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1]));
        pk[i / 2 + 1] = ptx::pack_bf16x2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
        const float x0 = ptx::bf16_lo(xw[i / 2]), x1 = ptx::bf16_hi(xw[i / 2]), x2 = ptx::bf16_lo(xw[i / 2 + 1]), x3 = ptx::bf16_hi(xw[i / 2 + 1]);
        const float d0 = (valid && fmaf(x0, sc.x, sh.x) > 0.f) ? ptx::bf16_lo(pk[i / 2]) : 0.f;
        const float d1 = (valid && fmaf(x1, sc.y, sh.y) > 0.f) ? ptx::bf16_hi(pk[i / 2]) : 0.f;
        const float d2 = (valid && fmaf(x2, sc.z, sh.z) > 0.f) ? ptx::bf16_lo(pk[i / 2 + 1]) : 0.f;
        const float d3 = (valid && fmaf(x3, sc.w, sh.w) > 0.f) ? ptx::bf16_hi(pk[i / 2 + 1]) : 0.f;
        f[i] = d0; f[i + 1] = d1; f[i + 2] = d2; f[i + 3] = d3;
        f2[i] = d0 * (x0 - mu.x) * is.x; f2[i + 1] = d1 * (x1 - mu.y) * is.y; f2[i + 2] = d2 * (x2 - mu.z) * is.z; f2[i + 3] = d3 * (x3 - mu.w) * is.w;
      }
      if (valid) {
#pragma unroll
        for (int i = 0; i < 16; i += 8)
          ptx::st_global_v8(out + c0 + 2 * i, pk[i], pk[i + 1], pk[i + 2], pk[i + 3], pk[i + 4], pk[i + 5], pk[i + 6], pk[i + 7]);
      }
      const float s1 = warp_colsum32(f, lane);
      const float s2 = warp_colsum32(f2, lane);
      atomicAdd(p.stats + col0 + c0 + lane, (double)s1);
      atomicAdd(p.stats + p.Nc + col0 + c0 + lane, (double)s2);
    }
  } else if (EPI == EPI_CONV_F32) {
    float* out = reinterpret_cast<float*>(p.out) + (((size_t)n_img * p.H + h) * p.Wd + w) * p.Nc + col0;
#pragma unroll 1
    for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(tbase + c0, v);
      ptx::tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<uint4*>(out + c0 + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    }
  } else if (EPI == EPI_RELU_POOL22_T || EPI == EPI_RELU_POOL12_T) {
    // Training variants: max over the pooling window carried as an integer key
    //   key = (bf16 bits of relu(x) << 2) | (3 - window_index)
    // post-ReLU bf16 bit patterns are monotone as unsigned integers, so max(key) picks the largest value and, among
    // equal values, the FIRST window position (row-major (dy,dx), the tie-break of TF/torch max-pool gradients).
    constexpr bool P22 = (EPI == EPI_RELU_POOL22_T);
    size_t off;
    if (P22) off = (((size_t)n_img * (p.H >> 1) + (h >> 1)) * (p.Wd >> 1) + (w >> 1)) * p.Nc;
    else off = (((size_t)n_img * p.H + h) * (p.Wd >> 1) + (w >> 1)) * p.Nc;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out) + off + col0;
    uint8_t* amx = p.argmax + off + col0;
    const uint32_t kidx = P22 ? (uint32_t)((((lane >> 4) & 1) << 1) | (lane & 1)) : (uint32_t)(lane & 1);
    const uint32_t kinv = (P22 ? 3u : 1u) - kidx;
#pragma unroll 1
    for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(tbase + c0, v);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + c0 + i));
        const uint32_t p0 = ptx::pack_bf16x2(fmaxf(__uint_as_float(v[i]) + b.x, 0.f), fmaxf(__uint_as_float(v[i + 1]) + b.y, 0.f));
        const uint32_t p1 = ptx::pack_bf16x2(fmaxf(__uint_as_float(v[i + 2]) + b.z, 0.f), fmaxf(__uint_as_float(v[i + 3]) + b.w, 0.f));
        v[i] = ((p0 & 0xFFFFu) << 2) | kinv;
        v[i + 1] = ((p0 >> 16) << 2) | kinv;
        v[i + 2] = ((p1 & 0xFFFFu) << 2) | kinv;
        v[i + 3] = ((p1 >> 16) << 2) | kinv;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        v[i] = max(v[i], __shfl_xor_sync(0xffffffffu, v[i], 1));
        if (P22) v[i] = max(v[i], __shfl_xor_sync(0xffffffffu, v[i], 16));
      }
      // each lane of the window stores its share of the 32 columns: 8 (2x2 window) or 16 (1x2 window)
      constexpr int NS = P22 ? 4 : 2, PER = 32 / NS;
      const int sub = P22 ? ((lane & 1) | ((lane >> 3) & 2)) : (lane & 1);
      uint32_t sel[PER];
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        uint32_t x = v[i];
#pragma unroll
        for (int j = 1; j < NS; ++j) x = (sub == j) ? v[j * PER + i] : x;
        sel[i] = x;
      }
      if (valid) {
#pragma unroll
        for (int i = 0; i < PER; i += 8) {
          uint4 o;
          o.x = ((sel[i] >> 2) & 0xFFFFu) | ((sel[i + 1] >> 2) << 16);
          o.y = ((sel[i + 2] >> 2) & 0xFFFFu) | ((sel[i + 3] >> 2) << 16);
          o.z = ((sel[i + 4] >> 2) & 0xFFFFu) | ((sel[i + 5] >> 2) << 16);
          o.w = ((sel[i + 6] >> 2) & 0xFFFFu) | ((sel[i + 7] >> 2) << 16);
          *reinterpret_cast<uint4*>(out + c0 + sub * PER + i) = o;
          const uint32_t km = P22 ? 3u : 1u;
          uint2 a;
          a.x = (km - (sel[i] & km)) | ((km - (sel[i + 1] & km)) << 8) | ((km - (sel[i + 2] & km)) << 16) | ((km - (sel[i + 3] & km)) << 24);
          a.y = (km - (sel[i + 4] & km)) | ((km - (sel[i + 5] & km)) << 8) | ((km - (sel[i + 6] & km)) << 16) | ((km - (sel[i + 7] & km)) << 24);
          *reinterpret_cast<uint2*>(amx + c0 + sub * PER + i) = a;
        }
      }
    }
  } else if (EPI == EPI_STATS) {
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out) + (((size_t)n_img * p.H + h) * p.Wd + w) * p.Nc + col0;
#pragma unroll 1
    for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(tbase + c0, v);
      ptx::tmem_ld_wait();
      float f[32], f2[32];
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + c0 + i));
        f[i] = __uint_as_float(v[i]) + b.x;
        f[i + 1] = __uint_as_float(v[i + 1]) + b.y;
        f[i + 2] = __uint_as_float(v[i + 2]) + b.z;
        f[i + 3] = __uint_as_float(v[i + 3]) + b.w;
        pk[i / 2] = ptx::pack_bf16x2(f[i], f[i + 1]);
        pk[i / 2 + 1] = ptx::pack_bf16x2(f[i + 2], f[i + 3]);
      }
      if (valid) {
#pragma unroll
        for (int i = 0; i < 16; i += 8)
          ptx::st_global_v8(out + c0 + 2 * i, pk[i], pk[i + 1], pk[i + 2], pk[i + 3], pk[i + 4], pk[i + 5], pk[i + 6], pk[i + 7]);
      }
      // statistics of the values the next layer will actually read (bf16-rounded), masked to valid rows
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a = valid ? ptx::bf16_lo(pk[i]) : 0.f, b = valid ? ptx::bf16_hi(pk[i]) : 0.f;
        f[2 * i] = a; f[2 * i + 1] = b;
        f2[2 * i] = a * a; f2[2 * i + 1] = b * b;
      }
      const float s1 = warp_colsum32(f, lane);
      const float s2 = warp_colsum32(f2, lane);
      atomicAdd(p.stats + col0 + c0 + lane, (double)s1);
      atomicAdd(p.stats + p.Nc + col0 + c0 + lane, (double)s2);
    }
  } else if (EPI == EPI_LSTM) {
    // row = sample within the direction-stacked batch; tile columns = [i(64) j(64) f(64) o(64)] of 64 units
    const int grow = m_blk * BLOCK_M + row;
    const int dir = (m_blk >= p.m_tiles_per_dir) ? 1 : 0;
    const int n = grow - dir * p.Npad;
    const bool okn = n < p.Nimg;
    const int len = okn ? min(max(__ldg(p.seq_len + n), 0), p.T) : 0;
    const bool active = p.step < len;
    const int t = active ? (dir ? (len - 1 - p.step) : p.step) : p.step;
    const size_t rt = (size_t)n * p.H + t;
    const __nv_bfloat16* xp = p.xproj + rt * 2048 + dir * 1024 + n_blk * 256;
    float* cst = p.c_state + ((size_t)dir * p.Npad + n) * 256 + n_blk * 64;
    __nv_bfloat16* hn = p.h_next + ((size_t)dir * p.Npad + n) * 256 + n_blk * 64;
    __nv_bfloat16* lo = p.lstm_out + rt * 512 + dir * 256 + n_blk * 64;
#pragma unroll 1
    for (int u0 = (c_lo >> 2); u0 < (c_hi >> 2); u0 += 16) {
      uint32_t gi[16], gj[16], gf[16], go[16];
      ptx::tmem_ld_32x32b_x16(tbase + u0, gi);
      ptx::tmem_ld_32x32b_x16(tbase + 64 + u0, gj);
      ptx::tmem_ld_32x32b_x16(tbase + 128 + u0, gf);
      ptx::tmem_ld_32x32b_x16(tbase + 192 + u0, go);
      ptx::tmem_ld_wait();
      uint32_t hp[8];
      if (active) {
        uint4 xi[2], xj[2], xf[2], xo[2];
        float4 cp[4];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          xi[i] = __ldg(reinterpret_cast<const uint4*>(xp + u0) + i);
          xj[i] = __ldg(reinterpret_cast<const uint4*>(xp + 64 + u0) + i);
          xf[i] = __ldg(reinterpret_cast<const uint4*>(xp + 128 + u0) + i);
          xo[i] = __ldg(reinterpret_cast<const uint4*>(xp + 192 + u0) + i);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) cp[i] = *(reinterpret_cast<const float4*>(cst + u0) + i);
        const uint32_t* xiw = reinterpret_cast<const uint32_t*>(xi);
        const uint32_t* xjw = reinterpret_cast<const uint32_t*>(xj);
        const uint32_t* xfw = reinterpret_cast<const uint32_t*>(xf);
        const uint32_t* xow = reinterpret_cast<const uint32_t*>(xo);
        float* cpf = reinterpret_cast<float*>(cp);
        float hv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float zi = __uint_as_float(gi[i]) + ((i & 1) ? ptx::bf16_hi(xiw[i >> 1]) : ptx::bf16_lo(xiw[i >> 1]));
          const float zj = __uint_as_float(gj[i]) + ((i & 1) ? ptx::bf16_hi(xjw[i >> 1]) : ptx::bf16_lo(xjw[i >> 1]));
          const float zf = __uint_as_float(gf[i]) + ((i & 1) ? ptx::bf16_hi(xfw[i >> 1]) : ptx::bf16_lo(xfw[i >> 1]));
          const float zo = __uint_as_float(go[i]) + ((i & 1) ? ptx::bf16_hi(xow[i >> 1]) : ptx::bf16_lo(xow[i >> 1]));
          // forget_bias (+1.0) is folded into the projected bias at weight-prep time
          const float c = ptx::fast_sigmoid(zf) * cpf[i] + ptx::fast_sigmoid(zi) * ptx::fast_tanh(zj);
          cpf[i] = c;
          hv[i] = ptx::fast_sigmoid(zo) * ptx::fast_tanh(c);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) *(reinterpret_cast<float4*>(cst + u0) + i) = cp[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) hp[i] = ptx::pack_bf16x2(hv[2 * i], hv[2 * i + 1]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) hp[i] = 0u;    // zero output past sequence_length; state no longer used
      }
      if (okn) {
        *reinterpret_cast<uint4*>(hn + u0) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
        *reinterpret_cast<uint4*>(hn + u0 + 8) = make_uint4(hp[4], hp[5], hp[6], hp[7]);
        *reinterpret_cast<uint4*>(lo + u0) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
        *reinterpret_cast<uint4*>(lo + u0 + 8) = make_uint4(hp[4], hp[5], hp[6], hp[7]);
      }
    }
  }

}

// KIND 0: bf16 operands (kind::f16), 64 elements per 128 B K-block.  KIND 1: f32 words read as tf32 (kind::tf32), 32 elements per
// K-block (forward_x3.cu, compute_dtype 3); tensor maps are FLOAT32 with 32-element boxes, everything else is shared.
template <int BLOCK_N, int AMODE, int EPI, int STAGES, int KIND = 0>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
  static_assert(BLOCK_N == 64 || BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N");
  constexpr int B_STAGE_BYTES = BLOCK_N * 128;
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;        // power of two >= 32
  constexpr uint32_t IDESC = KIND ? ptx::make_idesc_tf32(BLOCK_M, BLOCK_N) : ptx::make_idesc_bf16(BLOCK_M, BLOCK_N);
  constexpr int KELEMS = KIND ? 32 : BLOCK_K;        // operand elements per 128 B K-block

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Smem<BLOCK_N, STAGES>::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      ptx::mbar_init(&tmem_empty[s], NUM_EPI_WARPS);   // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp_idx == 1) {
    ptx::tmem_alloc(tmem_ptr, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    // One lane per TMA box: a single thread issuing 5 boxes per K-block was the bottleneck of the conv mainloop
    // (~200 cycles per cp.async.bulk.tensor issue vs 512 MMA cycles per K-block).  Lanes 0..nA-1 load the A sub-boxes,
    // lane nA loads B; lane 0 also arms the transaction count.  conv: when the tile's 4 sub-boxes are contiguous rows of
    // one image (p.merged), a single 128-position box replaces them.
    const int nA = (AMODE == A_CONV3 && !p.merged) ? 4 : 1;
    if (lane <= nA) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = p.m_tile0 + tile / p.num_n_tiles, n_blk = tile % p.num_n_tiles;
        int b_row = n_blk * BLOCK_N;
        if (EPI == EPI_LSTM) b_row += (m_blk >= p.m_tiles_per_dir) ? 1024 : 0;
        // per-tile coordinates of this lane's A box (conv): image n, first H row h0
        int cn = 0, ch0 = 0;
        if (AMODE == A_CONV3 && lane < nA) {
          const int g = m_blk * 4 + lane;
          cn = g / p.sb_per_img;
          ch0 = (g - cn * p.sb_per_img) * p.bh;
        }
        int tap = 0, cb = 0;                    // conv K-block decomposition, advanced incrementally
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          if (p.debug_skip_tma) {
            if (lane == 0) ptx::mbar_arrive(&full_bar[stage]);
          } else {
            if (lane == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], A_STAGE_BYTES + B_STAGE_BYTES);
            if (lane < nA) {
              uint8_t* a_dst = smem_a + stage * A_STAGE_BYTES;
              if (AMODE == A_PLAIN) {
                const int rs = kb / p.kb_per_shift;
                int kc = kb - rs * p.kb_per_shift;
                if (p.kb_phys > 0 && kc >= p.kb_phys) kc -= p.kb_phys;
                ptx::tma_load_2d(&tmA, &full_bar[stage], a_dst, kc * KELEMS, m_blk * BLOCK_M + rs * p.row_shift_mul);
              } else {
                const int r = tap / 3, sx = tap - 3 * r;
                const int cbp = (p.cin_phys > 0 && cb >= p.cin_phys) ? cb - p.cin_phys : cb;
                ptx::tma_load_4d(&tmA, &full_bar[stage], a_dst + lane * 4096, cbp * KELEMS, sx - 1, ch0 + r - 1, cn);
              }
            } else {
              ptx::tma_load_2d(&tmB, &full_bar[stage], smem_b + stage * B_STAGE_BYTES, kb * KELEMS, b_row);
            }
          }
          if (++cb == p.cin_blocks) { cb = 0; ++tap; }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint64_t a_desc = ptx::make_desc_k_sw128(ptx::smem_u32(smem_a + stage * A_STAGE_BYTES));
          const uint64_t b_desc = ptx::make_desc_k_sw128(ptx::smem_u32(smem_b + stage * B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 bf16 (8 tf32) = 32 B along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
            if (KIND) ptx::mma_tf32_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, IDESC, (kb | k) != 0);
            else ptx::mma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, IDESC, (kb | k) != 0);
          }
          ptx::tc_commit(&empty_bar[stage]);         // frees the smem slot once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::tc_commit(&tmem_full[acc]);             // accumulator complete -> epilogue
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps =====================
    const int q = warp_idx & 3;                      // TMEM lane quadrant accessible to this warp
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = p.m_tile0 + tile / p.num_n_tiles, n_blk = tile % p.num_n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N;
      const int chalf = (warp_idx - 2) >> 2;           // warps 2..5 take columns [0, N/2), warps 6..9 take [N/2, N)
      run_epilogue<BLOCK_N, EPI>(p, tbase, m_blk, n_blk, q, lane, chalf * (BLOCK_N / 2), (chalf + 1) * (BLOCK_N / 2));
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace gemm
