// CTC loss (+ gradient) and greedy decode for C = 64 classes, sm_100a.
//
// Replaces warpctc_tensorflow.ctc at lib/networks/network.py:653-654 and the decode at
// lib/networks/network.py:656-657 (+ zero stripping, lib/lstm/utils/training.py:32).
//
// ctc_loss_kernel: one CTA (4 warps) per utterance; warps 0/1 run the two recursions, all 4 share the frame-parallel phases.
//   phase 0  both warps, rows interleaved: log2-softmax normaliser per frame and the S <= 32*KS
//            emission scores e[t][s] = log2 y_t(l'_s), kept in shared memory (HBM read #1, coalesced 256 B rows)
//   phase 1  warp 0 runs the alpha recursion forward while warp 1 runs the beta recursion backward --
//            the two 63-step dependency chains overlap; each step is a warp-shuffle scan over the states
//            (lane owns KS consecutive states, neighbours via __shfl_up/__shfl_down), all in log2 space
//   phase 2  both warps, rows interleaved: re-read the logits row (L2 hit), y = softmax, per-class
//            sum of alpha*beta/y via shared-memory accumulators, write grad row (HBM write, coalesced)
// No tensor cores: the dynamic program is a scan, not a contraction.
#include "common.cuh"

namespace {

constexpr int CTC_C = 64;
constexpr int CTC_WARPS = 4;                 // phases 0/2 are parallel over frames: 4 warps x 16 rows covers T <= 64 in one batch
constexpr int CTC_THREADS = 32 * CTC_WARPS;
constexpr int CTC_RB = 16;                   // rows per warp per batch: all 16 loads are issued before any use (MLP)
constexpr int CTC_RG = 4;                    // rows reduced in lockstep (ILP) within a batch
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
#define NEG_INF (-INFINITY)

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// log2(2^a + 2^b + 2^c) with -inf handling
__device__ __forceinline__ float lse3(float a, float b, float c) {
  float m = fmaxf(a, fmaxf(b, c));
  if (m == NEG_INF) return NEG_INF;
  return m + ptx::lg2(ptx::ex2(a - m) + ptx::ex2(b - m) + ptx::ex2(c - m));
}

template <int KS>
__global__ void __launch_bounds__(CTC_THREADS, 8) ctc_loss_kernel(const float* __restrict__ logits, float* __restrict__ grad,
                                                      const int* __restrict__ flat_labels,
                                                      const int* __restrict__ label_len,
                                                      const int* __restrict__ input_len, int T, int N, int blank,
                                                      float grad_scale, float* __restrict__ costs) {
  constexpr int SP = 32 * KS;  // padded state count
  extern __shared__ float sm[];
  float* s_lse = sm;                    // [T]
  float* s_e = s_lse + T;               // [T][SP]
  float* s_alpha = s_e + (size_t)T * SP;
  float* s_beta = s_alpha + (size_t)T * SP;
  float* s_acc = s_beta + (size_t)T * SP;   // [4 warps][4 rows][64]
  __shared__ int s_off;
  __shared__ int s_ext[SP];
  __shared__ int s_repeats;

  const int n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = label_len[n];
  int Tn = input_len[n];
  Tn = max(0, min(Tn, T));
  const int S = 2 * L + 1;

  // label offset = sum(label_len[0..n))
  if (warp == 0) {
    int acc = 0;
    for (int i = lane; i < n; i += 32) acc += label_len[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_off = acc;
  }
  __syncthreads();
  const bool too_long = (S > SP) || (L < 0);
  if (!too_long) {
    int rep = 0;
    for (int s = threadIdx.x; s < SP; s += CTC_THREADS) {
      int v = blank;
      if (s < S && (s & 1)) {
        v = flat_labels[s_off + (s >> 1)];
        if (s >= 3 && v == flat_labels[s_off + (s >> 1) - 1]) rep++;
      }
      s_ext[s] = v;
    }
    // count repeats (tiny): block reduce through smem atomics
    if (threadIdx.x == 0) s_repeats = 0;
    __syncthreads();
    if (rep) atomicAdd(&s_repeats, rep);
  }
  __syncthreads();

  const bool feasible = !too_long && Tn > 0 && (L + s_repeats <= Tn);
  if (!feasible) {
    if (threadIdx.x == 0) costs[n] = too_long ? __int_as_float(0x7fc00000) : 0.0f;
    if (grad != nullptr) {
      for (int t = warp; t < T; t += CTC_WARPS)
        *reinterpret_cast<float2*>(grad + ((size_t)t * N + n) * CTC_C + 2 * lane) = make_float2(0.f, 0.f);
    }
    return;
  }

  // ---------------- phase 0: normalisers + emission gather ----------------
  int cls[KS];
#pragma unroll
  for (int k = 0; k < KS; ++k) cls[k] = s_ext[lane * KS + k];
  for (int t0 = warp; t0 < Tn; t0 += CTC_WARPS * CTC_RB) {
    float2 xr[CTC_RB];
#pragma unroll
    for (int u = 0; u < CTC_RB; ++u) {
      const int t = t0 + CTC_WARPS * u;
      xr[u] = (t < Tn) ? __ldg(reinterpret_cast<const float2*>(logits + ((size_t)t * N + n) * CTC_C) + lane)
                       : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int g = 0; g < CTC_RB; g += CTC_RG) {
      if (t0 + CTC_WARPS * g >= Tn) break;               // warp-uniform
      float x0[CTC_RG], x1[CTC_RG], m[CTC_RG], sum[CTC_RG];
#pragma unroll
      for (int u = 0; u < CTC_RG; ++u) {
        x0[u] = xr[g + u].x * LOG2E; x1[u] = xr[g + u].y * LOG2E;
        m[u] = fmaxf(x0[u], x1[u]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int u = 0; u < CTC_RG; ++u) m[u] = fmaxf(m[u], __shfl_xor_sync(0xffffffffu, m[u], o));
#pragma unroll
      for (int u = 0; u < CTC_RG; ++u) sum[u] = ptx::ex2(x0[u] - m[u]) + ptx::ex2(x1[u] - m[u]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int u = 0; u < CTC_RG; ++u) sum[u] += __shfl_xor_sync(0xffffffffu, sum[u], o);
#pragma unroll
      for (int u = 0; u < CTC_RG; ++u) {
        const int t = t0 + CTC_WARPS * (g + u);
        const float lse = m[u] + ptx::lg2(sum[u]);
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          const int c = cls[k];
          const float v0 = __shfl_sync(0xffffffffu, x0[u], c >> 1);
          const float v1 = __shfl_sync(0xffffffffu, x1[u], c >> 1);
          if (t < Tn) s_e[(size_t)t * SP + lane * KS + k] = ((c & 1) ? v1 : v0) - lse;
        }
        if (lane == 0 && t < Tn) s_lse[t] = lse;
      }
    }
  }
  __syncthreads();

  // ---------------- phase 1: alpha (warp 0) || beta (warp 1) ----------------
  {
    // skip transition allowed into state s from s-2 (alpha) / from s into s+2 (beta)
    bool skip_in[KS], skip_out[KS];
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      int s = lane * KS + k;
      skip_in[k] = (s >= 2) && (s < S) && (s_ext[s] != blank) && (s_ext[s] != s_ext[s - 2]);
      skip_out[k] = (s + 2 < S) && (s_ext[s + 2] != blank) && (s_ext[s + 2] != s_ext[s]);
    }
    float a[KS];
    if (warp >= 2) {
      // idle during the two recursions (frame-parallel phases only)
    } else if (warp == 0) {
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        int s = lane * KS + k;
        a[k] = (s < 2 && s < S) ? s_e[s] : NEG_INF;
        s_alpha[s] = a[k];
      }
      for (int t = 1; t < Tn; ++t) {
        float up1 = __shfl_up_sync(0xffffffffu, a[KS - 1], 1);
        float up2 = (KS >= 2) ? __shfl_up_sync(0xffffffffu, a[(KS >= 2) ? KS - 2 : 0], 1)
                              : __shfl_up_sync(0xffffffffu, a[0], 2);
        if (lane == 0) { up1 = NEG_INF; up2 = NEG_INF; }
        if (KS == 1 && lane == 1) up2 = NEG_INF;
        float nw[KS];
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          float p1 = (k >= 1) ? a[(k >= 1) ? k - 1 : 0] : up1;
          float p2 = (k >= 2) ? a[(k >= 2) ? k - 2 : 0] : ((k == 1) ? up1 : up2);
          if (!skip_in[k]) p2 = NEG_INF;
          int s = lane * KS + k;
          float v = lse3(a[k], p1, p2) + s_e[(size_t)t * SP + s];
          nw[k] = (s < S) ? v : NEG_INF;
        }
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          a[k] = nw[k];
          s_alpha[(size_t)t * SP + lane * KS + k] = a[k];
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        int s = lane * KS + k;
        a[k] = (s < S && s >= S - 2) ? s_e[(size_t)(Tn - 1) * SP + s] : NEG_INF;
        s_beta[(size_t)(Tn - 1) * SP + s] = a[k];
      }
      for (int t = Tn - 2; t >= 0; --t) {
        float dn1 = __shfl_down_sync(0xffffffffu, a[0], 1);
        float dn2 = (KS >= 2) ? __shfl_down_sync(0xffffffffu, a[(KS >= 2) ? 1 : 0], 1)
                              : __shfl_down_sync(0xffffffffu, a[0], 2);
        if (lane == 31) { dn1 = NEG_INF; dn2 = NEG_INF; }
        if (KS == 1 && lane == 30) dn2 = NEG_INF;
        float nw[KS];
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          float p1 = (k + 1 < KS) ? a[(k + 1 < KS) ? k + 1 : 0] : dn1;
          // state s+2: own register, else neighbour's first (dn1) or second / lane+2's (dn2)
          float p2 = (k + 2 < KS) ? a[(k + 2 < KS) ? k + 2 : 0] : ((k + 1 < KS) ? dn1 : dn2);
          if (!skip_out[k]) p2 = NEG_INF;
          int s = lane * KS + k;
          float v = lse3(a[k], p1, p2) + s_e[(size_t)t * SP + s];
          nw[k] = (s < S) ? v : NEG_INF;
        }
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          a[k] = nw[k];
          s_beta[(size_t)t * SP + lane * KS + k] = a[k];
        }
      }
    }
  }
  __syncthreads();

  // log2-likelihood from the last alpha row
  const float aS1 = s_alpha[(size_t)(Tn - 1) * SP + (S - 1)];
  const float aS2 = (S >= 2) ? s_alpha[(size_t)(Tn - 1) * SP + (S - 2)] : NEG_INF;
  const float ll2 = lse3(aS1, aS2, NEG_INF);
  if (threadIdx.x == 0) costs[n] = -ll2 * LN2;
  if (grad == nullptr) return;

  // ---------------- phase 2: gradient rows ----------------
  float* acc = s_acc + warp * CTC_RG * CTC_C;         // [CTC_RG rows][64] per warp
  for (int t0 = warp; t0 < T; t0 += CTC_WARPS * CTC_RB) {
    float2 xr[CTC_RB];
#pragma unroll
    for (int u = 0; u < CTC_RB; ++u) {
      const int t = t0 + CTC_WARPS * u;
      xr[u] = ((t < Tn) && (ll2 != NEG_INF)) ? __ldg(reinterpret_cast<const float2*>(logits + ((size_t)t * N + n) * CTC_C) + lane)
                                             : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int g = 0; g < CTC_RB; g += CTC_RG) {
      if (t0 + CTC_WARPS * g >= T) break;                 // warp-uniform
      bool live[CTC_RG];
#pragma unroll
      for (int u = 0; u < CTC_RG; ++u) {
        const int t = t0 + CTC_WARPS * (g + u);
        live[u] = (t < Tn) && (ll2 != NEG_INF);
        acc[u * CTC_C + 2 * lane] = 0.f;
        acc[u * CTC_C + 2 * lane + 1] = 0.f;
      }
      __syncwarp();
      float blank_sum[CTC_RG];
#pragma unroll
      for (int u = 0; u < CTC_RG; ++u) {
        const int t = t0 + CTC_WARPS * (g + u);
        blank_sum[u] = 0.f;
        if (live[u]) {
#pragma unroll
          for (int k = 0; k < KS; ++k) {
            const int st = lane * KS + k;
            if (st < S) {
              const size_t i = (size_t)t * SP + st;
              const float w = ptx::ex2(s_alpha[i] + s_beta[i] - s_e[i] - ll2);   // alpha*beta / y / p(l|x)
              if (st & 1) atomicAdd(&acc[u * CTC_C + cls[k]], w);
              else blank_sum[u] += w;
            }
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int u = 0; u < CTC_RG; ++u) blank_sum[u] += __shfl_xor_sync(0xffffffffu, blank_sum[u], o);
      if (lane == 0) {
#pragma unroll
        for (int u = 0; u < CTC_RG; ++u) atomicAdd(&acc[u * CTC_C + blank], blank_sum[u]);
      }
      __syncwarp();
#pragma unroll
      for (int u = 0; u < CTC_RG; ++u) {
        const int t = t0 + CTC_WARPS * (g + u);
        if (t >= T) continue;
        float2* gp = reinterpret_cast<float2*>(grad + ((size_t)t * N + n) * CTC_C) + lane;
        if (!live[u]) {
          *gp = make_float2(0.f, 0.f);
        } else {
          const float lse = s_lse[t];
          const float y0 = ptx::ex2(xr[g + u].x * LOG2E - lse), y1 = ptx::ex2(xr[g + u].y * LOG2E - lse);
          *gp = make_float2(grad_scale * (y0 - acc[u * CTC_C + 2 * lane]), grad_scale * (y1 - acc[u * CTC_C + 2 * lane + 1]));
        }
      }
      __syncwarp();
    }
  }
}

// Greedy decode: one warp per utterance, lane = frame (chunks of 32 frames).
__global__ void __launch_bounds__(128) ctc_greedy_kernel(const float* __restrict__ logits,
                                                         const int* __restrict__ input_len, int T, int N,
                                                         int tf_blank, int strip, int* __restrict__ out,
                                                         int* __restrict__ out_len) {
  const int n = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  int Tn = max(0, min(input_len[n], T));
  int count = 0;
  int prev_carry = -1;
  for (int t0 = 0; t0 < Tn; t0 += 32) {
    int t = t0 + lane;
    int best = -1;
    if (t < Tn) {
      const float4* row = reinterpret_cast<const float4*>(logits + ((size_t)t * N + n) * CTC_C);
      float bv = -INFINITY;
      best = 0;
#pragma unroll
      for (int q = 0; q < CTC_C / 4; ++q) {
        float4 v = __ldg(row + q);
        if (v.x > bv) { bv = v.x; best = 4 * q; }
        if (v.y > bv) { bv = v.y; best = 4 * q + 1; }
        if (v.z > bv) { bv = v.z; best = 4 * q + 2; }
        if (v.w > bv) { bv = v.w; best = 4 * q + 3; }
      }
      // NaN rows: comparisons false -> best stays 0 (lowest index), matching argmax-on-ties
    }
    int prev = __shfl_up_sync(0xffffffffu, best, 1);
    if (lane == 0) prev = prev_carry;
    bool keep = (t < Tn) && (best != tf_blank) && (best != prev) && (best != strip);
    unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) out[(size_t)n * T + count + __popc(m & ((1u << lane) - 1))] = best;
    count += __popc(m);
    prev_carry = __shfl_sync(0xffffffffu, best, 31);
  }
  for (int i = count + lane; i < T; i += 32) out[(size_t)n * T + i] = 0;
  if (lane == 0) out_len[n] = count;
}

size_t ctc_smem_bytes(int T, int KS) { return sizeof(float) * ((size_t)T + 3 * (size_t)T * 32 * KS + CTC_WARPS * 4 * CTC_C); }

template <int KS>
int launch_ctc(const float* logits, float* grad, const int* flat_labels, const int* label_len, const int* input_len,
               int T, int N, int blank, float grad_scale, float* costs, cudaStream_t st) {
  size_t smem = ctc_smem_bytes(T, KS);
  if (smem > 200 * 1024) return CRNN_UNSUPPORTED;
  CUDA_TRY(cudaFuncSetAttribute(ctc_loss_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ctc_loss_kernel<KS><<<N, CTC_THREADS, smem, st>>>(logits, grad, flat_labels, label_len, input_len, T, N, blank, grad_scale,
                                           costs);
  CUDA_TRY(cudaGetLastError());
  return CRNN_OK;
}

}  // namespace

extern "C" int crnn_ctc_workspace_size(int T, int N, int C, int max_label_len, size_t* bytes) {
  if (!bytes || T <= 0 || N <= 0 || max_label_len < 0) return crnn_fail(CRNN_INVALID_VALUE, "ctc_workspace_size: bad args");
  if (C != CTC_C) return crnn_fail(CRNN_UNSUPPORTED, "ctc: C must be 64");
  *bytes = 0;   // alpha/beta live in shared memory; kept for warp-ctc call-shape compatibility
  return CRNN_OK;
}

extern "C" int crnn_ctc_loss(const float* logits, float* grad, const int* flat_labels, const int* label_len,
                             const int* input_len, int T, int N, int C, int blank, int max_label_len,
                             float grad_scale, float* costs, void* workspace, size_t workspace_bytes,
                             crnn_stream_t stream) {
  (void)workspace; (void)workspace_bytes;
  if (!logits || !flat_labels || !label_len || !input_len || !costs) return crnn_fail(CRNN_INVALID_VALUE, "ctc_loss: null pointer");
  if (T <= 0 || N <= 0 || blank < 0 || blank >= C || max_label_len < 0) return crnn_fail(CRNN_INVALID_VALUE, "ctc_loss: bad shape");
  if (C != CTC_C) return crnn_fail(CRNN_UNSUPPORTED, "ctc: C must be 64");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int S = 2 * max_label_len + 1;
  if (S <= 32) return launch_ctc<1>(logits, grad, flat_labels, label_len, input_len, T, N, blank, grad_scale, costs, st);
  if (S <= 64) return launch_ctc<2>(logits, grad, flat_labels, label_len, input_len, T, N, blank, grad_scale, costs, st);
  if (S <= 128) return launch_ctc<4>(logits, grad, flat_labels, label_len, input_len, T, N, blank, grad_scale, costs, st);
  return crnn_fail(CRNN_UNSUPPORTED, "ctc: max_label_len > 63");
}

extern "C" int crnn_ctc_greedy(const float* logits, const int* input_len, int T, int N, int C, int tf_blank,
                               int strip, int* out, int* out_len, crnn_stream_t stream) {
  if (!logits || !input_len || !out || !out_len || T <= 0 || N <= 0) return crnn_fail(CRNN_INVALID_VALUE, "ctc_greedy: bad args");
  if (C != CTC_C) return crnn_fail(CRNN_UNSUPPORTED, "ctc: C must be 64");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ctc_greedy_kernel<<<(N + 3) / 4, 128, 0, st>>>(logits, input_len, T, N, tf_blank, strip, out, out_len);
  CUDA_TRY(cudaGetLastError());
  return CRNN_OK;
}
