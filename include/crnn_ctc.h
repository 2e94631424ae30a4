/*
 * crnn_ctc.h -- C ABI of libcrnnctc.so: the B200 (sm_100a) CRNN+CTC hot path that stands
 * behind the model/solver API of ilovin/lstm_ctc_ocr.
 *
 * The reference has no C ABI of its own (it is pure Python on TensorFlow 1.0.1 + the
 * warp-ctc TF binding).  Each entry point below names the reference call site it replaces
 * (paths relative to the reference checkout).  Conventions follow warp-ctc's ctc.h:
 * status-code returns, no exceptions across the ABI, caller-owned device buffers and
 * workspace (size queried first), every call asynchronous on the caller's stream, no
 * hidden host synchronisation.  All pointers are DEVICE pointers unless marked host.
 * One host thread per handle (thread-compatible, not thread-safe).
 */
#ifndef CRNN_CTC_H_
#define CRNN_CTC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* crnn_stream_t;   /* == cudaStream_t */
typedef struct crnn_model crnn_model;

enum crnn_status {
  CRNN_OK = 0,
  CRNN_INVALID_VALUE = 1,     /* bad shape / length / null pointer */
  CRNN_CUDA_ERROR = 2,        /* a CUDA runtime/driver call failed; see crnn_last_error() */
  CRNN_NOT_BOUND = 3,         /* model used before crnn_model_bind() */
  CRNN_UNSUPPORTED = 4,       /* shape outside what the sm_100a kernels implement */
  CRNN_WORKSPACE_TOO_SMALL = 5
};

int         crnn_version(void);
const char* crnn_status_string(int status);
const char* crnn_last_error(void);           /* host string, valid until the next failing call */

/* ------------------------------------------------------------------------------------------
 * CTC operator boundary.
 * Replaces warpctc_tensorflow.ctc(activations, flat_labels, label_lengths, input_lengths)
 * at lib/networks/network.py:653-654 (warp-ctc compute_ctc_loss semantics: softmax inside,
 * blank label 0 by default, cost = -log p(l|x), zero gradient for frames >= input_length,
 * infeasible alignment -> cost 0 and zero gradient).
 *   logits      [T,N,C] f32, unnormalised, time-major (lib/networks/network.py:126-128)
 *   grad        [T,N,C] f32 or NULL; receives grad_scale * d costs[n] / d logits
 *   flat_labels [sum(label_len)] i32, label_len [N] i32, input_len [N] i32
 *   costs       [N] f32
 * max_label_len is a host-side upper bound on label_len[] (chooses the states-per-lane
 * variant; labels longer than it make that sample's cost NaN).  C must be 64.
 * Validation: a sample whose labels contain an id outside [0,C) or equal to `blank`, or whose label_len is
 * negative or above max_label_len, gets cost NaN and an all-zero gradient -- the kernels never index with such an id.
 * input_len is clamped to [0,T].  flat_labels MUST hold sum(label_len) entries (its length is not passed and cannot be
 * checked on the device; the Python wrappers check it).
 * ---------------------------------------------------------------------------------------- */
int crnn_ctc_workspace_size(int T, int N, int C, int max_label_len, size_t* bytes);
int crnn_ctc_loss(const float* logits, float* grad, const int* flat_labels, const int* label_len,
                  const int* input_len, int T, int N, int C, int blank, int max_label_len,
                  float grad_scale, float* costs, void* workspace, size_t workspace_bytes,
                  crnn_stream_t stream);

/* Greedy decode.  Replaces tf.nn.ctc_*_decoder(merge_repeated=True) + sparse_tensor_to_dense
 * at lib/networks/network.py:656-657 and the zero stripping of lib/lstm/utils/training.py:32:
 * per frame argmax (lowest index on ties) for t < input_len; emit iff != tf_blank and != the
 * previous raw argmax; drop `strip`.  out [N,T] i32 zero padded, out_len [N] i32. */
int crnn_ctc_greedy(const float* logits, const int* input_len, int T, int N, int C, int tf_blank,
                    int strip, int* out, int* out_len, crnn_stream_t stream);

/* Beam-search decode on the HOST.  Replaces tf.nn.ctc_beam_search_decoder(logits, seq_len, merge_repeated=True)
 * (beam_width 100, top_paths 1, blank = C-1) + sparse_tensor_to_dense(default 0) at lib/networks/network.py:656-657 and
 * lib/lstm/test.py:30-31.  The reference's op is a CPU-only TensorFlow kernel used at validation / evaluation time; so is
 * this one: ALL pointers are HOST pointers (copy the logits back once), utterances are spread over `num_threads` host
 * threads (0 = hardware concurrency).  logits [T,N,C] f32 unnormalised; out [N,T] i32 zero padded (labels equal to `strip`
 * dropped, lib/lstm/utils/training.py:32); out_len [N]; neg_log_prob [N] or NULL (-log P of the best prefix). */
int crnn_ctc_beam_search(const float* logits_host, const int* input_len_host, int T, int N, int C, int beam_width,
                         int merge_repeated, int strip, int* out_host, int* out_len_host, float* neg_log_prob_host,
                         int num_threads);

/* 1 when `host_ptr` lies in page-locked (cudaHostAlloc / cudaHostRegister) memory, else 0.  The Python feed path uses it to
 * decide whether a fed numpy batch can be DMA'd in place (crnn_forward_host) or has to be staged. */
int crnn_host_is_pinned(const void* host_ptr);

/* ------------------------------------------------------------------------------------------
 * Model boundary.  Replaces the graph built by lib/networks/LSTM_train.py:22-38 through
 * lib/networks/network.py (conv_single :160-191, max_pool :343-350, reshape_squeeze_layer
 * :361-368, bi_lstm :97-129) and executed by sess.run at lib/lstm/train.py:129-130.
 * ---------------------------------------------------------------------------------------- */
typedef struct crnn_config {
  int   img_height;     /* cfg.IMG_HEIGHT = 32        (lib/lstm/config.py:19) */
  int   nclasses;       /* cfg.NCLASSES   = 64        (lib/lstm/config.py:23) */
  int   num_hid;        /* cfg.TRAIN.NUM_HID = 512    (lib/lstm/config.py:48) */
  float bn_eps;         /* 1e-3  tf.contrib.layers.batch_norm default */
  float weight_decay;   /* cfg.TRAIN.WEIGHT_DECAY (lstm/lstm.yml:13 -> 1e-5) */
  int   compute_dtype;  /* 1 = bf16 operands / f32 accumulate (tcgen05 kind::f16): the throughput path, forward + backward.
                         * 2 = f32-class: every operand split into bf16 hi + bf16 lo, three tcgen05 products per term, f32
                         *     accumulate and f32 elementwise math (the reference computes in fp32, LSTM_train.py:10);
                         *     forward + CTC only (BASELINE configs[1])
                         * 3 = tf32: the same forward-only orchestration on tcgen05 kind::tf32 operands (f32 tensors, rounded to
                         *     nearest tf32 where produced; 10-bit mantissa, one pass over K at half the kind::f16 rate) */
} crnn_config;

int     crnn_model_create(const crnn_config* cfg, crnn_model** out);
int     crnn_model_destroy(crnn_model* m);

/* Trainable tensors, addressable by TF variable name and laid out exactly as the reference
 * checkpoint has them (HWIO conv kernels, [768,1024] LSTM matrices with gate order i,j,f,o;
 * SURVEY §8(a)).  All live in ONE flat f32 buffer owned by the caller. */
int     crnn_num_tensors(const crnn_model* m);
int64_t crnn_param_count(const crnn_model* m);
int     crnn_param_info(const crnn_model* m, int index, const char** tf_name, int64_t* offset,
                        int64_t shape[4], int* ndim);
/* Bind caller-owned flat buffers (each crnn_param_count() f32).  grads/adam_* may be NULL
 * for inference.  Marks the derived bf16 operand copies dirty. */
int     crnn_model_bind(crnn_model* m, float* params, float* grads, float* adam_m, float* adam_v);
int     crnn_model_params_changed(crnn_model* m);   /* caller wrote params in place */

int     crnn_model_workspace_size(const crnn_model* m, int N, int W, int train, size_t* bytes);

/* data [N,W,32] f32 (width-major rows, lib/lstm/utils/gen.py:62-64), time_step_len [N] i32
 * (nw//4-1, gen.py:54) -> logits_out [T=W/4-1, N, 64] f32. */
int     crnn_forward(crnn_model* m, const float* data, const int* time_step_len, int N, int W,
                     float* logits_out, void* workspace, size_t workspace_bytes,
                     crnn_stream_t stream);

/* Same forward, fed from PAGE-LOCKED HOST memory (the reference feeds host numpy arrays through feed_dict every iteration,
 * lib/lstm/train.py:121-130).  The batch is copied in `chunks` image ranges on `copy_stream` into the caller-owned device
 * tensor `data_staging` [N,W,32] while the batch-independent front end (conv1 .. conv3_2) of the previous range runs on
 * `stream`; from the first batch-statistics BatchNorm on the batch is processed whole.  chunks <= 1 (or a batch that does not
 * split on tile boundaries) degenerates to copy-then-compute.  `data_staging` holds the whole batch on return order of
 * `stream` (crnn_backward reads it). */
int     crnn_forward_host(crnn_model* m, const float* host_data, float* data_staging, const int* time_step_len,
                          int N, int W, float* logits_out, void* workspace, size_t workspace_bytes, int chunks,
                          crnn_stream_t stream, crnn_stream_t copy_stream);

/* Same forward, fed from ORDINARY (pageable) host memory -- the reference's solver builds a fresh np.array(...) batch every
 * iteration (lib/lstm/train.py:119-125).  Every image range is first moved into the caller's page-locked `pinned_staging`
 * [N,W,32] by `host_threads` host threads (a persistent pool inside the library), then DMA'd and processed as in
 * crnn_forward_host; the host moves range c+1 while the GPU copies / computes range c.  The caller must not touch
 * `pinned_staging` until the copies issued on `copy_stream` have completed. */
int     crnn_forward_pageable(crnn_model* m, const float* pageable_data, float* pinned_staging, float* data_staging,
                              const int* time_step_len, int N, int W, float* logits_out, void* workspace,
                              size_t workspace_bytes, int chunks, int host_threads, crnn_stream_t stream,
                              crnn_stream_t copy_stream);

/* The host-side copy crnn_forward_pageable uses, on its own: `bytes` from `src` to `dst` (plain host pointers, non-overlapping) split
 * over `threads` threads of the library's persistent pool (the caller's thread included).  No CUDA call is made. */
int     crnn_host_copy(void* dst, const void* src, size_t bytes, int threads);

/* loss = mean_n(costs) + weight_decay * 0.5 * sum(w^2) over conv kernels + logits matrix
 * (lib/networks/network.py:655,660-662).  loss_out: 1 f32 on device. */
int     crnn_total_loss(crnn_model* m, const float* costs, int N, float* loss_out,
                        crnn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Training.  Replaces tf.gradients + tf.clip_by_global_norm(., 10.0) + AdamOptimizer.apply_gradients
 * (lib/lstm/train.py:73-83).  Protocol per step:
 *   crnn_model_set_training(m, 1) once; crnn_forward (saves what the backward needs in the workspace);
 *   crnn_ctc_loss with grad != NULL and grad_scale = 1/N  -> dlogits;  crnn_backward -> flat `grads` buffer;
 *   [data parallel: all-reduce(SUM) `grads` across ranks];  crnn_clip_adam_step.
 * ---------------------------------------------------------------------------------------- */
int     crnn_model_set_training(crnn_model* m, int flag);
int     crnn_backward(crnn_model* m, const float* data, const int* time_step_len, const float* dlogits,
                      int N, int W, void* workspace, size_t workspace_bytes, crnn_stream_t stream);
/* grads += weight_decay*wd_mul*w on the L2-regularised tensors (network.py:660-662); g *= grad_mul;
 * g *= clip/max(||g||, clip); TF Adam: lr_t = lr*sqrt(1-b2^step)/(1-b1^step), theta -= lr_t*m/(sqrt(v)+1e-8).
 * Single GPU: grad_mul = wd_mul = 1.  Data parallel after a SUM all-reduce: grad_mul = 1/world, wd_mul = world. */
int     crnn_clip_adam_step(crnn_model* m, float lr, float clip, int step, float grad_mul, float wd_mul,
                            crnn_stream_t stream);
int     crnn_last_grad_norm(crnn_model* m, float grad_mul, float* out_host, crnn_stream_t stream);   /* syncs */

/* ------------------------------------------------------------------------------------------
 * Data parallelism (one process per GPU).  The reference is single-device: its BatchNorm sees the whole batch
 * (lib/networks/network.py:177-178) and its optimizer the whole-batch gradient (lib/lstm/train.py:81-83).  With the batch
 * sharded over `world` ranks the same function needs (1) the BN sums of conv4_1 / conv4_2 -- forward [sum x, sum x^2] and
 * backward [sum dy, sum dy*xhat], 2 x 512 f64 each -- summed over ranks, and (2) the SUM of the flat gradient buffers before
 * crnn_clip_adam_step(grad_mul = 1/world, wd_mul = world).
 *
 * (1) crnn_model_set_data_parallel switches the BN layers to global-batch statistics.  The exchange runs INSIDE the BN
 *     finalize kernel over NVLink peer memory when crnn_model_set_peers was called (every rank stores its 8 KB of sums into
 *     every peer's inbox, release/acquire flags at system scope, fixed-order f64 summation: bit-identical on all ranks, no
 *     NCCL launch on the forward path); otherwise through `allreduce` (e.g. an NCCL all-reduce issued by the caller).
 * (2) crnn_model_set_grad_ready_callback: crnn_backward calls `fn(user, offset, count, stream)` as soon as the gradients of a
 *     contiguous range [offset, offset+count) of the flat buffer are final (LSTM+logits first, conv1+conv2 last, 7 ranges
 *     covering the buffer exactly once), so the caller can all-reduce each range on a side stream while the rest of the
 *     backward pass still runs.
 * ---------------------------------------------------------------------------------------- */
typedef int  (*crnn_allreduce_fn)(void* user, void* dev_ptr, size_t count, int is_f64, crnn_stream_t stream);   /* in-place SUM */
typedef void (*crnn_grad_ready_fn)(void* user, int64_t offset, int64_t count, crnn_stream_t stream);
int     crnn_model_set_data_parallel(crnn_model* m, int rank, int world, crnn_allreduce_fn allreduce, void* user);
int     crnn_model_set_grad_ready_callback(crnn_model* m, crnn_grad_ready_fn fn, void* user);
/* Leave `sms` SMs free in the persistent kernels of crnn_backward (their grids are num_sms - sms CTAs), so that the collective
 * kernels the caller launches from the grad-ready callback find free SMs instead of delaying the tail of a full-GPU grid. */
int     crnn_model_set_backward_sm_reserve(crnn_model* m, int sms);
/* Peer-memory inboxes (cudaMalloc + CUDA IPC): create one per rank, exchange the 64-byte handles through the host language
 * (e.g. torch.distributed.all_gather_object), open the peers', hand all `world` pointers (own at [rank]) to the model. */
size_t  crnn_peer_inbox_bytes(void);
int     crnn_peer_inbox_create(void** dev_ptr, unsigned char handle[64]);
int     crnn_peer_inbox_open(const unsigned char handle[64], void** dev_ptr);
int     crnn_peer_inbox_close(void* dev_ptr);      /* a pointer obtained from crnn_peer_inbox_open */
int     crnn_peer_inbox_destroy(void* dev_ptr);    /* a pointer obtained from crnn_peer_inbox_create */
int     crnn_model_set_peers(crnn_model* m, int rank, int world, void* const* inbox_ptrs_host);
int     crnn_peer_error(crnn_model* m, int* err_host);   /* 1 if an exchange timed out waiting for a peer (syncs the device) */

/* Debug/parity taps: copy a named intermediate of the last crnn_forward() as f32 into dst.
 * names: "conv1" "conv2" "conv3_1" "conv3_2" "conv4_1" "conv4_2" "conv5" "lstm_out"
 * (pooled / post-activation, NHWC, as the reference's layers dict holds them). */
int     crnn_debug_tap(crnn_model* m, const char* name, float* dst, size_t dst_elems,
                       void* workspace, crnn_stream_t stream);

/* Per-stage timing of crnn_forward with CUDA events recorded on the caller's stream (bench.py's roofline line).
 * crnn_profile_begin arms the next `max_forwards` forward calls; crnn_profile_read synchronises on the events and
 * returns ms_out[forwards][crnn_profile_num_stages()]. */
int         crnn_profile_begin(crnn_model* m, int max_forwards);
int         crnn_profile_num_stages(void);
const char* crnn_profile_stage_name(int stage);
int         crnn_profile_read(crnn_model* m, float* ms_out, int* forwards);
/* same for crnn_backward (armed by the same crnn_profile_begin) */
int         crnn_profile_bwd_num_stages(void);
const char* crnn_profile_bwd_stage_name(int stage);
int         crnn_profile_bwd_read(crnn_model* m, float* ms_out, int* backwards);

/* Stand-alone bf16 GEMM test entry (tests only): D[M,Nc] f32 = A[M,K] * B[Nc,K]^T, bf16 in. */
int     crnn_test_gemm_bf16(const void* A, const void* B, float* D, int M, int Nc, int K,
                            int block_n, crnn_stream_t stream);

/* MN-major ("TN") GEMM test entry (tests only): D[M,N] f32 (caller-zeroed) += A[K,M]^T * B[K,N], bf16 in. */
int     crnn_test_gemm_tn_bf16(const void* A, const void* B, float* D, int M, int N, int K, int block_n,
                               int k_splits, crnn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif  /* CRNN_CTC_H_ */
