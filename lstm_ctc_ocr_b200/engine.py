"""Device-side engine behind the reference-facing API.  PyTorch tensors are used purely as
device-memory containers and for the current CUDA stream; every kernel on this path lives in
libcrnnctc.so (hand-written sm_100a CUDA, see csrc/)."""
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from ._lib import CrnnConfig, CrnnError, check

NCLASSES = 64
TF_BLANK = NCLASSES - 1


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


class CrnnModel:
    """Owns the flat f32 parameter buffer (TF variable names/layouts) and the C model handle."""

    def __init__(self, weight_decay=1e-5, bn_eps=1e-3, device=None, compute_dtype="bf16"):
        if not torch.cuda.is_available():
            raise CrnnError("lstm_ctc_ocr_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        torch.cuda.set_device(self.device)
        # "bf16": bf16 operands / f32 accumulate (throughput path); "f32": split-bf16 operands, f32-class (BASELINE configs[1]);
        # "tf32": kind::tf32 operands, same forward-only orchestration as "f32"
        self.compute_dtype = {"bf16": 1, "f32": 2, "tf32": 3, 1: 1, 2: 2, 3: 3}[compute_dtype]
        cfg = CrnnConfig(32, NCLASSES, 512, bn_eps, weight_decay, self.compute_dtype)
        h = _lib.c_void_p()
        check(self.lib.crnn_model_create(cfg, h))
        self.handle = h
        self.weight_decay = weight_decay
        self.total = int(self.lib.crnn_param_count(h))
        self.table = OrderedDict()
        for i in range(self.lib.crnn_num_tensors(h)):
            name = _lib.c_char_p(); off = _lib.c_int64(); shp = (_lib.c_int64 * 4)(); nd = _lib.c_int()
            check(self.lib.crnn_param_info(h, i, name, off, shp, nd))
            self.table[name.value.decode()] = (int(off.value), tuple(int(shp[k]) for k in range(nd.value)))
        self.params = torch.zeros(self.total, dtype=torch.float32, device=self.device)
        self.grads = None
        self.adam_m = None
        self.adam_v = None
        self._bind()
        self._ws = None
        self._ws_key = None
        self.training = False

    # ---- training --------------------------------------------------------------------------
    def set_training(self, flag=True):
        """Allocate the gradient / Adam-slot buffers (flat f32, same layout as params) and switch the forward to the
        variant that saves what the backward pass needs."""
        if flag and self.grads is None:
            self.grads = torch.zeros_like(self.params)
            self.adam_m = torch.zeros_like(self.params)
            self.adam_v = torch.zeros_like(self.params)
            self._bind()
        check(self.lib.crnn_model_set_training(self.handle, 1 if flag else 0))
        self.training = bool(flag)
        self._ws_key = None

    def grad_tensor(self, name):
        off, shp = self.table[name]
        return self.grads[off:off + int(np.prod(shp))].view(*shp)

    def backward(self, data, time_step_len, dlogits):
        """dlogits [T,N,64] f32 = d loss / d logits (e.g. the CTC gradient scaled by 1/N) -> fills self.grads."""
        N, W, _ = data.shape
        ws, nbytes = self._workspace(N, W)
        check(self.lib.crnn_backward(self.handle, data.data_ptr(), time_step_len.data_ptr(), dlogits.data_ptr(), N, W, ws, nbytes,
                                     _stream()))

    def clip_adam_step(self, lr, step, clip=10.0, grad_mul=1.0, wd_mul=1.0):
        check(self.lib.crnn_clip_adam_step(self.handle, float(lr), float(clip), int(step), float(grad_mul), float(wd_mul), _stream()))

    def last_grad_norm(self, grad_mul=1.0):
        out = _lib.c_float()
        check(self.lib.crnn_last_grad_norm(self.handle, float(grad_mul), out, _stream()))
        return float(out.value)

    def _bind(self):
        check(self.lib.crnn_model_bind(self.handle, _ptr(self.params), _ptr(self.grads), _ptr(self.adam_m), _ptr(self.adam_v)))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                torch.cuda.synchronize(self.device)
                self.lib.crnn_model_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- parameters, addressable by TF variable name --------------------------------------
    def tensor(self, name):
        off, shp = self.table[name]
        n = int(np.prod(shp))
        return self.params[off:off + n].view(*shp)

    def load_params(self, params):
        for name in self.table:
            v = params[name]
            v = torch.as_tensor(np.asarray(v, dtype=np.float32)) if not torch.is_tensor(v) else v.detach().float().cpu()
            self.tensor(name).copy_(v.to(self.device))
        check(self.lib.crnn_model_params_changed(self.handle))

    def state_dict(self):
        return OrderedDict((k, self.tensor(k).detach().cpu().numpy().copy()) for k in self.table)

    # ---- forward ----------------------------------------------------------------------------
    def _workspace(self, N, W):
        key = (N, W, self.training)
        if self._ws_key != key:
            nbytes = _lib.c_size_t()
            check(self.lib.crnn_model_workspace_size(self.handle, N, W, 1 if self.training else 0, nbytes))
            self._ws = None
            self._ws = torch.empty(nbytes.value + 1024, dtype=torch.uint8, device=self.device)
            self._ws_key = key
        base = self._ws.data_ptr()
        aligned = (base + 1023) // 1024 * 1024
        return aligned, self._ws.numel() - (aligned - base)

    def forward(self, data, time_step_len, out=None):
        """data [N,W,32] f32 cuda, time_step_len [N] i32 cuda -> logits [T,N,64] f32 (time-major)."""
        assert data.is_cuda and data.dtype == torch.float32 and data.is_contiguous()
        assert time_step_len.is_cuda and time_step_len.dtype == torch.int32
        N, W, Hh = data.shape
        if Hh != 32:
            raise CrnnError("data must be [N, W, 32] (cfg.NUM_FEATURES = 32)")
        T = W // 4 - 1
        if out is None:
            out = torch.empty((T, N, NCLASSES), dtype=torch.float32, device=self.device)
        ws, nbytes = self._workspace(N, W)
        check(self.lib.crnn_forward(self.handle, data.data_ptr(), time_step_len.data_ptr(), N, W, out.data_ptr(), ws,
                                    nbytes, _stream()))
        return out

    def forward_host(self, host_data, time_step_len, chunks=4, out=None, wait_copy=True):
        """host_data: C-contiguous f32 numpy array [N,W,32] in PAGE-LOCKED memory.  The H2D copy is cut into `chunks` image
        ranges on a side stream and overlapped with the conv front end (crnn_forward_host).  Returns (logits, device data)."""
        N, W, Hh = host_data.shape
        if Hh != 32:
            raise CrnnError("data must be [N, W, 32] (cfg.NUM_FEATURES = 32)")
        assert host_data.dtype == np.float32 and host_data.flags.c_contiguous
        T = W // 4 - 1
        if out is None:
            out = torch.empty((T, N, NCLASSES), dtype=torch.float32, device=self.device)
        if getattr(self, "_stage", None) is None or self._stage.shape != (N, W, 32):
            self._stage = torch.empty((N, W, 32), dtype=torch.float32, device=self.device)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        ws, nbytes = self._workspace(N, W)
        check(self.lib.crnn_forward_host(self.handle, host_data.ctypes.data, self._stage.data_ptr(), time_step_len.data_ptr(), N, W,
                                         out.data_ptr(), ws, nbytes, int(chunks), _stream(), self._copy_stream.cuda_stream))
        if wait_copy:
            # the caller may rewrite `host_data` as soon as this returns (a feeder recycling its ring slot): wait for the DMA --
            # not for the compute, which keeps running on the main stream
            self._copy_stream.synchronize()
        return out, self._stage

    def forward_pageable(self, host_data, pinned, time_step_len, chunks=4, host_threads=8, out=None):
        """host_data: C-contiguous f32 numpy array [N,W,32] in ORDINARY memory (the reference's np.array(...) per step); `pinned`: a
        page-locked f32 torch tensor with at least N*W*32 elements.  Range by range the library's host threads move the batch into
        `pinned`, DMA it and run the conv front end (crnn_forward_pageable).  Returns (logits, device data, copy stream)."""
        N, W, Hh = host_data.shape
        if Hh != 32:
            raise CrnnError("data must be [N, W, 32] (cfg.NUM_FEATURES = 32)")
        assert host_data.dtype == np.float32 and host_data.flags.c_contiguous and pinned.is_pinned() and pinned.numel() >= host_data.size
        T = W // 4 - 1
        if out is None:
            out = torch.empty((T, N, NCLASSES), dtype=torch.float32, device=self.device)
        if getattr(self, "_stage", None) is None or self._stage.shape != (N, W, 32):
            self._stage = torch.empty((N, W, 32), dtype=torch.float32, device=self.device)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        ws, nbytes = self._workspace(N, W)
        check(self.lib.crnn_forward_pageable(self.handle, host_data.ctypes.data, pinned.data_ptr(), self._stage.data_ptr(),
                                             time_step_len.data_ptr(), N, W, out.data_ptr(), ws, nbytes, int(chunks), int(host_threads),
                                             _stream(), self._copy_stream.cuda_stream))
        return out, self._stage, self._copy_stream

    def tap(self, name, N, W):
        """Intermediate of the last forward as f32 NHWC (tests only)."""
        H1, H2 = W // 2, W // 4
        shapes = {"conv1": (N, H1, 16, 64), "conv2": (N, H2, 8, 128), "conv3_1": (N, H2, 8, 256),
                  "conv3_2": (N, H2, 4, 256), "conv4_1": (N, H2, 4, 512), "conv4_2": (N, H2, 2, 512),
                  "conv5": (N, H2, 512), "lstm_out": (N, H2, 512), "xproj": (N, H2, 2048)}
        shp = shapes[name]
        dst = torch.empty(shp, dtype=torch.float32, device=self.device)
        ws, _ = self._workspace(N, W)
        check(self.lib.crnn_debug_tap(self.handle, name.encode(), dst.data_ptr(), dst.numel(), ws, _stream()))
        return dst

    # ---- loss / decode ------------------------------------------------------------------------
    def total_loss(self, costs):
        loss = torch.empty(1, dtype=torch.float32, device=self.device)
        check(self.lib.crnn_total_loss(self.handle, costs.data_ptr(), costs.numel(), loss.data_ptr(), _stream()))
        return loss


def ctc_loss(logits, flat_labels, label_len, input_len, blank=0, want_grad=False, grad_scale=1.0, max_label_len=None,
             costs=None, grad=None, validate=False):
    """logits [T,N,64] f32 cuda; integer tensors i32 cuda.  Returns (costs [N], grad [T,N,64] | None)."""
    lib = _lib.load()
    assert logits.is_cuda and logits.dtype == torch.float32 and logits.is_contiguous()
    T, N, C = logits.shape
    if label_len.numel() != N or input_len.numel() != N:
        raise CrnnError("ctc_loss: label_len and input_len must hold one entry per utterance")
    if validate:         # host-synchronising checks (the kernel itself rejects bad ids per sample: cost NaN, zero gradient)
        ll = label_len.cpu()
        if int(ll.min()) < 0 or int(ll.sum()) != flat_labels.numel():
            raise CrnnError("ctc_loss: label_len must be non-negative and sum to len(flat_labels)")
        if flat_labels.numel() and (int(flat_labels.min()) < 0 or int(flat_labels.max()) >= C or bool((flat_labels == blank).any())):
            raise CrnnError(f"ctc_loss: label ids must lie in [0, {C}) and differ from the blank ({blank})")
    if max_label_len is None:
        max_label_len = int(label_len.max().item()) if label_len.numel() else 0      # host sync; pass it to avoid
    if costs is None:
        costs = torch.empty(N, dtype=torch.float32, device=logits.device)
    if want_grad and grad is None:
        grad = torch.empty_like(logits)
    check(lib.crnn_ctc_loss(logits.data_ptr(), _ptr(grad) if want_grad else 0, flat_labels.data_ptr(),
                            label_len.data_ptr(), input_len.data_ptr(), T, N, C, blank, int(max_label_len),
                            float(grad_scale), costs.data_ptr(), 0, 0, _stream()))
    return costs, (grad if want_grad else None)


def ctc_greedy(logits, input_len, tf_blank=TF_BLANK, strip=0):
    """Returns (out [N,T] i32 zero padded, out_len [N] i32) on device."""
    lib = _lib.load()
    T, N, C = logits.shape
    out = torch.empty((N, T), dtype=torch.int32, device=logits.device)
    out_len = torch.empty(N, dtype=torch.int32, device=logits.device)
    check(lib.crnn_ctc_greedy(logits.data_ptr(), input_len.data_ptr(), T, N, C, tf_blank, strip, out.data_ptr(),
                              out_len.data_ptr(), _stream()))
    return out, out_len


def ctc_beam_search(logits, input_len, beam_width=100, merge_repeated=True, strip=0, num_threads=0):
    """The reference's decoder (network.py:656: ctc_beam_search_decoder, width 100, blank C-1, merge_repeated) on the HOST, as
    the TF op is.  logits: [T,N,C] f32 (cuda tensor -> copied back once, or numpy); returns (out [N,T] i32, out_len [N] i32,
    neg_log_prob [N] f32) as numpy arrays."""
    lib = _lib.load()
    x = logits.detach().float().cpu().numpy() if torch.is_tensor(logits) else np.asarray(logits, dtype=np.float32)
    x = np.ascontiguousarray(x)
    il = input_len.detach().cpu().numpy() if torch.is_tensor(input_len) else np.asarray(input_len)
    il = np.ascontiguousarray(il, dtype=np.int32)
    T, N, C = x.shape
    out = np.zeros((N, T), np.int32); out_len = np.zeros(N, np.int32); nlp = np.zeros(N, np.float32)
    check(lib.crnn_ctc_beam_search(x.ctypes.data, il.ctypes.data, T, N, C, int(beam_width), 1 if merge_repeated else 0, int(strip),
                                   out.ctypes.data, out_len.ctypes.data, nlp.ctypes.data, int(num_threads)))
    return out, out_len, nlp


def dense_decoded(out, out_len):
    """sparse_tensor_to_dense(default 0) shape [N, max_len] (network.py:657); one D2H sync."""
    m = int(out_len.max().item()) if out_len.numel() else 0
    return out[:, :m].contiguous()


def test_gemm_tn_bf16(A, B, block_n, k_splits=0):
    """D[M,N] = A[K,M]^T @ B[K,N] through the MN-major tcgen05 path (tests only)."""
    lib = _lib.load()
    K, M = A.shape
    Nc = B.shape[1]
    D = torch.zeros((M, Nc), dtype=torch.float32, device=A.device)
    check(lib.crnn_test_gemm_tn_bf16(A.data_ptr(), B.data_ptr(), D.data_ptr(), M, Nc, K, block_n, k_splits, _stream()))
    return D


def test_gemm_bf16(A, B, block_n):
    lib = _lib.load()
    M, K = A.shape
    Nc = B.shape[0]
    D = torch.empty((M, Nc), dtype=torch.float32, device=A.device)
    check(lib.crnn_test_gemm_bf16(A.data_ptr(), B.data_ptr(), D.data_ptr(), M, Nc, K, block_n, _stream()))
    return D
