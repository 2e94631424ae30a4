"""Synthetic batches with the data layer's tensor contract and the reference's initialisers.

* ``synth_batch``: what ``groupBatch`` hands the solver (reference lib/lstm/utils/gen.py:41-67): data [N,W,32] f32 in
  [0,1) with exact-zero right padding beyond each sample's width, flat labels in 1..62, label_len U{4..6}
  (config.py:24-25), ``time_step_len = nw//4 - 1`` (gen.py:54).  Generator: numpy PCG64.
* ``init_params``: the reference's initialisers by TF variable name (network.py:168-169,119-120; TF defaults for the
  LSTM cell): xavier-uniform conv kernels, zero biases, BN gamma 1 / beta 0, glorot-uniform LSTM matrices,
  variance-scaling(0.01, FAN_AVG, truncated normal) logits matrix."""
import math
from collections import OrderedDict

import numpy as np

CHARSET_LEN = 62
CONVS = [("conv1", 3, 3, 1, 64, False), ("conv2", 3, 3, 64, 128, False), ("conv3_1", 3, 3, 128, 256, False),
         ("conv3_2", 3, 3, 256, 256, False), ("conv4_1", 3, 3, 256, 512, True), ("conv4_2", 3, 3, 512, 512, True),
         ("conv5", 2, 2, 512, 512, False)]


def synth_batch(N, W, seed=3, widths=None, min_len=4, max_len=6):
    assert W % 4 == 0
    rng = np.random.Generator(np.random.PCG64(seed))
    widths = np.full((N,), W, dtype=np.int64) if widths is None else np.asarray(widths, dtype=np.int64)
    data = rng.random((N, W, 32)).astype(np.float32)
    data = np.where(np.arange(W)[None, :, None] < widths[:, None, None], data, 0).astype(np.float32)
    label_len = rng.integers(min_len, max_len + 1, size=N).astype(np.int32)
    labels = rng.integers(1, CHARSET_LEN + 1, size=int(label_len.sum())).astype(np.int32)
    time_step_len = (widths // 4 - 1).astype(np.int32)
    return data, labels, label_len, time_step_len


def synth_bucket_batch(N, W, seed=3, buckets=(80, 160, 256), min_len=4, max_len=6):
    """One width-bucketed batch (BASELINE configs[3]): true widths uniform in (previous bucket, W], at least one
    sample of width exactly W, everything else as ``synth_batch``."""
    lo = max([b for b in buckets if b < W] or [0])
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    widths = rng.integers(max(lo + 1, 8), W + 1, size=N)
    widths[rng.integers(0, N)] = W
    return synth_batch(N, W, seed=seed, widths=widths, min_len=min_len, max_len=max_len)


def init_params(seed=3, logits_scale=1.0):
    rng = np.random.Generator(np.random.PCG64(seed))
    p = OrderedDict()
    for name, kh, kw, ci, co, bn in CONVS:
        lim = math.sqrt(6.0 / (kh * kw * ci + kh * kw * co))
        p[f"{name}/weights"] = rng.uniform(-lim, lim, size=(kh, kw, ci, co)).astype(np.float32)
        p[f"{name}/biases"] = np.zeros((co,), np.float32)
        if bn:
            p[f"{name}/{name}/beta"] = np.zeros((co,), np.float32)
            p[f"{name}/{name}/gamma"] = np.ones((co,), np.float32)
    for d in ("fw", "bw"):
        lim = math.sqrt(6.0 / (768 + 1024))
        p[f"logits/bidirectional_rnn/{d}/lstm_cell/weights"] = rng.uniform(-lim, lim, size=(768, 1024)).astype(np.float32)
        p[f"logits/bidirectional_rnn/{d}/lstm_cell/biases"] = np.zeros((1024,), np.float32)
    std = math.sqrt(1.3 * 0.01 / ((512 + 64) / 2.0))
    p["logits/weights"] = (np.clip(rng.normal(0.0, std, size=(512, 64)), -2 * std, 2 * std) * logits_scale).astype(np.float32)
    p["logits/biases"] = np.zeros((64,), np.float32)
    return p
