"""CTC loss+gradient kernel alone at the C3 shape (T=63, N=1024, C=64): CUDA-event time per launch for the fast (S <= 32)
(tma = tensor-map tile load/store, fast = per-thread bulk row copies) and the generic kernel, rotating over input sets larger than L2.  Usage: python tools/ctc_bench.py [N] [T]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lstm_ctc_ocr_b200 import engine, synthetic  # noqa: E402


def run(N, T, max_len, kernel, reps=200, sets=10):
    os.environ["CRNN_CTC_KERNEL"] = kernel.split("-")[0]          # fast (default) | tma | generic
    os.environ["CRNN_CTC_RECUR"] = "me" if kernel.endswith("-me") else "log"
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    ll = rng.integers(min(4, max_len), max_len + 1, size=N).astype(np.int32)
    lab = rng.integers(1, 63, size=int(ll.sum())).astype(np.int32)
    il = np.full(N, T, dtype=np.int32)
    xs = [torch.randn(T, N, 64, device=dev) * 2 for _ in range(sets)]
    gs = [torch.empty(T, N, 64, device=dev) for _ in range(sets)]
    costs = torch.empty(N, device=dev)
    d_lab, d_ll, d_il = (torch.tensor(a, device=dev) for a in (lab, ll, il))
    mll = int(ll.max())
    for i in range(10):
        engine.ctc_loss(xs[i % sets], d_lab, d_ll, d_il, want_grad=True, grad_scale=1.0 / N, max_label_len=mll, costs=costs, grad=gs[i % sets])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        engine.ctc_loss(xs[i % sets], d_lab, d_ll, d_il, want_grad=True, grad_scale=1.0 / N, max_label_len=mll, costs=costs, grad=gs[i % sets])
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    nbytes = 2 * T * N * 64 * 4 + 4 * (lab.size + 2 * N)
    return {"kernel": kernel, "N": N, "T": T, "max_label_len": mll, "us": round(us, 2), "GBps": round(nbytes / us / 1e3, 1),
            "cost_sum": float(costs.sum().item())}


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 63
    for max_len in (6, 15):
        for k in ("fast", "fast-me", "tma", "tma-me", "generic"):
            print(json.dumps(run(N, T, max_len, k)))
