"""Host beam-search decoder (crnn_ctc_beam_search, csrc/beam.cpp) throughput on this machine's cores: lines/s at the benchmark
shape (T = 63 frames, 64 classes, beam width 100) on peaked, soft and flat frames.  CPU only -- the reference's decoder
(tf.nn.ctc_beam_search_decoder, network.py:656) is a host op too.  Usage: python tools/beam_bench.py [N] [threads ...]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lstm_ctc_ocr_b200 import engine  # noqa: E402


def frames(kind, T, N, rng):
    if kind == "flat":
        return (rng.standard_normal((T, N, 64)) * 0.3).astype(np.float32)
    path = rng.choice(64, size=(T, N), p=np.r_[0.25, np.full(62, 0.65 / 62), 0.10])
    x = rng.standard_normal((T, N, 64)).astype(np.float32)
    x[np.arange(T)[:, None], np.arange(N)[None, :], path] += 10.0 if kind == "peaked" else 4.0
    return x


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    threads = [int(a) for a in sys.argv[2:]] or sorted({1, min(8, os.cpu_count() or 1), os.cpu_count() or 1})
    T = 63
    rng = np.random.default_rng(0)
    il = np.full(N, T, np.int32)
    res = {"T": T, "C": 64, "beam_width": 100, "lines": N, "host_cpus": os.cpu_count(), "lines_per_s": {}}
    for kind in ("peaked", "soft", "flat"):
        x = frames(kind, T, N, rng)
        engine.ctc_beam_search(x[:, :32], il[:32])                 # warm the allocator
        for nt in threads:
            t0 = time.perf_counter()
            engine.ctc_beam_search(x, il, num_threads=nt)
            res["lines_per_s"][f"{kind}/{nt}t"] = round(N / (time.perf_counter() - t0), 1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
