"""e2e probe: Session.run(loss, feed_dict=host numpy) per-step wall time for several H2D chunk counts, after the rotating host
buffers have been page-locked (3 sightings).  Usage: python tools/e2e_probe.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lstm_ctc_ocr_b200 import engine, synthetic  # noqa: E402
from lstm_ctc_ocr_b200.lib.networks.factory import get_network  # noqa: E402
from lstm_ctc_ocr_b200.session import Session  # noqa: E402

N, W = 1024, 256
dev = torch.device("cuda:0")
model = engine.CrnnModel(weight_decay=1e-5, device=dev)
model.load_params(synthetic.init_params(3))
batches = [synthetic.synth_batch(N, W, seed=3 + i) for i in range(5)]
net = get_network("LSTM_train")
sess = Session(device=dev)
sess._engines[id(net)] = model
loss_h, _ = net.build_loss()


def run(i):
    data, lab, ll, tsl = batches[i % 5]
    return sess.run(loss_h, feed_dict={net.data: data, net.labels: lab, net.time_step_len: tsl, net.labels_len: ll, net.keep_prob: 0.5})


for chunks in (1, 1, 2, 4, 8, 4, 1):
    sess.h2d_chunks = chunks
    for i in range(15):
        run(i)
    torch.cuda.synchronize()
    ts = []
    for i in range(20):
        t0 = time.perf_counter()
        run(i)
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"chunks={chunks}: median {np.median(ts):.3f} ms  min {min(ts):.3f}  max {max(ts):.3f}  -> {N / np.median(ts) * 1e3:.0f} img/s")

# host-side share: the same call with the GPU work already queued is bounded below by pure Python/ctypes time
data, lab, ll, tsl = batches[0]
d, t = torch.tensor(data, device=dev), torch.tensor(tsl, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    model.forward(d, t)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"device-resident forward: host launch time {(t1 - t0) / 20 * 1e3:.3f} ms/step, total {(t2 - t0) / 20 * 1e3:.3f} ms/step")
