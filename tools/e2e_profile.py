"""Host-side profile of the end-to-end step (PrefetchFeeder -> Session.run(loss)) at the c3 workload: where the time between two
steps goes on the host.  Usage: python tools/e2e_profile.py [steps]"""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lstm_ctc_ocr_b200 import engine, synthetic                      # noqa: E402
from lstm_ctc_ocr_b200.lib.lstm.utils import gen as datagen          # noqa: E402
from lstm_ctc_ocr_b200.lib.networks.factory import get_network       # noqa: E402
from lstm_ctc_ocr_b200.session import Session                        # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    N, W = 1024, 256
    dev = torch.device("cuda:0")
    model = engine.CrnnModel(weight_decay=1e-5, device=dev)
    model.load_params(synthetic.init_params(3))
    net = get_network("LSTM_train")
    sess = Session(device=dev)
    sess._engines[id(net)] = model
    loss_h, _ = net.build_loss()
    arg_fn = lambda k: dict(k=k, batch_size=N, render=False, seed=3, rank=0, world=1, width=W, cache=4)
    feeder = datagen.PrefetchFeeder(arg_fn, num_workers=8, depth=4, max_width=W, batch_size=N, keep=2, warm=[arg_fn(k) for k in range(4)])
    marks = []

    def step():
        t0 = time.perf_counter()
        view, lab, ll, tsl = next(feeder)
        t1 = time.perf_counter()
        out = sess.run(loss_h, feed_dict={net.data: view, net.labels: np.asarray(lab, np.int32), net.time_step_len: np.asarray(tsl, np.int32),
                                          net.labels_len: np.asarray(ll, np.int32), net.keep_prob: 0.5})
        t2 = time.perf_counter()
        marks.append((t1 - t0, t2 - t1))
        return out
    try:
        sess.attach_feeder(feeder)
        for _ in range(64):
            step()
        torch.cuda.synchronize()
        del marks[:]
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / steps
        print("wall ms/step %.3f   next(feeder) ms %.3f   Session.run ms %.3f" % (wall * 1e3, 1e3 * np.mean([m[0] for m in marks]), 1e3 * np.mean([m[1] for m in marks])))
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(steps):
            step()
        pr.disable()
        buf = io.StringIO()
        pstats.Stats(pr, stream=buf).sort_stats("tottime").print_stats(28)
        print(buf.getvalue()[:6000])
    finally:
        sess.attach_feeder(None)
        feeder.close()


if __name__ == "__main__":
    main()
