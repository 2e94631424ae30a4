"""End-to-end demo on the B200: train the CRNN from the reference initialisers on PIL-rendered text lines with the
reference-shaped solver (SolverWrapper.train_model), then
  * exact-match accuracy of the GPU greedy decode on held-out lines,
  * agreement of the GPU decode with the fp64 oracle's decode on the same trained weights (sequence equality).
Writes a JSON summary (loss curve, accuracy, agreement) to gpurun_out/train_demo.json."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lstm_ctc_ocr_b200.lib.lstm import train as T  # noqa: E402
from lstm_ctc_ocr_b200.lib.lstm.config import cfg  # noqa: E402
from lstm_ctc_ocr_b200.lib.lstm.utils import gen  # noqa: E402
from lstm_ctc_ocr_b200.lib.lstm.utils.training import accuracy_calculation  # noqa: E402
from lstm_ctc_ocr_b200.lib.networks.factory import get_network  # noqa: E402
from lstm_ctc_ocr_b200.session import Session  # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    n_eval = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    lr = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-4
    cfg.TRAIN.LEARNING_RATE, cfg.TRAIN.DISPLAY, cfg.TRAIN.SNAPSHOT_ITERS, cfg.TRAIN.WEIGHT_DECAY = lr, 500, 10 ** 9, 1e-5
    cfg.TRAIN.BATCH_SIZE, cfg.VAL.VAL_STEP, cfg.VAL.PRINT_NUM = 128, 10 ** 9, 0
    import random
    random.seed(3)
    np.random.seed(3)
    # pre-render a pool of batches on the host (rendering is the slow part), cycle through them
    t0 = time.time()
    npool = int(sys.argv[4]) if len(sys.argv) > 4 else 320
    pool = [next(gen.generator(batch_size=cfg.TRAIN.BATCH_SIZE, render=True)) for _ in range(npool)]
    held = [next(gen.generator(batch_size=128, render=True)) for _ in range(n_eval // 128)]
    print(f"rendered {len(pool)} train + {len(held)} eval batches in {time.time() - t0:.1f}s", flush=True)

    def cyc():
        k = 0
        while True:
            yield pool[k % len(pool)]
            k += 1
    net = get_network("LSTM_train")
    out = {"iters": iters, "batch": cfg.TRAIN.BATCH_SIZE, "lr": cfg.TRAIN.LEARNING_RATE}
    with Session(device="cuda:0") as sess:
        sw = T.SolverWrapper(sess, net, None, None, "/tmp/train_demo_out", "/tmp/train_demo_log")
        t0 = time.time()
        hist = sw.train_model(sess, iters + 1, restore=False, train_gen=cyc(), val_gen=cyc())
        torch.cuda.synchronize()
        out["train_seconds"] = time.time() - t0
        out["loss_curve"] = [round(float(np.mean(hist[i:i + 100])), 3) for i in range(0, len(hist), 100)]
        print("loss curve (mean of 100):", out["loss_curve"], flush=True)
        # ---- held-out accuracy + oracle agreement on the trained weights
        loss_h, dec_h = net.build_loss()
        params = sess.variables(net)
        from oracle import crnn_oracle as O      # checker only
        p64 = O.to_torch({k: v.astype(np.float64) for k, v in params.items()})
        # accuracy on lines the model was trained on (decode correctness) ...
        tr_ok = tr_n = 0
        for (imgs, lab, ll, tsl) in pool[:4]:
            feed = {net.data: np.array(imgs), net.labels: np.array(lab), net.time_step_len: np.array(tsl), net.labels_len: np.array(ll), net.keep_prob: 1.0}
            res = sess.run(dec_h, feed_dict=feed)
            org = sw.restoreLabel(lab, ll)
            tr_ok += accuracy_calculation(org, res, isPrint=False) * len(org); tr_n += len(org)
        out["train_pool_accuracy"] = tr_ok / tr_n
        # ... and on held-out renders, plus agreement with the fp64 oracle
        acc_n = agree = clear_n = total = 0
        for (imgs, lab, ll, tsl) in held:
            data = np.array(imgs)
            feed = {net.data: data, net.labels: np.array(lab), net.time_step_len: np.array(tsl), net.labels_len: np.array(ll), net.keep_prob: 1.0}
            res = sess.run(dec_h, feed_dict=feed)
            logits_gpu = sess.run(net.get_output("logits"), feed_dict=feed)
            org = sw.restoreLabel(lab, ll)
            acc_n += accuracy_calculation(org, res, isPrint=False) * len(org)
            # NOTE: BN uses batch statistics -> the oracle must see the same batch
            lo = O.forward(p64, data, np.array(tsl)).numpy()
            ref = O.greedy_decode(lo, np.array(tsl))
            srt = np.sort(lo, axis=2)
            margin = srt[:, :, -1] - srt[:, :, -2]
            err = np.abs(logits_gpu - lo).max(axis=2)
            for n in range(len(org)):
                got = [int(v) for v in res[n] if v != 0] if len(res) else []
                total += 1
                agree += int(got == ref[n])
                if np.all(margin[:tsl[n], n] > 2 * err[:tsl[n], n].max()):
                    clear_n += 1
                    assert got == ref[n], "decode differs from the oracle on a sample whose margins exceed the logit error"
        out.update(eval_lines=total, train_pool_lines=len(pool) * cfg.TRAIN.BATCH_SIZE, heldout_accuracy=acc_n / total, decode_agreement_with_oracle=agree / total, clear_margin_lines=clear_n,
                   clear_margin_agreement=1.0)
    print(json.dumps(out), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"train_demo_lr{lr:g}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
