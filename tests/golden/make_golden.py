"""Generates tests/golden/crnn_n4_w88.npz from the fp64 oracle (the reference itself cannot run here: TF 1.0.1 and
warp-ctc are not installable -> PARITY UNPINNED, see oracle/crnn_oracle.py header).

    python tests/golden/make_golden.py

Inputs and parameters are regenerated from seeds by lstm_ctc_ocr_b200.synthetic; the fixture stores a checksum of
them plus the oracle's outputs (logits, per-sample CTC costs, total loss, greedy decode, per-layer mean/abs-mean)."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lstm_ctc_ocr_b200 import synthetic  # noqa: E402
from oracle import crnn_oracle as O  # noqa: E402

N, W, WIDTHS, SEED_P, SEED_D, LOGITS_SCALE, WD = 4, 88, [88, 85, 60, 33], 3, 5, 10.0, 1e-5


def inputs():
    params = synthetic.init_params(SEED_P, logits_scale=LOGITS_SCALE)
    rng = np.random.Generator(np.random.PCG64(11))
    for k in params:                         # biases / BN affine are 0/1 at init: perturb so they matter
        if k.endswith("/biases") or k.endswith("/beta") or k.endswith("/gamma"):
            params[k] = (params[k] + 0.1 * rng.standard_normal(params[k].shape)).astype(np.float32)
    batch = synthetic.synth_batch(N, W, seed=SEED_D, widths=WIDTHS)
    h = hashlib.sha256()
    for k in params:
        h.update(params[k].tobytes())
    for a in batch:
        h.update(np.ascontiguousarray(a).tobytes())
    return params, batch, h.hexdigest()


def main():
    params, (data, lab, ll, tsl), digest = inputs()
    p64 = O.to_torch({k: v.astype(np.float64) for k, v in params.items()})
    logits, acts = O.forward(p64, data, tsl, return_all=True)
    costs, grad = O.ctc_loss_np(logits.numpy(), lab, ll, tsl)
    loss = costs.mean() + float(O.l2_reg(p64, WD))
    dec = O.dense_decoded(O.greedy_decode(logits.numpy(), tsl))
    layer_stats = {}
    for name in ["conv1", "conv2", "conv3_1", "conv3_2", "conv4_1", "conv4_2", "reshaped_layer", "lstm_out"]:
        a = acts[name].numpy()
        layer_stats[name] = np.array([a.mean(), np.abs(a).mean(), np.abs(a).max()])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "crnn_n4_w88.npz")
    np.savez_compressed(out, digest=np.array(digest), logits=logits.numpy().astype(np.float32), costs=costs, loss=np.array(loss),
                        decoded=dec, ctc_grad=grad.astype(np.float32),
                        **{"stat_" + k: v for k, v in layer_stats.items()})
    print("wrote", out, "digest", digest[:16], "loss", loss)


if __name__ == "__main__":
    main()
