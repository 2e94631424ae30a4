"""GPU diagnostic for the backward pass: per-tensor gradient error vs the fp64 oracle (autograd)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lstm_ctc_ocr_b200 import engine  # noqa: E402
from oracle import crnn_oracle as O  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    # ---- TN GEMM unit checks
    for bn, (K, M, Nc, ks) in [(64, (1000, 512, 64, 0)), (128, (300, 64, 128, 1)), (256, (5000, 256, 512, 0)), (256, (64, 128, 256, 1))]:
        g = torch.Generator().manual_seed(K)
        A = (torch.randn(K, M, generator=g) * 0.5).to(torch.bfloat16).to(dev)
        B = (torch.randn(K, Nc, generator=g) * 0.5).to(torch.bfloat16).to(dev)
        D = engine.test_gemm_tn_bf16(A, B, bn, ks)
        torch.cuda.synchronize()
        ref = A.float().t() @ B.float()
        e = float((D - ref).abs().max() / ref.abs().max())
        print(f"gemm_tn bn={bn} K={K} M={M} N={Nc} splits={ks} rel_err={e:.3e}", flush=True)
    # ---- full backward
    for (N, W, widths) in [(4, 88, [88, 85, 60, 33]), (130, 40, None)]:
        pn = O.randomize_params(O.init_params(3, dtype=np.float32, logits_scale=10.0))
        batch = O.synth_batch(N, W, seed=5, widths=widths)
        data, lab, ll, tsl = batch
        out = O.train_step({k: v.astype(np.float64) for k, v in pn.items()}, batch, wd=0.0)
        m = engine.CrnnModel(weight_decay=0.0)
        m.load_params(pn)
        m.set_training(True)
        t = lambda a: torch.tensor(a, device=dev)
        d_data, d_tsl = t(data), t(tsl)
        logits = m.forward(d_data, d_tsl)
        costs, grad = engine.ctc_loss(logits, t(lab), t(ll), d_tsl, want_grad=True, grad_scale=1.0 / N, max_label_len=int(ll.max()))
        m.backward(d_data, d_tsl, grad)
        torch.cuda.synchronize()
        print(f"N={N} W={W}: loss gpu {float(costs.mean()):.5f} oracle {out['loss']:.5f}")
        for name in reversed(list(m.table)):
            g = m.grad_tensor(name).cpu().numpy().astype(np.float64)
            go = out["grads"][name].numpy()
            rel = np.linalg.norm(g - go) / max(np.linalg.norm(go), 1e-30)
            cos = float((g * go).sum() / max(np.linalg.norm(g) * np.linalg.norm(go), 1e-30))
            print(f"  {name:55s} |g|={np.linalg.norm(go):.3e} rel_l2={rel:.3e} cos={cos:.5f}", flush=True)
        del m


if __name__ == "__main__":
    main()
