"""A few training steps at the bench shape (for ncu launch lists of the backward pass)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lstm_ctc_ocr_b200 import engine, synthetic  # noqa: E402

N, W, steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 256, 3
dev = torch.device("cuda:0")
m = engine.CrnnModel(device=dev)
m.load_params(synthetic.init_params(3))
m.set_training(True)
data, lab, ll, tsl = synthetic.synth_batch(N, W, seed=3)
t = lambda a: torch.tensor(a, device=dev)
d, dl, dll, dt = t(data), t(lab), t(ll), t(tsl)
for s in range(steps):
    logits = m.forward(d, dt)
    costs, grad = engine.ctc_loss(logits, dl, dll, dt, want_grad=True, grad_scale=1.0 / N, max_label_len=int(ll.max()))
    m.backward(d, dt, grad)
    m.clip_adam_step(1e-4, s + 1)
torch.cuda.synchronize()
print("loss", float(costs.mean()))
