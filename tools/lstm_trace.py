"""Timeline of the persistent LSTM kernel (debug): CRNN_LSTM_TRACE=1 makes crnn_forward print clock64 stamps of CTAs 0 and 5
for steps 8..11 to stderr (events: 0 step top, 1 TMA issued, 2 first K-block landed, 3 last K-block landed, 4 MMAs committed,
5 xproj loads issued, 6 accumulator ready, 7 first TMEM half loaded, 8 cell + stores issued, 9 fence.proxy.async done,
10 cluster arrive, 11 cluster wait done).  Usage: python tools/lstm_trace.py [N] [W]"""
import os
import sys

os.environ["CRNN_LSTM_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from lstm_ctc_ocr_b200 import engine, synthetic  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
W = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda:0")
m = engine.CrnnModel(device=dev)
m.load_params(synthetic.init_params(3))
data, lab, ll, tsl = synthetic.synth_batch(N, W, seed=3)
d, t = torch.tensor(data, device=dev), torch.tensor(tsl, device=dev)
for _ in range(2):
    m.forward(d, t)
torch.cuda.synchronize()
