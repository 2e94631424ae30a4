"""GPU kernels against numbers held by the third-party projects behind the reference's loss / decode call sites
(lib/networks/network.py:653-657): TensorFlow's ctc_loss_op_test.py::testBasic (== warp-ctc's options_test) and
ctc_decoder_ops_test.py::testCTCGreedyDecoder, embedded in the 64-class layout the kernels take.  The vectors and their
provenance are in tests/golden/third_party_kats.py; the CPU suite checks the oracle (and the host beam-search decoder) against
the same file.  Everything goes through the C ABI (crnn_ctc_loss / crnn_ctc_greedy) via ctypes.

Tolerances: costs are published with 6 significant digits and the kernel recursion is f32 log2-space (ex2/lg2.approx), bound
rel 1e-4 as in test_gpu_parity.py; gradient abs 2e-4 (same bound as there for T <= 63); decode exact."""
import importlib.util
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"


def _kats():
    spec = importlib.util.spec_from_file_location("third_party_kats", os.path.join(ROOT, "tests", "golden", "third_party_kats.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("kernel", ["fast", "generic"])
def test_ctc_kernels_reproduce_the_tensorflow_and_warpctc_known_answer(kernel, monkeypatch):
    """warp-ctc numbering (blank 0, TF label k -> k+1), classes 6..63 at probability < 1e-26."""
    from lstm_ctc_ocr_b200 import engine
    K = _kats()
    monkeypatch.setenv("CRNN_CTC_KERNEL", kernel)
    monkeypatch.setenv("CRNN_CTC_RECUR", "log")
    x, flat, ll, il, cost, grad = K.ctc_case(num_classes=64, blank=0)
    t = lambda a: torch.tensor(a, device=DEV)
    c, g = engine.ctc_loss(t(x.astype(np.float32)), t(flat.astype(np.int32)), t(ll.astype(np.int32)), t(il.astype(np.int32)),
                           want_grad=True)
    c = c.cpu().numpy(); g = g.cpu().numpy()
    assert np.allclose(c, cost, rtol=1e-4, atol=0), (kernel, c)
    assert np.abs(g - grad).max() < 2e-4, (kernel, float(np.abs(g - grad).max()))


def test_greedy_kernel_reproduces_the_tensorflow_known_answer():
    """TF's blank (class 3 of 4) moved to class 63 (network.py:656 numbering); frames past seq_len are ignored; strip = -1 keeps
    class 0, which is an ordinary label to TF's decoder (the solver strips it afterwards, training.py:32)."""
    from lstm_ctc_ocr_b200 import engine
    K = _kats()
    x, il, want = K.greedy_case(num_classes=64, blank=63)
    out, out_len = engine.ctc_greedy(torch.tensor(x.astype(np.float32), device=DEV), torch.tensor(il.astype(np.int32), device=DEV),
                                     tf_blank=63, strip=-1)
    out = out.cpu().numpy(); out_len = out_len.cpu().numpy()
    assert [out[n, :out_len[n]].tolist() for n in range(2)] == want
    assert not out[0, out_len[0]:].any() and not out[1, out_len[1]:].any()          # zero padded (sparse_tensor_to_dense default)
    out, out_len = engine.ctc_greedy(torch.tensor(x.astype(np.float32), device=DEV), torch.tensor(il.astype(np.int32), device=DEV),
                                     tf_blank=63, strip=0)
    out = out.cpu().numpy(); out_len = out_len.cpu().numpy()
    assert [out[n, :out_len[n]].tolist() for n in range(2)] == [[1], [1, 1]]


def test_c_abi_from_plain_c_creates_the_model_on_the_gpu(tmp_path):
    """tests/c_abi/abi_smoke.c (C99, includes only include/crnn_ctc.h) run with --gpu: besides the GPU-free checks of the CPU
    suite it creates the model through the C ABI -- no Python, no torch in that process -- and reads back 24 trainable tensors
    / 7 158 592 parameters."""
    import subprocess
    from lstm_ctc_ocr_b200 import _lib
    _lib.load()
    libdir = os.path.join(ROOT, "lstm_ctc_ocr_b200")
    exe = str(tmp_path / "abi_smoke")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_abi", "abi_smoke.c"), "-o", exe, "-L" + libdir, "-lcrnnctc", "-lm", "-Wl,-rpath," + libdir]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-3000:]
    p = subprocess.run([exe, "--gpu"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "FAIL" not in p.stdout, p.stdout + p.stderr
