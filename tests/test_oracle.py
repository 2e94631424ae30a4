"""Oracle self-consistency (CPU only): every restated op is cross-checked against an
independent implementation, since the reference ships no golden vectors."""
import itertools
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import crnn_oracle as O


def brute_force_ctc(logits_tc, lab, blank=0):
    """-log sum over all alignments that collapse to lab (T<=5, C small)."""
    T, C = logits_tc.shape
    lp = logits_tc - np.log(np.exp(logits_tc).sum(1, keepdims=True))
    tot = 0.0
    for path in itertools.product(range(C), repeat=T):
        col, prev = [], -1
        for a in path:
            if a != prev and a != blank:
                col.append(a)
            prev = a
        if col == list(lab):
            tot += math.exp(sum(lp[t, a] for t, a in enumerate(path)))
    return -math.log(tot) if tot > 0 else float("inf")


@pytest.mark.parametrize("lab", [[1], [1, 2], [2, 2], [1, 2, 1]])
def test_ctc_brute_force(lab):
    rng = np.random.default_rng(0)
    T, C = 5, 4
    x = rng.standard_normal((T, 1, C))
    costs, _ = O.ctc_loss_np(x, lab, [len(lab)], [T])
    assert abs(costs[0] - brute_force_ctc(x[:, 0], lab)) < 1e-10


def test_ctc_infeasible_returns_zero():
    x = np.random.default_rng(1).standard_normal((2, 1, 4))
    costs, g = O.ctc_loss_np(x, [2, 2], [2], [2])      # needs 3 frames
    assert costs[0] == 0.0 and np.all(g == 0)


def test_ctc_vs_torch_and_grad():
    rng = np.random.default_rng(2)
    T, N, C = 21, 6, 64
    x = rng.standard_normal((T, N, C)) * 2
    ll = np.array([4, 5, 6, 4, 6, 5]); il = np.array([21, 20, 13, 21, 17, 12])
    lab = rng.integers(1, 63, size=ll.sum()); lab[1] = lab[0]      # a repeat
    costs, grad = O.ctc_loss_np(x, lab, ll, il)
    xt = torch.tensor(x, requires_grad=True)
    ct = O.ctc_loss_torch(xt, lab, ll, il)
    ct.sum().backward()
    assert np.allclose(costs, ct.detach().numpy(), atol=1e-9)
    assert np.allclose(grad, xt.grad.numpy(), atol=1e-9)
    # frames >= input_len get zero gradient
    assert np.all(grad[13:, 2] == 0)


def test_lstm_vs_torch_nn_lstm():
    rng = np.random.default_rng(3)
    N, T, D, H = 3, 7, 512, 256
    x = torch.tensor(rng.standard_normal((N, T, D)) * 0.5)
    w = torch.tensor(rng.standard_normal((D + H, 4 * H)) * 0.05)
    b = torch.tensor(rng.standard_normal((4 * H,)) * 0.1)
    lens = np.array([7, 4, 1])
    for reverse in (False, True):
        y = O.lstm_direction(x, lens, w, b, reverse)
        # torch gate order i,f,g,o ; TF i,j(g),f,o ; forget bias +1
        perm = torch.cat([torch.arange(0, H), torch.arange(2 * H, 3 * H), torch.arange(H, 2 * H), torch.arange(3 * H, 4 * H)])
        lstm = torch.nn.LSTM(D, H, batch_first=True).double()
        with torch.no_grad():
            lstm.weight_ih_l0.copy_(w[:D, perm].t()); lstm.weight_hh_l0.copy_(w[D:, perm].t())
            bb = b.clone(); bb[2 * H:3 * H] += 1.0
            lstm.bias_ih_l0.copy_(bb[perm]); lstm.bias_hh_l0.zero_()
        for n in range(N):
            L = int(lens[n])
            xs = x[n:n + 1, :L]
            if reverse:
                xs = xs.flip(1)
            ref, _ = lstm(xs)
            if reverse:
                ref = ref.flip(1)
            assert torch.allclose(y[n, :L], ref[0], atol=1e-10)
            assert torch.all(y[n, L:] == 0)


def test_forward_shapes_and_padding_contract():
    p = O.to_torch(O.randomize_params(O.init_params(3)))
    data, lab, ll, tsl = O.synth_batch(2, 88, seed=5, widths=[85, 60])
    assert tsl.tolist() == [85 // 4 - 1, 60 // 4 - 1]
    assert np.all(data[1, 60:] == 0)
    logits, acts = O.forward(p, data, tsl, return_all=True)
    assert logits.shape == (88 // 4 - 1, 2, 64)
    assert acts["conv1"].shape == (2, 64, 44, 16)
    assert acts["conv3_2"].shape == (2, 256, 22, 4)
    assert acts["conv4_2"].shape == (2, 512, 22, 2)
    assert acts["reshaped_layer"].shape == (2, 21, 512)
    # frames past len: LSTM output zero -> logits == bias
    b = p["logits/biases"]
    assert torch.allclose(logits[tsl[1]:, 1], b.expand(21 - tsl[1], 64))


def test_bn_matches_functional():
    rng = np.random.default_rng(4)
    x = torch.tensor(rng.standard_normal((3, 8, 5, 4)))
    w = torch.tensor(rng.standard_normal((3, 3, 8, 6)) * 0.2)
    b = torch.tensor(rng.standard_normal(6)); beta = torch.tensor(rng.standard_normal(6)); gamma = torch.tensor(rng.standard_normal(6))
    y, _ = O.conv_single(x, w, b, (beta, gamma), relu=False)
    z = F.conv2d(x, w.permute(3, 2, 0, 1), b, padding=1)
    ref = F.batch_norm(z, None, None, gamma, beta, training=True, eps=1e-3)
    assert torch.allclose(y, ref, atol=1e-12)


def test_greedy_rule_table():
    C = 64
    def onehot(seq):
        x = np.zeros((len(seq), 1, C)); 
        for t, a in enumerate(seq): x[t, 0, a] = 5.0
        return x
    dec = lambda seq, L=None: O.greedy_decode(onehot(seq), [L or len(seq)])[0]
    assert dec([5, 5, 63, 5, 0, 7]) == [5, 5, 7]          # repeat merged, 63 splits, 0 stripped
    assert dec([63, 63, 63]) == []
    assert dec([0, 0, 3, 3, 0, 3]) == [3, 3]              # 0 acts as separator (raw prev), then stripped
    assert dec([1, 2, 3, 4], 2) == [1, 2]                 # honours input_len
    x = np.zeros((1, 1, C)); assert O.greedy_decode(x, [1])[0] == []   # tie -> lowest index 0 -> stripped


def _peaked_lines(n, T, seed, margin=6.0):
    """Frames like the 10k-line GPU decode test: peaked at a path with CTC blanks (0), decoder blanks (63) and repeats."""
    rng = np.random.default_rng(seed)
    path = rng.choice(64, size=(T, n), p=np.r_[0.25, np.full(62, 0.65 / 62), 0.10])
    rep = rng.random((T, n)) < 0.3
    for t in range(1, T):
        path[t] = np.where(rep[t], path[t - 1], path[t])
    x = rng.standard_normal((T, n, 64))
    x[np.arange(T)[:, None], np.arange(n)[None, :], path] += margin
    return x


def test_beam_search_restatement_rule_table_and_defined_deviation():
    """What the reference calls (beam search, width 100, merge_repeated=True) vs the greedy rule the product implements:
    identical on peaked frames EXCEPT that the beam decoder also collapses consecutive equal labels of the decoded sequence
    (a genuine double letter 5,<63>,5 comes out as a single 5) -- the defined deviation recorded in DESIGN.md section 2."""
    C = 64
    def onehot(seq):
        x = np.zeros((len(seq), 1, C))
        for t, a in enumerate(seq): x[t, 0, a] = 8.0
        return x
    beam = lambda seq, **kw: O.beam_search_decode(onehot(seq), [len(seq)], **kw)[0]
    assert beam([1, 2, 3, 4]) == [1, 2, 3, 4]
    assert beam([63, 63, 63]) == []
    assert beam([5, 5, 63, 5, 0, 7], merge_repeated=False) == [5, 5, 7] == O.greedy_decode(onehot([5, 5, 63, 5, 0, 7]), [6])[0]
    assert beam([5, 5, 63, 5, 0, 7]) == [5, 7]                       # the double 5 is merged away by merge_repeated=True
    assert beam([3, 63, 3, 63, 4]) == [3, 4]
    # peaked random lines: beam without the output merge == greedy; with it == greedy followed by the same collapse
    x = _peaked_lines(12, 19, seed=2)
    il = [19] * 12
    g = O.greedy_decode(x, il, strip=-1)                             # keep class 0: it is an ordinary label to TF's decoder
    b_plain = O.beam_search_decode(x, il, merge_repeated=False, strip=-1)
    b_merge = O.beam_search_decode(x, il, merge_repeated=True, strip=-1)
    collapse = lambda s: [v for i, v in enumerate(s) if i == 0 or v != s[i - 1]]
    assert b_plain == g
    assert b_merge == [collapse(s) for s in g]


def test_split_bf16_products_reproduce_the_f32_conv1():
    """The arithmetic csrc/conv1_tc.cuh puts on the bf16 tensor pipe -- x = xh + xl, w = wh + wl (bf16 high part + bf16
    remainder), products xh*wh + xl*wh + xh*wl accumulated in f32 -- emulated on the CPU: it reproduces the f32 conv1 to ~2^-16
    of max |out|, while plain bf16 operands are ~50x worse (which is why the kernel splits)."""
    import torch.nn.functional as F
    pn = O.randomize_params(O.init_params(3, dtype=np.float32))
    data, _, _, _ = O.synth_batch(3, 100, seed=5)
    x = torch.tensor(data, dtype=torch.float64)[:, None]
    w = torch.tensor(pn["conv1/weights"], dtype=torch.float64).permute(3, 2, 0, 1)
    bf = lambda t: t.float().bfloat16().double()
    ref = F.conv2d(x, w, None, padding=1)
    xh, wh = bf(x), bf(w)
    xl, wl = bf(x - xh), bf(w - wh)
    split = F.conv2d(xh, wh, None, padding=1) + F.conv2d(xl, wh, None, padding=1) + F.conv2d(xh, wl, None, padding=1)
    plain = F.conv2d(xh, wh, None, padding=1)
    rel = lambda a: float((a - ref).abs().max() / ref.abs().max())
    assert rel(split) < 3e-5
    assert rel(plain) > 20 * rel(split)


def test_accuracy_calculation():
    assert O.accuracy_calculation([[1, 2, 0], [3]], [[1, 2], [3, 0, 0]]) == 1.0
    assert O.accuracy_calculation([[1, 2], [3]], [[1, 2], [4]]) == 0.5
    assert O.accuracy_calculation([[1]], [[1], [2]]) == 0


def test_adam_and_clip_hand_computed():
    p = {"a": torch.tensor([1.0, 2.0], dtype=torch.float64)}
    g = {"a": torch.tensor([30.0, 40.0], dtype=torch.float64)}          # norm 50 -> scale 0.2
    cg, gn = O.clip_by_global_norm(g, 10.0)
    assert abs(gn - 50) < 1e-12 and torch.allclose(cg["a"], torch.tensor([6.0, 8.0], dtype=torch.float64))
    m = {"a": torch.zeros(2, dtype=torch.float64)}; v = {"a": torch.zeros(2, dtype=torch.float64)}
    p2, m, v = O.adam_step(dict(p), cg, m, v, 1, lr=1e-4)
    lr_t = 1e-4 * math.sqrt(1 - 0.999) / (1 - 0.9)
    exp0 = 1.0 - lr_t * 0.6 / (math.sqrt(0.001 * 36) + 1e-8)
    assert abs(float(p2["a"][0]) - exp0) < 1e-15


def test_finite_difference_full_graph():
    """d loss / d param via autograd == central differences (fp64), tiny batch."""
    pn = O.randomize_params(O.init_params(3, logits_scale=20.0))
    batch = O.synth_batch(2, 24, seed=9, min_len=1, max_len=2)
    out = O.train_step(pn, batch, wd=1e-2)
    for name, idx in [("conv5/biases", (3,)), ("logits/weights", (7, 5)), ("conv4_1/conv4_1/gamma", (10,)),
                      (O.LSTM_BW + "/weights", (600, 300)), ("conv2/weights", (1, 1, 3, 4))]:
        eps = 1e-5
        vals = []
        for sgn in (+1, -1):
            q = {k: v.copy() for k, v in pn.items()}
            q[name][idx] += sgn * eps
            loss, *_ = O.build_loss(O.to_torch(q), *[batch[0], batch[1], batch[2], batch[3]], wd=1e-2)
            vals.append(float(loss))
        fd = (vals[0] - vals[1]) / (2 * eps)
        an = float(out["grads"][name][idx])
        assert abs(fd - an) <= 1e-6 * max(1.0, abs(an)), (name, fd, an)


# ---------------------------------------------------------------------------------------------------------------------------
# Known answers held by the third-party projects behind network.py:653-657 (TensorFlow's and warp-ctc's own unit tests):
# the one place the oracle is pinned to numbers it did not produce itself.  tests/golden/third_party_kats.py has provenance.
# ---------------------------------------------------------------------------------------------------------------------------
def _kats():
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "third_party_kats.py")
    spec = importlib.util.spec_from_file_location("third_party_kats", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("blank,C", [(5, 6), (0, 6), (0, 64)])
def test_third_party_known_answers_ctc_loss_and_gradient(blank, C):
    """tf.nn.ctc_loss testBasic == warp-ctc options_test: costs to the 6 published digits, gradient w.r.t. the unnormalised
    activations to 1e-6, in TF's class numbering (blank 5), in warp-ctc's default numbering (blank 0, network.py:653) and
    embedded in the 64-class layout the product kernels take."""
    K = _kats()
    x, flat, ll, il, cost, grad = K.ctc_case(num_classes=C, blank=blank)
    costs, g = O.ctc_loss_np(x, flat, ll, il, blank=blank)
    assert np.allclose(costs, cost, rtol=0, atol=6e-6), costs          # published with 6 significant digits
    assert np.abs(g - grad).max() < 2e-6
    # the differentiable restatement used for whole-graph autograd agrees too
    xt = torch.tensor(x, requires_grad=True)
    ct = O.ctc_loss_torch(xt, flat, ll, il, blank=blank)
    ct.sum().backward()
    assert np.allclose(ct.detach().numpy(), cost, rtol=0, atol=6e-6)
    assert np.abs(xt.grad.numpy() - grad).max() < 2e-6


def test_third_party_known_answers_greedy_decoder():
    """tf.nn.ctc_greedy_decoder testCTCGreedyDecoder (merge_repeated=True, frames past seq_len ignored), with TF's blank in its
    own place (class 3 of 4) and moved to class 63 of 64 (network.py:656 numbering)."""
    K = _kats()
    for C, blank in ((4, 3), (64, 63)):
        x, il, want = K.greedy_case(num_classes=C, blank=blank)
        assert O.greedy_decode(x, il, tf_blank=blank, strip=-1) == want
    # the solver's zero stripping (training.py:32) on top of it
    x, il, want = K.greedy_case()
    assert O.greedy_decode(x, il, tf_blank=3, strip=0) == [[1], [1, 1]]


def test_third_party_known_answers_beam_search_decoder():
    """tf.nn.ctc_beam_search_decoder testCTCDecoderBeamSearch: at beam_width 2 TF's top path is [1, 0] -- NOT the most
    probable labelling [0, 1, 0], which every other width returns -- so the vector pins the candidate ordering and the
    eviction rule of the restatement, not just its probabilities."""
    K = _kats()
    x, il = K.beam_case()
    assert O.beam_search_decode(x, il, beam_width=K.BEAM_WIDTH, merge_repeated=True, strip=-1) == [K.BEAM_TOP_PATHS[0]]
    for bw in (1, 3, 100):
        assert O.beam_search_decode(x, il, beam_width=bw, merge_repeated=True, strip=-1) == [K.BEAM_TOP_PATHS[1]]


def test_third_party_known_answer_lstm_cell():
    """TensorFlow's rnn_cell_test.py::testBasicLSTMCell: c' = sigmoid(f + 1) c + sigmoid(i) tanh(j), h' = sigmoid(o) tanh(c')
    with [x, h] @ W -- the cell lstm_direction steps (network.py:104-107), to the published f32 digits."""
    K = _kats()
    w = torch.full((4, 8), K.LSTM_WEIGHT, dtype=torch.float64)
    b = torch.zeros(8, dtype=torch.float64)
    s0 = torch.full((1, 2), K.LSTM_STATE0, dtype=torch.float64)
    c1, h1 = O.lstm_cell(torch.tensor(K.LSTM_X), s0, s0, w, b)
    c2, h2 = O.lstm_cell(h1, s0, s0, w, b)
    assert np.allclose(h2.numpy(), K.LSTM_OUT, rtol=0, atol=2e-7)
    assert np.allclose(torch.cat([c1, h1, c2, h2], 1).numpy(), K.LSTM_STATE, rtol=0, atol=2e-7)


def test_third_party_known_answers_conv_pool_layouts_and_global_norm_clip():
    """TensorFlow's conv2d (NHWC x HWIO, VALID: conv5's kind and the checkpoint's kernel layout), max_pool and
    clip_by_global_norm known answers against the oracle's conv_single / pooling call / clip_by_global_norm."""
    K = _kats()
    x = torch.tensor(K.CONV_INPUT_NHWC).permute(0, 3, 1, 2)                      # the oracle computes in NCHW
    zero = torch.zeros(3, dtype=torch.float64)
    for w, want in ((K.CONV_1X1_FILTER_HWIO, K.CONV_1X1_EXPECTED), (K.CONV_2X2_FILTER_HWIO, K.CONV_2X2_EXPECTED)):
        y, _ = O.conv_single(x, torch.tensor(w), zero, bn=None, relu=False, padding="VALID")
        assert y.permute(0, 2, 3, 1).reshape(-1).tolist() == want
    xp = torch.tensor(K.POOL_INPUT_NHWC).permute(0, 3, 1, 2)
    assert F.max_pool2d(xp, (2, 2), (2, 2)).permute(0, 2, 3, 1).reshape(-1).tolist() == K.POOL_2X2_S2_VALID_EXPECTED
    g = {"x0": torch.tensor(K.CLIP_X0), "x1": torch.tensor(K.CLIP_X1)}
    for clip, want in ((4.0, K.CLIP_AT_4), (6.0, K.CLIP_AT_6)):
        out, gn = O.clip_by_global_norm(g, clip=clip)
        assert gn == K.CLIP_GLOBAL_NORM
        assert np.allclose(out["x0"].numpy(), want[0], rtol=0, atol=1e-15) and np.allclose(out["x1"].numpy(), want[1], rtol=0, atol=1e-15)
