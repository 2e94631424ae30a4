"""Host-side data-parallel logic on CPU with the gloo backend, world_size 2 (tests only; NCCL is the runtime backend)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lstm_ctc_ocr_b200 import parallel, synthetic
    assert parallel.world_size() == world and parallel.rank() == rank
    # gradient exchange: SUM all-reduce of the flat buffer, then the 1/world factor applied by the optimizer kernel
    g = torch.full((1000,), float(rank + 1))
    parallel.allreduce_sum_(g)
    assert torch.all(g == 3.0)
    # parameter broadcast from rank 0
    p = torch.arange(10, dtype=torch.float32) * (rank + 1)
    parallel.broadcast_(p)
    assert torch.equal(p, torch.arange(10, dtype=torch.float32))
    # batch sharding keeps the flat-label bookkeeping consistent
    data, lab, ll, tsl = synthetic.synth_batch(8, 40, seed=2, widths=[40, 33, 17, 40, 8, 24, 40, 12])
    d, l, n, t = parallel.shard_batch(data, lab, ll, tsl, rank, world)
    assert d.shape[0] == 4 and int(n.sum()) == l.size
    offs = np.concatenate([[0], np.cumsum(ll)])
    assert np.array_equal(l, lab[offs[4 * rank]:offs[4 * rank + 4]]) and np.array_equal(t, tsl[4 * rank:4 * rank + 4])
    # mean-of-shard-means == global mean for equal shards (the loss / gradient scaling rule)
    x = torch.tensor(np.arange(8, dtype=np.float64))
    local = x[4 * rank:4 * rank + 4].mean().reshape(1)
    parallel.allreduce_sum_(local)
    assert abs(float(local) / world - float(x.mean())) < 1e-12
    ret[rank] = 1
    dist.destroy_process_group()


def test_data_parallel_host_logic_gloo_world2():
    mp.set_start_method("spawn", force=True)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    procs = [mp.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(ret) == {0: 1, 1: 1}


def test_shard_batch_rejects_uneven_split():
    import pytest
    from lstm_ctc_ocr_b200 import parallel, synthetic
    data, lab, ll, tsl = synthetic.synth_batch(6, 24, seed=1)
    with pytest.raises(ValueError):
        parallel.shard_batch(data, lab, ll, tsl, 0, 4)


def test_bucket_ranges_tile_the_flat_gradient_buffer():
    """parallel.bucket_ranges mirrors the 7 notifications of csrc/backward.cu: announcement order LSTM+logits first, conv1+conv2
    last; the ranges tile [0, 7 158 592) exactly once (24 tensors, 7 158 592 parameters)."""
    from collections import OrderedDict
    from lstm_ctc_ocr_b200 import parallel, synthetic
    p = synthetic.init_params(3)
    order = []
    for name in ("conv1", "conv2", "conv3_1", "conv3_2"):
        order += [f"{name}/weights", f"{name}/biases"]
    for name in ("conv4_1", "conv4_2"):
        order += [f"{name}/weights", f"{name}/biases", f"{name}/{name}/beta", f"{name}/{name}/gamma"]
    order += ["conv5/weights", "conv5/biases"]
    for d in ("fw", "bw"):
        order += [f"logits/bidirectional_rnn/{d}/lstm_cell/weights", f"logits/bidirectional_rnn/{d}/lstm_cell/biases"]
    order += ["logits/weights", "logits/biases"]
    assert sorted(order) == sorted(p)
    table, off = OrderedDict(), 0
    for k in order:
        table[k] = (off, p[k].shape)
        off += p[k].size
    assert off == 7158592
    r = parallel.bucket_ranges(table, off)
    assert len(r) == 7 and sum(c for _, c in r) == off
    assert r[0] == (table["logits/bidirectional_rnn/fw/lstm_cell/weights"][0], 1574912 + 32832)        # LSTM + logits, announced first
    assert r[-1] == (0, 640 + 73856)                                                                    # conv1 + conv2, announced last
    covered = sorted(r)
    assert covered[0][0] == 0 and all(a[0] + a[1] == b[0] for a, b in zip(covered, covered[1:])) and covered[-1][0] + covered[-1][1] == off
