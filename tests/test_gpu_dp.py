"""Data-parallel training step on 2 GPUs over NCCL (skipped when fewer than 2 devices are visible).
Checks the §8(e) contract: one SUM all-reduce of the flat gradient buffer, clip after the reduction, replicas stay
bit-identical, and with equal shards the reduced gradient equals the mean of the per-shard gradients."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from lstm_ctc_ocr_b200 import engine, parallel, synthetic
    dev = torch.device("cuda", rank)
    params = synthetic.init_params(3, logits_scale=10.0)
    data, lab, ll, tsl = synthetic.synth_batch(16, 88, seed=31, widths=[88, 85, 60, 33] * 4)
    d, l, n, t = parallel.shard_batch(data, lab, ll, tsl, rank, world)
    m = engine.CrnnModel(weight_decay=1e-5, device=dev)
    m.load_params(params)
    m.set_training(True)
    tt = lambda a: torch.tensor(a, device=dev)
    dd, dt = tt(d), tt(t)
    logits = m.forward(dd, dt)
    costs, grad = engine.ctc_loss(logits, tt(l), tt(n), dt, want_grad=True, grad_scale=1.0 / d.shape[0], max_label_len=int(n.max()))
    m.backward(dd, dt, grad)
    local = m.grads.clone()
    parallel.allreduce_sum_(m.grads)
    # the reduced buffer is the sum of both ranks' local gradients
    other = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(other, local)
    assert torch.allclose(m.grads, other[0] + other[1], rtol=1e-6, atol=1e-7)
    m.clip_adam_step(lr=1e-3, step=1, clip=10.0, grad_mul=1.0 / world, wd_mul=float(world))
    # replicas remain bit-identical after the step
    mine = m.params.clone()
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    assert torch.equal(gathered[0], gathered[1])
    assert not torch.equal(mine.cpu(), torch.tensor(np.concatenate([params[k].ravel() for k in m.table])))
    ret[rank] = float(m.last_grad_norm(1.0 / world))
    dist.destroy_process_group()


def test_two_gpu_training_step_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    mp.set_start_method("spawn", force=True)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29700 + (os.getpid() % 1000)
    procs = [mp.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert abs(ret[0] - ret[1]) < 1e-6 * max(1.0, ret[0])
