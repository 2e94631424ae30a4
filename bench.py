#!/usr/bin/env python
"""bench.py -- images/sec of the CRNN forward + CTC loss hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2shape|c1shape] [--impl ours|reference]

One "step" = conv stack -> BiLSTM -> logits -> CTC loss (+gradient, as warp-ctc's forward op computes it)
-> mean + L2, over one synthetic batch.  Default workload = BASELINE.json configs[2] (1xB200 bf16 tcgen05 path,
batch 1024, 32x256): the configuration the north-star targets are quoted on; under torchrun every rank runs the
same per-GPU batch (weak scaling, batch-sharded, no data-path collective for the forward).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (per-GPU batch, padded width, description)
    "c3": (1024, 256, "BASELINE configs[2]: bf16 tcgen05 conv+LSTM path, batch 1024, 32x256, fwd+CTC"),
    "c2shape": (256, 160, "BASELINE configs[1] shapes (batch 256, 32x160) on the bf16 path"),
    "c1shape": (32, 100, "BASELINE configs[0] shapes (batch 32, 32x100)"),
}
GFLOP_PER_IMG = lambda W: (12357632 * W - 2097152 + 3145728 * (W // 4 - 1) + 65536 * (W // 4 - 1)) / 1e9


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        # median of the upper half: idle samples before/after the region would drag a plain median down
        sm.sort()
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (float(np.median(load)) if load else None), "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference(N, W, steps, warmup, seed=3):
    """The reference's CPU path: the op-for-op fp32 restatement (oracle port) on all host cores."""
    import torch
    from oracle import crnn_oracle as O
    p32 = O.to_torch(O.init_params(seed, dtype=np.float32), torch.float32)
    data, lab, ll, tsl = O.synth_batch(N, W, seed=seed)
    # "all the host threads it can use": oneDNN on these small convs is fastest well below the core count of a
    # 128-core host, so try a ladder of thread counts once and keep the best (reported as `cores`)
    ncpu = os.cpu_count() or 1
    best = (None, 1e30)
    for nt in sorted({min(ncpu, c) for c in (8, 16, 32, 64, ncpu)}):
        torch.set_num_threads(nt)
        O.fwd_ctc_fp32(p32, data[:max(1, N // 4)], lab[:int(ll[:max(1, N // 4)].sum())], ll[:max(1, N // 4)], tsl[:max(1, N // 4)])
        t0 = time.perf_counter()
        O.fwd_ctc_fp32(p32, data[:max(1, N // 4)], lab[:int(ll[:max(1, N // 4)].sum())], ll[:max(1, N // 4)], tsl[:max(1, N // 4)])
        dt = time.perf_counter() - t0
        if dt < best[1]:
            best = (nt, dt)
    torch.set_num_threads(best[0])
    times = []
    loss = None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss, _ = O.fwd_ctc_fp32(p32, data, lab, ll, tsl)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    t = float(np.median(times))
    return dict(value=N / t, ms_per_step=t * 1e3, loss=loss, cores=torch.get_num_threads())


def run_reference_arm(args, rank):
    N_full, W, desc = WORKLOADS[args.workload]
    if rank != 0:
        return
    sample_n = min(N_full, 32)        # bounded sample of the same workload: 32 lines of 32xW per step
    r = cpu_reference(sample_n, W, args.steps, max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "text-line images/sec (fwd+CTC loss)", "value": r["value"], "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "batch_per_step": sample_n, "width": W,
                   "note": "TF1/warp-ctc not installable (py3.12, no network): op-for-op fp32 restatement on torch-CPU"},
        "cpu_baseline": {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": "port",
                         "sample": f"{sample_n} lines of 32x{W} per step, median of {args.steps} steps"},
        "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=32, help="lines per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the extra training-step measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from lstm_ctc_ocr_b200 import engine
    from lstm_ctc_ocr_b200._lib import c_int, check
    from lstm_ctc_ocr_b200.lib.networks.factory import get_network
    from lstm_ctc_ocr_b200.session import Session
    from lstm_ctc_ocr_b200 import synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, W, desc = WORKLOADS[args.workload]
    T = W // 4 - 1
    K, Wm = args.steps, args.warmup
    peaks = load_peaks()

    # ---- model with reference initialisers (random init; no checkpoints offline), identical on every rank
    model = engine.CrnnModel(weight_decay=1e-5, device=dev)
    model.load_params(synthetic.init_params(3))

    # ---- rotating set of distinct input batches > L2 (8 x 33.5 MB at c3), resident in HBM
    nrot = max(2, int(np.ceil(160e6 / (N * W * 32 * 4))))
    batches = []
    for i in range(nrot):
        data, lab, ll, tsl = synthetic.synth_batch(N, W, seed=3 + 1000 * rank + i)
        batches.append((torch.tensor(data, device=dev), torch.tensor(lab, device=dev), torch.tensor(ll, device=dev),
                        torch.tensor(tsl, device=dev), int(ll.max()), (data, lab, ll, tsl)))
    logits = torch.empty((T, N, 64), dtype=torch.float32, device=dev)
    costs = torch.empty(N, dtype=torch.float32, device=dev)
    grad = torch.empty_like(logits)

    def step(i):
        d, lab, ll, tsl, mll, _ = batches[i % nrot]
        model.forward(d, tsl, out=logits)
        engine.ctc_loss(logits, lab, ll, tsl, want_grad=True, grad_scale=1.0 / N, max_label_len=mll, costs=costs, grad=grad)
        return model.total_loss(costs)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for i in range(Wm):
        loss = step(i)
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    # ---- timed region: K steps, CUDA events on the launching stream, per-stage events inside
    check(model.lib.crnn_profile_begin(model.handle, K))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctc_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sync_all()
    e0.record()
    for i in range(K):
        d, lab, ll, tsl, mll, _ = batches[i % nrot]
        model.forward(d, tsl, out=logits)
        ctc_ev[i][0].record()
        engine.ctc_loss(logits, lab, ll, tsl, want_grad=True, grad_scale=1.0 / N, max_label_len=mll, costs=costs, grad=grad)
        ctc_ev[i][1].record()
        loss = model.total_loss(costs)
    e1.record()
    sync_all()
    ms_total = e0.elapsed_time(e1)
    loss_val = float(loss.item())
    nst = model.lib.crnn_profile_num_stages()
    buf = (np.zeros((K, nst), dtype=np.float32))
    nf = c_int()
    check(model.lib.crnn_profile_read(model.handle, buf.ctypes.data, nf))
    stage_ms = buf[:nf.value].mean(axis=0)
    stage_names = [model.lib.crnn_profile_stage_name(i).decode() for i in range(nst)]
    ctc_ms = float(np.mean([a.elapsed_time(b) for a, b in ctc_ev]))

    # ---- e2e: the reference-facing call (Session.run on HOST numpy buffers; H2D + D2H inside the timed region)
    net = get_network("LSTM_train")
    sess = Session(device=dev)
    sess._engines[id(net)] = model            # same weights / same engine instance
    loss_h, _ = net.build_loss()

    def e2e_step(i):
        data, lab, ll, tsl = batches[i % nrot][5]
        return sess.run(loss_h, feed_dict={net.data: data, net.labels: lab, net.time_step_len: tsl, net.labels_len: ll,
                                           net.keep_prob: 0.5})
    for i in range(max(3, 3 * nrot)):      # every rotating host buffer is fed three times: page-locked in place on its second
                                           # sighting, fed through the chunked crnn_forward_host path from the third on
        e2e_step(i)
    sync_all()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    Ke = max(3, min(K, 10))
    f0.record()
    for i in range(Ke):
        e2e_loss = e2e_step(i)
    f1.record()
    sync_all()
    ms_e2e = f0.elapsed_time(f1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- BASELINE configs[4] companion: full training step (fwd + CTC + backward + [NCCL grad all-reduce] + clip + Adam)
    ms_train = None
    if not args.no_train:
        from lstm_ctc_ocr_b200 import parallel
        model.set_training(True)
        Kt = max(3, min(K, 10))

        def train_step(i, stepno):
            d, lab, ll, tsl, mll, _ = batches[i % nrot]
            model.forward(d, tsl, out=logits)
            engine.ctc_loss(logits, lab, ll, tsl, want_grad=True, grad_scale=1.0 / N, max_label_len=mll, costs=costs, grad=grad)
            model.backward(d, tsl, grad)
            if world > 1:
                parallel.allreduce_sum_(model.grads)
            model.clip_adam_step(lr=1e-4, step=stepno, clip=10.0, grad_mul=1.0 / world, wd_mul=float(world))
        for i in range(3):
            train_step(i, i + 1)
        sync_all()
        check(model.lib.crnn_profile_begin(model.handle, Kt))
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for i in range(Kt):
            train_step(i, 4 + i)
        g1.record()
        sync_all()
        ms_train = g0.elapsed_time(g1) / Kt
        nb = model.lib.crnn_profile_bwd_num_stages()
        bbuf = np.zeros((Kt, nb), dtype=np.float32)
        nfb = c_int()
        check(model.lib.crnn_profile_bwd_read(model.handle, bbuf.ctypes.data, nfb))
        bwd_stage_ms = {model.lib.crnn_profile_bwd_stage_name(i).decode(): round(float(bbuf[:nfb.value, i].mean()), 4) for i in range(nb)}
        fbuf = np.zeros((Kt, nst), dtype=np.float32)
        check(model.lib.crnn_profile_read(model.handle, fbuf.ctypes.data, nfb))
        bwd_stage_ms["forward_total(train mode)"] = round(float(fbuf[:nfb.value].sum(axis=1).mean()), 4)

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([ms_total, ms_e2e, ms_train or 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e = float(t[0]), float(t[1])
        ms_train = float(t[2]) if ms_train is not None else None
    ms_step = ms_total / K
    value = world * N / (ms_step / 1e3)
    e2e_value = world * N / (ms_e2e / Ke / 1e3)

    if rank == 0:
        # dominant kernel = the stage with the largest share of the step
        flops = {"conv2_pool2": 2.0 * N * (W // 2) * 16 * 576 * 128, "conv3_1": 2.0 * N * (W // 4) * 8 * 1152 * 256,
                 "conv3_2_pool": 2.0 * N * (W // 4) * 8 * 2304 * 256, "conv4_1_gemm": 2.0 * N * (W // 4) * 4 * 2304 * 512,
                 "conv4_2_gemm": 2.0 * N * (W // 4) * 4 * 4608 * 512, "conv5": 2.0 * N * T * 2048 * 512,
                 "lstm_xproj": 2.0 * N * T * 512 * 2048, "lstm_recurrence": 2.0 * N * T * 256 * 1024 * 2, "logits": 2.0 * N * T * 512 * 64,
                 "conv1_pool1": 2.0 * N * W * 32 * 9 * 64}
        stages = {n: {"ms": round(float(m), 4), "share": round(float(m) / ms_step, 4),
                      **({"tflops": round(flops[n] / (float(m) * 1e-3) / 1e12, 1)} if n in flops and m > 0 else {})}
                  for n, m in zip(stage_names, stage_ms)}
        ctc_bytes = 2 * T * N * 64 * 4 + 4 * (int(batches[0][1].numel()) + 2 * N)
        stages["ctc_loss_grad"] = {"ms": round(ctc_ms, 4), "share": round(ctc_ms / ms_step, 4),
                                   "gbs": round(ctc_bytes / (ctc_ms * 1e-3) / 1e9, 1),
                                   "hbm_frac": round(ctc_bytes / (ctc_ms * 1e-3) / 1e9 / peaks["hbm"], 3)}
        dom = max((n for n in stage_names if n in flops and n != "conv1_pool1" and n != "lstm_recurrence"),
                  key=lambda n: stages[n]["ms"])
        ach = flops[dom] / (stages[dom]["ms"] * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")
        if os.path.exists(tp) and args.workload == "c3":
            traffic = json.load(open(tp))["kernels"].get(dom, {}).get("traffic_bytes")
        roofline = {"kernel": f"gemm_kernel<{dom}>", "bound": "tensor", "achieved": round(ach, 1), "peak": peaks["bf16_sustained"],
                    "unit": "TFLOP/s", "frac": round(ach / peaks["bf16_sustained"], 3), "traffic": traffic,
                    "traffic_note": "DRAM read+write bytes of one launch from the committed ncu --set full capture (profiles/r1_ncu_traffic.json); "
                                    "algorithmic bytes of conv4_2 = 268 MB in + 4.7 MB weights + 268 MB out",
                    "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['src']}); kernel timed inside a long step",
                    "whole_step_tflops": round(N * GFLOP_PER_IMG(W) / ms_step, 1)}
        line = {
            "metric": "text-line images/sec (fwd+CTC loss)", "value": round(value, 1), "unit": "images/s", "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": desc, "batch_per_gpu": N, "global_batch": N * world, "width": W, "T": T,
                       "parallelism": f"dp{world} (batch-sharded replicas, per-replica BN statistics, no forward collective)",
                       "l2": f"rotating {nrot} distinct input batches ({nrot * N * W * 32 * 4 / 1e6:.0f} MB > 126 MB L2); "
                             f"per-step activation traffic ~2.5 GB"},
            "loss": round(loss_val, 5),
            "e2e": {"value": round(e2e_value, 1), "unit": "images/s", "h2d_bytes_per_step": int(sess.h2d_bytes),
                    "d2h_bytes_per_step": int(sess.d2h_bytes), "steps": Ke, "api": "Session.run(loss, feed_dict=host numpy)",
                    "loss": float(e2e_loss)},
            "gpu_launches": K * 16,      # per step: conv1, 8 tcgen05 GEMMs, 2x(bn finalize + apply), persistent LSTM, CTC, loss
            "roofline": roofline, "stages": stages, "clocks": clocks,
        }
        if ms_train is not None:
            line["train_step"] = {"ms_per_step": round(ms_train, 4), "images_per_s": round(world * N / (ms_train / 1e3), 1),
                                  "what": "fwd + CTC loss/grad + backward + " + ("NCCL all-reduce(28.6 MB f32) + " if world > 1 else "") +
                                          "global-norm clip + Adam (BASELINE configs[4] per-GPU shape)",
                                  "stages_ms": bwd_stage_ms}
        if world == 1 and not args.no_cpu_baseline:
            sn = args.cpu_sample
            r = cpu_reference(sn, W, steps=8, warmup=2)
            line["cpu_baseline"] = {"value": round(r["value"], 2), "unit": "images/s", "cores": r["cores"], "kind": "port",
                                    "sample": f"{sn} lines of 32x{W} per step (same shapes, fp32 torch-CPU restatement), median of 8"}
            # BASELINE metric, second half ("CTC-loss delta vs ref"): the same sample (same seeds, same initialisers) through the GPU path
            try:
                sd, slab, sll, stsl = synthetic.synth_batch(sn, W, seed=3)
                t_ = lambda a: torch.tensor(a, device=dev)
                lg = model.forward(t_(sd), t_(stsl))
                cs, _ = engine.ctc_loss(lg, t_(slab), t_(sll), t_(stsl), max_label_len=int(sll.max()))
                gpu_loss = float(model.total_loss(cs).item())
                cpu_loss = float(r["loss"])
                line["ctc_loss_delta"] = {"gpu": round(gpu_loss, 5), "cpu_port_fp32": round(cpu_loss, 5),
                                          "rel": round(abs(gpu_loss - cpu_loss) / abs(cpu_loss), 6),
                                          "sample": f"{sn} lines of 32x{W}, seed 3, reference initialisers (mean CTC NLL + L2 term)"}
            except Exception as e:      # never lose the bench line over the side statistic
                line["ctc_loss_delta"] = {"error": repr(e)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
