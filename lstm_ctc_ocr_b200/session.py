"""``Session.run(fetches, feed_dict)`` -- the call the reference's solver makes every iteration
(lib/lstm/train.py:129-130,160; lib/lstm/test.py:77), evaluated by the sm_100a engine.

Per run: feed_dict numpy arrays -> pinned host staging -> async H2D on the current stream ->
crnn_forward -> crnn_ctc_loss / crnn_total_loss / crnn_ctc_greedy as the fetches require -> D2H of
exactly the fetched values."""
import numpy as np
import torch

from . import engine
from ._lib import CrnnError
from .lib.networks.network import Fetch, Placeholder

_NP2T = {np.dtype("float32"): torch.float32, np.dtype("int32"): torch.int32}


class _Pinned(object):
    """Host->device staging.  Default: copy into reusable pinned staging buffers, then async DMA.  A LARGE C-contiguous
    numpy input that is fed again from the same buffer (a feeder reusing its batch buffers) is page-locked IN PLACE on its
    second sighting (cudaHostRegister) and DMA'd directly from then on.  The registry keeps a reference to every registered
    array, so its memory cannot be freed or reused while it is pinned; the oldest entry is unregistered on overflow."""
    REGISTER_MIN_BYTES = 8 << 20
    MAX_REGISTERED = 16

    def __init__(self):
        self.bufs = {}
        self.events = {}              # staging buffer name -> event recorded after its last H2D copy
        self.pending = []             # events of copies that read the CALLER's memory in place
        self.registered = {}          # (ptr, nbytes) -> array (keeps the memory alive), insertion-ordered
        self.seen_once = {}           # (ptr, nbytes) -> True, bounded
        from . import _lib
        self._is_pinned = _lib.load().crnn_host_is_pinned

    def _registered(self, arr):
        key = (arr.ctypes.data, arr.nbytes)
        if key in self.registered:
            return True
        if key not in self.seen_once:
            if len(self.seen_once) > 64:
                self.seen_once.pop(next(iter(self.seen_once)))
            self.seen_once[key] = True
            return False
        rt = torch.cuda.cudart()
        if len(self.registered) >= self.MAX_REGISTERED:
            old = next(iter(self.registered))
            rt.cudaHostUnregister(old[0])
            del self.registered[old]
        if int(rt.cudaHostRegister(arr.ctypes.data, arr.nbytes, 0)) != 0:
            return False
        self.registered[key] = arr
        self.seen_once.pop(key, None)
        return True

    def is_page_locked(self, arr):
        """True when `arr` lives in page-locked memory: registered in place by stage(), or allocated pinned by the caller
        (e.g. a PrefetchFeeder ring slot) -- asked of the driver through crnn_host_is_pinned."""
        return (arr.ctypes.data, arr.nbytes) in self.registered or bool(self._is_pinned(arr.ctypes.data))

    def wait_pending(self):
        for ev in self.pending:
            ev.synchronize()
        self.pending = []

    def close(self):
        rt = torch.cuda.cudart()
        for (ptr, _n) in list(self.registered):
            rt.cudaHostUnregister(ptr)
        self.registered.clear()

    def stage_ints(self, arrays, device, name="_ints"):
        """The small int32 feeds of one run (time_step_len, labels, labels_len) through ONE pinned staging buffer and ONE async
        copy; returns {name: device view}.  Sub-arrays start on 256-byte boundaries."""
        offs, tot = {}, 0
        for k, a in arrays.items():
            offs[k] = tot
            tot += (max(a.size, 1) + 63) // 64 * 64
        t = self.bufs.get(name)
        if t is None or t.numel() < tot:
            t = self.bufs[name] = torch.empty(max(tot, 4096), dtype=torch.int32).pin_memory()
            self.events.pop(name, None)
        ev = self.events.get(name)
        if ev is not None:
            ev.synchronize()          # the previous run's DMA out of this staging buffer must be done before it is rewritten
        tn = t.numpy()
        for k, a in arrays.items():
            tn[offs[k]:offs[k] + a.size] = a.reshape(-1)
        d = t[:tot].to(device, non_blocking=True)
        if ev is None:
            ev = self.events[name] = torch.cuda.Event()
        ev.record()
        return {k: d[offs[k]:offs[k] + a.size].view(a.shape) for k, a in arrays.items()}

    def staging_for(self, name, numel):
        """Page-locked f32 staging buffer `name` with room for `numel` elements, safe to overwrite (the previous DMA out of it has
        completed), and the event the caller must record after issuing the next DMA out of it."""
        t = self.bufs.get(name)
        if t is None or t.numel() < numel or t.dtype != torch.float32:
            t = self.bufs[name] = torch.empty(max(numel, 1), dtype=torch.float32).pin_memory()
            self.events.pop(name, None)
        ev = self.events.get(name)
        if ev is not None:
            ev.synchronize()
        else:
            ev = self.events[name] = torch.cuda.Event()
        return t, ev

    def stage(self, name, arr, device):
        arr = np.ascontiguousarray(arr)
        tdt = _NP2T[arr.dtype]
        if arr.nbytes >= self.REGISTER_MIN_BYTES and arr.flags.owndata and self._registered(arr):
            d = torch.from_numpy(arr).to(device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self.pending.append(ev)   # DMA straight out of the caller's array: run() waits for it before returning
            return d
        t = self.bufs.get(name)
        if t is None or t.numel() < arr.size or t.dtype != tdt:
            t = torch.empty(max(arr.size, 1), dtype=tdt).pin_memory()
            self.bufs[name] = t
            self.events.pop(name, None)
        ev = self.events.get(name)
        if ev is not None:
            ev.synchronize()          # the previous run's DMA out of this staging buffer must be done before it is rewritten
        v = t[:arr.size].view(arr.shape)
        v.copy_(torch.from_numpy(arr))
        d = v.to(device, non_blocking=True)
        if ev is None:
            ev = self.events[name] = torch.cuda.Event()
        ev.record()
        return d


class Session(object):
    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise CrnnError("Session needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self._engines = {}
        self._pinned = _Pinned()
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.last_feed_path = None    # "page-locked in place" (chunked crnn_forward_host) | "staged" (copy into pinned staging first)
        import os
        self.h2d_chunks = int(os.environ.get("CRNN_H2D_CHUNKS", "4"))   # image ranges of the overlapped host->device feed (1 = copy, then compute)
        self.pageable_pool = not os.environ.get("CRNN_NO_PAGEABLE_POOL")   # large pageable batches through crnn_forward_pageable (else: torch copy into staging, then copy-then-compute)
        self.host_copy_threads = int(os.environ.get("CRNN_HOST_COPY_THREADS", str(min(8, os.cpu_count() or 1))))   # pageable -> pinned staging copies
        # device prefetch (attach_feeder): the NEXT batch of a PrefetchFeeder is copied host->device on a side stream while the
        # current step computes -- what tf.data's prefetch_to_device does for a TF input pipeline
        self._feeder = None
        self._ahead = None                      # (feeder sequence number, host ptr, nbytes, device tensor, copy-done event, buffer index)
        self._ahead_bufs = [None, None]
        self._ahead_free = [None, None]         # per buffer: event recorded on the compute stream after the last step that read it
        self._ahead_idx = 0
        self._ahead_stream = None
        self.ahead_hits = 0

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        torch.cuda.synchronize(self.device)
        self._feeder = None
        self._ahead = None
        self._ahead_bufs = [None, None]
        self._pinned.close()
        self._engines.clear()

    # ---- device prefetch ------------------------------------------------------------------
    def attach_feeder(self, feeder):
        """``feeder``: a PrefetchFeeder (anything with ``peek()`` and ``delivered``) whose batches are fed to ``run`` in order.
        From then on every ``run`` starts the host->device copy of the feeder's NEXT batch on a side stream before it waits for
        its own results, and the next ``run`` finds its input already resident (the copy is still one H2D per step, issued
        from the page-locked ring slot; it just overlaps the previous step instead of preceding its own).  ``None`` detaches."""
        self._feeder = feeder if (feeder is not None and hasattr(feeder, "peek") and hasattr(feeder, "delivered")) else None
        self._ahead = None

    def _stage_ahead(self):
        f = self._feeder
        if f is None or self._ahead is not None:
            return
        try:
            nxt = f.peek()
        except StopIteration:
            return
        data = nxt[0]
        if not (isinstance(data, np.ndarray) and data.dtype == np.float32 and data.flags.c_contiguous and data.ndim == 3
                and self._pinned.is_page_locked(data)):
            return
        i = self._ahead_idx
        buf = self._ahead_bufs[i]
        if buf is None or tuple(buf.shape) != tuple(data.shape):
            buf = self._ahead_bufs[i] = torch.empty(data.shape, dtype=torch.float32, device=self.device)
            self._ahead_free[i] = None
        if self._ahead_stream is None:
            self._ahead_stream = torch.cuda.Stream(device=self.device)
        st = self._ahead_stream
        if self._ahead_free[i] is not None:
            st.wait_event(self._ahead_free[i])   # the step that last read this device buffer has finished with it
        # the small integer feeds of that batch too (validated here, off the critical path); run() re-uses them when the arrays it
        # is fed compare equal
        ints = None
        try:
            h = {"labels": np.ascontiguousarray(nxt[1], dtype=np.int32), "llen": np.ascontiguousarray(nxt[2], dtype=np.int32),
                 "tsl": np.ascontiguousarray(nxt[3], dtype=np.int32)}
            self.validate_feed(data, h["tsl"], h["labels"], h["llen"])
        except Exception:
            h = None                            # run() will validate what it is actually fed and raise there
        with torch.cuda.stream(st):
            buf.copy_(torch.from_numpy(data), non_blocking=True)
            if h is not None:
                ints = (h, self._pinned.stage_ints(h, self.device, name="_ints_ahead%d" % i))
            ev = torch.cuda.Event()
            ev.record(st)
        self._ahead = (f.delivered, data.ctypes.data, data.nbytes, buf, ev, i, ints)   # f.delivered == sequence number of the peeked batch
        self._ahead_idx = i ^ 1

    def _take_ahead(self, data):
        """Device copy of `data` if it is the batch staged ahead (same feeder sequence number, same ring slot), else None."""
        a, f = self._ahead, self._feeder
        if a is None or f is None:
            return None
        self._ahead = None
        seq, ptr, nbytes, buf, ev, i, ints = a
        if seq != f.delivered - 1 or ptr != data.ctypes.data or nbytes != data.nbytes or tuple(buf.shape) != tuple(data.shape):
            return None
        torch.cuda.current_stream(self.device).wait_event(ev)
        self._pinned.pending.append(ev)        # the ring slot must not be recycled before this DMA is done (it is, long before)
        return buf, i, ints

    # ---- variables ------------------------------------------------------------------------
    def engine_for(self, net):
        eng = self._engines.get(id(net))
        if eng is None:
            from .lib.lstm.config import cfg
            eng = engine.CrnnModel(weight_decay=float(getattr(net, "_wd", cfg.TRAIN.WEIGHT_DECAY)), device=self.device)
            self._engines[id(net)] = eng
        return eng

    def assign(self, net, state_dict, ignore_missing=False):
        eng = self.engine_for(net)
        full = eng.state_dict()
        for k in full:
            if k in state_dict:
                full[k] = np.asarray(state_dict[k], dtype=np.float32).reshape(full[k].shape)
            elif not ignore_missing:
                raise KeyError(k)
        eng.load_params(full)

    def variables(self, net):
        return self.engine_for(net).state_dict()

    # ---- run ------------------------------------------------------------------------------
    @staticmethod
    def validate_feed(data, tsl, labels, labels_len):
        """Host-side checks the C ABI cannot do without a device sync (invalid lengths)."""
        if data.ndim != 3 or data.shape[2] != 32:
            raise ValueError(f"data must be [N, W, 32], got {data.shape}")
        N, W, _ = data.shape
        if W % 4 != 0 or W < 8:
            raise ValueError("padded width must be a multiple of POOL_SCALE=4 (gen.py:58) and >= 8")
        T = W // 4 - 1
        if tsl.shape != (N,):
            raise ValueError("time_step_len must be [N]")
        if tsl.min() < 0 or tsl.max() > T:
            raise ValueError(f"time_step_len must lie in [0, {T}] (conv output has W/4-1 frames)")
        if labels is not None:
            if labels_len.shape != (N,) or labels_len.min() < 0:
                raise ValueError("labels_len must be [N], non-negative")
            if int(labels_len.sum()) != labels.size:
                raise ValueError("sum(labels_len) != len(labels)")
            if labels.size and (labels.min() < 1 or labels.max() > 62):
                raise ValueError("label ids must lie in 1..62 (0 is the CTC blank, 63 the decoder blank)")

    def run(self, fetches, feed_dict=None):
        single = not isinstance(fetches, (list, tuple))
        flist = [fetches] if single else list(fetches)
        feed_dict = feed_dict or {}
        net = None
        for f in flist:
            if isinstance(f, Fetch):
                net = f.net
        if net is None:
            raise ValueError("nothing to run: fetches must come from a network (build_loss / get_output)")
        feeds = {}
        for k, v in feed_dict.items():
            if not isinstance(k, Placeholder):
                raise TypeError("feed_dict keys must be the network's placeholders")
            feeds[k.name] = v
        kinds = [f.kind for f in flist]
        need_labels = any(k in ("loss", "ctc_costs", "train_op", "ctc_grad") for k in kinds)
        data = np.asarray(feeds["data"], dtype=np.float32)
        tsl = np.asarray(feeds["time_step_len"], dtype=np.int32)
        labels = np.asarray(feeds["labels"], dtype=np.int32) if need_labels else None
        llen = np.asarray(feeds["labels_len"], dtype=np.int32) if need_labels else None
        eng = self.engine_for(net)
        dev = self.device
        # training mode is sticky: its forward is a superset (it also saves what the backward needs), and switching back
        # and forth would re-plan the multi-GB workspace
        if any(k == "train_op" for k in kinds) and not eng.training:
            eng.set_training(True)
        data = np.ascontiguousarray(data)
        ahead = self._take_ahead(data)
        d_ints = None
        if ahead is not None and ahead[2] is not None:
            h, dv = ahead[2]
            if np.array_equal(h["tsl"], tsl) and (not need_labels or (np.array_equal(h["labels"], labels) and np.array_equal(h["llen"], llen))):
                d_ints = dv                     # validated and copied while the previous step was running
        if d_ints is None:
            self.validate_feed(data, tsl, labels, llen)
            ints = {"tsl": tsl}
            if need_labels:
                ints["labels"], ints["llen"] = labels, llen
            d_ints = self._pinned.stage_ints(ints, dev)
        d_tsl = d_ints["tsl"]
        self.h2d_bytes = data.nbytes + tsl.nbytes
        used_ahead = None
        if ahead is not None:
            d_data, used_ahead = ahead[0], ahead[1]
            logits = eng.forward(d_data, d_tsl)
            self.last_feed_path = "page-locked in place, copied during the previous step (device prefetch)"
            self.ahead_hits += 1
        elif self.h2d_chunks > 1 and (self._pinned.is_page_locked(data) or
                                      (data.nbytes >= self._pinned.REGISTER_MIN_BYTES and data.flags.owndata and self._pinned._registered(data))):
            # page-locked batch buffer (a feeder's ring slot, or a large array fed a second time and registered in place now):
            # chunked H2D overlapped with the conv front end
            logits, d_data = eng.forward_host(data, d_tsl, chunks=self.h2d_chunks)
            self.last_feed_path = "page-locked in place"
        elif self.pageable_pool and self.h2d_chunks > 1 and data.nbytes >= self._pinned.REGISTER_MIN_BYTES:
            # large batch in ordinary memory (the reference's np.array(...) per step): the library's host threads move it into
            # page-locked staging range by range while the GPU copies / computes the previous range (crnn_forward_pageable)
            pin, ev = self._pinned.staging_for("data", data.size)
            logits, d_data, cst = eng.forward_pageable(data, pin, d_tsl, chunks=self.h2d_chunks, host_threads=self.host_copy_threads)
            ev.record(cst)
            self.last_feed_path = "staged"
        else:
            d_data = self._pinned.stage("data", data, dev)
            logits = eng.forward(d_data, d_tsl)
            self.last_feed_path = "staged"
        costs = grad = loss = None
        if need_labels:
            d_lab, d_ll = d_ints["labels"], d_ints["llen"]
            self.h2d_bytes += labels.nbytes + llen.nbytes
            N = data.shape[0]
            # warp-ctc computes the gradient inside its forward op; so does this kernel (one launch)
            costs, grad = engine.ctc_loss(logits, d_lab, d_ll, d_tsl, want_grad=True, grad_scale=1.0 / N,
                                          max_label_len=int(llen.max()) if llen.size else 0)
            loss = eng.total_loss(costs)
        has_train = any(k == "train_op" for k in kinds)
        if not has_train:
            self._stage_ahead()                # everything of this step is enqueued: start the next batch's copy before waiting for results
        out = []
        self.d2h_bytes = 0
        for f in flist:
            k = f.kind
            if k == "loss":
                v = loss.cpu().numpy()[0]
            elif k == "ctc_costs":
                v = costs.cpu().numpy()
            elif k == "ctc_grad":
                v = grad.cpu().numpy()
            elif k == "logits":
                v = logits.cpu().numpy()
            elif k == "dense_decoded":
                from .lib.lstm.config import cfg
                if str(cfg.get("DECODER", "greedy")) == "beam":
                    # the reference's own decoder (network.py:656): host-side prefix beam search, width 100, blank 63
                    o, ol, _ = engine.ctc_beam_search(logits, tsl, beam_width=int(cfg.get("BEAM_WIDTH", 100)), merge_repeated=True)
                    m_ = int(ol.max()) if ol.size else 0
                    v = np.ascontiguousarray(o[:, :m_])
                else:
                    o, ol = engine.ctc_greedy(logits, d_tsl)
                    v = engine.dense_decoded(o, ol).cpu().numpy()
            elif k == "train_op":
                v = f.step_fn(eng, logits, grad, d_data, d_tsl)
                self._stage_ahead()
            elif k.startswith("layer:"):
                name = k.split(":", 1)[1]
                tapname = {"pool1": "conv1", "pool2": "conv3_2", "pool3": "conv4_2", "reshaped_layer": "conv5"}.get(name, name)
                v = eng.tap(tapname, data.shape[0], data.shape[1]).cpu().numpy()
            else:
                raise ValueError(f"unknown fetch {f}")
            if isinstance(v, np.ndarray):
                self.d2h_bytes += v.nbytes
            elif isinstance(v, (np.floating, float)):
                self.d2h_bytes += 4
            out.append(v)
        if used_ahead is not None:
            ev = self._ahead_free[used_ahead]
            if ev is None:
                ev = self._ahead_free[used_ahead] = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
        self._pinned.wait_pending()
        return out[0] if single else out
