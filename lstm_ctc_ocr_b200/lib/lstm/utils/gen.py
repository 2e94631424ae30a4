"""Batch source with the reference data layer's contract (lib/lstm/utils/gen.py:41-67,112-128).

``get_batch(num_workers, batch_size)`` returns a generator of
``(images [N x [W,32] f32], flat_labels, label_len, time_steps)`` exactly as ``groupBatch`` does.

What replaces what:
  * ``generateImg`` (captcha package, gen.py:31-37)  -> ``render_line``: PIL + the same TTF (``fonts/Ubuntu-M.ttf``), fresh random
    text / jitter / shades per line (falls back to Pillow's embedded scalable font at the same size; only without FreeType are
    the contract-identical random batches of ``synthetic`` used).
  * ``groupBatch`` (gen.py:41-67)                    -> ``groupBatch`` (resize to height 32, ``time_step = nw//4 - 1``, zero
    right-padding to a multiple of 4, /255, transpose to [W, 32]); ``pad_to`` pads to a fixed bucket width instead of the batch max.
  * nothing in the reference                          -> ``BucketSampler``: width-bucketed batches (BASELINE configs[3]:
    W in {80,160,256}); every batch comes from ONE bucket and is padded to the bucket width, so the engine keeps three
    workspace plans / TMA maps instead of re-planning for every new batch-max width.
  * ``GeneratorEnqueuer`` + ``multiprocessing.Queue`` (gen.py:112-128, lib/utils/data_util.py) -> ``PrefetchFeeder``:
    ``num_workers`` render processes write finished batches straight into a ring of PAGE-LOCKED shared-memory slots that are
    handed out as numpy views, so ``Session.run`` DMAs from the slot (chunked ``crnn_forward_host``) with no staging copy.

Data-parallel runs: batch ``k`` of rank ``r`` is generated from seed ``base + k*world + r`` -- every rank sees a different
stream (the reference is single-process and has no such concern)."""
import math
import os
import random

import numpy as np

from ..config import cfg, get_encode_decode_dict
from .... import synthetic

encode_maps, decode_maps = get_encode_decode_dict()

BUCKETS = (80, 160, 256)                       # BASELINE configs[3]
# characters per line that make the rendered width (height 32, ~13.4 px per glyph + margin) fall into each bucket
BUCKET_CHARS = {80: (2, 4), 160: (5, 10), 256: (11, 15)}


def gen_rand(rng=random, min_len=None, max_len=None):
    n = rng.randint(cfg.MIN_LEN if min_len is None else min_len, cfg.MAX_LEN if max_len is None else max_len)
    return "".join(rng.choice(cfg.CHARSET) for _ in range(n))


def _font_path():
    if os.environ.get("CRNN_FONT", "") == "default":      # force Pillow's embedded scalable font (identical on every box)
        return None
    for p in (cfg.FONT, os.path.join(cfg.ROOT_DIR, cfg.FONT), os.path.join(os.path.dirname(os.path.abspath(__file__)), "Ubuntu-M.ttf"),
              os.path.join("/root/reference", cfg.FONT)):
        if os.path.exists(p):
            return p
    for p in ("/usr/share/fonts/truetype/dejavu/DejaVuSans-Bold.ttf", "/usr/share/fonts/truetype/dejavu/DejaVuSans.ttf"):
        if os.path.exists(p):
            return p
    try:                                       # PIL ships a copy of DejaVuSans for its own default font in some builds
        import PIL
        p = os.path.join(os.path.dirname(PIL.__file__), "fonts", "DejaVuSans.ttf")
        if os.path.exists(p):
            return p
    except Exception:
        pass
    return None


_FONT_CACHE = {}


def can_render():
    """True when PIL can draw 42-px glyphs: the reference's TTF at cfg.FONT, a system font, or Pillow's embedded scalable
    default (Pillow >= 10.1 with FreeType).  Round 1 fell back to the 10-px bitmap default on the GPU box -- glyphs ~5 px tall
    after the resize to height 32, T ~ 7 frames for 4-6 characters -- which no model can read."""
    try:
        f = _font(42)
        return hasattr(f, "getlength") and f.getlength("W") >= 20
    except Exception:
        return False


def _font(size=42):
    from PIL import ImageFont
    key = (_font_path(), size)
    if key not in _FONT_CACHE:
        try:
            _FONT_CACHE[key] = ImageFont.truetype(key[0], size) if key[0] else ImageFont.load_default(size)
        except Exception:
            _FONT_CACHE[key] = ImageFont.load_default()
    return _FONT_CACHE[key]


_GLYPHS = {}          # font cache key -> {"adv": {ch: int}, "mask": {ch: (core mask, (ox, oy))}, "fast": True | False}


def _glyphs(font):
    """Per-font cache of what ImageDraw.text recomputes for every character it draws: the advance and the anti-aliased glyph
    mask + offset (FreeType rasterisation is ~85 % of a rendered line).  The cached path draws with the same primitive
    ImageDraw.text ends in (`draw_bitmap(xy + offset, mask, ink)`), and is switched on only after it reproduced ImageDraw.text
    byte for byte on a probe line in this process -- otherwise (another Pillow, no getmask2) the plain path stays."""
    g = _GLYPHS.get(id(font))
    if g is None:
        g = _GLYPHS[id(font)] = {"adv": {}, "mask": {}, "fast": False, "font": font}
        try:
            from PIL import Image, ImageDraw
            probe = "Wg0jQy8"
            a = Image.new("L", (260, 60), color=200); da = ImageDraw.Draw(a)
            b = Image.new("L", (260, 60), color=200); db = ImageDraw.Draw(b)
            x = 3
            for i, ch in enumerate(probe):
                da.text((x, i), ch, font=font, fill=10 * i)
                _draw_glyph(db, g, font, ch, x, i, 10 * i)
                x += 33
            g["fast"] = a.tobytes() == b.tobytes()
        except Exception:
            g["fast"] = False
    return g


def _draw_glyph(d, g, font, ch, x, y, fill):
    m = g["mask"].get(ch)
    if m is None:
        m = g["mask"][ch] = font.getmask2(ch, d.fontmode, anchor="la", start=(0.0, 0.0))
    mask, off = m
    d.draw.draw_bitmap((x + off[0], y + off[1]), mask, d.draw.draw_ink(fill))


def render_line(chars, height=60, width=None, rng=random):
    """Gray uint8 HxW image of the text (stand-in for ImageCaptcha.generate_image + gray conversion, gen.py:31-37,79)."""
    from PIL import Image, ImageDraw
    font = _font(42)
    g = _glyphs(font)
    advc = g["adv"]
    adv = []
    for c in chars:
        a = advc.get(c)
        if a is None:
            a = advc[c] = int(font.getlength(c))
        adv.append(a)
    if width is None:                      # wide enough for the text: batches then mix widths (exercises the padding contract)
        width = sum(adv) + 28
    img = Image.new("L", (width, height), color=rng.randint(180, 255))
    d = ImageDraw.Draw(img)
    x = rng.randint(2, 12)
    fast = g["fast"]
    for ch, a in zip(chars, adv):
        y = rng.randint(0, 10)
        fill = rng.randint(0, 90)
        if fast:
            _draw_glyph(d, g, font, ch, x, y, fill)
        else:
            d.text((x, y), ch, font=font, fill=fill)
        x += a + rng.randint(-2, 3)
    return np.asarray(img, dtype=np.uint8)


def generateImg(rng=random):
    """(gray uint8 image, its characters): the reference's per-sample entry (gen.py:29-35; it returns the captcha's RGB array and
    converts to gray in the generator, gen.py:79 -- here the line is rendered gray directly)."""
    theChars = gen_rand(rng)
    return render_line(theChars, rng=rng), theChars


def groupBatch(imgs, labels, pad_to=None):
    """Resize to height 32 keeping aspect, time_step = nw//4 - 1, right-pad with 0 to a multiple of 4 (or to ``pad_to``), /255,
    transpose to [W, 32] (gen.py:41-67)."""
    from PIL import Image
    nh = cfg.IMG_HEIGHT
    resized, time_steps, label_len, label_vec = [], [], [], []
    max_w = 0
    for img, lab in zip(imgs, labels):
        h, w = img.shape[:2]
        nw = int(nh / h * w)
        max_w = max(max_w, nw)
        resized.append(np.asarray(Image.fromarray(img).resize((nw, nh), Image.BILINEAR), dtype=np.float32))
        time_steps.append(nw // cfg.POOL_SCALE + cfg.OFFSET_TIME_STEP)
        label_vec.extend(encode_maps[c] for c in lab)
        label_len.append(len(lab))
    max_w = int(math.ceil(max_w / cfg.POOL_SCALE) * cfg.POOL_SCALE)
    if pad_to is not None:
        if max_w > pad_to:
            raise ValueError(f"line of width {max_w} does not fit the bucket width {pad_to}")
        max_w = int(pad_to)
    # one zero-filled [N, W, 32] block, every line written transposed into its rows; the list holds its N contiguous [W, 32] views
    block = np.zeros((len(resized), max_w, nh), np.float32)
    for i, im in enumerate(resized):
        np.divide(im.T, np.float32(255.0), out=block[i, :im.shape[1], :])
    return list(block), label_vec, label_len, time_steps


def batch_seed(k, seed=None, rank=0, world=1):
    """Seed of batch k on rank `rank`: distinct across ranks and iterations (rank-independent seeds made every
    data-parallel replica train on the same batch)."""
    return int(cfg.RNG_SEED if seed is None else seed) + k * int(world) + int(rank)


def make_batch(k, batch_size=32, render=True, seed=None, rank=0, world=1, bucket=None, width=None):
    """Batch k of a deterministic stream (picklable entry point of the feeder's worker processes).
    ``bucket`` = None: the reference's 4-6 character lines padded to the batch max width; else one of BUCKETS."""
    s = batch_seed(k, seed, rank, world)
    if not render:
        if width is not None:                  # full-width synthetic lines (the throughput workloads)
            data, lab, ll, tsl = synthetic.synth_batch(batch_size, int(width), seed=s)
            return data, lab.tolist(), ll.tolist(), tsl.tolist()
        if bucket is None:
            data, lab, ll, tsl = synthetic.synth_batch(batch_size, 88, seed=s, widths=[85] * batch_size)
        else:
            data, lab, ll, tsl = synthetic.synth_bucket_batch(batch_size, bucket, seed=s, buckets=BUCKETS)
        return list(data), lab.tolist(), ll.tolist(), tsl.tolist()
    rng = random.Random(s)
    if bucket is None:
        labels = [gen_rand(rng) for _ in range(batch_size)]
        return groupBatch([render_line(l, rng=rng) for l in labels], labels)
    lo = max([b for b in BUCKETS if b < bucket] or [0])
    cmin, cmax = BUCKET_CHARS[bucket]
    imgs, labels = [], []
    while len(imgs) < batch_size:
        text = gen_rand(rng, cmin, cmax)
        im = render_line(text, rng=rng)
        nw = int(cfg.IMG_HEIGHT / im.shape[0] * im.shape[1])
        if lo < nw <= bucket:                      # rejection: the resized width must fall into (previous bucket, bucket]
            imgs.append(im); labels.append(text)
    return groupBatch(imgs, labels, pad_to=bucket)


def generator(batch_size=32, vis=False, render=None, seed=None, rank=None, world=None):
    if render is None:
        render = can_render()
    if rank is None or world is None:
        rank, world = _dist_rank_world()
    k = 0
    while True:
        yield make_batch(k, batch_size, render, seed, rank, world)
        k += 1


class BucketSampler(object):
    """Width-bucketed batch stream (BASELINE configs[3]): batch k comes from bucket ``order[k % len(order)]`` and is padded to
    that bucket's width.  Iterating yields data-layer tuples; ``.bucket_of(k)`` tells which width batch k has."""

    def __init__(self, batch_size=512, buckets=BUCKETS, render=None, seed=None, rank=None, world=None, order=None):
        self.batch_size, self.buckets = batch_size, tuple(buckets)
        self.render = can_render() if render is None else render
        self.seed = seed
        if rank is None or world is None:
            rank, world = _dist_rank_world()
        self.rank, self.world = rank, world
        self.order = tuple(order) if order is not None else self.buckets

    def bucket_of(self, k):
        return self.order[k % len(self.order)]

    def args(self, k):
        return dict(k=k, batch_size=self.batch_size, render=self.render, seed=self.seed, rank=self.rank, world=self.world,
                    bucket=self.bucket_of(k))

    def batch(self, k):
        return make_batch(**self.args(k))

    def __iter__(self):
        k = 0
        while True:
            yield self.batch(k)
            k += 1


def _dist_rank_world():
    try:
        from .... import parallel
        return parallel.rank(), parallel.world_size()
    except Exception:
        return 0, 1


_SHM_CACHE = {}
_SYNTH_CACHE = {}


def _attach(name):
    """Attach to a ring slot created by the parent (cached per worker process)."""
    shm = _SHM_CACHE.get(name)
    if shm is None:
        from multiprocessing import shared_memory
        # spawn children share the parent's resource tracker, so attaching registers nothing new: the parent unlinks the segment
        shm = shared_memory.SharedMemory(name=name)
        _SHM_CACHE[name] = shm
    return shm


def _fill(buf, kwargs):
    """Produce batch `kwargs` and write it into `buf` as [N, W, 32] f32; returns (N, W, labels, label_len, time_steps)."""
    cache = kwargs.pop("cache", 0)
    if cache and not kwargs.get("render", True):
        # synthetic stream for throughput runs: `cache` distinct batches per producer, generated once, then re-written into the
        # slot every time (the per-step work that remains is the copy into page-locked memory a real decoder would do)
        key = (kwargs["batch_size"], kwargs.get("width"), kwargs.get("seed"), kwargs.get("rank"), kwargs["k"] % cache)
        if key not in _SYNTH_CACHE:
            _SYNTH_CACHE[key] = make_batch(**dict(kwargs, k=kwargs["k"] % cache))
        imgs, lab, ll, tsl = _SYNTH_CACHE[key]
    else:
        imgs, lab, ll, tsl = make_batch(**kwargs)
    N, W = len(imgs), imgs[0].shape[0]
    if N * W * cfg.NUM_FEATURES * 4 > len(buf):
        raise ValueError(f"batch [{N},{W}] does not fit the feeder's ring slot")
    view = np.ndarray((N, W, cfg.NUM_FEATURES), np.float32, buffer=buf)
    if isinstance(imgs, np.ndarray):
        np.copyto(view, imgs)
    else:
        for i, im in enumerate(imgs):
            view[i] = im
    # the integer feeds as int32 arrays: what the solver's np.array(...) would make of the lists, built on the PRODUCER side
    # (turning a 10 000-element label list into an array costs the consumer ~0.2 ms per step otherwise)
    return N, W, np.asarray(lab, np.int32), np.asarray(ll, np.int32), np.asarray(tsl, np.int32)


def _worker(shm_name, kwargs):
    return _fill(_attach(shm_name).buf, kwargs)


def _warm_synth_cache(kwargs_list):
    """Pool initializer: every producer process generates its cached synthetic batches up front (a cache miss inside a timed
    region would stall the consumer for the ~0.4 s it takes to draw 8.4 M random pixels)."""
    for kw in kwargs_list:
        kw = dict(kw)
        cache = kw.pop("cache", 0)
        key = (kw["batch_size"], kw.get("width"), kw.get("seed"), kw.get("rank"), kw["k"] % max(cache, 1))
        if cache and key not in _SYNTH_CACHE:
            _SYNTH_CACHE[key] = make_batch(**dict(kw, k=kw["k"] % cache))


class PrefetchFeeder(object):
    """Prefetching feeder in front of the solver: a ring of PAGE-LOCKED shared-memory slots filled by producer processes.

    ``arg_fn(k)`` -> kwargs of ``make_batch`` for batch k.  ``num_workers`` > 0: producer processes (``spawn`` context: they
    import numpy/PIL only, never CUDA) render batch k straight INTO ring slot ``k % slots`` -- a POSIX shared-memory segment
    the parent has page-locked with cudaHostRegister -- so no pickling of pixels and no parent-side copy; at most ``depth``
    batches are in flight / ready ahead of the consumer, delivered in order as ``(ndarray view [N,W,32], labels, label_len,
    time_steps)`` (the three integer feeds as int32 arrays).  ``Session.run`` recognises the view as page-locked (crnn_host_is_pinned) and DMAs straight from it (chunked
    crnn_forward_host).  The ring has ``depth + keep`` slots: the views of the last ``keep`` delivered batches are never
    rewritten, so the consumer may still be DMA-ing from batch j while batches j+1 .. j+depth are produced."""

    def __init__(self, arg_fn, num_workers=4, depth=3, max_width=256, batch_size=None, pinned=True, keep=3, warm=None):
        from multiprocessing import shared_memory
        self.arg_fn, self.depth, self.keep = arg_fn, max(1, int(depth)), max(1, int(keep))
        self.num_workers = int(num_workers)
        self.batch_size = batch_size if batch_size is not None else arg_fn(0)["batch_size"]
        self.max_width = int(max_width)
        self.slot_bytes = self.batch_size * self.max_width * cfg.NUM_FEATURES * 4
        self._shm, self._registered = [], []
        self._rt = None
        if pinned:
            try:
                import torch
                if torch.cuda.is_available():
                    self._rt = torch.cuda.cudart()
            except Exception:
                self._rt = None
        for _ in range(self.depth + self.keep):
            shm = shared_memory.SharedMemory(create=True, size=self.slot_bytes)
            self._shm.append(shm)
            if self._rt is not None:
                ptr = np.ndarray((1,), np.uint8, buffer=shm.buf).ctypes.data
                if int(self._rt.cudaHostRegister(ptr, self.slot_bytes, 0)) == 0:
                    self._registered.append(ptr)
        self.pinned = len(self._registered) == len(self._shm)
        self._pool = None
        self._pending = {}
        self._next_submit = 0
        self._next_yield = 0
        self._peeked = None
        self.delivered = 0                     # batches handed to the consumer so far (the consumer holds batch delivered - 1)
        if self.num_workers > 0:
            import multiprocessing as mp
            # `warm`: kwargs of the batches every producer should pre-generate (synthetic `cache` streams)
            self._pool = (mp.get_context("spawn").Pool(self.num_workers, initializer=_warm_synth_cache, initargs=(list(warm),))
                          if warm else mp.get_context("spawn").Pool(self.num_workers))

    def _slot(self, k):
        return self._shm[k % len(self._shm)]

    def _submit(self):
        while self._pool is not None and self._next_submit < self._next_yield + self.depth:
            k = self._next_submit
            self._pending[k] = self._pool.apply_async(_worker, (self._slot(k).name, self.arg_fn(k)))
            self._next_submit += 1

    def __iter__(self):
        return self

    def peek(self):
        """The batch the NEXT ``next()`` will deliver, without delivering it -- lets ``Session.attach_feeder`` start its host->device
        copy while the step on the current batch is still running.  Needs ``keep >= 2``: the peeked batch counts as handed out."""
        if self._peeked is None:
            self._peeked = self._take()
        return self._peeked

    def __next__(self):
        if self._peeked is not None:
            b, self._peeked = self._peeked, None
        else:
            b = self._take()
        self.delivered += 1
        return b

    def _take(self):
        k = self._next_yield
        self._next_yield += 1                  # batch k is being handed out: slot k+depth (== batch k-keep's) may be refilled
        if self._pool is not None:
            self._submit()
            N, W, lab, ll, tsl = self._pending.pop(k).get()
            self._submit()
        else:
            N, W, lab, ll, tsl = _fill(self._slot(k).buf, self.arg_fn(k))
        view = np.ndarray((N, W, cfg.NUM_FEATURES), np.float32, buffer=self._slot(k).buf)
        return view, lab, ll, tsl

    def close(self):
        if self._pool is not None:
            self._pool.terminate()
            self._pool.join()
            self._pool = None
        for ptr in self._registered:
            try:
                self._rt.cudaHostUnregister(ptr)
            except Exception:
                pass
        self._registered = []
        for shm in self._shm:
            try:
                shm.close()
            except BufferError:                # a consumer still holds a view: the segment is unmapped when the view dies
                pass
            try:
                shm.unlink()
            except Exception:
                pass
        self._shm = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def get_batch(num_workers, **kwargs):
    """Reference entry point (gen.py:112-128): ``get_batch(num_workers=12, batch_size=64, vis=False)``.  ``num_workers`` render
    processes feed a page-locked ring (PrefetchFeeder); ``num_workers <= 1`` renders in-process."""
    kwargs.pop("vis", None)
    batch_size = kwargs.pop("batch_size", 32)
    render = kwargs.pop("render", None)
    if render is None:
        render = can_render()
    seed = kwargs.pop("seed", None)
    rank, world = kwargs.pop("rank", None), kwargs.pop("world", None)
    if rank is None or world is None:
        rank, world = _dist_rank_world()
    bucket = kwargs.pop("bucket", None)

    def arg_fn(k):
        return dict(k=k, batch_size=batch_size, render=render, seed=seed, rank=rank, world=world, bucket=bucket)
    if num_workers is None or num_workers <= 1 or not render:
        return (make_batch(**arg_fn(k)) for k in _count())
    return PrefetchFeeder(arg_fn, num_workers=num_workers, depth=3, max_width=kwargs.pop("max_width", 256), batch_size=batch_size)


def _count():
    k = 0
    while True:
        yield k
        k += 1
