"""Batch source with the reference data layer's contract (lib/lstm/utils/gen.py:41-67,112-128).

``get_batch(num_workers, batch_size)`` returns a generator of
``(images [N x [W,32] f32], flat_labels, label_len, time_steps)`` exactly as ``groupBatch`` does.

What replaces what:
  * ``generateImg`` (captcha package, gen.py:31-37)  -> ``render_line``: PIL + the same TTF (``fonts/Ubuntu-M.ttf``), fresh random
    text / jitter / shades per line; without the font the contract-identical random batches of ``synthetic`` are used.
  * ``groupBatch`` (gen.py:41-67)                    -> ``groupBatch`` (resize to height 32, ``time_step = nw//4 - 1``, zero
    right-padding to a multiple of 4, /255, transpose to [W, 32]); ``pad_to`` pads to a fixed bucket width instead of the batch max.
  * nothing in the reference                          -> ``BucketSampler``: width-bucketed batches (BASELINE configs[3]:
    W in {80,160,256}); every batch comes from ONE bucket and is padded to the bucket width, so the engine keeps three
    workspace plans / TMA maps instead of re-planning for every new batch-max width.
  * ``GeneratorEnqueuer`` + ``multiprocessing.Queue`` (gen.py:112-128, lib/utils/data_util.py) -> ``PrefetchFeeder``:
    ``num_workers`` render processes; finished batches are copied into a ring of PAGE-LOCKED slots and handed out as numpy
    views, so ``Session.run`` DMAs straight from the slot (chunked ``crnn_forward_host``) instead of staging a pageable copy.

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


def _font(size=42):
    from PIL import ImageFont
    key = (_font_path(), size)
    if key not in _FONT_CACHE:
        try:
            _FONT_CACHE[key] = ImageFont.truetype(key[0], size) if key[0] else ImageFont.load_default(size)
        except Exception:
            _FONT_CACHE[key] = ImageFont.load_default()
    return _FONT_CACHE[key]


def render_line(chars, height=60, width=None, rng=random):
    """Gray uint8 HxW image of the text (stand-in for ImageCaptcha.generate_image + gray conversion, gen.py:31-37,79)."""
    from PIL import Image, ImageDraw
    font = _font(42)
    adv = [int(font.getlength(c)) for c in chars]
    if width is None:                      # wide enough for the text: batches then mix widths (exercises the padding contract)
        width = sum(adv) + 28
    img = Image.new("L", (width, height), color=rng.randint(180, 255))
    d = ImageDraw.Draw(img)
    x = rng.randint(2, 12)
    for ch, a in zip(chars, adv):
        d.text((x, rng.randint(0, 10)), ch, font=font, fill=rng.randint(0, 90))
        x += a + rng.randint(-2, 3)
    return np.asarray(img, dtype=np.uint8)


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
    batch = []
    for im in resized:
        pad = np.zeros((nh, max_w), np.float32)
        pad[:, :im.shape[1]] = im / 255.0
        batch.append(np.ascontiguousarray(pad.swapaxes(0, 1)).reshape(-1, cfg.NUM_FEATURES))
    return batch, label_vec, label_len, time_steps


def batch_seed(k, seed=None, rank=0, world=1):
    """Seed of batch k on rank `rank`: distinct across ranks and iterations (rank-independent seeds made every
    data-parallel replica train on the same batch)."""
    return int(cfg.RNG_SEED if seed is None else seed) + k * int(world) + int(rank)


def make_batch(k, batch_size=32, render=True, seed=None, rank=0, world=1, bucket=None):
    """Batch k of a deterministic stream (picklable entry point of the feeder's worker processes).
    ``bucket`` = None: the reference's 4-6 character lines padded to the batch max width; else one of BUCKETS."""
    s = batch_seed(k, seed, rank, world)
    if not render:
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
        render = _font_path() is not None
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
        self.render = (_font_path() is not None) if render is None else render
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


def _worker(kwargs):
    return make_batch(**kwargs)


class PrefetchFeeder(object):
    """Double-buffered (``depth``-deep) page-locked feeder in front of the solver.

    ``arg_fn(k)`` -> kwargs of ``make_batch`` for batch k.  ``num_workers`` > 0: batches are rendered by a process pool
    (``spawn`` context: the children import numpy/PIL only, never CUDA), at most ``depth + num_workers`` in flight, delivered
    in order.  Each delivered batch is copied into ring slot ``k % depth`` -- one page-locked [N, Wmax, 32] f32 allocation per
    slot -- and handed out as ``(ndarray view [N,W,32], labels, label_len, time_steps)``.  A slot is rewritten ``depth`` batches
    later; ``Session.run`` has finished DMA-ing from it by then (it synchronises the copy stream before returning)."""

    def __init__(self, arg_fn, num_workers=4, depth=3, max_width=256, batch_size=None, pinned=True):
        self.arg_fn, self.depth = arg_fn, max(2, int(depth))
        self.num_workers = int(num_workers)
        self.batch_size = batch_size if batch_size is not None else arg_fn(0)["batch_size"]
        self.max_width = int(max_width)
        self._slots, self._keep = [], []
        n = self.batch_size * self.max_width * cfg.NUM_FEATURES
        for _ in range(self.depth):
            buf = None
            if pinned:
                try:
                    import torch
                    if torch.cuda.is_available():
                        t = torch.empty(n, dtype=torch.float32).pin_memory()
                        self._keep.append(t)
                        buf = t.numpy()
                except Exception:
                    buf = None
            if buf is None:
                buf = np.empty(n, np.float32)
            self._slots.append(buf)
        self.pinned = len(self._keep) == self.depth
        self._pool = None
        self._pending = {}
        self._next_submit = 0
        self._next_yield = 0
        if self.num_workers > 0:
            import multiprocessing as mp
            self._pool = mp.get_context("spawn").Pool(self.num_workers)

    def _submit(self):
        while self._pool is not None and self._next_submit < self._next_yield + self.depth + self.num_workers:
            k = self._next_submit
            self._pending[k] = self._pool.apply_async(_worker, (self.arg_fn(k),))
            self._next_submit += 1

    def __iter__(self):
        return self

    def __next__(self):
        k = self._next_yield
        if self._pool is not None:
            self._submit()
            imgs, lab, ll, tsl = self._pending.pop(k).get()
        else:
            imgs, lab, ll, tsl = make_batch(**self.arg_fn(k))
        self._next_yield += 1
        if self._pool is not None:
            self._submit()
        N = len(imgs)
        W = imgs[0].shape[0]
        if N > self.batch_size or W > self.max_width:
            raise ValueError(f"batch [{N},{W}] exceeds the feeder's slot [{self.batch_size},{self.max_width}]")
        view = self._slots[k % self.depth][:N * W * cfg.NUM_FEATURES].reshape(N, W, cfg.NUM_FEATURES)
        for i, im in enumerate(imgs):
            view[i] = im
        return view, lab, ll, tsl

    def close(self):
        if self._pool is not None:
            self._pool.terminate()
            self._pool.join()
            self._pool = None

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
        render = _font_path() is not None
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
