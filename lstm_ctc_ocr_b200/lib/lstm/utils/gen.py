"""Batch source with the reference data layer's contract (lib/lstm/utils/gen.py:41-67,112-128).

``get_batch(num_workers, batch_size)`` returns a generator of
``(img_list [N x [W,32] f32], flat_labels, label_len, time_steps)`` exactly as ``groupBatch`` does.
The reference renders captchas with the ``captcha`` package (not installable here); this module renders text lines
with PIL and the same TTF when the font is available, and otherwise falls back to contract-identical random batches
(``lstm_ctc_ocr_b200.synthetic``).  ``num_workers`` is accepted for call-shape compatibility: batches are produced
in-process (the GPU step is far faster than the reference's 12-process enqueuer was built for)."""
import math
import os
import random

import numpy as np

from ..config import cfg, get_encode_decode_dict
from .... import synthetic

encode_maps, decode_maps = get_encode_decode_dict()


def gen_rand(rng=random):
    n = rng.randint(cfg.MIN_LEN, cfg.MAX_LEN)
    return "".join(rng.choice(cfg.CHARSET) for _ in range(n))


def _font_path():
    for p in (cfg.FONT, os.path.join(cfg.ROOT_DIR, cfg.FONT), os.path.join("/root/reference", cfg.FONT)):
        if os.path.exists(p):
            return p
    return None


def render_line(chars, height=60, width=None):
    """Gray uint8 HxW image of the text (stand-in for ImageCaptcha.generate_image + gray conversion, gen.py:31-37,79)."""
    from PIL import Image, ImageDraw, ImageFont
    fp = _font_path()
    font = ImageFont.truetype(fp, 42) if fp else ImageFont.load_default()
    if width is None:                      # wide enough for the text: batches then mix widths (exercises the padding contract)
        width = int(sum(font.getlength(c) for c in chars)) + 28
    img = Image.new("L", (width, height), color=random.randint(180, 255))
    d = ImageDraw.Draw(img)
    x = random.randint(2, 12)
    for ch in chars:
        d.text((x, random.randint(0, 10)), ch, font=font, fill=random.randint(0, 90))
        x += int(font.getlength(ch)) + random.randint(-2, 3)
    return np.asarray(img, dtype=np.uint8)


def groupBatch(imgs, labels):
    """Resize to height 32 keeping aspect, time_step = nw//4 - 1, right-pad with 0 to a multiple of 4, /255,
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
    batch = []
    for im in resized:
        pad = np.zeros((nh, max_w), np.float32)
        pad[:, :im.shape[1]] = im / 255.0
        batch.append(np.ascontiguousarray(pad.swapaxes(0, 1)).reshape(-1, cfg.NUM_FEATURES))
    return batch, label_vec, label_len, time_steps


def generator(batch_size=32, vis=False, render=None, seed=None):
    if render is None:
        render = _font_path() is not None
    k = 0
    while True:
        if render:
            labels = [gen_rand() for _ in range(batch_size)]
            yield groupBatch([render_line(l) for l in labels], labels)
        else:
            data, lab, ll, tsl = synthetic.synth_batch(batch_size, 88, seed=(seed or cfg.RNG_SEED) + k, widths=[85] * batch_size)
            yield list(data), lab.tolist(), ll.tolist(), tsl.tolist()
        k += 1


def get_batch(num_workers, **kwargs):
    return generator(**kwargs)
