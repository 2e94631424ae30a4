"""Evaluation solver with the reference's surface (lib/lstm/test.py): ``SolverWrapper.test_model(sess, testDir, restore)`` and
``test_net(network, imgdb, testDir, output_dir, log_dir, pretrained_model, restore)``.

Per file (test.py:57-88): read as gray, right-pad the width to a multiple of POOL_SCALE with 0, /255, transpose to
[1, W, 32], decode, map ids -> chars, exact match against the label encoded in the file name (``<idx>_<chars>.png``).
Deviation from the reference, per SURVEY §3.4: ``time_step_len`` is fed as W//4 - 1 (the data layer's convention,
gen.py:54), not the off-by-one W//4 of test.py:74 which exceeds the number of conv frames."""
import math
import os

import numpy as np

from ...session import Session
from .config import cfg, get_encode_decode_dict
from .utils.timer import Timer


def load_line_image(path):
    """uint8 gray HxW (cv2.imread(path, 0) in the reference; PIL here when cv2 is unavailable)."""
    try:
        import cv2
        img = cv2.imread(path, 0)
        if img is not None:
            return img
    except Exception:
        pass
    from PIL import Image
    return np.asarray(Image.open(path).convert("L"), dtype=np.uint8)


def prepare_line(img):
    """[H=32, W] uint8 -> ([1, Wpad, 32] f32, time_step_len) exactly as test.py:65-70 lays the tensor out."""
    if img.shape[0] != cfg.IMG_HEIGHT:
        from PIL import Image
        nw = max(1, int(cfg.IMG_HEIGHT / img.shape[0] * img.shape[1]))
        img = np.asarray(Image.fromarray(img).resize((nw, cfg.IMG_HEIGHT), Image.BILINEAR), dtype=np.uint8)
    w = img.shape[1]
    width = max(8, int(math.ceil(w / cfg.POOL_SCALE) * cfg.POOL_SCALE))
    pad = np.zeros((cfg.IMG_HEIGHT, width), np.float32)
    pad[:, :w] = img.astype(np.float32) / 255.0
    data = np.ascontiguousarray(pad.swapaxes(0, 1)).reshape(1, width, cfg.NUM_FEATURES)
    return data, np.array([max(w // cfg.POOL_SCALE + cfg.OFFSET_TIME_STEP, 0)], np.int32)


def decodeRes(nums, ignore=0):
    _, decode_maps = get_encode_decode_dict()
    return [decode_maps[int(i)] for i in nums if i != ignore]


class SolverWrapper(object):
    def __init__(self, sess, network, imgdb, output_dir, logdir, pretrained_model=None):
        self.net = network
        self.imgdb = imgdb
        self.output_dir = output_dir
        self.pretrained_model = pretrained_model
        print("done")

    def test_model(self, sess, testDir=None, restore=True):
        dense_decoded = self.net.get_output("logits").net  # noqa: F841  (handle kept for symmetry with the reference)
        from ..networks.network import Fetch
        dense_decoded = Fetch(self.net, "dense_decoded")
        if restore:
            from .train import SolverWrapper as TrainSolver
            ts = TrainSolver.__new__(TrainSolver)
            ts.net, ts.output_dir = self.net, self.output_dir
            path = self.pretrained_model or ts._latest_checkpoint()
            try:
                print("Restoring from {}...".format(path), end=" ")
                sess.engine_for(self.net)
                ts.restore(sess, path)
                print("done")
            except Exception:
                raise Exception("Check your pretrained {:s}".format(str(path)))
        timer = Timer()
        total = correct = 0
        for file in sorted(os.listdir(testDir)):
            timer.tic()
            total += 1
            img = load_line_image(os.path.join(testDir, file))
            print(file, end=" ")
            data, tsl = prepare_line(img)
            feed_dict = {self.net.data: data, self.net.time_step_len: tsl, self.net.keep_prob: 1.0}
            res = sess.run(fetches=dense_decoded, feed_dict=feed_dict)
            res = res[0] if len(res) else []
            org = file.split(".")[0].split("_")[1]
            res = "".join(decodeRes(res))
            if org == res:
                correct += 1
            _diff_time = timer.toc(average=False)
            print("cost time: {:.3f},\n    res: {}".format(_diff_time, res))
        print("total acc:{}/{}={:.4f}".format(correct, total, correct / max(total, 1)))
        return correct, total


def test_net(network, imgdb, testDir, output_dir, log_dir, pretrained_model=None, restore=True):
    with Session() as sess:
        sw = SolverWrapper(sess, network, imgdb, output_dir, logdir=log_dir, pretrained_model=pretrained_model)
        print("Solving...")
        sw.test_model(sess, testDir=testDir, restore=restore)
        print("done solving")
