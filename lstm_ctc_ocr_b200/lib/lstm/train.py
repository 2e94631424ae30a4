"""Training solver with the reference's surface (lib/lstm/train.py): ``SolverWrapper(sess, network, imgdb, pre_train,
output_dir, logdir)``, ``snapshot``, ``restoreLabel``, ``train_model(sess, max_iters, restore)``, ``train_net(...)``.

Every iteration runs ``sess.run([loss, train_op], feed_dict)`` (train.py:129-130): forward, CTC loss + gradient, backward,
global-norm clip 10.0, Adam -- all on the GPU engine.  Checkpoints are ``.npz`` files holding the TF variable names/layouts,
Adam slots, lr and the step, named ``<prefix>_ctc_iter_<k>.ckpt.npz`` with a TF-style ``checkpoint`` index file."""
import os
import re

import numpy as np

from ...session import Session
from ..networks.network import Fetch
from .config import cfg
from .utils.gen import get_batch
from .utils.timer import Timer
from .utils.training import accuracy_calculation


class Variable(object):
    """Scalar host variable (the reference keeps lr / global_step as tf.Variables, train.py:73,78)."""

    def __init__(self, value):
        self.value = value

    def eval(self):
        return self.value

    def assign(self, v):
        self.value = v
        return self


class TrainOp(Fetch):
    """What ``opt.apply_gradients(zip(clip_by_global_norm(tf.gradients(loss, tvars), 10.0), tvars), global_step)`` returns."""

    def __init__(self, net, lr, global_step, clip=10.0):
        super().__init__(net, "train_op")
        self.lr, self.global_step, self.clip = lr, global_step, clip

    def step_fn(self, eng, logits, grad, d_data, d_tsl):
        eng.backward(d_data, d_tsl, grad)             # data parallel: announces gradient buckets, reduced on a side stream meanwhile
        self.global_step.assign(self.global_step.eval() + 1)
        dp = getattr(eng, "_dp", None)
        if dp is not None:
            dp.step(self.lr.eval(), self.global_step.eval(), clip=self.clip)       # waits for the buckets; clip + Adam on the SUM / world
        else:
            eng.clip_adam_step(self.lr.eval(), self.global_step.eval(), clip=self.clip)
        return None


class SolverWrapper(object):
    def __init__(self, sess, network, imgdb, pre_train, output_dir, logdir):
        self.net = network
        self.imgdb = imgdb
        self.pre_train = pre_train
        self.output_dir = output_dir
        self.logdir = logdir
        self.sess = sess
        print("done")

    # ---- checkpoints (train.py:23-37, 96-106) ---------------------------------------------------------------
    def snapshot(self, sess, iter):
        os.makedirs(self.output_dir, exist_ok=True)
        infix = ("_" + cfg.TRAIN.SNAPSHOT_INFIX) if cfg.TRAIN.SNAPSHOT_INFIX != "" else ""
        filename = cfg.TRAIN.SNAPSHOT_PREFIX + "_ctc" + infix + "_iter_{:d}".format(iter + 1) + ".ckpt"
        path = os.path.join(self.output_dir, filename)
        eng = sess.engine_for(self.net)
        blob = dict(eng.state_dict())
        if eng.adam_m is not None:
            for k, (off, shp) in eng.table.items():
                n = int(np.prod(shp))
                blob["adam_m/" + k] = eng.adam_m[off:off + n].view(*shp).cpu().numpy()
                blob["adam_v/" + k] = eng.adam_v[off:off + n].view(*shp).cpu().numpy()
        blob["global_step"] = np.array(getattr(self, "_global_step", Variable(0)).eval())
        blob["lr"] = np.array(getattr(self, "_lr", Variable(cfg.TRAIN.LEARNING_RATE)).eval())
        np.savez(path + ".npz", **blob)
        with open(os.path.join(self.output_dir, "checkpoint"), "w") as f:
            f.write('model_checkpoint_path: "{}"\n'.format(filename))
        print("Wrote snapshot to: {:s}".format(path))
        return path

    def _latest_checkpoint(self):
        idx = os.path.join(self.output_dir, "checkpoint")
        if not os.path.exists(idx):
            return None
        m = re.search(r'model_checkpoint_path: "(.*)"', open(idx).read())
        return os.path.join(self.output_dir, m.group(1)) if m else None

    def restore(self, sess, path):
        blob = np.load(path + ".npz")
        eng = sess.engine_for(self.net)
        eng.load_params({k: blob[k] for k in eng.table})
        if "adam_m/" + next(iter(eng.table)) in blob.files and eng.adam_m is not None:
            for k, (off, shp) in eng.table.items():
                n = int(np.prod(shp))
                eng.adam_m[off:off + n].copy_(_to_dev(blob["adam_m/" + k], eng))
                eng.adam_v[off:off + n].copy_(_to_dev(blob["adam_v/" + k], eng))
        return blob

    def restoreLabel(self, label_vec, label_len):
        labels = []
        for l_len in label_len:
            labels.append(label_vec[:l_len])
            label_vec = label_vec[l_len:]
        return labels

    # ---- the loop (train.py:63-162) -----------------------------------------------------------------------------
    def _feed(self, batch, keep_prob):
        """feed_dict for one data-layer tuple (train.py:119-127)."""
        imgs, flat_labels, label_len, time_steps = batch
        net = self.net
        # a PrefetchFeeder hands out an ndarray view of a page-locked ring slot: keep it (np.array would copy it to pageable memory)
        data = imgs if isinstance(imgs, np.ndarray) and imgs.ndim == 3 else np.array(imgs)
        return {net.data: data, net.labels: np.array(flat_labels), net.time_step_len: np.array(time_steps),
                net.labels_len: np.array(label_len), net.keep_prob: keep_prob}

    def _prepare(self, sess, restore, lr, global_step):
        """Variable initialisation, data-parallel broadcast and the optional resume (train.py:88-106)."""
        from ... import parallel, synthetic
        eng = sess.engine_for(self.net)
        if not getattr(eng, "_initialised", False):
            eng.load_params(synthetic.init_params(cfg.RNG_SEED))       # global_variables_initializer
            eng._initialised = True
        eng.set_training(True)
        if parallel.world_size() > 1 and getattr(eng, "_dp", None) is None:
            # parameter broadcast, global-batch BatchNorm over peer memory, overlapped gradient buckets (parallel.DataParallel)
            eng._dp = parallel.DataParallel(eng, sync_bn=bool(cfg.TRAIN.get("SYNC_BN", True)))
        if not restore:
            return 1
        path = self._latest_checkpoint()
        try:
            print("Restoring from {}...".format(path), end=" ")
            blob = self.restore(sess, path)
            first_iter = int(os.path.splitext(os.path.basename(path))[0].split("_")[-1])   # iteration from the file name
            global_step.assign(first_iter)
            lr.assign(float(blob["lr"]))
            print("done")
            return first_iter
        except Exception:
            raise Exception("Check your pretrained {:s}".format(str(path)))

    def _validate(self, sess, dense_decoded, val_gen, cache):
        """Accuracy on ONE cached validation batch (train.py:145-162)."""
        if "batch" not in cache:
            cache["batch"] = next(val_gen)
            cache["org"] = self.restoreLabel(cache["batch"][1], cache["batch"][2])
        res = sess.run(fetches=dense_decoded, feed_dict=self._feed(cache["batch"], 1.0))
        acc = accuracy_calculation(cache["org"], res, ignore_value=0)
        print("accuracy: {:.5f}".format(acc))
        return acc

    def train_model(self, sess, max_iters, restore=False, train_gen=None, val_gen=None):
        from ... import parallel
        if cfg.TRAIN.SOLVER != "Adam":
            raise NotImplementedError("only the Adam solver of lstm/lstm.yml is implemented (RMS/Momentum are unused upstream)")
        train_gen = train_gen or get_batch(num_workers=12, batch_size=cfg.TRAIN.BATCH_SIZE, vis=False)
        val_gen = val_gen or get_batch(num_workers=1, batch_size=cfg.VAL.BATCH_SIZE, vis=False)
        loss, dense_decoded = self.net.build_loss()
        lr, global_step = Variable(cfg.TRAIN.LEARNING_RATE), Variable(0)
        self._lr, self._global_step = lr, global_step
        train_op = TrainOp(self.net, lr, global_step, clip=10.0)
        first_iter = self._prepare(sess, restore, lr, global_step)
        if hasattr(sess, "attach_feeder"):
            sess.attach_feeder(train_gen)                  # PrefetchFeeder: the next batch's H2D copy overlaps the current step (no-op for plain generators)
        timer, history, val_cache = Timer(), [], {}
        loss_min = 0.015                                   # best-loss snapshot threshold (train.py:109)
        is_chief = parallel.rank() == 0
        for iter in range(first_iter, max_iters):
            timer.tic()
            if iter != 0 and iter % cfg.TRAIN.STEPSIZE == 0:
                lr.assign(lr.eval() * cfg.TRAIN.GAMMA)
            ctc_loss, _ = sess.run(fetches=[loss, train_op], feed_dict=self._feed(next(train_gen), 0.5))
            history.append(float(ctc_loss))
            step_seconds = timer.toc(average=False)
            if iter % cfg.TRAIN.DISPLAY == 0:
                print("iter: %d / %d, total loss: %.7f, lr: %.7f" % (iter, max_iters, ctc_loss, lr.eval()), end=" ")
                print("speed: {:.3f}s / iter".format(step_seconds))
            new_best = ctc_loss < loss_min
            if new_best or (iter + 1) % cfg.TRAIN.SNAPSHOT_ITERS == 0:
                if is_chief:
                    if new_best:
                        print("loss: ", ctc_loss, end=" ")
                        self.snapshot(sess, 1)             # the reference always names the best-loss snapshot iter_2
                    else:
                        self.snapshot(sess, iter)
                if new_best:
                    loss_min = ctc_loss
            if new_best or (iter + 1) % cfg.VAL.VAL_STEP == 0:
                self._validate(sess, dense_decoded, val_gen, val_cache)
        return history


def _to_dev(arr, eng):
    import torch
    return torch.as_tensor(np.asarray(arr, dtype=np.float32).reshape(-1), device=eng.device)


def train_net(network, imgdb, pre_train, output_dir, log_dir, max_iters=40000, restore=False):
    with Session() as sess:
        sw = SolverWrapper(sess, network, imgdb, pre_train, output_dir, logdir=log_dir)
        print("Solving...")
        sw.train_model(sess, max_iters, restore=restore)
        print("done solving")
