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
        from ... import parallel
        eng.backward(d_data, d_tsl, grad)
        world = parallel.world_size()
        if world > 1:
            parallel.allreduce_sum_(eng.grads)
        self.global_step.assign(self.global_step.eval() + 1)
        eng.clip_adam_step(self.lr.eval(), self.global_step.eval(), clip=self.clip, grad_mul=1.0 / world, wd_mul=float(world))
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
    def train_model(self, sess, max_iters, restore=False, train_gen=None, val_gen=None):
        from ... import parallel
        train_gen = train_gen or get_batch(num_workers=12, batch_size=cfg.TRAIN.BATCH_SIZE, vis=False)
        val_gen = val_gen or get_batch(num_workers=1, batch_size=cfg.VAL.BATCH_SIZE, vis=False)
        loss, dense_decoded = self.net.build_loss()
        if cfg.TRAIN.SOLVER != "Adam":
            raise NotImplementedError("only the Adam solver of lstm/lstm.yml is implemented (RMS/Momentum are unused upstream)")
        lr = Variable(cfg.TRAIN.LEARNING_RATE)
        global_step = Variable(0)
        self._lr, self._global_step = lr, global_step
        train_op = TrainOp(self.net, lr, global_step, clip=10.0)
        eng = sess.engine_for(self.net)
        if not getattr(eng, "_initialised", False):
            from ... import synthetic
            eng.load_params(synthetic.init_params(cfg.RNG_SEED))       # global_variables_initializer
            eng._initialised = True
        eng.set_training(True)
        if parallel.world_size() > 1:
            parallel.broadcast_(eng.params)
            eng.lib.crnn_model_params_changed(eng.handle)
        restore_iter = 1
        if restore:
            path = self._latest_checkpoint()
            try:
                print("Restoring from {}...".format(path), end=" ")
                blob = self.restore(sess, path)
                stem = os.path.splitext(os.path.basename(path))[0]
                restore_iter = int(stem.split("_")[-1])
                global_step.assign(restore_iter)
                lr.assign(float(blob["lr"]))
                print("done")
            except Exception:
                raise Exception("Check your pretrained {:s}".format(str(path)))
        timer = Timer()
        loss_min = 0.015
        first_val = True
        history = []
        for iter in range(restore_iter, max_iters):
            timer.tic()
            if iter != 0 and iter % cfg.TRAIN.STEPSIZE == 0:
                lr.assign(lr.eval() * cfg.TRAIN.GAMMA)
            img_Batch, label_Batch, label_len_Batch, time_step_Batch = next(train_gen)
            feed_dict = {
                self.net.data: np.array(img_Batch),
                self.net.labels: np.array(label_Batch),
                self.net.time_step_len: np.array(time_step_Batch),
                self.net.labels_len: np.array(label_len_Batch),
                self.net.keep_prob: 0.5,
            }
            ctc_loss, _ = sess.run(fetches=[loss, train_op], feed_dict=feed_dict)
            history.append(float(ctc_loss))
            _diff_time = timer.toc(average=False)
            if iter % cfg.TRAIN.DISPLAY == 0:
                print("iter: %d / %d, total loss: %.7f, lr: %.7f" % (iter, max_iters, ctc_loss, lr.eval()), end=" ")
                print("speed: {:.3f}s / iter".format(_diff_time))
            if (iter + 1) % cfg.TRAIN.SNAPSHOT_ITERS == 0 or ctc_loss < loss_min:
                if parallel.rank() == 0:
                    if ctc_loss < loss_min:
                        print("loss: ", ctc_loss, end=" ")
                        self.snapshot(sess, 1)
                        loss_min = ctc_loss
                    else:
                        self.snapshot(sess, iter)
            if (iter + 1) % cfg.VAL.VAL_STEP == 0 or loss_min == ctc_loss:
                if first_val:
                    val_img_Batch, val_label_Batch, val_label_len_Batch, val_time_step_Batch = next(val_gen)
                    org = self.restoreLabel(val_label_Batch, val_label_len_Batch)
                    first_val = False
                feed_dict = {
                    self.net.data: np.array(val_img_Batch),
                    self.net.labels: np.array(val_label_Batch),
                    self.net.time_step_len: np.array(val_time_step_Batch),
                    self.net.labels_len: np.array(val_label_len_Batch),
                    self.net.keep_prob: 1.0,
                }
                res = sess.run(fetches=dense_decoded, feed_dict=feed_dict)
                acc = accuracy_calculation(org, res, ignore_value=0)
                print("accuracy: {:.5f}".format(acc))
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
