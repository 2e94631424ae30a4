"""Network object with the reference's attribute surface (lib/networks/network.py:40-95,647-664).

The reference builds a TF1 graph through a chaining DSL; here the graph is fixed (it is the one
LSTM_train.setup/LSTM_test.setup build, LSTM_train.py:22-38) and runs as hand-written sm_100a
kernels in libcrnnctc.so.  What is preserved is the *surface* the solver touches: placeholders
``data/labels/time_step_len/labels_len/keep_prob``, the ``layers`` dict, ``feed``/``get_output``,
``build_loss() -> (loss, dense_decoded)``; handles are evaluated by ``Session.run``."""
from ..lstm.config import cfg


class Placeholder(object):
    """Stand-in for tf.placeholder: a named, typed slot filled through ``feed_dict``."""

    def __init__(self, name, dtype, shape):
        self.name, self.dtype, self.shape = name, dtype, shape

    def __repr__(self):
        return f"<Placeholder {self.name} {self.dtype} {self.shape}>"


class Fetch(object):
    """Stand-in for a TF tensor/op handle: something ``Session.run`` can evaluate."""

    def __init__(self, net, kind, name=None):
        self.net, self.kind, self.name = net, kind, name or kind

    def __repr__(self):
        return f"<Fetch {self.name}>"


# layer names of LSTM_train.setup in order (LSTM_train.py:24-38); 'pool2' is overwritten by the
# second max_pool of the same name exactly as in the reference (network.py:33).
LAYER_NAMES = ["conv1", "pool1", "conv2", "pool2", "conv3_1", "conv3_2", "conv4_1", "conv4_2", "pool3", "conv5",
               "reshaped_layer", "logits"]


class Network(object):
    def __init__(self, inputs, trainable=True):
        self.inputs = []
        self.layers = dict(inputs)
        self.trainable = trainable
        self.setup()

    def setup(self):
        raise NotImplementedError("Must be subclassed.")

    def feed(self, *args):
        assert len(args) != 0
        self.inputs = []
        for layer in args:
            if isinstance(layer, str):
                try:
                    layer = self.layers[layer]
                except KeyError:
                    print(list(self.layers.keys()))
                    raise KeyError("Unknown layer name fed: %s" % layer)
            self.inputs.append(layer)
        return self

    def get_output(self, layer):
        try:
            return self.layers[layer]
        except KeyError:
            print(list(self.layers.keys()))
            raise KeyError("Unknown layer name fed: %s" % layer)

    def get_unique_name(self, prefix):
        n = sum(t.startswith(prefix) for t in self.layers) + 1
        return "%s_%d" % (prefix, n)

    def _declare_graph(self):
        for name in LAYER_NAMES:
            self.layers[name] = Fetch(self, "layer:" + name, name)
        self.layers["logits"] = Fetch(self, "logits")

    def load(self, data_path, session, ignore_missing=False):
        """npy dict {scope: {var: array}} loader (network.py:50-63)."""
        import numpy as np
        data = np.load(data_path, allow_pickle=True, encoding="latin1").item()
        sd = {}
        for scope, sub in data.items():
            for var, arr in sub.items():
                sd[f"{scope}/{var}"] = arr
        session.assign(self, sd, ignore_missing=ignore_missing)

    def build_loss(self):
        """(loss, dense_decoded) handles; semantics of network.py:647-664 with the greedy decode of
        SURVEY §8(c) in place of the width-100 beam search."""
        if "labels" not in self.layers:
            raise KeyError("Unknown layer name fed: labels")
        self._wd = float(cfg.TRAIN.WEIGHT_DECAY)
        return Fetch(self, "loss"), Fetch(self, "dense_decoded")
