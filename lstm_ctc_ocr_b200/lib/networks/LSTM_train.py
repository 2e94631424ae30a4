"""Training-time network (reference lib/networks/LSTM_train.py:6-38)."""
from ..lstm.config import cfg
from .network import Network, Placeholder


class LSTM_train(Network):
    def __init__(self, trainable=True):
        self.inputs = []
        self.data = Placeholder("data", "float32", [None, None, cfg.NUM_FEATURES])     # N x W x 32
        self.labels = Placeholder("labels", "int32", [None])
        self.time_step_len = Placeholder("time_step_len", "int32", [None])
        self.labels_len = Placeholder("labels_len", "int32", [None])
        self.keep_prob = Placeholder("keep_prob", "float32", [])                       # accepted and ignored
        self.layers = dict({"data": self.data, "labels": self.labels, "time_step_len": self.time_step_len,
                            "labels_len": self.labels_len})
        self.trainable = trainable
        self.setup()

    def setup(self):
        # conv1..conv5 -> reshaped_layer -> bi_lstm('logits'): fixed graph, see csrc/model.cu crnn_forward
        self._declare_graph()
