"""Inference-time network (reference lib/networks/LSTM_test.py:6-34): no label placeholders."""
from ..lstm.config import cfg
from .network import Network, Placeholder


class LSTM_test(Network):
    def __init__(self, trainable=True):
        self.inputs = []
        self.data = Placeholder("data", "float32", [None, None, cfg.NUM_FEATURES])
        self.time_step_len = Placeholder("time_step_len", "int32", [None])
        self.keep_prob = Placeholder("keep_prob", "float32", [])
        self.layers = dict({"data": self.data, "time_step_len": self.time_step_len})
        self.trainable = trainable
        self.setup()

    def setup(self):
        self._declare_graph()
