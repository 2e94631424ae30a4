"""name -> network object (reference lib/networks/factory.py:13-21)."""
from .LSTM_test import LSTM_test
from .LSTM_train import LSTM_train

_REGISTRY = {"train": LSTM_train, "test": LSTM_test}


def get_network(name):
    parts = name.split("_")
    if parts[0] == "LSTM":
        if len(parts) > 1 and parts[1] in _REGISTRY:
            return _REGISTRY[parts[1]]()
        raise KeyError("Unknown dataset: {}".format(name))
    return None     # the reference falls through and returns None for non-LSTM names


def list_networks():
    return ["LSTM_" + k for k in _REGISTRY]
