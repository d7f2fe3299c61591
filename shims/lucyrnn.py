"""Shim: `from lucyrnn import LucyRNN` (model.py:7) -> the sm_100a-backed module."""
from statecatcher_b200.lucyrnn import LucyRNN, LucyRNNCell  # noqa: F401
from statecatcher_b200.lucyrnn_conf import LucyRNNConfig  # noqa: F401
