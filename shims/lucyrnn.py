"""Shim: `from lucyrnn import LucyRNN` (model.py:7) -> the sm_100a-backed module."""
from statecatcher_b200.lucyrnn import LucyRNN, LucyRNNCell  # noqa: F401
from statecatcher_b200.lucyrnn_conf import LucyRNNConfig  # noqa: F401
import _sc_route  # noqa: F401,E402  routes nn.CTCLoss (train.py:142) to the CUDA kernels; SC_SHIM_CTC=0 disables
