"""Shim: lets the reference's model.py:8 (`from lucyrnn_conf import LucyRNNConfig`) pick up the
B200 package.  See INTEGRATION.md."""
from statecatcher_b200.lucyrnn_conf import LucyRNNConfig  # noqa: F401
import _sc_route  # noqa: F401,E402  routes nn.CTCLoss (train.py:142) to the CUDA kernels; SC_SHIM_CTC=0 disables
