"""Shim: `from lucyrnn_triton import LucyRNNtriton` (model.py:9, constructed at model.py:310).
The Triton network of that name is retired (BASELINE.json north_star); this is the LucyRNN of
lucyrnn.py on the CUDA kernels."""
from statecatcher_b200.lucyrnn import LucyRNNtriton  # noqa: F401
import _sc_route  # noqa: F401,E402  routes nn.CTCLoss (train.py:142) to the CUDA kernels; SC_SHIM_CTC=0 disables
