"""Shim: `from warp_rnnt import RNNTLoss` (train.py:39) -> the sm_100a wavefront loss.
Parity with the real warp_rnnt is unpinned."""
from statecatcher_b200.rnnt import RNNTLoss, rnnt_loss  # noqa: F401
