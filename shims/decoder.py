"""Shim: `from decoder import ctc_greedy_decoder` (train.py) -> the GPU decoder."""
from statecatcher_b200.decoder import ctc_greedy_decoder  # noqa: F401
