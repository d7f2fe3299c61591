"""Shim: `from lion_pytorch import Lion` (train.py:126) -> the fused sm_100a Lion step.
The real package is absent and unpinned upstream; the published rule is implemented."""
from statecatcher_b200.optim import Lion  # noqa: F401
