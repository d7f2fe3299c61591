"""Shim: lets the reference's model.py:8 (`from lucyrnn_conf import LucyRNNConfig`) pick up the
B200 package.  See INTEGRATION.md."""
from statecatcher_b200.lucyrnn_conf import LucyRNNConfig  # noqa: F401
