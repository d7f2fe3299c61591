"""Test helper: the import environment INTEGRATION.md section 1 describes — `shims/` in front of the
reference directory on the module path — set up and torn down around a test.  With the reference
tree present (the authoring container) the reference's own model.py is importable through the shims
(`xlstm`, its one missing dependency, stubbed as in tests/golden/make_glue_golden.py); on the GPU box
it is not there and only the shim modules themselves resolve."""
import contextlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = os.path.join(ROOT, "shims")
REFERENCE = "/root/reference"
_NAMES = ("lucyrnn", "lucyrnn_conf", "lucyrnn_triton", "decoder", "warp_rnnt", "lion_pytorch", "model", "_sc_route",
          "xlstm", "xlstm.xlstm_large", "xlstm.xlstm_large.model")


def have_reference():
    return os.path.isfile(os.path.join(REFERENCE, "model.py"))


@contextlib.contextmanager
def shim_imports(with_reference=True):
    import torch.nn as nn
    saved = {k: sys.modules.pop(k) for k in _NAMES if k in sys.modules}
    saved_path = list(sys.path)
    sys.path.insert(0, SHIMS)
    use_ref = with_reference and have_reference()
    if use_ref:
        sys.path.insert(1, REFERENCE)
        xm = types.ModuleType("xlstm.xlstm_large.model")
        xm.xLSTMLargeConfig = type("xLSTMLargeConfig", (), {})
        xm.xLSTMLarge = type("xLSTMLarge", (nn.Module,), {})
        sys.modules.update({"xlstm": types.ModuleType("xlstm"), "xlstm.xlstm_large": types.ModuleType("xlstm.xlstm_large"),
                            "xlstm.xlstm_large.model": xm})
    try:
        yield use_ref
    finally:
        route = sys.modules.get("_sc_route")
        if route is not None:
            route.uninstall()
        sys.path[:] = saved_path
        for k in _NAMES:
            sys.modules.pop(k, None)
        sys.modules.update(saved)
