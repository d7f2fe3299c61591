"""CPU: "train.py runs unchanged" at the import level.  With `shims/` in front of
the reference on the module path, the reference's own model.py (model.py:7-9, 310) builds the
CUDA-backed encoder, and the criterion train.py:142 builds — `nn.CTCLoss(blank=blank_id,
zero_infinity=True)` — is the routed class that sends CUDA inputs to the K3 kernels.  The GPU half is
tests/test_gpu_shim_route.py."""
import importlib

import pytest
import torch
import torch.nn as nn

from shim_env import have_reference, shim_imports


def test_ctc_criterion_is_routed_by_importing_a_shim(monkeypatch):
    monkeypatch.delenv("SC_SHIM_CTC", raising=False)
    torch_cls = nn.CTCLoss
    with shim_imports(with_reference=False):
        importlib.import_module("lucyrnn")                         # what model.py:7 does
        routed = nn.CTCLoss
        assert routed is not torch_cls and issubclass(routed, torch_cls) and routed.__name__ == "CTCLoss"
        crit = nn.CTCLoss(blank=0, zero_infinity=True)             # train.py:142
        assert isinstance(crit, torch_cls) and crit.blank == 0 and crit.zero_infinity and crit.reduction == "mean"
        # CPU tensors keep torch's implementation (the kernels have no CPU path)
        g = torch.Generator().manual_seed(0)
        logp = torch.randn(12, 2, 7, generator=g).log_softmax(-1)
        tok = torch.randint(1, 7, (2, 3), generator=g)
        want = torch.nn.functional.ctc_loss(logp, tok, [12, 9], [3, 2], blank=0, zero_infinity=True)
        assert torch.equal(crit(logp, tok, [12, 9], [3, 2]), want)
    assert nn.CTCLoss is torch_cls                                  # the helper restores torch's class


def test_ctc_route_switch_off(monkeypatch):
    monkeypatch.setenv("SC_SHIM_CTC", "0")
    torch_cls = nn.CTCLoss
    with shim_imports(with_reference=False):
        importlib.import_module("lucyrnn_conf")
        assert nn.CTCLoss is torch_cls


@pytest.mark.skipif(not have_reference(), reason="reference tree not mounted (GPU box)")
def test_reference_model_py_builds_the_cuda_encoder_through_shims(monkeypatch):
    """The reference's OWN model.py, imported unmodified through the shims: `ASRModel(encoder=LucyRNNConfig)`
    (model.py:308-311) constructs our module under the reference's class name, with the reference
    module's state_dict keys and shapes; its own `compute_loss` reaches our forward (which refuses CPU
    tensors loudly — there is no CPU fallback to fall into)."""
    import importlib.util
    import statecatcher_b200 as sb
    monkeypatch.delenv("SC_SHIM_CTC", raising=False)
    with shim_imports() as use_ref:
        assert use_ref
        model = importlib.import_module("model")                    # /root/reference/model.py
        assert model.__file__.startswith("/root/reference/")
        assert model.LucyRNN is sb.LucyRNN and model.LucyRNNConfig is sb.LucyRNNConfig
        assert issubclass(model.LucyRNNtriton, sb.LucyRNN)
        import argparse
        args = argparse.Namespace(encoder="lucyrnn", input_proj_dim=-1, hidden_size=16, num_layers=2)   # train.py:608-655
        cfg = model.build_encoder(args, 33)                         # model.py:202, 231-245
        assert isinstance(cfg, sb.LucyRNNConfig) and cfg.fused_ops and not cfg.layer_norm and cfg.kernel_impl == "triton"
        asr = model.ASRModel(None, cfg, cfg.vocab_size, 80, -1, debug=False)
        assert type(asr.encoder) is model.LucyRNNtriton and isinstance(asr.encoder, sb.LucyRNN)
        # same parameter names / shapes / order as the reference's native module (checkpoints interchange)
        spec = importlib.util.spec_from_file_location("_ref_lucyrnn_native", "/root/reference/lucyrnn.py")
        ref_native = importlib.util.module_from_spec(spec)
        with shim_free_names():
            spec.loader.exec_module(ref_native)
        ref_cfg = ref_native.LucyRNNConfig(**{**cfg.__dict__, "kernel_impl": "native"})
        ref_sd = ref_native.LucyRNN(ref_cfg).state_dict()
        ours = asr.encoder.state_dict()
        assert list(ours.keys()) == list(ref_sd.keys())
        assert [tuple(v.shape) for v in ours.values()] == [tuple(v.shape) for v in ref_sd.values()]
        # train.py:142 + model.py:60-71, on CPU tensors: the reference's compute_loss calls OUR forward
        crit = nn.CTCLoss(blank=0, zero_infinity=True)
        assert getattr(type(crit), "_statecatcher_b200_routed", False)
        feats = torch.randn(2, 11, 80)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            model.compute_loss("ctc", crit, asr, feats, torch.ones(2, 11, dtype=torch.bool), torch.ones(2, 3, dtype=torch.long),
                               [11, 11], [3, 2], blank_id=0)


import contextlib  # noqa: E402


@contextlib.contextmanager
def shim_free_names():
    """Import the reference's lucyrnn.py with ITS OWN lucyrnn_conf / lucyrnn_triton (not the shims)."""
    import sys
    names = ("lucyrnn_conf", "lucyrnn_triton")
    saved = {k: sys.modules.pop(k) for k in names if k in sys.modules}
    path = list(sys.path)
    sys.path.insert(0, "/root/reference")
    try:
        yield
    finally:
        sys.path[:] = path
        for k in names:
            sys.modules.pop(k, None)
        sys.modules.update(saved)
