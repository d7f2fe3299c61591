"""CPU: host logic of the LucyRNN module with the C-ABI call replaced by a recorder — which kernels a
forward/backward enqueues, with which problem sizes, and the state-tuple contract of lucyrnn.py:101-107 /
188-191 (values are garbage here; the arithmetic is checked on the GPU against the golden vectors)."""
import collections

import pytest
import torch


@pytest.fixture
def rec(monkeypatch):
    from statecatcher_b200 import ops, _lib
    calls = []
    monkeypatch.setattr(ops, "call", lambda name, *a: calls.append((name, a)))
    monkeypatch.setattr(ops, "stream", lambda: 0)
    monkeypatch.setattr(_lib, "require_cuda", lambda t, name: None)
    return calls


def _model(compute_dtype=None, **kw):
    import statecatcher_b200 as sb
    base = dict(input_dim=80, hidden_dim=64, num_layers=2, vocab_size=33, fused_ops=True, layer_norm=False,
                is_training=True)
    base.update(kw)
    cfg = sb.LucyRNNConfig(**base)
    return sb.LucyRNN(cfg, compute_dtype=compute_dtype), cfg


def _gemm_shapes(calls, name="sc_gemm_fwd"):
    return [(a[7], a[8], a[9]) for n, a in calls if n == name]


def test_fp32_training_path_launch_sequence(rec):
    m, cfg = _model()
    B, T, H, V = 3, 10, 64, 33
    logits, (h, s) = m(torch.randn(B, T, 80))
    assert logits.shape == (B, T, V) and logits.dtype == torch.float32
    # every fp32 projection = two operand splits (sc_split6_bf16) + ONE tensor-core GEMM over a six-fold reduction
    # (V = 33 is not a multiple of 8: the output projection stays on the fp32-FMA kernel)
    gemm6 = ["sc_split6_bf16", "sc_split6_bf16", "sc_gemm_fwd"]
    assert [c[0] for c in rec] == (gemm6 * 2 + ["sc_lucy_scan_fwd"]) * 2 + ["sc_gemm_fwd"]
    # (M, N, K): input projection, gate projection with the dead r gate left out (5H, not 6H), output projection
    assert _gemm_shapes(rec) == [(B * T, H, 6 * 80), (B * T, 5 * H, 6 * H), (B * T, H, 6 * H), (B * T, 5 * H, 6 * H), (B * T, V, H)]
    n = len(rec)
    logits.sum().backward()
    cnt = collections.Counter(c[0] for c in rec[n:])
    assert cnt["sc_lucy_scan_bwd"] == 2 and cnt["sc_gemm_wgrad"] == 5 and cnt["sc_colsum"] == 3
    assert cnt["sc_gemm_dgrad"] == 4                         # none into the features (no gradient asked for)
    assert all(p.grad is not None for k, p in m.named_parameters())


def test_bf16_path_folds_the_input_projection(rec):
    m, _ = _model(torch.bfloat16)
    B, T, H = 2, 8, 64
    logits, _ = m(torch.randn(B, T, 80))
    assert logits.dtype == torch.bfloat16
    # one projection GEMM per layer straight from the layer input (W_fused . W_in folded by a small GEMM)
    assert _gemm_shapes(rec) == [(B * T, 5 * H, 80), (B * T, 5 * H, H), (B * T, 33, H)]
    assert sum(c[0] == "sc_lucy_scan_fwd" for c in rec) == 2


def test_state_tuple_contract(rec):
    m, cfg = _model()
    x = torch.randn(2, 6, 80)
    h0 = [torch.zeros(2, 64) for _ in range(2)]
    s0 = [torch.zeros(2, 64) for _ in range(2)]
    s_ids = [id(t) for t in s0]
    _, (h1, s1) = m(x, (h0, s0))
    assert h1 is h0 and s1 is s0                            # the caller's lists come back (lucyrnn.py:188-191)
    assert [id(t) for t in s1] == s_ids                     # training path: s passes through untouched
    assert all(t.dtype == torch.float32 and t.shape == (2, 64) for t in h1)
    ms, _ = _model(is_training=False)
    s_before = [id(t) for t in s0]
    _, (h2, s2) = ms(x, (h0, s0))
    assert s2 is s0 and [id(t) for t in s2] != s_before     # step path: fresh s written into the caller's list
    assert ms(x)[1][0][0].shape == (2, 64)                  # no state given: zeros are created


def test_return_last_states_false_and_stacking(rec):
    m, _ = _model(return_last_states=False, stack_order=3)   # layer 0 takes 80 * 3 features
    out = m(torch.randn(2, 11, 80))
    assert torch.is_tensor(out) and out.shape == (2, 3, 33)  # 11 // 3 frames, remainder trimmed (lucyrnn.py:92-99)
    assert _gemm_shapes(rec)[0] == (2 * 3, 64, 6 * 240)       # six bf16 blocks per fp32 operand (ops.F32_GEMM)


@pytest.mark.parametrize("kw,kernels", [
    (dict(layer_norm=True), {"sc_layernorm_fwd", "sc_lucy_sscan_fwd", "sc_lucy_hscan_fwd"}),
    (dict(fused_ops=False), {"sc_lucy_sscan_fwd", "sc_lucy_hscan_fwd"}),
    (dict(decay_mode="prefix_sum"), {"sc_lucy_sscan_fwd"}),
])
def test_general_configurations_take_the_split_scans(rec, kw, kernels):
    m, _ = _model(**kw)
    logits, _ = m(torch.randn(2, 5, 80))
    n = len(rec)
    assert kernels <= {c[0] for c in rec}
    logits.sum().backward()
    assert any(c[0] == "sc_lucy_sscan_bwd" for c in rec[n:])
    used = {k for k, p in m.named_parameters() if p.grad is not None}
    dead = {k for k, p in m.named_parameters() if p.grad is None}
    assert all(".W_r." in k or "layernorm_r" in k for k in dead), dead     # only the dead r gate gets no gradient
    assert "output_proj.weight" in used


def test_reference_errors(rec):
    import statecatcher_b200 as sb
    with pytest.raises(ValueError):
        sb.LucyRNN(sb.LucyRNNConfig(80, 8, 1, 5, kernel_impl="cuda"))
    m, _ = _model(decay_mode="cumsum")
    with pytest.raises(ValueError):
        m(torch.randn(1, 4, 80))
    m, _ = _model()
    with pytest.raises((ValueError, NotImplementedError)):
        m(torch.randn(1, 4, 80), None, torch.ones(1, 4))
    assert not rec
