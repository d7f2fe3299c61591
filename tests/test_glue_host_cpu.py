"""CPU: host logic of glue.compute_loss / LucyASRModel (model.py:37-110, 318-398 mirror) with the C-ABI call
replaced by a recorder: the 4-tuple, the truthiness gate on the carried state, the detach at the segment
boundary, the fused CTC head against a foreign criterion, and the mask multiply."""
import types

import pytest
import torch


@pytest.fixture
def rec(monkeypatch):
    from statecatcher_b200 import ops, ctc, _lib
    calls = []
    for mod in (ops, ctc):
        monkeypatch.setattr(mod, "call", lambda name, *a: calls.append((name, a)))
        monkeypatch.setattr(mod, "stream", lambda: 0)
    monkeypatch.setattr(ctc, "ptr", lambda t: t)
    monkeypatch.setattr(_lib, "require_cuda", lambda t, name: None)
    return calls


def _setup(**kw):
    import statecatcher_b200 as sb
    cfg = sb.LucyRNNConfig(input_dim=80, hidden_dim=32, num_layers=2, vocab_size=17, fused_ops=True, layer_norm=False)
    model = sb.LucyASRModel(cfg, **kw)
    feats = torch.randn(3, 12, 80)
    mask = torch.ones(3, 12, dtype=torch.bool)
    mask[2, 7:] = False
    tokens = torch.tensor([[1, 2, 3], [4, 5, 0], [6, 0, 0]])
    return sb, model, feats, mask, tokens, [12, 12, 7], [3, 2, 1]


def test_four_tuple_and_state_handoff(rec):
    sb, model, feats, mask, tokens, inl, tgl = _setup()
    crit = sb.CTCLoss(blank=0, zero_infinity=True)
    loss, st_a, enc_out, st_b = sb.compute_loss("ctc", crit, model, feats, mask, tokens, inl, tgl, 0)
    assert st_a is st_b and enc_out.shape == (3, 12, 17) and loss.dim() == 0 and loss.requires_grad
    names = [c[0] for c in rec]
    assert names[0] == "sc_mask_rows"                                   # model.py:377 as our kernel
    assert names[-2:] == ["sc_ctc_emissions", "sc_ctc_lattice"]
    em = rec[-2][1]
    assert em[0].data_ptr() == enc_out.data_ptr() and (em[1], em[2]) == (12 * 17, 17)   # (B,T,V) read in place
    # segment 2: the state of segment 1 seeds it, detached at the boundary, h tensors identical in storage
    h_ptrs = [t.data_ptr() for t in st_a[0]]
    st_a[0][0].requires_grad_(True)
    seen = {}
    orig = model.encoder.forward

    def spy(x, hs=None, masks=None):       # the module updates the lists in place: look at them on the way in
        seen.update(hs=hs, ptrs=[t.data_ptr() for t in hs[0]] if hs else None,
                    grad=any(t.requires_grad for t in hs[0] + hs[1]) if hs else None)
        return orig(x, hs, masks)
    model.encoder.forward = spy
    args = types.SimpleNamespace(debug=True)
    sb.compute_loss("ctc", crit, model, feats, mask, tokens, inl, tgl, 0, input_state=st_a, args=args)
    assert seen["ptrs"] == h_ptrs and seen["grad"] is False
    # no state: the encoder is called without one (model.py:384-391)
    seen.clear()
    sb.compute_loss("ctc", crit, model, feats, mask, tokens, inl, tgl, 0, input_state=None)
    assert seen["hs"] is None
    # an EMPTY tuple passes the truthiness gate of model.py:60 undetached, is "not None" at model.py:381 and
    # fails to unpack at lucyrnn.py:107 — same exception here
    with pytest.raises(ValueError):
        sb.compute_loss("ctc", crit, model, feats, mask, tokens, inl, tgl, 0, input_state=())


def test_foreign_criterion_gets_the_reference_call(rec):
    sb, model, feats, mask, tokens, inl, tgl = _setup()
    got = {}

    def crit(logp, tok, il, tl):
        got.update(shape=tuple(logp.shape), norm=logp.exp().sum(-1), tok=tok, il=il, tl=tl)
        return logp.sum() * 0

    sb.compute_loss("ctc", crit, model, feats, mask, tokens, inl, tgl, 0)
    assert got["shape"] == (12, 3, 17) and got["il"] == inl and got["tl"] == tgl and got["tok"] is tokens
    assert not any(c[0].startswith("sc_ctc") for c in rec)


def test_mask_with_input_projection_keeps_autograd(rec):
    sb, model, feats, mask, tokens, inl, tgl = _setup(proj_dim=80)
    crit = sb.CTCLoss(blank=0, zero_infinity=True)
    sb.compute_loss("ctc", crit, model, feats, mask, tokens, inl, tgl, 0)
    assert "sc_mask_rows" not in [c[0] for c in rec]        # projected features carry grad: torch multiply
    assert "proj.weight" in dict(model.named_parameters())


def test_modes(rec):
    sb, model, feats, mask, tokens, inl, tgl = _setup()
    with pytest.raises(ValueError):
        sb.compute_loss("attention", None, model, feats, mask, tokens, inl, tgl, 0)
    with pytest.raises(AssertionError):
        sb.compute_loss("rnnt", None, model, feats, mask, tokens, inl, tgl, 0)
    seen = {}

    def joiner(enc_out, prefix, *lens):
        seen.update(prefix=prefix, lens=lens)
        return torch.zeros(3, 12, 4, 17, requires_grad=True)

    def crit(**kw):
        seen.update(kw=kw)
        return kw["log_probs"].sum()

    sb.compute_loss("rnnt", crit, model, feats, mask, tokens, inl, tgl, 0, use_rnnt_joiner=joiner)
    assert seen["prefix"].tolist() == [[0, 1, 2, 3], [0, 4, 5, 0], [0, 6, 0, 0]] and seen["lens"] == ()
    assert set(seen["kw"]) == {"log_probs", "labels", "frames_lengths", "labels_lengths", "blank_id", "compact", "gather"}
    assert seen["kw"]["gather"] is True and seen["kw"]["compact"] is False
    sb.compute_loss("rnnt", crit, model, feats, mask, tokens, inl, tgl, 0, use_rnnt_joiner=joiner,
                    args=types.SimpleNamespace(compact_rnnt=True), compact=True)
    assert seen["lens"] == (inl, tgl) and seen["kw"]["compact"] is True
