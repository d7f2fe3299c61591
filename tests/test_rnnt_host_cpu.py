"""CPU: host logic of rnnt.rnnt_loss / RNNTLoss (keyword call of model.py:97-105) with the C-ABI call
replaced by a recorder that sees the tensors: padded and compact layouts, offsets, workspace shapes, errors."""
import pytest
import torch


@pytest.fixture
def rec(monkeypatch):
    from statecatcher_b200 import rnnt, _lib
    calls = []
    monkeypatch.setattr(rnnt, "call", lambda name, *a: calls.append((name, a)))
    monkeypatch.setattr(rnnt, "ptr", lambda t: t)
    monkeypatch.setattr(rnnt, "stream", lambda: 0)
    monkeypatch.setattr(_lib, "require_cuda", lambda t, name: None)
    return calls


def test_padded_keyword_call(rec):
    import statecatcher_b200 as sb
    B, T, U, V = 3, 6, 4, 9
    lp = torch.randn(B, T, U + 1, V, requires_grad=True)
    labels = torch.randint(1, V, (B, U))
    loss = sb.RNNTLoss(log_probs=lp, labels=labels, frames_lengths=[6, 5, 6], labels_lengths=[4, 2, 0],
                       blank_id=0, compact=False, gather=True)
    assert loss.dim() == 0
    name, a = rec[0]
    assert name == "sc_rnnt_fwd"
    assert a[2] == U and a[3].tolist() == [6, 5, 6] and a[4].tolist() == [4, 2, 0]
    assert a[5:10] == (B, T, U + 1, V, 0) and a[10] is None          # no row offsets: padded layout
    assert a[11].shape == (B, T + U + 1, 8)                          # skewed [B][T+U1][U1 padded to 4]
    loss.backward()
    name, b = rec[1]
    assert name == "sc_rnnt_bwd" and b[10] == 0 and b[-2].shape == lp.shape
    assert b[-3].shape == (B,) and torch.allclose(b[-3], torch.full((B,), 1.0 / B))   # d mean / d nll_b


def test_compact_layout_offsets(rec):
    import statecatcher_b200 as sb
    fl, ll, V = [4, 0, 3], [2, 1, 0], 7
    rows = sum(t * (u + 1) for t, u in zip(fl, ll))                    # 12 + 0 + 3
    lp = torch.randn(rows, V)
    out = sb.rnnt_loss(lp, torch.zeros(3, 2, dtype=torch.int32), fl, ll, blank=0, reduction="none", compact=True)
    assert out.shape == (3,)
    a = rec[0][1]
    assert a[10].tolist() == [0, 12, 12] and (a[6], a[7]) == (4, 3)    # T = max T_b, U1 = max U_b + 1
    assert a[1].dtype == torch.int64
    with pytest.raises(ValueError):
        sb.rnnt_loss(torch.randn(rows + 1, V), torch.zeros(3, 2, dtype=torch.int64), fl, ll, compact=True)
    with pytest.raises(ValueError):
        sb.rnnt_loss(torch.randn(3, 4, 3, V), torch.zeros(3, 2, dtype=torch.int64), fl, ll, compact=True)


def test_argument_errors(rec):
    import statecatcher_b200 as sb
    lp = torch.randn(2, 5, 4, 6)
    with pytest.raises(ValueError):
        sb.rnnt_loss(lp, torch.zeros(2, 2, dtype=torch.int64), [5, 5], [2, 2])       # labels narrower than U
    with pytest.raises(ValueError):
        sb.rnnt_loss(lp, torch.zeros(2, 3, dtype=torch.int64), [5], [2, 2])          # batch mismatch
    with pytest.raises(ValueError):
        sb.rnnt_loss(lp[0], torch.zeros(2, 3, dtype=torch.int64), [5, 5], [2, 2])    # not 4-D
    with pytest.raises(ValueError):
        sb.rnnt_loss(lp, torch.zeros(2, 5, dtype=torch.int64), [5, 5], [2, 4])       # a transcript longer than the lattice (U = 3)
    assert not rec
    with pytest.raises(ValueError):
        sb.rnnt_loss(lp, torch.zeros(2, 3, dtype=torch.int64), [5, 5], [2, 2], reduction="avg")
