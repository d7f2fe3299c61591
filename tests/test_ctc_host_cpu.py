"""CPU: host logic of ctc.ctc_loss (argument forms of torch.nn.functional.ctc_loss: padded or concatenated
targets, list or tensor lengths, unbatched input, strided views) with the C-ABI call replaced by a
recorder that sees the tensors themselves."""
import pytest
import torch


@pytest.fixture
def rec(monkeypatch):
    from statecatcher_b200 import ctc, _lib
    calls = []
    monkeypatch.setattr(ctc, "call", lambda name, *a: calls.append((name, a)))
    monkeypatch.setattr(ctc, "ptr", lambda t: t)
    monkeypatch.setattr(ctc, "stream", lambda: 0)
    monkeypatch.setattr(_lib, "require_cuda", lambda t, name: None)
    return calls


def _em(calls):
    name, a = calls[0]
    assert name == "sc_ctc_emissions"
    return dict(x=a[0], stride_b=a[1], stride_t=a[2], targets=a[4], ldt=a[5], in_lens=a[6], tgt_lens=a[7],
                B=a[8], T=a[9], V=a[10], Umax=a[11], blank=a[12], lplat=a[14])


def test_padded_targets_and_list_lengths(rec):
    from statecatcher_b200.ctc import ctc_loss
    x = torch.randn(4, 20, 7)                                   # (B,T,V) encoder output
    tok = torch.tensor([[1, 2, 3, 0, 0, 0], [4, 0, 0, 0, 0, 0], [5, 6, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0]])
    ctc_loss(x.transpose(0, 1), tok, [20, 20, 11, 20], [3, 1, 2, 0], blank=0, zero_infinity=True)
    e = _em(rec)
    assert (e["B"], e["T"], e["V"]) == (4, 20, 7)
    assert e["x"].data_ptr() == x.data_ptr()                    # the transposed view is read in place
    assert (e["stride_b"], e["stride_t"]) == (20 * 7, 7)
    assert e["Umax"] == 3 and e["ldt"] == 6                     # lattice sized by the longest transcript
    from statecatcher_b200 import _lib
    assert e["lplat"].shape == (4, 20, _lib.load().sc_ctc_lplat_pitch(3))   # >= 8 (2*3+1 = 7 nodes, 16-byte rows): 36, the fp64 kernel's row pitch
    assert e["in_lens"].dtype == torch.int64 and e["in_lens"].tolist() == [20, 20, 11, 20]
    assert e["tgt_lens"].tolist() == [3, 1, 2, 0]
    assert [c[0] for c in rec] == ["sc_ctc_emissions", "sc_ctc_lattice"]
    assert rec[1][1][-3] == 1                                   # reduction code: mean (then ws, stream)


def test_concatenated_targets_are_padded_with_blank(rec):
    from statecatcher_b200.ctc import ctc_loss
    x = torch.randn(9, 3, 5)                                    # (T,B,V)
    flat = torch.tensor([3, 4, 1, 2, 2, 4])
    ctc_loss(x, flat, torch.tensor([9, 8, 9]), torch.tensor([2, 0, 4]), blank=0, reduction="sum")
    e = _em(rec)
    assert e["targets"].tolist() == [[3, 4, 0, 0], [0, 0, 0, 0], [1, 2, 2, 4]]
    assert e["Umax"] == 4 and e["ldt"] == 4
    assert rec[1][1][-3] == 2


def test_unbatched_input(rec):
    from statecatcher_b200.ctc import ctc_loss
    x = torch.randn(12, 6)                                      # (T,V)
    ctc_loss(x, torch.tensor([1, 2, 3]), [12], [3])
    e = _em(rec)
    assert (e["B"], e["T"], e["V"]) == (1, 12, 6) and e["targets"].tolist() == [[1, 2, 3]]


def test_dtype_and_layout_normalisation(rec):
    from statecatcher_b200.ctc import ctc_loss
    x = torch.randn(5, 2, 8, dtype=torch.float64)
    ctc_loss(x, torch.tensor([[1], [2]], dtype=torch.int32), [5, 5], [1, 1])
    e = _em(rec)
    assert e["x"].dtype == torch.float32 and e["targets"].dtype == torch.int64
    rec.clear()
    xs = torch.randn(5, 2, 16)[:, :, ::2]                       # vocabulary axis strided: copied
    ctc_loss(xs, torch.tensor([[1], [2]]), [5, 5], [1, 1])
    assert _em(rec)["x"].stride(2) == 1


def test_argument_errors(rec):
    from statecatcher_b200.ctc import ctc_loss, CTCLoss
    x = torch.randn(5, 2, 8)
    with pytest.raises(ValueError):
        ctc_loss(x, torch.tensor([[1], [2]]), [5], [1, 1])                      # batch-size mismatch
    with pytest.raises(ValueError):
        ctc_loss(x, torch.tensor([[1], [2]]), [5, 5], [1, 2])                   # transcript longer than the tensor
    with pytest.raises(ValueError):
        ctc_loss(x, torch.tensor([[1], [2]]), [5, 5], [1, 1], reduction="avg")
    with pytest.raises(ValueError):
        ctc_loss(torch.randn(2, 2, 2, 2), torch.tensor([[1], [2]]), [5, 5], [1, 1])
    assert not rec
    crit = CTCLoss(blank=3, reduction="none", zero_infinity=True)
    out = crit(x, torch.tensor([[1], [2]]), [5, 5], [1, 1])
    assert out.shape == (2,) and _em(rec)["blank"] == 3 and rec[1][1][-3] == 0
