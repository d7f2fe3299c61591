"""GPU: sc_ctc_head (loss and gradient in one call, the V-wide passes on side streams under recursions that run as
several launches over frame ranges) against the three passes one after the other — same kernels, so everything
is compared bit for bit — and against torch's fp64 CTC (model.py:70-71, train.py:142)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(B, T, V, U, seed, dtype, ragged=True):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(B, T, V, generator=g) * 1.5).to(dtype).cuda()
    tok = torch.randint(1, V, (B, max(U, 1)), generator=g).cuda()
    il = torch.full((B,), T, dtype=torch.int64)
    tl = torch.randint(max(U // 2, 0), U + 1, (B,), generator=g)
    if ragged and B >= 4:
        il[1] = T - 37                     # ends inside a chunk
        il[2] = max(T // 2, 1)             # ends at / near a chunk edge
        il[3] = 3                          # shorter than its transcript can be: infeasible when tl[3] > 3
        tl[3] = min(U, 5)
    if B >= 5:
        tl[4] = 0                          # empty transcript
    if B >= 6:
        il[5] = 0                          # no frames at all
    return x, tok, il.cuda(), tl.cuda()


def _run(x, tok, il, tl, U, phases, overlapped, red=1, out_dtype=None):
    """-> dict of everything the passes produce.  overlapped=False: emissions, lattice, gradient (grad_out = 1)."""
    from statecatcher_b200._lib import call, dt, ptr, stream, load
    from statecatcher_b200 import ctc
    B, T, V = x.shape
    S = (2 * U + 1 + 7) & ~7
    f32 = dict(dtype=torch.float32, device="cuda")
    lse, lplat, csh = torch.zeros(B, T, **f32), torch.zeros(B, T, load().sc_ctc_lplat_pitch(U), **f32), torch.zeros(B, T, **f32)
    alpha, beta = torch.zeros(B, T, S, **f32), torch.zeros(B, T, S, **f32)
    nll, loss = torch.zeros(B, **f32), torch.zeros((), **f32)
    ws = torch.zeros(load().sc_ctc_workspace_bytes(B, T, U) // 8 + 1, dtype=torch.float64, device="cuda")
    dx = torch.full((B, T, V), float("nan"), dtype=out_dtype or x.dtype, device="cuda")
    if overlapped:
        sl, se, sg = ctc._head_streams(x.device)
        call("sc_ctc_head", ptr(x), x.stride(0), x.stride(1), dt(x), ptr(tok), tok.stride(0), ptr(il), ptr(tl),
             B, T, V, U, 0, ptr(lse), ptr(lplat), ptr(csh), ptr(alpha), ptr(beta), ptr(nll), ptr(loss), red, ptr(ws),
             ptr(dx), dx.stride(0), dx.stride(1), dt(dx), phases, stream(), sl, se, sg)
    else:
        one = torch.ones((), **f32)
        call("sc_ctc_fwd", ptr(x), x.stride(0), x.stride(1), dt(x), ptr(tok), tok.stride(0), ptr(il), ptr(tl),
             B, T, V, U, 0, ptr(lse), ptr(lplat), ptr(csh), ptr(alpha), ptr(beta), ptr(nll), ptr(loss), red, ptr(ws), stream())
        call("sc_ctc_bwd", ptr(x), x.stride(0), x.stride(1), dt(x), ptr(tok), tok.stride(0), ptr(il), ptr(tl),
             B, T, V, U, 0, ptr(lse), ptr(alpha), ptr(beta), ptr(nll), ptr(one), red,
             ptr(dx), dx.stride(0), dx.stride(1), dt(dx), ptr(ws), stream())
    torch.cuda.synchronize()
    return dict(lse=lse, csh=csh, nll=nll, loss=loss, dx=dx, alpha=alpha.view(torch.int32), beta=beta.view(torch.int32),
                il=il, tl=tl)


def _same(a, b, what):
    """Bit for bit where the sequential passes define the content: rows of live frames of valid utterances."""
    for k in ("nll", "loss"):
        assert torch.equal(a[k], b[k]), (what, k)
    assert torch.equal(a["dx"].view(torch.int16 if a["dx"].dtype == torch.bfloat16 else torch.int32),
                       b["dx"].view(torch.int16 if b["dx"].dtype == torch.bfloat16 else torch.int32)), (what, "dx")
    B, T = a["lse"].shape
    for bi in range(B):
        n = min(int(a["il"][bi]), T)
        for k in ("lse", "csh", "alpha", "beta"):
            assert torch.equal(a[k][bi, :n], b[k][bi, :n]), (what, k, bi)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("V", [96, 256], ids=["V96", "V256"])
@pytest.mark.parametrize("T,phases", [(700, 2), (700, 3), (1000, 5), (1000, 8), (450, 0), (3000, 0), (130, 2), (64, 4)])
def test_head_equals_the_three_passes(cuda_device, dtype, T, phases, V):
    from statecatcher_b200._lib import load
    B, U = 7, 40
    x, tok, il, tl = _case(B, T, V, U, 100 + T + phases, dtype)
    P = load().sc_ctc_head_phases(T, U, phases)
    assert P >= 2 or T <= 64 or (phases == 0 and T < 768)
    ref = _run(x, tok, il, tl, U, phases, overlapped=False)
    got = _run(x, tok, il, tl, U, phases, overlapped=True)
    _same(ref, got, f"T={T} P={P}")
    assert torch.isinf(got["nll"][3]) and not got["dx"][3].any()          # infeasible: zero rows
    assert not got["dx"][1, int(il[1]):].any()                            # frames past the utterance's end: zero rows
    assert not torch.isnan(got["dx"].float()).any()


def test_head_flagged_utterances_are_redone(cuda_device, monkeypatch):
    """Every utterance sent down the log-domain recomputation (SC_CTC_FORCE_LOSSY=1): the speculative gradient rows,
    formed from the fp64 rows before the check ran, are replaced by the fix-up launch."""
    B, T, V, U = 6, 900, 64, 30
    x, tok, il, tl = _case(B, T, V, U, 7, torch.float32)
    monkeypatch.setenv("SC_CTC_FORCE_LOSSY", "1")
    ref = _run(x, tok, il, tl, U, 4, overlapped=False)
    got = _run(x, tok, il, tl, U, 4, overlapped=True)
    _same(ref, got, "forced")
    monkeypatch.delenv("SC_CTC_FORCE_LOSSY")
    lin = _run(x, tok, il, tl, U, 4, overlapped=True)
    assert not torch.equal(lin["alpha"], got["alpha"])                      # the two row formats really differ
    np.testing.assert_allclose(lin["dx"].cpu().numpy(), got["dx"].cpu().numpy(), rtol=1e-3, atol=1e-6)


def test_head_wide_lattice_and_sum_reduction(cuda_device):
    """U up to 255 (8 pairs per lane, the shorter staging ring), reduction 'sum'."""
    B, T, V, U = 4, 800, 300, 255
    x, tok, il, tl = _case(B, T, V, U, 11, torch.bfloat16, ragged=False)
    tl[0] = 255
    ref = _run(x, tok, il, tl, U, 6, overlapped=False, red=2)
    got = _run(x, tok, il, tl, U, 6, overlapped=True, red=2)
    _same(ref, got, "wide")
    # a lattice the fp64 kernel does not take: the head runs the passes one after the other
    U2 = 300
    x2, tok2, il2, tl2 = _case(3, 700, 40, U2, 12, torch.float32, ragged=False)
    _same(_run(x2, tok2, il2, tl2, U2, 4, overlapped=False), _run(x2, tok2, il2, tl2, U2, 4, overlapped=True), "log-domain")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_head_through_the_module_against_torch_fp64(cuda_device, dtype, monkeypatch):
    """ctc_loss with an input that wants a gradient takes the head; loss and gradient against torch's fp64 CTC on the
    same (rounded) logits; a non-unit upstream gradient; a second backward over a retained graph."""
    import statecatcher_b200 as sb
    from statecatcher_b200 import ctc as ctc_mod
    monkeypatch.setattr(ctc_mod, "_OVERLAP", True)                            # what SC_CTC_OVERLAP=1 selects
    B, T, V, U = 5, 1200, 128, 50
    x0, tok, il, tl = _case(B, T, V, U, 21, dtype)
    tl[3] = 2                                                                # feasible again (3 frames, 2 labels)... unless repeated
    tok[3, 1] = tok[3, 0] % (V - 1) + 1
    x = x0.clone().requires_grad_(True)
    calls = []
    from statecatcher_b200 import _lib
    orig = _lib.call
    ctc_mod.call = lambda name, *a: (calls.append(name), orig(name, *a))[1]
    try:
        loss = sb.ctc_loss_from_logits(x, tok, il, tl, zero_infinity=True)
        (3.0 * loss).backward(retain_graph=True)
        g3 = x.grad.clone()
        x.grad = None
        (3.0 * loss).backward()
        g3b = x.grad.clone()
    finally:
        ctc_mod.call = orig
    assert calls == ["sc_ctc_head", "sc_ctc_scale_grad", "sc_ctc_bwd"], calls
    xr = x0.double().requires_grad_(True)
    want = torch.nn.functional.ctc_loss(xr.log_softmax(-1).transpose(0, 1), tok, il.clamp(max=T), tl, blank=0,
                                        reduction="mean", zero_infinity=True)
    (3.0 * want).backward()
    np.testing.assert_allclose(loss.item(), want.item(), rtol=1e-5)
    w = xr.grad.cpu().numpy()
    tol = dict(rtol=1e-4, atol=2e-7) if dtype == torch.float32 else dict(rtol=8e-3, atol=1e-6)   # bf16: the gradient is stored in bf16
    np.testing.assert_allclose(g3.float().cpu().numpy(), w, **tol)
    np.testing.assert_allclose(g3b.float().cpu().numpy(), w, **tol)


def test_head_inside_a_stream_capture(cuda_device, monkeypatch):
    """Both side streams fork from and join the capturing stream: the head replays from a CUDA graph."""
    import statecatcher_b200 as sb
    from statecatcher_b200 import ctc as ctc_mod
    monkeypatch.setattr(ctc_mod, "_OVERLAP", True)
    B, T, V, U = 4, 900, 64, 20
    x0, tok, il, tl = _case(B, T, V, U, 31, torch.float32, ragged=False)
    x = x0.clone().requires_grad_(True)
    eager = sb.ctc_loss_from_logits(x, tok, il, tl, zero_infinity=True)
    eager.backward()
    want_g, want_l = x.grad.clone(), eager.detach().clone()
    xs = x0.clone().requires_grad_(True)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):                                                   # warm-up on the side stream (allocations, attributes)
            xs.grad = None
            sb.ctc_loss_from_logits(xs, tok, il, tl, zero_infinity=True).backward()
    torch.cuda.current_stream().wait_stream(s)
    xs.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        l = sb.ctc_loss_from_logits(xs, tok, il, tl, zero_infinity=True)
        l.backward()
    xs.grad.zero_()
    l.detach().zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(xs.grad, want_g) and torch.equal(l.detach(), want_l)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("V,U", [(256, 60), (1024, 150), (2048, 300), (128, 20)])
def test_gradient_with_repeated_labels(cuda_device, dtype, V, U):
    """Transcripts drawn from 5 symbols, so every label repeats many times and the scatter of the label occupancies
    sums into the same vocabulary entries; one transcript holds the blank index itself; all three lattice widths of
    the gradient kernel; against torch's fp64 CTC on the same (rounded) logits."""
    import statecatcher_b200 as sb
    g = torch.Generator().manual_seed(V + U)
    B, T = 4, 2 * U + 40
    x0 = (torch.randn(B, T, V, generator=g) * 1.5).to(dtype).cuda()
    tok = torch.randint(1, 6, (B, U), generator=g).cuda()
    tok[2, 3] = 0                                                            # the blank as a label (torch tolerates it)
    il = torch.tensor([T, T - 7, T, T // 2 + U], device="cuda")
    tl = torch.tensor([U, U // 2, U - 1, U // 3], device="cuda")
    x = x0.clone().requires_grad_(True)
    loss = sb.ctc_loss_from_logits(x, tok, il, tl, reduction="sum", zero_infinity=True)
    loss.backward()
    xr = x0.double().requires_grad_(True)
    want = torch.nn.functional.ctc_loss(xr.log_softmax(-1).transpose(0, 1), tok, il, tl, blank=0, reduction="sum",
                                        zero_infinity=True)
    want.backward()
    np.testing.assert_allclose(loss.item(), want.item(), rtol=1e-5)
    tol = dict(rtol=1e-4, atol=2e-6) if dtype == torch.float32 else dict(rtol=8e-3, atol=2e-3)
    np.testing.assert_allclose(x.grad.float().cpu().numpy(), xr.grad.cpu().numpy(), **tol)
