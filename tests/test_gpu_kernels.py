"""GPU parity tests, kernel level: every C-ABI kernel family against the CPU oracle
(oracle/, fp64) on the same seeded inputs.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest
import torch

from conftest import assert_grad_close
from oracle import ctc_oracle, lucy_oracle as LO

pytestmark = pytest.mark.gpu

# fp32 contract: rtol 1e-4 (BASELINE.json north_star).  bf16: looser, stated bound 3e-2 of
# the tensor's max magnitude (the reference's own fp32->bf16-autocast drift is ~1e-2 rel-L2).
F32_RTOL, F32_ATOL = 1e-4, 2e-5
BF16_REL = 3e-2
# CTC gradient w.r.t. fp32 logits, elementwise rtol at atol 2e-7 against torch's fp64 CTC.  fp64 linear-domain
# lattice (lattices up to 255 labels): the north star's 1e-4.  Log-domain kernels (larger lattices): measured on a
# B200 (tests/test_gpu_configs1_parity.py, r02: 1.6e-5 needed at T=3000, U<=150), bound kept at 2e-4.
CTC_GRAD_RTOL = 1e-4
CTC_GRAD_RTOL_LOG = 2e-4


def _close(got, want, dtype, what=""):
    got = got.detach().double().cpu().numpy()
    want = np.asarray(want, dtype=np.float64)
    if dtype == torch.float32:
        np.testing.assert_allclose(got, want, rtol=F32_RTOL, atol=F32_ATOL * max(1.0, np.abs(want).max()), err_msg=what)
    else:
        scale = max(np.abs(want).max(), 1e-3)
        assert np.abs(got - want).max() <= BF16_REL * scale, (what, np.abs(got - want).max(), scale)


def _ops():
    from statecatcher_b200 import ops
    return ops


# ------------------------------------------------------------------ K1 projections ---
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (37, 50, 19), (300, 130, 80), (513, 257, 129), (2048, 256, 256)])
def test_gemm_family(cuda_device, dtype, M, N, K):
    ops = _ops()
    g = torch.Generator().manual_seed(M * 7 + N)
    a = torch.randn(M, K, generator=g).to(dtype)
    w = torch.randn(N, K, generator=g).to(dtype)
    dy = torch.randn(M, N, generator=g).to(dtype)
    bias = torch.randn(N, generator=g)
    A, W, DY = a.double(), w.double(), dy.double()
    y = ops.gemm_fwd(a.cuda(), w.cuda(), bias.cuda())
    _close(y, A @ W.T + bias.double(), dtype, "fwd")
    da = ops.gemm_dgrad(dy.cuda(), w.cuda())
    _close(da, DY @ W, dtype, "dgrad")
    dw = ops.gemm_wgrad(dy.cuda(), a.cuda())
    assert dw.dtype == torch.float32
    _close(dw, DY.T @ A, torch.float32 if dtype == torch.float32 else dtype, "wgrad")
    # accumulate flag
    base = torch.randn(N, K, generator=g).cuda()
    dw2 = ops.gemm_wgrad(dy.cuda(), a.cuda(), out=base.clone(), accumulate=True)
    _close(dw2, DY.T @ A + base.double().cpu(), torch.float32 if dtype == torch.float32 else dtype, "wgrad acc")


def test_gemm_strided_views(cuda_device):
    """Column-block views (ld != width) as the layer code passes them."""
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    big = torch.randn(70, 5 * 24, generator=g).cuda()
    w = torch.randn(5 * 24, 24, generator=g).cuda()
    blk = big[:, 24:48]
    y = ops.gemm_dgrad(blk, w[24:48])
    _close(y, blk.double().cpu() @ w[24:48].double().cpu(), torch.float32)
    out = torch.zeros(70, 3 * 16).cuda()
    ops.gemm_fwd(big[:, :24], torch.randn(16, 24, generator=g).cuda(), None, out=out[:, 16:32])
    assert out[:, :16].abs().max() == 0 and out[:, 32:].abs().max() == 0


# ------------------------------------------------------------------ row helpers ------
@pytest.mark.parametrize("H", [48, 256, 768, 1024, 1280])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_layernorm_cast_colsum(cuda_device, dtype, H):
    ops = _ops()
    g = torch.Generator().manual_seed(9)
    M = 77
    x = (torch.randn(M, H, generator=g) * 2 + 0.5).to(dtype)
    w = torch.randn(H, generator=g)
    b = torch.randn(H, generator=g)
    dy = torch.randn(M, H, generator=g).to(dtype)
    xd = x.double().requires_grad_(True)
    wd, bd = w.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xd, (H,), wd, bd, 1e-5)
    ref.backward(dy.double())
    y, mean, rstd = ops.layernorm_fwd(x.cuda(), w.cuda(), b.cuda())
    _close(y, ref.detach(), dtype, "ln fwd")
    dx, dw, db = ops.layernorm_bwd(dy.cuda(), x.cuda(), w.cuda(), mean, rstd)
    _close(dx, xd.grad, dtype, "ln dx")
    _close(dw, wd.grad, dtype, "ln dw")
    _close(db, bd.grad, dtype, "ln db")
    _close(ops.colsum(dy.cuda()), dy.double().sum(0), dtype, "colsum")
    c = ops.cast(x.float().cuda(), torch.bfloat16)
    assert torch.equal(c.cpu(), x.float().to(torch.bfloat16))       # round-to-nearest-even, bit exact


# ------------------------------------------------------------------ K2 fused scan ----
def _scan_reference(G, h0, s0, training, g_out):
    """fp64 closed form + autograd (oracle) for the fused scan stage."""
    B, T, H5 = G.shape
    H = H5 // 5
    Gd = G.double().requires_grad_(True)
    z, k, v, p, q = [Gd[..., i * H:(i + 1) * H] for i in range(5)]
    d = torch.sigmoid(q)
    kv = k * v
    S = LO._linear_scan(d, kv, torch.zeros(B, H, dtype=torch.float64) if training else s0.double())
    sp = d * S + kv if training else S
    c = torch.tanh(p + sp)
    zh = torch.sigmoid(z)
    Hout = LO._linear_scan(zh, (1 - zh) * c, h0.double())
    (Hout * g_out.double()).sum().backward()
    return Hout.detach(), S.detach(), Gd.grad


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("training", [True, False], ids=["train", "step"])
@pytest.mark.parametrize("B,T,H", [(1, 1, 8), (3, 17, 24), (2, 48, 64), (5, 100, 40), (2, 37, 12), (2, 37, 6), (3, 70, 520)])
def test_fused_scan_fwd_bwd(cuda_device, dtype, training, B, T, H):
    ops = _ops()
    g = torch.Generator().manual_seed(B * 1000 + T)
    G = torch.randn(B, T, 5 * H, generator=g).to(dtype)
    h0 = torch.randn(B, H, generator=g) * 0.5
    s0 = torch.randn(B, H, generator=g) * 0.5
    go = torch.randn(B, T, H, generator=g).to(dtype)
    Href, Sref, dGref = _scan_reference(G.float(), h0, s0, training, go.float())
    Gc = G.cuda().view(B * T, 5 * H)
    Hout, hT, sT, ck = ops.scan_fwd(Gc, B, T, H, h0.cuda(), s0.cuda(), training)
    _close(Hout.view(B, T, H), Href, dtype, "Hout")
    _close(hT, Href[:, -1], dtype, "hT")
    if not training:
        _close(sT, Sref[:, -1], dtype, "sT")
    dG, dbias = ops.scan_bwd(Gc, Hout, h0.cuda(), s0.cuda(), ck, go.cuda().view(B * T, H), B, T, H, training)
    _close(dG.view(B, T, 5 * H), dGref, dtype, "dG")
    # bias gradient = column sums of the (unrounded, fp32) gate gradients
    _close(dbias, dGref.reshape(B * T, 5 * H).sum(0), dtype, "dbias = column sums of dG")


def test_step_path_state_handoff_is_bit_exact(cuda_device):
    """Segment-boundary handoff: one 96-step segment == three 32-step segments with carried
    (h, s), bit for bit, in the step path (both states carried in fp32)."""
    ops = _ops()
    g = torch.Generator().manual_seed(4)
    B, T, H = 4, 96, 64
    G = torch.randn(B, T, 5 * H, generator=g).cuda()
    h0 = torch.randn(B, H, generator=g).cuda()
    s0 = torch.randn(B, H, generator=g).cuda()
    full, hT, sT, _ = ops.scan_fwd(G.view(B * T, 5 * H), B, T, H, h0, s0, False)
    h, s, parts = h0, s0, []
    for i in range(3):
        seg = G[:, 32 * i:32 * (i + 1)].contiguous().view(B * 32, 5 * H)
        o, h, s, _ = ops.scan_fwd(seg, B, 32, H, h, s, False)
        parts.append(o.view(B, 32, H))
    assert torch.equal(torch.cat(parts, 1), full.view(B, T, H))
    assert torch.equal(h, hT) and torch.equal(s, sT)


def test_fused_scan_large_properties(cuda_device):
    """cfg2-sized layer (B=64,T=3000,H=1024, bf16): no oracle at this size, so check
    size-independent facts: the training path equals the step path started from s0=0 with
    the second application removed... (not equal) -> instead: (a) determinism, (b) h bounded by
    construction (convex combination of tanh values and h0), (c) split-segment equality."""
    ops = _ops()
    B, T, H = 64, 3000, 1024
    g = torch.Generator(device="cuda").manual_seed(1)
    G = torch.randn(B * T, 5 * H, generator=g, device="cuda", dtype=torch.bfloat16)
    h0 = torch.zeros(B, H, device="cuda")
    s0 = torch.zeros(B, H, device="cuda")
    o1, hT1, _, ck = ops.scan_fwd(G, B, T, H, h0, s0, True)
    o2, hT2, _, _ = ops.scan_fwd(G, B, T, H, h0, s0, True)
    assert torch.equal(o1, o2) and torch.equal(hT1, hT2)
    assert o1.float().abs().max() <= 1.0 + 1e-2
    # step path: 3000 = 1504 + 1496 split must be bit exact
    G3 = G.view(B, T, 5 * H)
    f, hf, sf, _ = ops.scan_fwd(G, B, T, H, h0, s0, False)
    a, ha, sa, _ = ops.scan_fwd(G3[:, :1504].contiguous().view(-1, 5 * H), B, 1504, H, h0, s0, False)
    b, hb, sb, _ = ops.scan_fwd(G3[:, 1504:].contiguous().view(-1, 5 * H), B, 1496, H, ha, sa, False)
    assert torch.equal(hb, hf) and torch.equal(sb, sf)
    assert torch.equal(b.view(B, 1496, H), f.view(B, T, H)[:, 1504:])
    # backward runs at full size and is finite
    go = torch.randn(B * T, H, generator=g, device="cuda", dtype=torch.bfloat16)
    dG, db = ops.scan_bwd(G, o1, h0, s0, ck, go, B, T, H, True)
    assert torch.isfinite(dG.float()).all() and torch.isfinite(db).all()


@pytest.mark.parametrize("rows,cols,pad", [(5120, 1032, 0), (77, 88, 8), (33, 30, 2), (5, 7, 0)])
def test_split_bf16_hi_lo(cuda_device, rows, cols, pad):
    """fp32 -> [hi | lo] bf16 expansion (what carries fp32 weight-space products through the bf16 tensor cores): bit-exact
    against torch's round-to-nearest-even, the four-column kernel and the scalar one (odd widths / strides)."""
    ops = _ops()
    g = torch.Generator().manual_seed(rows + cols)
    big = torch.randn(rows, cols + pad, generator=g) * 3
    src = big.cuda()[:, :cols] if pad else big.cuda()
    out = ops.split_bf16(src)
    hi = big[:, :cols].to(torch.bfloat16)
    lo = (big[:, :cols] - hi.float()).to(torch.bfloat16)
    assert out.shape == (rows, 2 * cols)
    assert torch.equal(out[:, :cols].cpu().view(torch.int16), hi.view(torch.int16))
    assert torch.equal(out[:, cols:].cpu().view(torch.int16), lo.view(torch.int16))


@pytest.mark.parametrize("N", [1024, 1032, 264])
def test_colsum_bf16_many_rows(cuda_device, N):
    """Many rows (M = 10 001: ragged against the 4-row unroll and the row chunks of the column-sum kernel), widths that
    are not a multiple of its 128-column block, a column block of a wider tensor, and accumulation onto an existing
    vector."""
    ops = _ops()
    g = torch.Generator().manual_seed(23)
    M = 10001
    big = torch.randn(M, N + 16, generator=g).bfloat16()
    x = big.cuda()[:, 8:8 + N]                                       # row stride N + 16, 16-byte aligned start
    want = big[:, 8:8 + N].double().sum(0)
    got = ops.colsum(x)
    assert float((got.cpu().double() - want).abs().max()) <= 1e-4 * float(want.abs().max()) + 1e-3
    base = torch.randn(N, generator=g)
    out = base.cuda().clone()
    ops.colsum(x, out=out, accumulate=True)
    assert float((out.cpu().double() - (want + base.double())).abs().max()) <= 1e-4 * float(want.abs().max()) + 1e-3


@pytest.mark.parametrize("H", [256, 1024])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_layernorm_bwd_many_rows_and_dx_colsum(cuda_device, dtype, H):
    """Enough rows that every warp of the staged backward kernel walks its shared-memory ring several times
    (M = 14 321 over 592 blocks of 4 warps: 6-7 rows per warp), strided operands as the module passes them (blocks of a
    wider tensor), and the column sums of dx (`dxsum`, the bias gradient of the producing projection)."""
    ops = _ops()
    g = torch.Generator().manual_seed(19)
    M = 14321
    big_x = (torch.randn(M, 3 * H, generator=g) * 1.5 - 0.3).to(dtype)
    big_dy = torch.randn(M, 2 * H, generator=g).to(dtype)
    x, dy = big_x[:, H:2 * H], big_dy[:, H:]                       # row stride 3H / 2H
    w = torch.randn(H, generator=g)
    b = torch.randn(H, generator=g)
    xd = x.double().requires_grad_(True)
    wd, bd = w.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xd, (H,), wd, bd, 1e-5)
    ref.backward(dy.double())
    xg, dyg = big_x.cuda()[:, H:2 * H], big_dy.cuda()[:, H:]
    y, mean, rstd = ops.layernorm_fwd(xg, w.cuda(), b.cuda())
    _close(y, ref.detach(), dtype, "ln fwd")
    big_dx = torch.zeros(M, 2 * H, dtype=dtype, device="cuda")
    dxsum = torch.zeros(H, device="cuda")
    dx, dw, db = ops.layernorm_bwd(dyg, xg, w.cuda(), mean, rstd, dx=big_dx[:, :H], dxsum=dxsum)
    _close(dx, xd.grad, dtype, "ln dx")
    assert not big_dx[:, H:].any()                                  # nothing written beside the block
    for got, want, what in ((dw, wd.grad, "dw"), (db, bd.grad, "db"), (dxsum, xd.grad.sum(0), "dxsum")):
        want = want.float()
        tol = (2e-4 if dtype == torch.float32 else 2e-2) * max(1.0, float(want.abs().max()))
        assert float((got.cpu() - want).abs().max()) <= tol, (what, float((got.cpu() - want).abs().max()), tol)


# ------------------------------------------------------------------ K2' split scans --
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("mode", ["train", "step", "prefix"])
@pytest.mark.parametrize("H", [24, 22], ids=["tma", "generic"])        # 22 channels: rows TMA cannot address -> the simple kernels
def test_split_scans(cuda_device, dtype, mode, H):
    ops = _ops()
    g = torch.Generator().manual_seed(21)
    B, T = 3, 19
    training = mode != "step"
    dmode = 1 if mode == "prefix" else 0
    lam = 0.05
    k, v, q, add = [torch.randn(B, T, H, generator=g).to(dtype) for _ in range(4)]
    s0 = torch.randn(B, H, generator=g)
    h0 = torch.randn(B, H, generator=g)
    dA = torch.randn(B, T, H, generator=g).to(dtype)
    kd, vd, qd, ad = [t.double().requires_grad_(True) for t in (k, v, q, add)]
    d = torch.sigmoid(qd)
    kv = kd * vd
    if dmode:
        S = LO._prefix_sum_scan(kv, lam)
    else:
        S = LO._linear_scan(d, kv, torch.zeros(B, H, dtype=torch.float64) if training else s0.double())
    A = ad + (d * S + kv if training else S)
    (A * dA.double()).sum().backward()
    c = lambda t: t.cuda().view(B * T, H)                      # noqa: E731
    Ag, S_all, sT = ops.sscan_fwd(c(k), c(v), c(q), c(add), s0.cuda(), B, T, H, training, dmode, lam)
    _close(Ag.view(B, T, H), A.detach(), dtype, "A")
    if dmode:                                                  # prefix_sum: every S_t is saved
        _close(S_all.view(B, T, H), S.detach(), dtype, "S_all")
    else:                                                      # learned decay: S entering every 8-step interval
        S_in = torch.cat([(torch.zeros(B, 1, H, dtype=torch.float64) if training else s0.double()[:, None]), S.detach()[:, :-1]], 1)
        _close(S_all.view(B, -1, H), S_in[:, ::8], dtype, "S checkpoints")
    dk, dv, dq = [torch.empty(B * T, H, dtype=dtype, device="cuda") for _ in range(3)]
    ops.sscan_bwd(c(k), c(v), c(q), S_all, s0.cuda(), c(dA), dk, dv, dq, B, T, H, training, dmode, lam)
    _close(dk.view(B, T, H), kd.grad, dtype, "dk")
    _close(dv.view(B, T, H), vd.grad, dtype, "dv")
    _close(dq.view(B, T, H), qd.grad, dtype, "dq")
    # h scan
    An, Zn, go = [torch.randn(B, T, H, generator=g).to(dtype) for _ in range(3)]
    and_, znd = An.double().requires_grad_(True), Zn.double().requires_grad_(True)
    zh = torch.sigmoid(znd)
    Hout = LO._linear_scan(zh, (1 - zh) * torch.tanh(and_), h0.double())
    (Hout * go.double()).sum().backward()
    Hg, hT = ops.hscan_fwd(c(An), c(Zn), h0.cuda(), B, T, H)
    _close(Hg.view(B, T, H), Hout.detach(), dtype, "Hout")
    _close(hT, Hout[:, -1].detach(), dtype, "hT")
    dAn, dZn = [torch.empty(B * T, H, dtype=dtype, device="cuda") for _ in range(2)]
    ops.hscan_bwd(c(An), c(Zn), Hg, h0.cuda(), c(go), dAn, dZn, B, T, H)
    _close(dAn.view(B, T, H), and_.grad, dtype, "dAn")
    _close(dZn.view(B, T, H), znd.grad, dtype, "dZn")


# ------------------------------------------------------------------ K3 CTC -----------
CTC_CASES = ["basic", "repeats_tight", "infeasible", "empty_target", "zero_frames", "all_empty", "long"]


@pytest.mark.parametrize("layout", ["btv", "tbv"])
@pytest.mark.parametrize("case", CTC_CASES)
def test_ctc_golden(cuda_device, case, layout):
    from conftest import load_golden
    from statecatcher_b200 import ctc_loss
    G = load_golden("ctc_cases")
    logits = torch.tensor(G[case + "/logits"]).cuda()
    if layout == "tbv":
        x = logits.transpose(0, 1).contiguous().requires_grad_(True)
        inp = x
    else:
        x = logits.clone().requires_grad_(True)
        inp = x.transpose(0, 1)
    tokens = torch.tensor(G[case + "/tokens"]).cuda()
    in_lens, tgt_lens = G[case + "/in_lens"].tolist(), G[case + "/tgt_lens"].tolist()
    loss = ctc_loss(inp, tokens, in_lens, tgt_lens, blank=0, reduction="mean", zero_infinity=True)
    loss.backward()
    np.testing.assert_allclose(loss.item(), G[case + "/loss"], rtol=1e-4, atol=1e-6)
    grad = x.grad if layout == "btv" else x.grad.transpose(0, 1)
    np.testing.assert_allclose(grad.cpu().numpy(), G[case + "/grad"], rtol=1e-4, atol=2e-6)
    # bit-exact structural facts: zero beyond T_b, zero rows for infeasible utterances
    nll = ctc_loss(inp.detach(), tokens, in_lens, tgt_lens, reduction="none", zero_infinity=True)
    np.testing.assert_allclose(nll.cpu().numpy(), G[case + "/nll"], rtol=1e-4, atol=1e-5)
    for b, Tb in enumerate(in_lens):
        assert (grad[b, Tb:] == 0).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_ctc_random_vs_oracle(cuda_device, dtype):
    from statecatcher_b200 import CTCLoss
    g = torch.Generator().manual_seed(5)
    B, T, V, U = 6, 40, 29, 9
    logits = (torch.randn(B, T, V, generator=g) * 3).to(dtype)
    tgt_lens = [9, 0, 5, 9, 1, 7]
    in_lens = [40, 40, 33, 17, 40, 2]                     # last one infeasible (2 < 7)
    tokens = torch.randint(1, V, (B, U), generator=g)
    loss_ref, nll_ref, grad_ref = ctc_oracle.ctc_loss_and_grad(logits.float().numpy(), tokens.numpy(), in_lens, tgt_lens)
    x = logits.cuda().requires_grad_(True)
    crit = CTCLoss(blank=0, zero_infinity=True)
    loss = crit(x.transpose(0, 1), tokens.cuda(), in_lens, tgt_lens)
    (loss * 3.0).backward()                               # upstream scale must flow through
    np.testing.assert_allclose(loss.item(), loss_ref, rtol=2e-4)
    if dtype == torch.float32:
        np.testing.assert_allclose(x.grad.cpu().numpy(), 3.0 * grad_ref, rtol=2e-4, atol=1e-6)
    else:
        assert np.abs(x.grad.float().cpu().numpy() - 3.0 * grad_ref).max() <= BF16_REL * np.abs(3.0 * grad_ref).max()
    assert (x.grad[5] == 0).all() and (x.grad[2, 33:] == 0).all()
    # tensor lengths, int32 targets, sum reduction
    x2 = logits.cuda().requires_grad_(True)
    from statecatcher_b200 import ctc_loss
    l2 = ctc_loss(x2.transpose(0, 1), tokens.int().cuda(), torch.tensor(in_lens).cuda(), torch.tensor(tgt_lens),
                  reduction="sum", zero_infinity=True)
    np.testing.assert_allclose(l2.item(), np.where(np.isfinite(nll_ref), nll_ref, 0).sum(), rtol=2e-4)


def test_ctc_long_labels_general_path(cuda_device):
    """2U+1 > 1024 lattice nodes (several nodes per thread) and T not a multiple of the
    emission block: general recursion path vs torch CPU ctc_loss (library pin, fp64)."""
    from statecatcher_b200 import ctc_loss
    g = torch.Generator().manual_seed(12)
    B, T, V, U = 2, 1300, 11, 600
    logits = torch.randn(B, T, V, generator=g)
    tokens = torch.randint(1, V, (B, U), generator=g)
    inl, tgl = [T, 1237], [600, 520]
    xd = logits.double().requires_grad_(True)
    ref = torch.nn.functional.ctc_loss(xd.log_softmax(-1).transpose(0, 1), tokens, inl, tgl, zero_infinity=True)
    ref.backward()
    x = logits.cuda().requires_grad_(True)
    loss = ctc_loss(x.transpose(0, 1), tokens.cuda(), inl, tgl, zero_infinity=True)
    loss.backward()
    np.testing.assert_allclose(loss.item(), ref.item(), rtol=1e-4)
    assert_grad_close(x.grad.cpu().numpy(), xd.grad.numpy(), CTC_GRAD_RTOL_LOG, 2e-7, "ctc_long_labels_U600_T1300")


@pytest.mark.parametrize("T", [1, 2, 15, 16, 17, 31, 33, 100])
def test_ctc_emission_block_boundaries(cuda_device, T):
    """Sequence lengths around the 16-row emission block of the alpha/beta kernel."""
    from statecatcher_b200 import ctc_loss
    g = torch.Generator().manual_seed(T)
    B, V, U = 3, 9, 4
    logits = torch.randn(B, T, V, generator=g) * 2
    tokens = torch.randint(1, V, (B, U), generator=g)
    inl = [T, max(1, T - 1), max(1, T // 2)]
    tgl = [min(U, T), min(2, T), 1]
    _, _, grad_ref = ctc_oracle.ctc_loss_and_grad(logits.numpy(), tokens.numpy(), inl, tgl)
    loss_ref, _, _ = ctc_oracle.ctc_loss_and_grad(logits.numpy(), tokens.numpy(), inl, tgl)
    x = logits.cuda().requires_grad_(True)
    loss = ctc_loss(x.transpose(0, 1), tokens.cuda(), inl, tgl, zero_infinity=True)
    loss.backward()
    np.testing.assert_allclose(loss.item(), loss_ref, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(x.grad.cpu().numpy(), grad_ref, rtol=2e-4, atol=2e-6)


def test_ctc_large_properties(cuda_device):
    """cfg2-sized CTC (B=64,T=3000,V=1024,U<=150): rows of dlogits sum to ~0, exact zeros
    beyond T_b, loss finite and reproducible."""
    from statecatcher_b200 import ctc_loss_from_logits
    B, T, V, U = 64, 3000, 1024, 150
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(B, T, V, generator=g, device="cuda").requires_grad_(True)
    tokens = torch.randint(1, V, (B, U), generator=g, device="cuda")
    tgt = [75 + (7 * b) % 76 for b in range(B)]
    inl = [T] * B
    inl[3] = 1700
    inl[5], tgt[5] = 0, 0
    l1 = ctc_loss_from_logits(x, tokens, inl, tgt)
    l1.backward()
    l2 = ctc_loss_from_logits(x.detach(), tokens, inl, tgt)
    assert torch.isfinite(l1) and l1.item() == l2.item()
    assert (x.grad[3, 1700:] == 0).all() and (x.grad[5] == 0).all()
    assert x.grad.sum(-1).abs().max().item() < 1e-5


def test_ctc_large_vocab(cuda_device):
    """V=20000 (shared-memory row per warp no longer fits 8 warps per block)."""
    from statecatcher_b200 import ctc_loss
    g = torch.Generator().manual_seed(8)
    B, T, V, U = 2, 30, 20000, 5
    logits = torch.randn(B, T, V, generator=g)
    tokens = torch.randint(1, V, (B, U), generator=g)
    inl, tgl = [T, 22], [5, 3]
    xd = logits.double().requires_grad_(True)
    ref = torch.nn.functional.ctc_loss(xd.log_softmax(-1).transpose(0, 1), tokens, inl, tgl, zero_infinity=True)
    ref.backward()
    x = logits.cuda().requires_grad_(True)
    loss = ctc_loss(x.transpose(0, 1), tokens.cuda(), inl, tgl, zero_infinity=True)
    loss.backward()
    np.testing.assert_allclose(loss.item(), ref.item(), rtol=1e-4)
    assert_grad_close(x.grad.cpu().numpy(), xd.grad.numpy(), CTC_GRAD_RTOL, 1e-7, "ctc_large_vocab_V20000")


@pytest.mark.parametrize("U,T", [(3, 40), (63, 150), (64, 200), (127, 300), (128, 330), (159, 400), (191, 470),
                                 (192, 480), (255, 600), (256, 640), (320, 800)])
def test_ctc_lattice_width_variants(cuda_device, U, T):
    """1, 2, 3 and 4 warps per lattice (four nodes per thread), the boundaries between them, and
    the thread-strided kernel beyond 512 nodes; small vocabulary so that repeated labels (no
    skip transition) are common."""
    from statecatcher_b200 import ctc_loss
    g = torch.Generator().manual_seed(U)
    B, V = 3, 6
    logits = torch.randn(B, T, V, generator=g) * 1.5
    tokens = torch.randint(1, V, (B, U), generator=g)
    inl, tgl = [T, T - 7, T - 16], [U, max(U - 5, 0), max(U // 2, 1)]
    xd = logits.double().requires_grad_(True)
    ref = torch.nn.functional.ctc_loss(xd.log_softmax(-1).transpose(0, 1), tokens, inl, tgl, zero_infinity=True)
    ref.backward()
    x = logits.cuda().requires_grad_(True)
    loss = ctc_loss(x.transpose(0, 1), tokens.cuda(), inl, tgl, zero_infinity=True)
    loss.backward()
    np.testing.assert_allclose(loss.item(), ref.item(), rtol=1e-4)
    assert_grad_close(x.grad.cpu().numpy(), xd.grad.numpy(), CTC_GRAD_RTOL if U <= 255 else CTC_GRAD_RTOL_LOG, 2e-7,
                      f"ctc_lattice_width_U{U}_T{T}")


def test_ctc_mismatched_transcript(cuda_device):
    """A confident model (40-nat margins) whose frames follow a label sequence unrelated to the
    transcript: forward and backward masses barely overlap (the case a scaled linear-domain
    recursion cannot represent).  Loss and gradient must still match torch's fp64 CTC."""
    from statecatcher_b200 import ctc_loss
    g = torch.Generator().manual_seed(21)
    B, T, V, U = 4, 300, 40, 30
    tokens = torch.randint(1, V, (B, U), generator=g)
    logits = torch.randn(B, T, V, generator=g)
    for b in (1, 3):
        wrong = torch.randint(1, V, (T,), generator=g)
        logits[b] = -20.0
        logits[b, torch.arange(T), wrong] = 20.0
    inl, tgl = [T, T, T - 3, T - 40], [U, U, U - 4, U]
    xd = logits.double().requires_grad_(True)
    ref = torch.nn.functional.ctc_loss(xd.log_softmax(-1).transpose(0, 1), tokens, inl, tgl,
                                       reduction="sum", zero_infinity=True)
    ref.backward()
    x = logits.cuda().requires_grad_(True)
    loss = ctc_loss(x.transpose(0, 1), tokens.cuda(), inl, tgl, reduction="sum", zero_infinity=True)
    loss.backward()
    np.testing.assert_allclose(loss.item(), ref.item(), rtol=1e-4)
    # 40-nat margins against the transcript: the fp64 range checks send utterances 1 and 3 to the log-domain
    # recomputation, whose fp32 log2 values (magnitudes ~1e4 here) limit the gradient to 1.5e-4 (measured r02)
    assert_grad_close(x.grad.cpu().numpy(), xd.grad.numpy(), 3e-4, 2e-6, "ctc_mismatched_transcript")


@pytest.mark.parametrize("U", [6, 300], ids=["fp64-linear", "log-domain"])
def test_ctc_combined_entry_matches_split(cuda_device, U):
    """sc_ctc_fwd (one call) == sc_ctc_emissions + sc_ctc_lattice (what ctc.py binds), bit for bit, for both
    lattice representations (Umax <= 255: fp64 linear domain; larger: log domain)."""
    from statecatcher_b200._lib import call, dt, ptr, stream, load
    g = torch.Generator().manual_seed(31)
    B, T, V = 3, 50 if U < 100 else 700, 17
    x = torch.randn(B, T, V, generator=g).cuda()
    tok = torch.randint(1, V, (B, U), generator=g).cuda()
    il = torch.tensor([T, T - 5, 30], device="cuda")
    tl = torch.tensor([U, 4, 0], device="cuda")
    S = (2 * U + 1 + 7) & ~7
    outs = []
    for split in (False, True):
        f32 = dict(dtype=torch.float32, device="cuda")
        lse, lplat, csh = torch.zeros(B, T, **f32), torch.zeros(B, T, load().sc_ctc_lplat_pitch(U), **f32), torch.zeros(B, T, **f32)
        alpha, beta = torch.zeros(B, T, S, **f32), torch.zeros(B, T, S, **f32)
        nll, loss = torch.zeros(B, **f32), torch.zeros((), **f32)
        ws = torch.zeros(load().sc_ctc_workspace_bytes(B, T, U) // 8 + 1, dtype=torch.float64, device="cuda")
        if split:
            call("sc_ctc_emissions", ptr(x), x.stride(0), x.stride(1), dt(x), ptr(tok), tok.stride(0), ptr(il), ptr(tl),
                 B, T, V, U, 0, ptr(lse), ptr(lplat), ptr(csh), stream())
            call("sc_ctc_lattice", ptr(lplat), ptr(csh), ptr(tok), tok.stride(0), ptr(il), ptr(tl), B, T, U, 0,
                 ptr(alpha), ptr(beta), ptr(nll), ptr(loss), 1, ptr(ws), stream())
        else:
            call("sc_ctc_fwd", ptr(x), x.stride(0), x.stride(1), dt(x), ptr(tok), tok.stride(0), ptr(il), ptr(tl),
                 B, T, V, U, 0, ptr(lse), ptr(lplat), ptr(csh), ptr(alpha), ptr(beta), ptr(nll), ptr(loss), 1, ptr(ws), stream())
        outs.append((lse, lplat.view(torch.int32), csh, alpha.view(torch.int32), beta.view(torch.int32), nll, loss))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("train_mode", [False, True], ids=["step", "train"])
@pytest.mark.parametrize("B,T,H", [(1, 256, 64), (2, 300, 200), (1, 1000, 256), (1, 3000, 1024)])
def test_scan_fwd_chunked_matches_sequential(cuda_device, dtype, train_mode, B, T, H):
    """The time-parallel chunked forward scan (few live streams) against the sequential kernel on
    the same gates: outputs, carried states and the backward's checkpoints.  The two differ
    only by re-association across 64-step chunk boundaries."""
    from statecatcher_b200._lib import call, dt, ptr, stream, load
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + T + H)
    G = (torch.randn(B * T, 5 * H, generator=g, device="cuda") * 1.5).to(dtype)
    h0 = torch.randn(B, H, generator=g, device="cuda")
    s0 = torch.randn(B, H, generator=g, device="cuda")
    nck = (T + 7) // 8
    outs = []
    for chunked in (False, True):
        Hout = torch.empty(B * T, H, dtype=dtype, device="cuda")
        hT = torch.empty(B, H, device="cuda")
        sT = torch.zeros(B, H, device="cuda")
        ck = torch.empty(B, nck, H, device="cuda")
        if chunked:
            wb = 4 * B * ((T + 63) // 64) * H * 4
            work = torch.empty(wb // 4, device="cuda")
            call("sc_lucy_scan_fwd_chunked", ptr(G), 5 * H, ptr(h0), ptr(s0), ptr(Hout), H, ptr(hT), ptr(sT), ptr(ck),
                 ptr(work), B, T, H, dt(G), int(train_mode), stream())
        else:
            call("sc_lucy_scan_fwd", ptr(G), 5 * H, ptr(h0), ptr(s0), ptr(Hout), H, ptr(hT), ptr(sT), ptr(ck),
                 B, T, H, dt(G), int(train_mode), stream())
        outs.append((Hout.float(), hT, sT, ck))
    tol = dict(rtol=2e-5, atol=2e-6) if dtype == torch.float32 else dict(rtol=1.6e-2, atol=1e-3)
    torch.testing.assert_close(outs[1][0], outs[0][0], **tol)                 # Hout (bf16: one rounding step apart at most)
    st = dict(rtol=2e-5, atol=2e-5) if dtype == torch.float32 else dict(rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(outs[1][1], outs[0][1], **st)                  # hT
    if not train_mode:
        torch.testing.assert_close(outs[1][2], outs[0][2], **st)              # sT
    torch.testing.assert_close(outs[1][3], outs[0][3], **st)                  # checkpoints of S
    lib = load()
    assert lib.sc_lucy_scan_chunked_work_bytes(1, 3000, 1024) == 4 * 47 * 1024 * 4
    assert lib.sc_lucy_scan_chunked_work_bytes(64, 3000, 1024) == 0           # enough streams: sequential kernel
    assert lib.sc_lucy_scan_chunked_work_bytes(1, 100, 1024) == 0             # too short to cut


@pytest.mark.parametrize("T", [63, 64, 65, 127, 128, 129, 193])
def test_ctc_wavefront_block_meetings(cuda_device, T):
    """Sequence lengths around the 64-row emission blocks at which the wavefront alpha/beta kernel
    holds its block meeting (re-centring, slot recycling), with a lattice that spans three warps
    in the node-per-thread kernel and two in the pair-per-thread kernel."""
    from statecatcher_b200 import ctc_loss
    g = torch.Generator().manual_seed(1000 + T)
    B, V, U = 3, 12, 40
    logits = torch.randn(B, T, V, generator=g) * 2
    tokens = torch.randint(1, V, (B, U), generator=g)
    inl = [T, T - 1, max(1, T - 20)]
    tgl = [min(U, T // 2), 17, 1]
    loss_ref, _, grad_ref = ctc_oracle.ctc_loss_and_grad(logits.numpy(), tokens.numpy(), inl, tgl)
    x = logits.cuda().requires_grad_(True)
    loss = ctc_loss(x.transpose(0, 1), tokens.cuda(), inl, tgl, zero_infinity=True)
    loss.backward()
    np.testing.assert_allclose(loss.item(), loss_ref, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(x.grad.cpu().numpy(), grad_ref, rtol=2e-4, atol=2e-6)


def test_ctc_lattice_kernels_agree(cuda_device, monkeypatch):
    """The four ways a 301-node lattice can be walked give the same loss and gradient on a cfg2-like lattice:
    fp64 linear domain (default), the same with every utterance forced down the log-domain recomputation path
    (SC_CTC_FORCE_LOSSY=1: emission words converted back to log2), and the two log-domain kernels on their own
    emission format (SC_CTC_LIN=0: pair-per-thread wavefront; + SC_CTC_WAVE=0: block barrier per step) — and the
    emission-block size of the wavefront kernel does not matter.  (Subprocess-free: the switches are read per call,
    except SC_CTC_LIN which is latched — that variant runs in a child process.)"""
    from statecatcher_b200 import ctc_loss_from_logits
    g = torch.Generator().manual_seed(77)
    B, T, V, U = 5, 700, 64, 150
    logits = torch.randn(B, T, V, generator=g) * 2
    tokens = torch.randint(1, V, (B, U), generator=g)
    inl, tgl = [T, T, 650, 333, 1], [150, 75, 149, 0, 1]
    xd = logits.double().requires_grad_(True)
    ref = torch.nn.functional.ctc_loss(xd.log_softmax(-1).transpose(0, 1), tokens, inl, tgl, zero_infinity=True)
    ref.backward()
    outs = []
    for force, eb in [(None, None), ("1", None), ("1", "8")]:
        for k, v in (("SC_CTC_FORCE_LOSSY", force), ("SC_CTC_EB", eb)):
            if v is None:
                monkeypatch.delenv(k, raising=False)
            else:
                monkeypatch.setenv(k, v)
        x = logits.cuda().requires_grad_(True)
        loss = ctc_loss_from_logits(x, tokens.cuda(), inl, tgl, zero_infinity=True)
        loss.backward()
        np.testing.assert_allclose(loss.item(), ref.item(), rtol=1e-5)
        np.testing.assert_allclose(x.grad.cpu().numpy(), xd.grad.numpy(), rtol=CTC_GRAD_RTOL, atol=2e-7)
        outs.append(x.grad)
    for o in outs[1:]:
        assert (o - outs[0]).abs().max().item() < 2e-6


@pytest.mark.parametrize("wave", ["2", "0"], ids=["wavefront", "block-barrier"])
def test_ctc_log_domain_kernels_in_child_process(cuda_device, wave, tmp_path):
    """SC_CTC_LIN=0 (latched at first use, hence a child process): the log-domain kernels serve the same cfg2-like
    lattice through their own emission format; compared with torch's fp64 CTC inside the child."""
    import os
    import subprocess
    import sys
    code = """
import numpy as np, torch
from statecatcher_b200 import ctc_loss_from_logits
g = torch.Generator().manual_seed(77)
B, T, V, U = 5, 700, 64, 150
logits = torch.randn(B, T, V, generator=g) * 2
tokens = torch.randint(1, V, (B, U), generator=g)
inl, tgl = [T, T, 650, 333, 1], [150, 75, 149, 0, 1]
xd = logits.double().requires_grad_(True)
ref = torch.nn.functional.ctc_loss(xd.log_softmax(-1).transpose(0, 1), tokens, inl, tgl, zero_infinity=True)
ref.backward()
x = logits.cuda().requires_grad_(True)
loss = ctc_loss_from_logits(x, tokens.cuda(), inl, tgl, zero_infinity=True)
loss.backward()
np.testing.assert_allclose(loss.item(), ref.item(), rtol=1e-4)
np.testing.assert_allclose(x.grad.cpu().numpy(), xd.grad.numpy(), rtol=%g, atol=2e-7)
print("ok")
""" % CTC_GRAD_RTOL_LOG
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SC_CTC_LIN="0", SC_CTC_WAVE=wave, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_ctc_more_lattices_than_sms(cuda_device):
    """100 utterances = 200 (utterance, direction) CTAs on 148 SMs: the wavefront kernel sizes
    its emission blocks so that two CTAs share an SM."""
    from statecatcher_b200 import ctc_loss
    g = torch.Generator().manual_seed(9)
    B, T, V, U = 100, 150, 20, 30
    logits = torch.randn(B, T, V, generator=g)
    tokens = torch.randint(1, V, (B, U), generator=g)
    inl = [T - (b % 7) for b in range(B)]
    tgl = [U - (b % 11) for b in range(B)]
    loss_ref, _, grad_ref = ctc_oracle.ctc_loss_and_grad(logits.numpy(), tokens.numpy(), inl, tgl)
    x = logits.cuda().requires_grad_(True)
    loss = ctc_loss(x.transpose(0, 1), tokens.cuda(), inl, tgl, zero_infinity=True)
    loss.backward()
    np.testing.assert_allclose(loss.item(), loss_ref, rtol=1e-4)
    np.testing.assert_allclose(x.grad.cpu().numpy(), grad_ref, rtol=2e-4, atol=2e-6)


def test_ctc_masked_vocabulary_entries(cuda_device):
    """-inf logits: harmless off the lattice; on a needed label they make the utterance
    infeasible (loss 0, gradient row 0 with zero_infinity) without poisoning its batch mates."""
    from statecatcher_b200 import ctc_loss
    g = torch.Generator().manual_seed(10)
    B, T, V, U = 3, 90, 10, 5
    logits = torch.randn(B, T, V, generator=g)
    tokens = torch.randint(1, 8, (B, U), generator=g)
    logits[:, :, 9] = float("-inf")                         # never a target
    logits[1, :, int(tokens[1, 2])] = float("-inf")         # utterance 1 cannot emit one of its labels
    inl, tgl = [T, T, T - 3], [U, U, U]
    # torch's own gradient is NaN here (log_softmax backward of a -inf input); the loss is not, and the
    # numpy oracle gives the finite gradient (0 at the masked entries)
    ref = torch.nn.functional.ctc_loss(logits.double().log_softmax(-1).transpose(0, 1), tokens, inl, tgl, zero_infinity=True)
    with np.errstate(all="ignore"):
        loss_ref, nll_ref, want = ctc_oracle.ctc_loss_and_grad(logits.numpy(), tokens.numpy(), inl, tgl)
    assert nll_ref[1] == 0 and abs(loss_ref - ref.item()) < 1e-9
    x = logits.cuda().requires_grad_(True)
    loss = ctc_loss(x.transpose(0, 1), tokens.cuda(), inl, tgl, zero_infinity=True)
    loss.backward()
    np.testing.assert_allclose(loss.item(), ref.item(), rtol=1e-4)
    got = x.grad.cpu().numpy()
    assert np.isfinite(got).all() and (got[1] == 0).all() and (got[:, :, 9] == 0).all()
    assert_grad_close(got[[0, 2]], want[[0, 2]], CTC_GRAD_RTOL, 2e-7, "ctc_masked_vocabulary")


@pytest.mark.parametrize("rows,cols", [(7, 80), (33, 1024), (5, 30), (1, 4), (129, 5124)])
def test_cast_vectorised_and_scalar_paths_bit_exact(cuda_device, rows, cols):
    """sc_cast both ways == torch's conversion bit for bit, on the 4-wide path (cols % 4 == 0,
    aligned) and the scalar one, contiguous and with a row stride."""
    from statecatcher_b200 import ops
    g = torch.Generator().manual_seed(rows * 131 + cols)
    x = (torch.randn(rows, cols, generator=g) * 3).cuda()
    got = ops.cast(x, torch.bfloat16)
    assert torch.equal(got.view(torch.int16), x.to(torch.bfloat16).view(torch.int16))
    back = ops.cast(got, torch.float32)
    assert torch.equal(back, got.float())
    wide = (torch.randn(rows, cols + 8, generator=g)).cuda()
    view = wide[:, 4:4 + cols]                               # row stride cols+8, 16-byte aligned start
    assert torch.equal(ops.cast(view, torch.bfloat16).view(torch.int16), view.to(torch.bfloat16).view(torch.int16))
    odd = wide[:, 1:1 + cols]                                # misaligned start: scalar path
    assert torch.equal(ops.cast(odd, torch.bfloat16).view(torch.int16), odd.to(torch.bfloat16).view(torch.int16))
