"""GPU: fused Lion step (sc_lion_step, optim.Lion) vs oracle/optim_oracle.py over several steps.
lion_pytorch (train.py:125-131) is absent upstream: parity unpinned, the published rule is the oracle.
The sign of an interpolation that fp32 cannot resolve from 0 is excluded element-wise."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("wd", [0.0, 0.1])
@pytest.mark.parametrize("max_norm", [None, 0.5, 1e9])
def test_lion_matches_oracle(cuda_device, wd, max_norm):
    from statecatcher_b200.optim import Lion
    from oracle.optim_oracle import lion_step, clip_coef
    rng = np.random.default_rng(7)
    shapes = [(33, 17), (5,), (257, 64), (1,)]
    lr, betas = 3e-3, (0.9, 0.99)
    ref_p = [rng.standard_normal(s).astype(np.float32).astype(np.float64) for s in shapes]
    ref_m = [np.zeros(s) for s in shapes]
    my_p = [torch.from_numpy(p.astype(np.float32)).cuda().requires_grad_(True) for p in ref_p]
    opt = Lion(my_p, lr=lr, betas=betas, weight_decay=wd, max_grad_norm=max_norm)
    ok = [np.ones(s, bool) for s in shapes]
    for it in range(4):
        grads = [(rng.standard_normal(s) * (3.0 if it == 1 else 0.3)).astype(np.float32) for s in shapes]
        for q, gr in zip(my_p, grads):
            q.grad = torch.from_numpy(gr).cuda()
        coef = 1.0
        if max_norm is not None:
            total, coef = clip_coef(grads, max_norm)
        opt.step()
        if max_norm is not None:
            np.testing.assert_allclose(opt.grad_norm.item(), total, rtol=1e-5)
        for i, gr in enumerate(grads):
            ref_p[i], ref_m[i], u = lion_step(ref_p[i], gr.astype(np.float64) * coef, ref_m[i], lr, betas, wd)
            ok[i] &= np.abs(u) > 1e-6 * (np.abs(ref_m[i]) + np.abs(gr) + 1e-30)
        for i, q in enumerate(my_p):
            got = q.detach().cpu().numpy().astype(np.float64)
            assert ok[i].mean() > 0.99
            np.testing.assert_allclose(got[ok[i]], ref_p[i][ok[i]], rtol=2e-6, atol=2e-6)
            np.testing.assert_allclose(opt.state[q]["exp_avg"].cpu().numpy(), ref_m[i], rtol=2e-5, atol=1e-6)


def test_lion_first_step_moves_by_lr_exactly(cuda_device):
    """Zero momentum: every element with a non-zero gradient moves by exactly lr (fp32), zeros stay."""
    from statecatcher_b200.optim import Lion
    p = torch.zeros(1000, device="cuda", requires_grad=True)
    g = torch.randn(1000, generator=torch.Generator().manual_seed(0))
    g[::10] = 0.0
    p.grad = g.cuda()
    Lion([p], lr=0.25).step()
    assert torch.equal(p.detach().cpu(), -0.25 * torch.sign(g))


def test_lion_shim_and_errors(cuda_device):
    import os, sys, importlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "shims"))
    try:
        sys.modules.pop("lion_pytorch", None)
        from statecatcher_b200.optim import Lion
        assert importlib.import_module("lion_pytorch").Lion is Lion
    finally:
        sys.path.remove(os.path.join(root, "shims"))
        sys.modules.pop("lion_pytorch", None)
    with pytest.raises(ValueError):
        Lion([torch.zeros(2, device="cuda", requires_grad=True)], lr=0.0)
    h = torch.zeros(4, device="cuda", dtype=torch.float16, requires_grad=True)
    h.grad = torch.zeros_like(h)
    with pytest.raises(TypeError):
        Lion([h]).step()


@pytest.mark.parametrize("kind", ["adam", "lion"])
def test_vector_and_scalar_bodies_agree_bitwise(cuda_device, kind):
    """16-byte aligned tensors take the float4 body, a view that starts 4 bytes into an allocation
    takes the scalar one: same arithmetic per element, so three steps must agree bit for bit."""
    from statecatcher_b200.optim import FusedAdam, Lion
    n = 4099
    g = torch.Generator().manual_seed(11)
    p0 = torch.randn(n, generator=g)
    grads = [torch.randn(n, generator=g) for _ in range(3)]

    def run(offset):
        base = torch.zeros(n + 8, device="cuda")
        p = base[offset:offset + n]
        p.copy_(p0.cuda())
        p.requires_grad_(True)
        assert p.data_ptr() % 16 == (4 * offset) % 16
        opt = (FusedAdam([p], lr=1e-2, weight_decay=0.01) if kind == "adam"
               else Lion([p], lr=1e-2, weight_decay=0.01))
        for gr in grads:
            gb = torch.zeros(n + 8, device="cuda")
            p.grad = gb[offset:offset + n]
            p.grad.copy_(gr.cuda())
            opt.step()
        return p.detach().cpu().clone()

    assert torch.equal(run(0), run(1))


@pytest.mark.parametrize("offset", [0, 1, 2, 3])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 4099])
def test_clip_norm_of_unaligned_views(cuda_device, offset, n):
    """sc_sumsq_accum takes a gradient that starts anywhere on a 4-byte boundary (scalar head, float4
    body, scalar tail): the norm matches fp64 and the clip scales the whole view."""
    from statecatcher_b200.optim import clip_grad_norm_
    gr = torch.randn(n, generator=torch.Generator().manual_seed(n + offset))
    buf = torch.zeros(n + 8, device="cuda")
    p = buf[offset:offset + n].requires_grad_(True)
    gb = torch.zeros(n + 8, device="cuda")
    p.grad = gb[offset:offset + n]
    p.grad.copy_(gr.cuda())
    total = gr.double().norm().item()
    norm = clip_grad_norm_([p], 0.5 * total)
    np.testing.assert_allclose(norm.item(), total, rtol=1e-5)
    coef = min(1.0, 0.5 * total / (total + 1e-6))
    np.testing.assert_allclose(p.grad.cpu().numpy(), (gr * coef).numpy(), rtol=1e-5, atol=1e-8)
    assert gb[:offset].abs().sum().item() == 0 and gb[offset + n:].abs().sum().item() == 0


def _assorted_params(seed, count=70):
    """`count` tensors of assorted sizes: an empty one, scalars, n % 4 tails, one off a 16-byte boundary,
    one large enough for several grid strides; more than one 32-tensor chunk."""
    g = torch.Generator().manual_seed(seed)
    sizes = [0, 1, 2, 3, 5, 1024, 4099, 300_001] + [int(torch.randint(1, 3000, (1,), generator=g)) for _ in range(count - 8)]
    vals = [torch.randn(n, generator=g) for n in sizes]
    grads = [[torch.randn(n, generator=g) for n in sizes] for _ in range(3)]
    return sizes, vals, grads


@pytest.mark.parametrize("kind", ["adam", "adamw", "lion"])
@pytest.mark.parametrize("max_norm", [None, 5.0])
def test_multi_tensor_step_equals_per_tensor_step(cuda_device, kind, max_norm):
    """sc_*_multi (one launch per 32 tensors) against the per-tensor calls over three steps: bit-identical
    without clipping (same arithmetic per element); with clipping the two sum the squares in a different
    order, so the shared coefficient may differ in its last bit."""
    from statecatcher_b200.optim import FusedAdam, Lion
    sizes, vals, grads = _assorted_params(5)

    def run(multi):
        ps = []
        for i, v in enumerate(vals):
            if i == 6:                                   # a view that starts 4 bytes into its allocation
                base = torch.zeros(v.numel() + 4, device="cuda")
                p = base[1:1 + v.numel()]
                p.copy_(v.cuda())
            else:
                p = v.cuda()
            ps.append(p.requires_grad_(True))
        if kind == "lion":
            opt = Lion(ps, lr=1e-2, weight_decay=0.01, max_grad_norm=max_norm, multi_tensor=multi)
        else:
            opt = FusedAdam(ps, lr=1e-2, weight_decay=0.01, decoupled=(kind == "adamw"), max_grad_norm=max_norm,
                            multi_tensor=multi)
        norms = []
        for gs in grads:
            for p, gr in zip(ps, gs):
                p.grad = gr.cuda()
            opt.step()
            if max_norm is not None:
                norms.append(opt.grad_norm.item())
        return [p.detach().cpu() for p in ps], norms

    one, n1 = run(False)
    many, n2 = run(True)
    for a, b in zip(one, many):
        if max_norm is None:
            assert torch.equal(a, b)
        else:
            # Lion moves by lr*sign: a flipped sign of a near-zero interpolation shows as 2*lr on one element
            if kind == "lion":
                assert (a - b).abs().gt(1e-6).float().mean().item() < 1e-3 if a.numel() else True
            else:
                np.testing.assert_allclose(b.numpy(), a.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(n2, n1, rtol=1e-5)


def test_multi_tensor_norm_matches_fp64(cuda_device):
    from statecatcher_b200.optim import _global_sumsq
    sizes, vals, _ = _assorted_params(9)
    gs = [v.cuda() for v in vals]
    acc = _global_sumsq([(None, g) for g in gs], gs[0].device, multi_tensor=True)
    want = sum(float((v.double() ** 2).sum()) for v in vals)
    np.testing.assert_allclose(acc.item(), want, rtol=1e-5)
