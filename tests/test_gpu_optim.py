"""GPU: fused clip + Adam/AdamW vs torch's CPU clip_grad_norm_ + optim.Adam/AdamW (the library
calls the reference makes at train.py:112-137, 553) over several steps."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("decoupled,wd", [(True, 0.01), (False, 0.0), (False, 0.05)])
@pytest.mark.parametrize("max_norm", [None, 0.5, 1e9])
def test_fused_adam_matches_torch(cuda_device, decoupled, wd, max_norm):
    from statecatcher_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(3)
    shapes = [(33, 17), (5,), (257, 64), (1,)]
    ref_p = [torch.randn(s, generator=g, dtype=torch.float64).requires_grad_(True) for s in shapes]
    my_p = [p.detach().float().cuda().requires_grad_(True) for p in ref_p]
    ref_opt = (torch.optim.AdamW if decoupled else torch.optim.Adam)(ref_p, lr=3e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=wd)
    my_opt = FusedAdam(my_p, lr=3e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=wd, decoupled=decoupled, max_grad_norm=max_norm)
    for it in range(4):
        grads = [torch.randn(s, generator=g, dtype=torch.float64) * (3.0 if it == 1 else 0.3) for s in shapes]
        for p, q, gr in zip(ref_p, my_p, grads):
            p.grad = gr.clone()
            q.grad = gr.float().cuda()
        if max_norm is not None:
            norm_ref = torch.nn.utils.clip_grad_norm_(ref_p, max_norm)
        ref_opt.step()
        my_opt.step()
        if max_norm is not None:
            np.testing.assert_allclose(my_opt.grad_norm.item(), norm_ref.item(), rtol=1e-5)
        for p, q in zip(ref_p, my_p):
            np.testing.assert_allclose(q.detach().cpu().numpy(), p.detach().numpy(), rtol=2e-5, atol=2e-6)


def test_standalone_clip(cuda_device):
    from statecatcher_b200.optim import clip_grad_norm_
    g = torch.Generator().manual_seed(1)
    ps = [torch.zeros(1000, 37).cuda().requires_grad_(True), torch.zeros(13).cuda().requires_grad_(True)]
    gs = [torch.randn(1000, 37, generator=g), torch.randn(13, generator=g)]
    for p, gr in zip(ps, gs):
        p.grad = gr.cuda()
    total = torch.sqrt(sum((x.double() ** 2).sum() for x in gs)).item()
    norm = clip_grad_norm_(ps, 2.0)
    np.testing.assert_allclose(norm.item(), total, rtol=1e-5)
    coef = min(1.0, 2.0 / (total + 1e-6))
    np.testing.assert_allclose(ps[0].grad.cpu().numpy(), (gs[0] * coef).numpy(), rtol=1e-5, atol=1e-7)
    before = ps[1].grad.clone()
    clip_grad_norm_(ps, 1e9)                               # no-op branch: bit-exact
    assert torch.equal(ps[1].grad, before)


@pytest.mark.parametrize("flat_dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_grads_pack_unpack_multi(cuda_device, flat_dtype):
    """sc_grads_pack_multi / sc_grads_unpack_multi (the flat communication buffer of dp.StreamDataParallel): 40 tensors
    (two launches of <= 32), odd sizes, slices on 16-byte boundaries; pack == torch cast, unpack == scale * slice."""
    import ctypes
    from statecatcher_b200 import _lib
    from statecatcher_b200.dp import _aligned_offsets
    from statecatcher_b200.optim import _counts, _table
    g = torch.Generator().manual_seed(3)
    sizes = [1, 7, 8, 9, 1000, 4097] * 6 + [33, 65537, 12, 5]
    grads = [torch.randn(n, generator=g).cuda() for n in sizes]
    offs, total = _aligned_offsets(grads)
    flat = torch.full((total,), 7.0, dtype=flat_dtype, device="cuda")
    esz = flat.element_size()
    slices = (ctypes.c_void_p * len(grads))(*[flat.data_ptr() + o * esz for o in offs])
    _lib.call("sc_grads_pack_multi", _table(grads), slices, _counts(grads), len(grads), _lib.dt(flat), _lib.stream())
    for gr, o in zip(grads, offs):
        assert torch.equal(flat[o:o + gr.numel()], gr.to(flat_dtype))
    out = [torch.zeros_like(gr) for gr in grads]
    _lib.call("sc_grads_unpack_multi", _table(out), slices, _counts(out), len(out), _lib.dt(flat), 0.125, _lib.stream())
    for gr, o2 in zip(grads, out):
        assert torch.equal(o2, gr.to(flat_dtype).float() * 0.125)
