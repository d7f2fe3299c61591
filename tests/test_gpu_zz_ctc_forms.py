"""GPU: the argument forms of torch.nn.functional.ctc_loss that the reference does not use but the
drop-in accepts — concatenated 1-D targets, tensor lengths, unbatched (T,V) input — give the loss and
gradient of the padded form bit for bit, and agree with the fp64 oracle.  (Green on a B200 since r02.)"""
import os

import numpy as np
import pytest
import torch

from oracle import ctc_oracle

pytestmark = pytest.mark.gpu


def _case(seed=0, B=5, T=40, V=11, U=7):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(T, B, V, generator=g)
    tgt_lens = [U, 0, 3, 1, U - 2]
    in_lens = [T, T, T - 9, 5, T]
    tok = torch.zeros(B, U, dtype=torch.int64)
    for b, u in enumerate(tgt_lens):
        tok[b, :u] = torch.randint(1, V, (u,), generator=g)
    return x, tok, in_lens, tgt_lens


@pytest.mark.parametrize("reduction", ["mean", "sum", "none"])
def test_concatenated_and_tensor_length_forms(cuda_device, reduction):
    import statecatcher_b200 as sb
    x, tok, in_lens, tgt_lens = _case()

    def run(targets, il, tl):
        xx = x.cuda().requires_grad_(True)
        loss = sb.ctc_loss(xx, targets, il, tl, blank=0, reduction=reduction, zero_infinity=True)
        (loss.sum() if reduction == "none" else loss).backward()
        return loss.detach().cpu(), xx.grad.cpu()

    l0, g0 = run(tok.cuda(), in_lens, tgt_lens)
    flat = torch.cat([tok[b, :u] for b, u in enumerate(tgt_lens)])
    l1, g1 = run(flat.cuda(), torch.tensor(in_lens), torch.tensor(tgt_lens))
    l2, g2 = run(tok.int().cuda(), torch.tensor(in_lens, dtype=torch.int32).cuda(), torch.tensor(tgt_lens).cuda())
    assert torch.equal(l0, l1) and torch.equal(g0, g1)
    assert torch.equal(l0, l2) and torch.equal(g0, g2)
    loss, nll, grad = ctc_oracle.ctc_loss_and_grad(x.transpose(0, 1).numpy(), tok.numpy(), in_lens, tgt_lens, blank=0,
                                                   reduction=reduction, zero_infinity=True)
    np.testing.assert_allclose(l0.numpy(), nll if reduction == "none" else loss, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(g0.transpose(0, 1).numpy(), grad, rtol=2e-4, atol=2e-5)


def test_unbatched_input_equals_batch_of_one(cuda_device):
    import statecatcher_b200 as sb
    x, tok, _, _ = _case(seed=3)
    x1 = x[:, 0]                                               # (T,V), a strided view of the batch
    t1 = tok[0, :5]
    a = x1.cuda().requires_grad_(True)
    la = sb.ctc_loss(a, t1.cuda(), [x1.size(0)], [5])
    la.backward()
    b = x[:, :1].contiguous().cuda().requires_grad_(True)
    lb = sb.ctc_loss(b, t1[None].cuda(), [x1.size(0)], [5])
    lb.backward()
    assert torch.equal(la, lb) and torch.equal(a.grad, b.grad[:, 0])
    loss, _, grad = ctc_oracle.ctc_loss_and_grad(x1[None].numpy(), t1[None].numpy(), [x1.size(0)], [5], blank=0)
    np.testing.assert_allclose(la.item(), loss, rtol=1e-4)
    np.testing.assert_allclose(a.grad.cpu().numpy(), grad[0], rtol=2e-4, atol=2e-6)

