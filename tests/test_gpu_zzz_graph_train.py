"""GPU: GraphedTrainStep (whole training step replayed from one CUDA graph) against the eager step over
carried segments: loss, every parameter gradient and the carried state.  (Green on a B200 since r02.)"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _segments(n, B, T, F, V, U, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for k in range(n):
        x = torch.randn(B, T, F, generator=g)
        tgt = [int(v) for v in torch.randint(0, U + 1, (B,), generator=g)]
        tgt[0], tgt[-1] = U, 0
        tok = torch.zeros(B, U, dtype=torch.int64)
        for b, u in enumerate(tgt):
            tok[b, :u] = torch.randint(1, V, (u,), generator=g)
        inl = [T] * B
        inl[1] = T - 7 - k
        out.append((x, tok, inl, tgt))
    return out


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_graphed_train_step_matches_eager(cuda_device, dtype):
    import statecatcher_b200 as sb
    B, T, F, H, L, V, U = 4, 50, 80, 64, 2, 33, 6
    cfg = sb.LucyRNNConfig(input_dim=F, hidden_dim=H, num_layers=L, vocab_size=V, fused_ops=True, layer_norm=False,
                           is_training=True)
    cd = None if dtype == torch.float32 else dtype
    torch.manual_seed(0)
    eager = sb.LucyRNN(cfg, compute_dtype=cd).cuda()
    with torch.no_grad():
        eager.output_proj.weight.normal_(0, 0.05)
    graphed = sb.LucyRNN(cfg, compute_dtype=cd).cuda()
    graphed.load_state_dict(eager.state_dict())
    runner = sb.GraphedTrainStep(graphed, batch=B, frames=T, feat_dim=F, max_labels=U)
    crit = sb.CTCLoss(blank=0, zero_infinity=True)
    state = None
    tol = dict(rtol=1e-5, atol=1e-6) if dtype == torch.float32 else dict(rtol=2e-2, atol=1e-3)
    for x, tok, inl, tgt in _segments(3, B, T, F, V, U):
        eager.zero_grad(set_to_none=True)
        if state:
            state = sb.detach_states(state)
        logits, state = eager(x.cuda(), state) if state else eager(x.cuda())
        loss = crit(logits.transpose(0, 1), tok.cuda(), inl, tgt)
        loss.backward()
        got = runner.step(x, tok, inl, tgt)
        np.testing.assert_allclose(got.item(), loss.item(), **tol)
        for (k, p), q in zip(eager.named_parameters(), graphed.parameters()):
            a, b = q.grad.float().cpu().numpy(), p.grad.float().cpu().numpy()
            assert np.abs(a - b).max() <= (1e-5 if dtype == torch.float32 else 3e-2) * max(np.abs(b).max(), 1e-6), k
        for a, b in zip(runner.state[0], state[0]):
            np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), **tol)
    # a new stream: zeroed state gives the first segment's loss again
    first = _segments(1, B, T, F, V, U)[0]
    runner.reset()
    again = runner.step(*first).item()
    fresh = crit(eager(first[0].cuda())[0].transpose(0, 1), first[1].cuda(), first[2], first[3]).item()
    np.testing.assert_allclose(again, fresh, **tol)


def test_graphed_train_step_rejects(cuda_device):
    import statecatcher_b200 as sb
    cfg = sb.LucyRNNConfig(80, 16, 1, 9, fused_ops=True, layer_norm=False, is_training=False)
    with pytest.raises(ValueError):
        sb.GraphedTrainStep(sb.LucyRNN(cfg).cuda(), 2, 10, 80, 3)
    cfg = sb.LucyRNNConfig(80, 16, 1, 9, fused_ops=True, layer_norm=False)
    r = sb.GraphedTrainStep(sb.LucyRNN(cfg).cuda(), 2, 10, 80, 3)
    with pytest.raises(ValueError):
        r.step(torch.zeros(2, 10, 80), torch.zeros(2, 4, dtype=torch.int64), [10, 10], [1, 1])   # U > max_labels
    with pytest.raises(ValueError):
        r.step(torch.zeros(2, 10, 80), torch.zeros(2, 3, dtype=torch.int64), [10], [1, 1])
