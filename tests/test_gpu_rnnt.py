"""GPU: RNN-T wavefront loss (K4) vs the numpy oracle (oracle/rnnt_oracle.py, Graves 2012)
and the torchaudio golden vectors.  Parity with the reference's own warp_rnnt call is
UNPINNED; the tolerance contract is fp32 rtol 1e-4."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import rnnt_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["basic", "single", "wide"])
def test_rnnt_golden_torchaudio(cuda_device, case):
    import statecatcher_b200 as sb
    G = load_golden("rnnt_cases")
    logits = torch.tensor(G[case + "/logits"]).cuda().requires_grad_(True)
    lp = logits.log_softmax(-1)
    nll = sb.rnnt_loss(lp, torch.tensor(G[case + "/targets"]).long().cuda(), G[case + "/frame_lens"].tolist(),
                       G[case + "/label_lens"].tolist(), blank=0, reduction="none")
    np.testing.assert_allclose(nll.detach().cpu().numpy(), G[case + "/nll"], rtol=1e-4, atol=1e-5)
    nll.sum().backward()                      # through torch's log_softmax backward -> d/dlogits
    np.testing.assert_allclose(logits.grad.cpu().numpy(), G[case + "/grad"], rtol=2e-4, atol=2e-6)


def test_rnnt_published_warp_transducer_vector(cuda_device):
    """The known-answer vector of the warp-transducer / warp-rnnt test suites (oracle/rnnt_oracle.py, WARP_KAT_*)
    through the keyword call of model.py:97-105: `log_probs` = log_softmax(acts), blank 0."""
    import statecatcher_b200 as sb
    acts = torch.tensor(rnnt_oracle.WARP_KAT_ACTS, dtype=torch.float32).cuda().requires_grad_(True)
    loss = sb.RNNTLoss(log_probs=acts.log_softmax(-1), labels=torch.tensor(rnnt_oracle.WARP_KAT_LABELS).cuda(),
                       frames_lengths=[2], labels_lengths=[2], blank_id=0, compact=False, gather=True)
    loss.backward()
    np.testing.assert_allclose(loss.item(), rnnt_oracle.WARP_KAT_COST, rtol=0, atol=2e-6)
    np.testing.assert_allclose(acts.grad.cpu().numpy(), np.array(rnnt_oracle.WARP_KAT_GRADS), rtol=0, atol=2e-6)


@pytest.mark.parametrize("B,T,U,V", [(1, 1, 0, 3), (2, 5, 1, 4), (3, 37, 9, 11), (2, 70, 40, 6), (4, 33, 17, 29)])
def test_rnnt_random_vs_oracle(cuda_device, B, T, U, V):
    import statecatcher_b200 as sb
    g = torch.Generator().manual_seed(B * 100 + T)
    lp = torch.randn(B, T, U + 1, V, generator=g).log_softmax(-1)
    labels = torch.randint(1, V, (B, max(U, 1)), generator=g)
    fl = [T] + [int(torch.randint(1, T + 1, (1,), generator=g)) for _ in range(B - 1)]
    ll = [U] + [int(torch.randint(0, U + 1, (1,), generator=g)) for _ in range(B - 1)]
    nll_ref, grad_ref = rnnt_oracle.rnnt_loss_and_grad(lp.numpy(), labels.numpy(), fl, ll, blank=0)
    x = lp.cuda().requires_grad_(True)
    w = torch.rand(B, generator=g) + 0.5
    nll = sb.rnnt_loss(x, labels.cuda(), fl, ll, blank=0, reduction="none")
    (nll * w.cuda()).sum().backward()
    np.testing.assert_allclose(nll.detach().cpu().numpy(), nll_ref, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(x.grad.cpu().numpy(), grad_ref * w.numpy()[:, None, None, None], rtol=2e-4, atol=2e-6)
    # structure: only blank/label entries of live nodes are non-zero; dead frames all zero
    for b in range(B):
        assert (x.grad[b, fl[b]:] == 0).all()
        assert (x.grad[b, :, ll[b] + 1:] == 0).all()


def test_rnnt_reference_call_signature_and_zero_frames(cuda_device):
    """The keyword call of model.py:97-105 (blank_id=, compact=, gather=True); an utterance with
    no frames contributes zero loss and zero gradient."""
    import statecatcher_b200 as sb
    g = torch.Generator().manual_seed(1)
    B, T, U, V = 3, 12, 4, 8
    lp = torch.randn(B, T, U + 1, V, generator=g).log_softmax(-1).cuda().requires_grad_(True)
    labels = torch.randint(1, V, (B, U), generator=g).cuda()
    loss = sb.RNNTLoss(log_probs=lp, labels=labels, frames_lengths=[T, 0, 7], labels_lengths=[U, 0, 2],
                       blank_id=0, compact=False, gather=True)
    assert loss.dim() == 0
    loss.backward()
    assert (lp.grad[1] == 0).all() and torch.isfinite(lp.grad).all()


def test_rnnt_compact_packing_matches_padded(cuda_device):
    """compact=True (model.py:147-200 packing): same loss and the same gradients as the padded
    call, through the compact joiner."""
    import statecatcher_b200 as sb
    torch.manual_seed(3)
    B, T, U, V, J, E = 3, 14, 5, 16, 24, 8
    fl, ll = [14, 9, 11], [5, 2, 0]
    joiner = sb.RNNTPredictorJoiner(enc_out_dim=V, pred_emb_dim=E, join_dim=J, vocab_size=V).cuda()
    cj = sb.RNNTCompactPredictorJoiner(enc_out_dim=V, pred_emb_dim=E, join_dim=J, vocab_size=V).cuda()
    cj.load_state_dict(joiner.state_dict())
    enc_out = torch.randn(B, T, V, device="cuda")
    tokens = torch.randint(1, V, (B, U), device="cuda")
    prefix = torch.cat([torch.zeros(B, 1, dtype=torch.long, device="cuda"), tokens], 1)
    lp = joiner(enc_out, prefix).log_softmax(-1)
    loss_p = sb.RNNTLoss(log_probs=lp, labels=tokens, frames_lengths=fl, labels_lengths=ll, blank_id=0)
    loss_p.backward()
    lpc = cj(enc_out, prefix, fl, ll).log_softmax(-1)
    assert lpc.shape[0] == sum(t * (u + 1) for t, u in zip(fl, ll))
    loss_c = sb.RNNTLoss(log_probs=lpc, labels=tokens, frames_lengths=fl, labels_lengths=ll, blank_id=0, compact=True)
    loss_c.backward()
    np.testing.assert_allclose(loss_c.item(), loss_p.item(), rtol=1e-5)
    for (n, p), (_, q) in zip(joiner.named_parameters(), cj.named_parameters()):
        np.testing.assert_allclose(q.grad.cpu().numpy(), p.grad.cpu().numpy(), rtol=2e-4, atol=1e-6, err_msg=n)


def test_rnnt_joiner_and_large_lattice(cuda_device):
    """Joiner (model.py:112-145) -> log_softmax -> loss on a cfg4-shaped small batch:
    J=512, V=1024, E=64, T=300, U<=60: finite, reproducible, gradients reach every joiner
    parameter; loss equals the oracle on one utterance."""
    import statecatcher_b200 as sb
    torch.manual_seed(0)
    B, T, U, V, J, E = 2, 300, 60, 1024, 512, 64
    joiner = sb.RNNTPredictorJoiner(enc_out_dim=V, pred_emb_dim=E, join_dim=J, vocab_size=V).cuda()
    enc_out = torch.randn(B, T, V, device="cuda") * 0.1
    tokens = torch.randint(1, V, (B, U), device="cuda")
    prefix = torch.cat([torch.zeros(B, 1, dtype=torch.long, device="cuda"), tokens], 1)
    logits = joiner(enc_out, prefix)
    assert logits.shape == (B, T, U + 1, V)
    lp = logits.float().log_softmax(-1)
    fl, ll = [T, 211], [U, 37]
    nll = sb.rnnt_loss(lp, tokens, fl, ll, reduction="none")
    nll2 = sb.rnnt_loss(lp.detach(), tokens, fl, ll, reduction="none")
    assert torch.isfinite(nll).all() and torch.equal(nll.detach(), nll2)
    nll.mean().backward()
    for n, p in joiner.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
    ref, _ = rnnt_oracle.rnnt_loss_and_grad(lp[1:2].detach().cpu().numpy(), tokens[1:2].cpu().numpy(), [211], [37])
    np.testing.assert_allclose(nll[1].item(), ref[0], rtol=1e-4)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_rnnt_fused_head_matches_unfused(cuda_device, dtype):
    """RNNTFusedHead (chunked joint -> loss, logits never materialised) == joiner +
    log_softmax + RNNTLoss: loss, and gradients of every joiner parameter and of enc_out."""
    import statecatcher_b200 as sb
    torch.manual_seed(4)
    B, T, U, V, J, E, De = 3, 21, 6, 40, 32, 16, 24
    fl, ll = [21, 13, 17], [6, 3, 0]
    ref = sb.RNNTPredictorJoiner(enc_out_dim=De, pred_emb_dim=E, join_dim=J, vocab_size=V).cuda()
    head = sb.RNNTFusedHead(enc_out_dim=De, pred_emb_dim=E, join_dim=J, vocab_size=V, chunk_frames=8,
                            compute_dtype=dtype).cuda()
    head.load_state_dict(ref.state_dict())
    enc_a = torch.randn(B, T, De, device="cuda").requires_grad_(True)
    enc_b = enc_a.detach().clone().requires_grad_(True)
    tokens = torch.randint(1, V, (B, U), device="cuda")
    prefix = torch.cat([torch.zeros(B, 1, dtype=torch.long, device="cuda"), tokens], 1)
    loss_ref = sb.RNNTLoss(log_probs=ref(enc_a, prefix).float().log_softmax(-1), labels=tokens, frames_lengths=fl,
                           labels_lengths=ll, blank_id=0)
    loss_ref.backward()
    loss = head(enc_b, tokens, fl, ll, blank_id=0)
    loss.backward()
    tol = 2e-4 if dtype == torch.float32 else 4e-2
    assert abs(loss.item() - loss_ref.item()) <= tol * abs(loss_ref.item())
    rel = lambda a, b: (a - b).norm().item() / max(b.norm().item(), 1e-12)   # noqa: E731
    assert rel(enc_b.grad, enc_a.grad) <= tol * 2
    for (n, p), (_, q) in zip(ref.named_parameters(), head.named_parameters()):
        assert q.grad is not None, n
        assert rel(q.grad, p.grad) <= tol * 2, (n, rel(q.grad, p.grad))


def test_rnnt_fused_head_cfg4_shape(cuda_device):
    """configs[3] geometry at reduced batch: J=512, V=1024, T=600, U=150, B=8 in bf16 — a lattice
    whose materialised fp32 logits would be 3 GB; runs block-wise, finite, deterministic."""
    import statecatcher_b200 as sb
    torch.manual_seed(0)
    B, T, U, V, J, E = 8, 600, 150, 1024, 512, 64
    head = sb.RNNTFusedHead(enc_out_dim=V, pred_emb_dim=E, join_dim=J, vocab_size=V, chunk_frames=64,
                            compute_dtype=torch.bfloat16).cuda()
    enc = (torch.randn(B, T, V, device="cuda") * 0.1).requires_grad_(True)
    tokens = torch.randint(1, V, (B, U), device="cuda")
    fl = [T] * B
    ll = [75 + 9 * b for b in range(B)]
    l1 = head(enc, tokens, fl, ll)
    l1.backward()
    l2 = head(enc.detach(), tokens, fl, ll)
    assert torch.isfinite(l1) and l1.item() == l2.item()
    assert torch.isfinite(enc.grad).all() and enc.grad.abs().max() > 0
    for n, p in head.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n


def test_rnnt_fused_head_keep_vs_recompute(cuda_device):
    """Keeping every block's joint/logits in HBM and recomputing them in the backward are the
    same computation: identical loss, gradients equal up to atomic summation order; a second backward over kept
    blocks (already released) falls back to recomputation."""
    import statecatcher_b200 as sb
    torch.manual_seed(9)
    B, T, U, V, J, E, De = 2, 37, 5, 48, 64, 16, 32
    fl, ll = [37, 30], [5, 2]
    heads = [sb.RNNTFusedHead(enc_out_dim=De, pred_emb_dim=E, join_dim=J, vocab_size=V, chunk_frames=16,
                              compute_dtype=torch.bfloat16, keep_blocks=k).cuda() for k in (True, False)]
    heads[1].load_state_dict(heads[0].state_dict())
    enc = torch.randn(B, T, De, device="cuda")
    tokens = torch.randint(1, V, (B, U), device="cuda")
    outs = []
    for h in heads:
        e = enc.clone().requires_grad_(True)
        loss = h(e, tokens, fl, ll, blank_id=0)
        loss.backward(retain_graph=True)
        g1 = [e.grad.clone()] + [p.grad.clone() for p in h.parameters()]
        e.grad = None
        h.zero_grad(set_to_none=True)
        loss.backward()                                  # kept blocks were released: recomputed now
        g2 = [e.grad.clone()] + [p.grad.clone() for p in h.parameters()]
        for a, b in zip(g1, g2):                         # split-R weight gradients / bias sums reduce atomically:
            torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-7)   # same values, summation order may differ
        outs.append((loss.detach(), g1))
    assert torch.equal(outs[0][0], outs[1][0])
    for a, b in zip(outs[0][1], outs[1][1]):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
@pytest.mark.parametrize("B,Tc,U1,J", [(3, 13, 7, 512), (2, 64, 151, 512), (2, 9, 5, 64), (2, 8, 3, 96)])
def test_joint_bwd_against_fp64(cuda_device, dtype, B, Tc, U1, J):
    """sc_joint_bwd on its own (the fused head tests only see it through the whole backward): d_enc written for a frame
    block of a longer tensor, d_pred accumulated onto what earlier blocks left, both against an fp64 evaluation of
    d_pre = dJ (1 - tanh^2(enc + pred)); reproducible bit for bit; J = 96 takes the scalar kernels."""
    from statecatcher_b200 import _lib
    from statecatcher_b200._lib import call, ptr, stream
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + Tc)
    big_enc = torch.randn(B, Tc + 3, J, generator=g, device="cuda").to(dtype)
    enc = big_enc[:, 2:2 + Tc]                                                   # a frame block of a longer tensor
    pred = torch.randn(B, U1, J, generator=g, device="cuda").to(dtype)
    dJ = torch.randn(B * Tc * U1, J, generator=g, device="cuda").to(dtype)
    base = torch.randn(B, U1, J, generator=g, device="cuda")                     # d_pred accumulates across blocks
    outs = []
    for _ in range(2):
        d_enc = torch.zeros(B, Tc + 1, J, dtype=dtype, device="cuda")
        d_pred = base.clone()
        call("sc_joint_bwd", ptr(dJ), ptr(enc), enc.stride(0), enc.stride(1), ptr(pred), pred.stride(0), pred.stride(1),
             ptr(d_enc), d_enc.stride(0), d_enc.stride(1), ptr(d_pred), B, Tc, U1, J, _lib.dt(dJ), stream())
        assert not d_enc[:, Tc].any()                                            # nothing written past the block
        outs.append((d_enc, d_pred))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    jt = torch.tanh(enc.double()[:, :, None, :] + pred.double()[:, None, :, :])
    d_pre = dJ.double().view(B, Tc, U1, J) * (1 - jt * jt)
    tol = 2e-2 if dtype == torch.bfloat16 else 1e-4                              # bf16: tanh.approx, bf16 d_enc
    want_p, want_e = base.double() + d_pre.sum(1), d_pre.sum(2)
    assert float((outs[0][1].double() - want_p).abs().max()) <= tol * float(want_p.abs().max())
    assert float((outs[0][0][:, :Tc].double() - want_e).abs().max()) <= tol * float(want_e.abs().max())
