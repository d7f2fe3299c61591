"""GPU: the pieces composed in train.py's order (train.py:466-580) — waveform -> frontend ->
frame mask / in_lens -> compute_loss (detach carried state, encoder, fused CTC) -> backward ->
clip + AdamW -> next segment with the carried state — against the CPU composition of the
oracles (fp64 frontend oracle -> LucyRNN oracle + torch CTC -> torch clip_grad_norm_ + AdamW)."""
import numpy as np
import pytest
import torch

from oracle import frontend_oracle as FO, lucy_oracle as LO

pytestmark = pytest.mark.gpu


def test_training_loop_three_segments_matches_cpu_composition(cuda_device):
    import statecatcher_b200 as sb
    from statecatcher_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(2026)
    B, S, NSEG, V, U = 3, 8000, 3, 19, 6
    cfg = sb.LucyRNNConfig(input_dim=80, hidden_dim=48, num_layers=2, vocab_size=V, fused_ops=True, layer_norm=False,
                           is_training=True)
    ocfg = LO.OracleConfig(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
    P = LO.random_params(ocfg, 5, dtype=torch.float32)
    t = torch.arange(S) / 16000.0
    wavs, masks, toks, tgls = [], [], [], []
    for k in range(NSEG):
        w = 0.05 * torch.randn(B, S, generator=g)
        for b in range(B):
            w[b] += 0.3 * torch.sin(2 * np.pi * (180.0 + 90 * b + 40 * k) * t)
        m = torch.ones(B, S, dtype=torch.bool)
        m[1, 5000 + 500 * k:] = False                       # stream 1 runs out of audio inside every segment
        w[1, 5000 + 500 * k:] = 0.0
        wavs.append(w); masks.append(m)
        toks.append(torch.randint(1, V, (B, U), generator=g))
        tgls.append([U, 3, U - 1])

    # ---- CPU composition of the oracles ----
    Pd = {k: v.double().clone().requires_grad_(True) for k, v in P.items()}
    ref_opt = torch.optim.AdamW(list(Pd.values()), lr=2e-3, weight_decay=0.01)
    crit_ref = torch.nn.CTCLoss(blank=0, zero_infinity=True)
    state, ref_losses, ref_lens = None, [], []
    for k in range(NSEG):
        feats = torch.tensor(FO.mfcc(wavs[k].numpy()))
        fm, in_lens = FO.frame_mask_and_lens(masks[k].numpy(), feats.shape[1])
        feats = feats * torch.tensor(fm).unsqueeze(-1).double()
        if state is not None:
            state = ([h.detach() for h in state[0]], [s.detach() for s in state[1]])
        logits, state = LO.forward_closed(Pd, ocfg, feats, state)
        loss = crit_ref(logits.log_softmax(-1).transpose(0, 1), toks[k], in_lens, tgls[k])
        ref_opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(Pd.values()), 5.0)
        ref_opt.step()
        ref_losses.append(loss.item()); ref_lens.append(in_lens)
    ref_state = state

    # ---- the same loop on the GPU through the package's public surface ----
    frontend, _ = sb.make_frontend("mfcc", 16000)
    frontend = frontend.cuda()
    model = sb.LucyASRModel(cfg, frontend=frontend).cuda()
    model.encoder.load_state_dict(P)
    crit = sb.CTCLoss(blank=0, zero_infinity=True)
    opt = FusedAdam(model.parameters(), lr=2e-3, weight_decay=0.01, decoupled=True, max_grad_norm=5.0)
    state = None
    for k in range(NSEG):
        batch, mask_tensor = wavs[k].cuda(), masks[k].cuda()
        with torch.no_grad():
            feats = frontend(batch)                         # train.py:473
        feats = feats.transpose(1, 2).contiguous()          # train.py:475
        frame_mask, in_lens = sb.frame_mask_and_lens(mask_tensor, feats.size(1), model.cfg.stack_order)
        assert in_lens == ref_lens[k] and feats.size(1) == frame_mask.size(1)
        loss, state, enc_out, _ = sb.compute_loss("ctc", crit, model, feats, frame_mask, toks[k].cuda(), in_lens, tgls[k],
                                                  blank_id=0, input_state=state)
        opt.zero_grad()
        loss.backward()
        opt.step()
        # frontend fp32 noise (~1e-3 on MFCC values ~100) goes through two recurrent layers
        np.testing.assert_allclose(loss.item(), ref_losses[k], rtol=2e-3)
    for name, p in model.encoder.named_parameters():
        want = Pd[name].detach().numpy()
        assert np.abs(p.detach().cpu().numpy() - want).max() <= 2e-3 * max(1e-2, np.abs(want).max()), name
    ref_h = torch.stack([h.detach() for h in ref_state[0]]).numpy()
    np.testing.assert_allclose(torch.stack(state[0]).cpu().numpy(), ref_h, rtol=5e-3, atol=5e-4)
