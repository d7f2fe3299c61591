"""GPU parity at BASELINE.json configs[1] — the shape every headline number is quoted on
(6 layers x 1024, V=1024, T=3000 frames per segment, bf16 and fp32) — against the fp64 oracle
and torch's fp64 CTC, with the batch cut to what the CPU finishes in about a minute.  The batch
dimension never enters a kernel's arithmetic (every stream is an independent chain / lattice),
so B=2..4 at full T, H, V, U exercises exactly the code paths bench.py times:
`lucy_scan_{fwd,bwd}_tma_kernel<bf16,2,...>`, the folded tcgen05 projections, the wavefront CTC
recursion with its 64-row block meetings at T=3000.

Every test records the error it MEASURED in gpurun_out/parity_configs1.json (DESIGN.md section 4
quotes those numbers); the asserted bounds are stated next to each assert.
Reference behaviour matched: /root/reference/lucyrnn.py:109-170, model.py:68-71."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import lucy_oracle as LO

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REC = os.path.join(ROOT, "gpurun_out", "parity_configs1.json")


def _record(key, **vals):
    """Append measured errors to gpurun_out/parity_configs1.json (best effort)."""
    try:
        os.makedirs(os.path.dirname(_REC), exist_ok=True)
        cur = {}
        if os.path.exists(_REC):
            with open(_REC) as f:
                cur = json.load(f)
        cur[key] = {k: float(v) for k, v in vals.items()}
        with open(_REC, "w") as f:
            json.dump(cur, f, indent=1, sort_keys=True)
    except OSError:
        pass


def _rel_l2(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return float(np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-300))


def _max_rel_to_max(got, want):
    """max|err| / max|want| (the bound the module-level gradient checks use)."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-300))


def _needed_rtol(got, want, atol):
    """smallest rtol for which |got-want| <= atol + rtol*|want| holds elementwise."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    err = np.abs(got - want) - atol
    m = err > 0
    if not m.any():
        return 0.0
    return float((err[m] / np.maximum(np.abs(want[m]), 1e-300)).max())


# ------------------------------------------------------------------ (a) K2 scan ------
def _scan_reference(G, h0, s0, training, g_out):
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
def test_scan_at_configs1_width_and_length(cuda_device, dtype, training):
    """K2 forward + backward at H=1024, T=3000 (B=2) against the fp64 closed form + autograd."""
    from statecatcher_b200 import ops
    B, T, H = 2, 3000, 1024
    g = torch.Generator().manual_seed(31)
    G = torch.randn(B, T, 5 * H, generator=g).to(dtype)
    h0 = torch.randn(B, H, generator=g) * 0.5
    s0 = torch.randn(B, H, generator=g) * 0.5
    go = torch.randn(B, T, H, generator=g).to(dtype)
    Href, Sref, dGref = _scan_reference(G.float(), h0, s0, training, go.float())
    Gc = G.cuda().view(B * T, 5 * H)
    Hout, hT, sT, ck = ops.scan_fwd(Gc, B, T, H, h0.cuda(), s0.cuda(), training)
    dG, dbias = ops.scan_bwd(Gc, Hout, h0.cuda(), s0.cuda(), ck, go.cuda().view(B * T, H), B, T, H, training)
    Hn = Hout.view(B, T, H).float().cpu().numpy()
    dGn = dG.view(B, T, 5 * H).float().cpu().numpy()
    m = dict(Hout_maxabs=np.abs(Hn - Href.numpy()).max(), Hout_rel_l2=_rel_l2(Hn, Href.numpy()),
             hT_maxabs=np.abs(hT.cpu().numpy() - Href[:, -1].numpy()).max(),
             dG_rel_to_max=_max_rel_to_max(dGn, dGref.numpy()), dG_rel_l2=_rel_l2(dGn, dGref.numpy()),
             dbias_rel_to_max=_max_rel_to_max(dbias.cpu().numpy(), dGref.reshape(B * T, 5 * H).sum(0).numpy()))
    if not training:
        m["sT_maxabs"] = np.abs(sT.cpu().numpy() - Sref[:, -1].numpy()).max()
    _record(f"scan_{'train' if training else 'step'}_{'f32' if dtype == torch.float32 else 'bf16'}", **m)
    if dtype == torch.float32:
        # fp32 contract: rtol 1e-4 with an absolute floor of 2e-5 of the tensor's largest value
        np.testing.assert_allclose(Hn, Href.numpy(), rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(hT.cpu().numpy(), Href[:, -1].numpy(), rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(dGn, dGref.numpy(), rtol=1e-4, atol=2e-5 * np.abs(dGref.numpy()).max())
        assert m["dbias_rel_to_max"] <= 1e-4
        if not training:
            np.testing.assert_allclose(sT.cpu().numpy(), Sref[:, -1].numpy(), rtol=1e-4, atol=2e-5)
    else:
        # bf16 storage of h_t (|h| <= 1): half an ulp = 2^-9 absolute; gate gradients are stored in bf16
        assert m["Hout_maxabs"] <= 1.2e-2 and m["Hout_rel_l2"] <= 6e-3
        assert m["hT_maxabs"] <= 1e-2
        assert m["dG_rel_l2"] <= 1e-2 and m["dG_rel_to_max"] <= 1e-2
        assert m["dbias_rel_to_max"] <= 1e-2


# ------------------------------------------------------------------ (b) K3 CTC -------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("peaky", [False, True], ids=["random", "peaky"])
def test_ctc_at_configs1_shape(cuda_device, dtype, peaky):
    """T=3000, V=1024, U in [75,150], one short stream and one finished stream (T_b=0, U_b=0), bf16
    and fp32 logits against F.ctc_loss in fp64 on the same (rounded) logits.  'peaky' puts 8-nat
    margins on an alignment of the transcript for half of the frames (a trained model's posteriors),
    which moves the occupancies from ~1e-3 to ~1."""
    from statecatcher_b200 import ctc_loss_from_logits
    B, T, V, U = 4, 3000, 1024, 150
    g = torch.Generator().manual_seed(41 + int(peaky))
    logits = torch.randn(B, T, V, generator=g)
    tokens = torch.randint(1, V, (B, U), generator=g)
    tgl = [150, 75, 113, 0]
    inl = [T, T, 1700, 0]
    if peaky:
        for b in range(3):
            Tb, Ub = inl[b], tgl[b]
            per = Tb // (2 * Ub + 1)
            for s in range(2 * Ub + 1):
                lab = 0 if s % 2 == 0 else int(tokens[b, s // 2])
                logits[b, s * per:s * per + per // 2 + 1, lab] += 8.0
    logits = logits.to(dtype)
    xd = logits.double().requires_grad_(True)
    ref = torch.nn.functional.ctc_loss(xd.log_softmax(-1).transpose(0, 1), tokens, inl, tgl, zero_infinity=True)
    ref.backward()
    x = logits.cuda().requires_grad_(True)
    loss = ctc_loss_from_logits(x, tokens.cuda(), inl, tgl, zero_infinity=True)
    loss.backward()
    got, want = x.grad.float().cpu().numpy(), xd.grad.numpy()
    nll = ctc_loss_from_logits(x.detach(), tokens.cuda(), inl, tgl, reduction="none").cpu().numpy()
    nll_ref = torch.nn.functional.ctc_loss(xd.detach().log_softmax(-1).transpose(0, 1), tokens, inl, tgl,
                                           reduction="none", zero_infinity=True).numpy()
    atol = 2e-7 * (1.0 if dtype == torch.float32 else 1.0)
    m = dict(loss_rel=abs(loss.item() - ref.item()) / abs(ref.item()),
             nll_rel=np.abs(nll - nll_ref)[:3].max() / np.abs(nll_ref[:3]).max(),
             grad_needed_rtol_at_atol_2e7=_needed_rtol(got, want, atol),
             grad_rel_to_max=_max_rel_to_max(got, want), grad_rel_l2=_rel_l2(got, want),
             grad_max=np.abs(want).max())
    _record(f"ctc_{'peaky' if peaky else 'random'}_{'f32' if dtype == torch.float32 else 'bf16'}", **m)
    # structure: exact zeros beyond T_b and for the finished stream
    assert (x.grad[2, 1700:] == 0).all() and (x.grad[3] == 0).all()
    assert m["loss_rel"] <= 1e-5 and m["nll_rel"] <= 1e-5
    if dtype == torch.float32:
        # measured on a B200 (gpurun_out/parity_configs1.json -> DESIGN.md section 4); bound = 2x measured
        np.testing.assert_allclose(got, want, rtol=CTC_F32_GRAD_RTOL, atol=atol)
    else:
        # gradient leaves in bf16: half an ulp = 2^-9 relative, plus the fp32 figure
        np.testing.assert_allclose(got, want, rtol=2.0 ** -8 + CTC_F32_GRAD_RTOL, atol=atol)


# rtol of the CTC gradient wrt logits at T=3000 (north star: 1e-4).  Set from the measured figure, see
# DESIGN.md section 4.
CTC_F32_GRAD_RTOL = 1e-4


# ------------------------------------------------------------------ (c) module -------
_MODULE_CACHE = {}


def _module_case():
    """6 x 1024, V=1024, T=3000, B=2, two carried segments: inputs + the fp64 oracle's outputs."""
    if _MODULE_CACHE:
        return _MODULE_CACHE
    import statecatcher_b200 as sb
    cfg = sb.LucyRNNConfig(input_dim=80, hidden_dim=1024, num_layers=6, vocab_size=1024, fused_ops=True,
                           layer_norm=False, is_training=True)
    ocfg = LO.OracleConfig(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
    P = LO.reference_init_params(ocfg, 11, out_std=0.02)
    Pd = {k: v.double().requires_grad_(True) for k, v in P.items()}
    g = torch.Generator().manual_seed(5)
    B, T = 2, 3000
    xs = [torch.randn(B, T, 80, generator=g) for _ in range(2)]
    toks = [torch.randint(1, 1024, (B, 150), generator=g) for _ in range(2)]
    inl = [[T, 2100], [T, T]]
    tgl = [[150, 75], [101, 0]]
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    losses, logits, state = LO.train_segments(Pd, ocfg, [x.double() for x in xs], toks, inl, tgl, looped=False)
    _MODULE_CACHE.update(cfg=cfg, P=P, grads={k: v.grad.numpy() for k, v in Pd.items() if v.grad is not None},
                         xs=xs, toks=toks, inl=inl, tgl=tgl, losses=[l.item() for l in losses],
                         logits=[l.numpy() for l in logits],
                         h=torch.stack(state[0]).detach().numpy())
    return _MODULE_CACHE


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_module_at_configs1_depth_width_length(cuda_device, dtype):
    """The full encoder + CTC head at configs[1]'s L, H, V, T (B=2), two carried segments, against
    the fp64 oracle (closed form + autograd + torch CTC; about a minute of CPU, shared by both dtypes)."""
    import statecatcher_b200 as sb
    C = _module_case()
    model = sb.LucyRNN(C["cfg"], compute_dtype=dtype).cuda()
    model.load_state_dict(C["P"])
    crit = sb.CTCLoss(blank=0, zero_infinity=True)
    state = None
    m = {}
    for i in range(2):
        if state:
            state = sb.detach_states(state)
        logits, state = model(C["xs"][i].cuda(), state) if state else model(C["xs"][i].cuda())
        loss = crit(logits.transpose(0, 1), C["toks"][i].cuda(), C["inl"][i], C["tgl"][i])
        loss.backward()
        got = logits.detach().float().cpu().numpy()
        m[f"seg{i}_logits_rel_l2"] = _rel_l2(got, C["logits"][i])
        m[f"seg{i}_logits_needed_rtol_at_atol_1e5"] = _needed_rtol(got, C["logits"][i], 1e-5)
        m[f"seg{i}_loss_rel"] = abs(loss.item() - C["losses"][i]) / abs(C["losses"][i])
    hgot = torch.stack(state[0]).cpu().numpy()
    m["h_final_maxabs"] = np.abs(hgot - C["h"]).max()
    worst_l2, worst_max, worst_rtol = 0.0, 0.0, 0.0
    for k, p in model.named_parameters():
        want = C["grads"].get(k)
        if want is None or np.abs(want).max() == 0:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k          # dead r gate: exact zeros
            continue
        gk = p.grad.cpu().numpy()
        m["grad_rel_l2/" + k] = _rel_l2(gk, want)
        worst_l2 = max(worst_l2, m["grad_rel_l2/" + k])
        worst_max = max(worst_max, _max_rel_to_max(gk, want))
        if "W_fused" in k:                                                          # r rows are exact zeros in both
            H = C["cfg"].hidden_dim
            assert (p.grad[:H] == 0).all(), k
            gk, want = gk[H:], want[H:]
        worst_rtol = max(worst_rtol, _needed_rtol(gk, want, 2e-5 * np.abs(want).max()))
    m.update(grad_worst_rel_l2=worst_l2, grad_worst_rel_to_max=worst_max, grad_worst_needed_rtol_at_atol_2e5max=worst_rtol)
    _record(f"module_{'f32' if dtype == torch.float32 else 'bf16'}", **m)
    if dtype == torch.float32:
        for i in range(2):
            assert m[f"seg{i}_logits_needed_rtol_at_atol_1e5"] <= 1e-4        # rtol 1e-4, atol 1e-5 (|logits| ~ 0.3)
            assert m[f"seg{i}_loss_rel"] <= 1e-4
        assert m["h_final_maxabs"] <= 2e-5
        assert worst_rtol <= 1e-4                                              # elementwise rtol 1e-4, atol 2e-5 of max
    else:
        # stated bf16 bound: <= 2x the reference's own fp32->bf16 autocast drift
        for i in range(2):
            assert m[f"seg{i}_logits_rel_l2"] <= 2e-2
            assert m[f"seg{i}_loss_rel"] <= 2e-2
        assert worst_l2 <= 3e-2
