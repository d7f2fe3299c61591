"""CPU: pin the oracle (oracle/) against the golden vectors generated from the unmodified
reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import golden_cfg_kwargs, golden_lucy_names, load_golden
from oracle import ctc_oracle, lucy_oracle as LO, rnnt_oracle


def _params(G, dtype):
    return {k[len("param/"):]: torch.tensor(v, dtype=dtype).requires_grad_(True)
            for k, v in G.items() if k.startswith("param/")}


def _state(G, seg, dtype):
    if f"seg{seg}/h_in" not in G:
        return None
    return ([torch.tensor(a, dtype=dtype) for a in G[f"seg{seg}/h_in"]],
            [torch.tensor(a, dtype=dtype) for a in G[f"seg{seg}/s_in"]])


@pytest.mark.parametrize("looped", [True, False], ids=["looped", "closed"])
@pytest.mark.parametrize("name", golden_lucy_names())
def test_lucy_oracle_matches_reference(name, looped):
    G = load_golden("lucy_" + name)
    cfg = LO.OracleConfig(**golden_cfg_kwargs(G))
    if name.startswith("medium") and looped:
        pytest.skip("closed form only for the medium cases (time)")
    P = _params(G, torch.float64)
    fwd = LO.forward_looped if looped else LO.forward_closed
    crit = torch.nn.CTCLoss(blank=0, zero_infinity=True)
    for seg in range(2):
        for p in P.values():
            p.grad = None
        x = torch.tensor(G[f"seg{seg}/x"], dtype=torch.float64)
        logits, (h, s) = fwd(P, cfg, x, _state(G, seg, torch.float64))
        np.testing.assert_allclose(logits.detach().numpy(), G[f"seg{seg}/logits"], rtol=2e-4, atol=2e-5)
        if f"seg{seg}/h_out" in G:
            np.testing.assert_allclose(torch.stack(h).detach().numpy(), G[f"seg{seg}/h_out"], rtol=2e-4, atol=2e-5)
            np.testing.assert_allclose(torch.stack(s).detach().numpy(), G[f"seg{seg}/s_out"], rtol=2e-4, atol=2e-5)
        loss = crit(logits.log_softmax(-1).transpose(0, 1), torch.tensor(G[f"seg{seg}/tokens"]),
                    G[f"seg{seg}/in_lens"].tolist(), G[f"seg{seg}/tgt_lens"].tolist())
        np.testing.assert_allclose(loss.item(), G[f"seg{seg}/loss"], rtol=1e-4, atol=1e-6)
        loss.backward()
        for k, p in P.items():
            want = G[f"seg{seg}/grad/" + k]
            got = p.grad.numpy() if p.grad is not None else np.zeros_like(want)
            scale = max(1e-3, np.abs(want).max())
            assert np.abs(got - want).max() <= 3e-4 * scale, (name, seg, k)


def test_r_gate_is_dead_in_reference():
    """r is computed and never used -> its rows of W_fused get exact zero grads."""
    G = load_golden("lucy_train_fused_ln")
    H = int(G["cfg_hidden_dim"])
    for seg in range(2):
        assert np.all(G[f"seg{seg}/grad/layers.0.W_fused.weight"][:H] == 0)
        assert np.all(G[f"seg{seg}/grad/layers.0.W_fused.bias"][:H] == 0)
        assert np.all(G[f"seg{seg}/grad/layers.0.layernorm_r.weight"] == 0)


def test_training_path_returns_s_unchanged():
    G = load_golden("lucy_train_fused_noln")
    assert np.all(G["seg0/s_out"] == 0)                      # zeros in, zeros out (lucyrnn.py:165)
    np.testing.assert_array_equal(G["seg1/s_out"], G["seg1/s_in"])
    G = load_golden("lucy_step_fused_noln")
    assert np.abs(G["seg0/s_out"]).max() > 0                  # step path carries s (lucyrnn.py:179)


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("ln", [False, True])
def test_hand_backward_matches_autograd(training, ln):
    """scan_backward_closed (the spec of the CUDA backward) vs autograd, fp64."""
    torch.manual_seed(5)
    B, T, H = 2, 9, 6
    z, k, v, p, q = [torch.randn(B, T, H, dtype=torch.float64, requires_grad=True) for _ in range(5)]
    h0 = torch.randn(B, H, dtype=torch.float64)
    s0 = torch.randn(B, H, dtype=torch.float64)
    lnz = [torch.randn(H, dtype=torch.float64, requires_grad=True) for _ in range(2)] if ln else None
    lnh = [torch.randn(H, dtype=torch.float64, requires_grad=True) for _ in range(2)] if ln else None
    g = torch.randn(B, T, H, dtype=torch.float64)

    def lnf(x, wb):
        return x if wb is None else torch.nn.functional.layer_norm(x, (H,), wb[0], wb[1], LO.LN_EPS)
    d = torch.sigmoid(q)
    kv = k * v
    S = LO._linear_scan(d, kv, torch.zeros_like(s0) if training else s0)
    sp = d * S + kv if training else S
    c = torch.tanh(lnf(p + sp, lnh))
    zh = torch.sigmoid(lnf(z, lnz))
    Hout = LO._linear_scan(zh, (1 - zh) * c, h0)
    (Hout * g).sum().backward()
    with torch.no_grad():
        R = LO.scan_backward_closed(g, z, k, v, p, q, h0, s0, training, lnz, lnh)
    for nm, t in (("dz", z), ("dk", k), ("dv", v), ("dp", p), ("dq", q)):
        np.testing.assert_allclose(R[nm].numpy(), t.grad.numpy(), rtol=1e-10, atol=1e-12)
    if ln:
        np.testing.assert_allclose(R["ln_h"][0].numpy(), lnh[0].grad.numpy(), rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(R["ln_z"][1].numpy(), lnz[1].grad.numpy(), rtol=1e-10, atol=1e-12)


CTC_CASES = ["basic", "repeats_tight", "infeasible", "empty_target", "zero_frames", "all_empty", "long"]


@pytest.mark.parametrize("case", CTC_CASES)
def test_ctc_oracle_matches_torch_golden(case):
    G = load_golden("ctc_cases")
    loss, nll, grad = ctc_oracle.ctc_loss_and_grad(G[case + "/logits"], G[case + "/tokens"],
                                                   G[case + "/in_lens"], G[case + "/tgt_lens"])
    np.testing.assert_allclose(loss, G[case + "/loss"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(nll, G[case + "/nll"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(grad, G[case + "/grad"], rtol=1e-4, atol=1e-6)
    # structural facts of App. B: zero rows beyond T_b, rows sum to zero
    for b, Tb in enumerate(G[case + "/in_lens"]):
        assert np.all(grad[b, int(Tb):] == 0)
    assert np.abs(grad.sum(-1)).max() < 1e-12


@pytest.mark.parametrize("case", ["basic", "single", "wide"])
def test_rnnt_oracle_matches_torchaudio_golden(case):
    G = load_golden("rnnt_cases")
    nll, dx = rnnt_oracle.rnnt_loss_and_grad_logits(G[case + "/logits"], G[case + "/targets"],
                                                    G[case + "/frame_lens"], G[case + "/label_lens"])
    np.testing.assert_allclose(nll, G[case + "/nll"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dx, G[case + "/grad"], rtol=1e-4, atol=1e-6)


def test_rnnt_oracle_matches_published_warp_transducer_vector():
    """The loss VALUE and gradient the warp-transducer / warp-rnnt test suites publish (oracle/rnnt_oracle.py,
    WARP_KAT_*): the nearest thing to a pin the reference's absent `warp_rnnt` dependency offers."""
    acts = np.array(rnnt_oracle.WARP_KAT_ACTS, dtype=np.float64)
    nll, dx = rnnt_oracle.rnnt_loss_and_grad_logits(acts, np.array(rnnt_oracle.WARP_KAT_LABELS), [2], [2])
    np.testing.assert_allclose(nll, [rnnt_oracle.WARP_KAT_COST], rtol=0, atol=1e-6)        # printed to seven digits
    np.testing.assert_allclose(dx, np.array(rnnt_oracle.WARP_KAT_GRADS), rtol=0, atol=2e-7)   # printed from an fp32 run


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("reduction", ["mean", "sum"])
def test_ctc_oracle_matches_live_torch_on_random_ragged_batches(seed, reduction):
    """Beyond the fixed edge cases: seeded ragged batches (repeated labels, short inputs, empty and
    infeasible utterances mixed in) against the installed torch's CPU ctc_loss in fp64 — the very
    call the reference makes (train.py:142, model.py:70-71)."""
    rng = np.random.default_rng(1000 + seed)
    B, T, V = int(rng.integers(1, 6)), int(rng.integers(1, 40)), int(rng.integers(2, 12))
    Umax = max(1, int(rng.integers(1, 12)))
    logits = rng.normal(size=(B, T, V)) * 2.0
    in_lens = [int(rng.integers(0, T + 1)) for _ in range(B)]
    in_lens[0] = T
    tgt_lens = [int(rng.integers(0, Umax + 1)) for _ in range(B)]
    tokens = rng.integers(1, V, size=(B, Umax))
    if seed % 3 == 0:                                            # runs of one label: needs blanks between repeats
        tokens[:] = tokens[:, :1]
    x = torch.tensor(logits, requires_grad=True)
    want = torch.nn.functional.ctc_loss(x.log_softmax(-1).transpose(0, 1), torch.tensor(tokens), in_lens, tgt_lens,
                                        blank=0, reduction=reduction, zero_infinity=True)
    want.backward()
    loss, nll, grad = ctc_oracle.ctc_loss_and_grad(logits, tokens, in_lens, tgt_lens, reduction=reduction)
    np.testing.assert_allclose(loss, want.item(), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(grad, x.grad.numpy(), rtol=1e-7, atol=1e-10)


@pytest.mark.parametrize("seed", range(8))
def test_rnnt_oracle_matches_live_torchaudio_on_random_ragged_batches(seed):
    TAF = pytest.importorskip("torchaudio.functional")
    rng = np.random.default_rng(2000 + seed)
    B, T, U, V = int(rng.integers(1, 5)), int(rng.integers(1, 14)), int(rng.integers(0, 7)), int(rng.integers(2, 9))
    logits = rng.normal(size=(B, T, U + 1, V)).astype(np.float32)
    fl = [T] + [int(rng.integers(1, T + 1)) for _ in range(B - 1)]
    ll = [U] + [int(rng.integers(0, U + 1)) for _ in range(B - 1)]
    labels = rng.integers(1, V, size=(B, max(U, 1)))[:, :U] if U else np.zeros((B, 0), np.int64)
    x = torch.tensor(logits, requires_grad=True)
    want = TAF.rnnt_loss(x, torch.tensor(labels).int(), torch.tensor(fl).int(), torch.tensor(ll).int(), blank=0,
                         reduction="none")
    want.sum().backward()
    nll, dx = rnnt_oracle.rnnt_loss_and_grad_logits(logits, labels, fl, ll)
    np.testing.assert_allclose(nll, want.detach().numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dx, x.grad.numpy(), rtol=1e-4, atol=1e-6)


def test_lion_oracle_known_answers():
    """Hand-worked Lion steps (published rule; lion_pytorch itself is absent: parity unpinned)."""
    from oracle.optim_oracle import lion_step, clip_coef
    p = np.array([1.0, -2.0, 0.5, 3.0]); g = np.array([0.2, -0.1, 0.0, -4.0]); m = np.array([0.0, 1.0, 0.0, 1.0])
    # u = .9m + .1g = [.02, .89, 0, .5]; sign = [1, 1, 0, 1]
    p1, m1, u = lion_step(p, g, m, lr=0.1, betas=(0.9, 0.99), weight_decay=0.5)
    np.testing.assert_allclose(u, [0.02, 0.89, 0.0, 0.5], atol=1e-15)
    np.testing.assert_allclose(p1, [0.95 - 0.1, -1.9 - 0.1, 0.475, 2.85 - 0.1], atol=1e-15)
    np.testing.assert_allclose(m1, [0.002, 0.989, 0.0, 0.95], atol=1e-15)
    # first step from zero momentum moves every element with a non-zero gradient by exactly lr
    p2, _, _ = lion_step(p, g, np.zeros(4), lr=1e-4)
    np.testing.assert_allclose(p2 - p, -1e-4 * np.sign(g), atol=1e-18)
    total, coef = clip_coef([np.array([3.0, 0.0]), np.array([4.0])], 1.0)
    assert total == 5.0 and abs(coef - 1.0 / (5.0 + 1e-6)) < 1e-15
    assert clip_coef([np.array([0.3])], 1.0)[1] == 1.0
