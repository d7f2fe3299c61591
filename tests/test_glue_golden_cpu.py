"""CPU: the oracles for the code around the encoder — compute_loss / ASRModel glue (model.py:37-110,
282-398), RNN-T predictor+joiner (model.py:112-200), greedy decoder (decoder.py:3-30) — against
tests/golden/glue_cases.npz, which tests/golden/make_glue_golden.py produced by RUNNING the
reference's own functions and classes.  The GPU half is tests/test_gpu_zglue_golden.py."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ctc_oracle, decoder_oracle, joiner_oracle as JO, lucy_oracle as LO


@pytest.fixture(scope="module")
def G():
    return load_golden("glue_cases")


def _sub(G, prefix):
    return {k[len(prefix):]: v for k, v in G.items() if k.startswith(prefix)}


@pytest.mark.parametrize("tag", ["ctc_plain", "ctc_proj"])
@pytest.mark.parametrize("looped", [True, False], ids=["looped", "closed"])
def test_compute_loss_ctc_three_carried_segments(G, tag, looped):
    """detach gate -> (input projection) -> mask multiply -> encoder -> log_softmax/CTC, three
    segments with carried state: loss, enc_out, returned state and every gradient."""
    C = _sub(G, tag + "/")
    cfg = LO.OracleConfig(**{k[4:]: v.item() for k, v in C.items() if k.startswith("cfg_")})
    P = {k[len("param/"):]: torch.tensor(v, dtype=torch.float64).requires_grad_(True) for k, v in C.items()
         if k.startswith("param/")}
    proj = {k[len("proj/"):]: torch.tensor(v, dtype=torch.float64).requires_grad_(True) for k, v in C.items()
            if k.startswith("proj/")}
    fwd = LO.forward_looped if looped else LO.forward_closed
    crit = torch.nn.CTCLoss(blank=0, zero_infinity=True)
    state = None
    for seg in range(3):
        S = _sub(C, f"seg{seg}/")
        for p in list(P.values()) + list(proj.values()):
            p.grad = None
        feats = torch.tensor(S["feats"], dtype=torch.float64)
        if state:                                                    # model.py:60-61
            state = LO.detach_states(state)
        if proj:
            feats = feats @ proj["weight"].T + proj["bias"]          # model.py:305-309
        feats = feats * torch.tensor(S["mask"]).unsqueeze(-1).double()   # model.py:376-377
        enc_out, state = fwd(P, cfg, feats, state)
        np.testing.assert_allclose(enc_out.detach().numpy(), S["enc_out"], rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(torch.stack(state[0]).detach().numpy(), S["h_out"], rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(torch.stack(state[1]).detach().numpy(), S["s_out"], rtol=2e-4, atol=2e-5)
        in_lens, tgt_lens = S["in_lens"].tolist(), S["tgt_lens"].tolist()
        loss = crit(enc_out.log_softmax(-1).transpose(0, 1), torch.tensor(S["tokens"]), in_lens, tgt_lens)
        np.testing.assert_allclose(loss.item(), S["loss"], rtol=1e-4)
        # the numpy CTC oracle on the same logits: loss, and its dlogits pushed through autograd
        nloss, _, dlogits = ctc_oracle.ctc_loss_and_grad(enc_out.detach().numpy(), S["tokens"], in_lens, tgt_lens)
        np.testing.assert_allclose(nloss, S["loss"], rtol=1e-4)
        enc_out.backward(torch.tensor(dlogits))
        for k, p in list(P.items()) + [("proj." + k, v) for k, v in proj.items()]:
            key = "grad/" + (k if k.startswith("proj.") else "encoder." + k)
            want = S[key]
            got = p.grad.numpy() if p.grad is not None else np.zeros_like(want)
            assert np.abs(got - want).max() <= 3e-4 * max(1e-3, np.abs(want).max()), (tag, seg, k)


def test_golden_masked_frames_do_not_reach_the_loss(G):
    """in_lens cuts stream 1 where its mask ends (train.py:486-490): CTC gradient rows beyond
    are exact zeros, so the padded frames only act through the carried state."""
    S = _sub(G, "ctc_plain/seg0/")
    _, _, d = ctc_oracle.ctc_loss_and_grad(S["enc_out"], S["tokens"], S["in_lens"].tolist(), S["tgt_lens"].tolist())
    cut = int(S["in_lens"][1])
    assert not S["mask"][1, cut:].any() and S["mask"][1, :cut].all()
    assert (d[1, cut:] == 0).all() and np.abs(d[1, :cut]).max() > 0


@pytest.mark.parametrize("case", ["joiner", "joiner16"])
def test_joiner_oracle_matches_reference_classes(G, case):
    J = _sub(G, case + "/")
    P = _sub(J, "param/")
    assert list(P) == ["embedding.weight", "enc_proj.weight", "enc_proj.bias", "pred_proj.weight", "pred_proj.bias",
                       "joiner.weight", "joiner.bias"]                               # state_dict order, model.py:115-127
    np.testing.assert_array_equal(JO.blank_prefix(J["tokens"], 0), J["prefix"])      # model.py:76-83
    pad = JO.joiner_padded(P, J["enc_out"], J["prefix"])
    np.testing.assert_allclose(pad, J["logits_padded"], rtol=1e-5, atol=1e-5)
    comp = JO.joiner_compact(P, J["enc_out"], J["prefix"], J["in_lens"], J["tgt_lens"])
    assert comp.shape == J["logits_compact"].shape
    np.testing.assert_allclose(comp, J["logits_compact"], rtol=1e-5, atol=1e-5)
    # compact rows are the live nodes of the padded lattice in (b, t, u) order (model.py:174-196)
    rows = [pad[b, :int(T), :int(U) + 1].reshape(-1, pad.shape[-1]) for b, (T, U) in enumerate(zip(J["in_lens"], J["tgt_lens"]))]
    np.testing.assert_allclose(np.concatenate(rows), comp, rtol=0, atol=1e-12)


@pytest.mark.parametrize("case", ["joiner", "joiner16"])
def test_rnnt_compute_loss_oracle_matches_reference_composition(G, case):
    """compute_loss(mode='rnnt') of the reference around its padded joiner (criterion: torchaudio
    behind warp_rnnt's keyword signature): loss, d enc_out, every joiner gradient."""
    J = _sub(G, case + "/")
    loss, d_enc, g = JO.rnnt_head_loss_and_grads(_sub(J, "param/"), J["enc_out"], J["tokens"], J["in_lens"], J["tgt_lens"])
    np.testing.assert_allclose(loss, J["rnnt_loss"], rtol=1e-5)
    np.testing.assert_allclose(d_enc, J["rnnt_grad_enc_out"], rtol=1e-4, atol=1e-6)
    for k, want in _sub(J, "rnnt_grad/").items():
        assert np.abs(g[k] - want).max() <= 1e-4 * max(1e-3, np.abs(want).max()), k


@pytest.mark.parametrize("name", ["ties", "wide", "one_frame"])
@pytest.mark.parametrize("blank", [0, 2])
def test_decoder_oracle_matches_reference_decoder(G, name, blank):
    D = _sub(G, f"decoder/{name}/")
    got = decoder_oracle.ctc_greedy_decode(D["x"].astype(np.float32), D["lens"], blank=blank)
    assert [len(s) for s in got] == D[f"b{blank}/counts"].tolist()
    assert [t for s in got for t in s] == D[f"b{blank}/flat"].tolist()
