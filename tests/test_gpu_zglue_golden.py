"""GPU: the package's glue, RNN-T joiners / fused head and greedy decoder against golden vectors
produced by RUNNING the reference's own ``compute_loss`` / ``ASRModel`` / joiner classes /
``ctc_greedy_decoder`` (tests/golden/make_glue_golden.py -> glue_cases.npz).  fp32 path, the
north star's tolerance: rtol 1e-4 on outputs, states and loss; gradients within 2e-4 of their
maximum (same bound as tests/test_gpu_module.py).  Integer work (decoder) is exact."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    return load_golden("glue_cases")


def _sub(G, prefix):
    return {k[len(prefix):]: v for k, v in G.items() if k.startswith(prefix)}


def _close(got, want, rtol=1e-4, atol_rel=2e-5):
    want = np.asarray(want)
    np.testing.assert_allclose(got.detach().float().cpu().numpy(), want, rtol=rtol,
                               atol=atol_rel * max(1.0, float(np.abs(want).max())))


def _grad_close(named_params, want_by_name, tol=2e-4):
    for k, p in named_params:
        want = want_by_name[k]
        got = p.grad.detach().cpu().numpy() if p.grad is not None else np.zeros_like(want)
        scale = max(1e-3, float(np.abs(want).max()))
        assert np.abs(got - want).max() <= tol * scale, (k, float(np.abs(got - want).max()), scale)


@pytest.mark.parametrize("tag", ["ctc_plain", "ctc_proj"])
def test_compute_loss_ctc_matches_reference_compute_loss(cuda_device, G, tag):
    """sb.compute_loss + sb.LucyASRModel + sb.CTCLoss == model.compute_loss + ASRModel + nn.CTCLoss
    over three segments with carried state, a masked stream and (ctc_proj) the input projection."""
    import statecatcher_b200 as sb
    C = _sub(G, tag + "/")
    cfg = sb.LucyRNNConfig(**{k[4:]: v.item() for k, v in C.items() if k.startswith("cfg_")})
    proj = _sub(C, "proj/")
    F = C["seg0/feats"].shape[-1]
    model = sb.LucyASRModel(cfg, frontend=None, feat_dim=F, proj_dim=cfg.input_dim if proj else -1).cuda()
    model.encoder.load_state_dict({k: torch.tensor(v) for k, v in _sub(C, "param/").items()}, strict=True)
    if proj:
        model.proj.load_state_dict({k: torch.tensor(v) for k, v in proj.items()}, strict=True)
    crit = sb.CTCLoss(blank=0, zero_infinity=True)
    state = None
    for seg in range(3):
        S = _sub(C, f"seg{seg}/")
        model.zero_grad(set_to_none=True)
        loss, new_state, enc_out, again = sb.compute_loss(
            "ctc", crit, model, torch.tensor(S["feats"]).cuda(), torch.tensor(S["mask"]).cuda(),
            torch.tensor(S["tokens"]).cuda(), S["in_lens"].tolist(), S["tgt_lens"].tolist(), blank_id=0, input_state=state)
        assert again is new_state                                   # model.py:110 returns the state twice
        _close(enc_out, S["enc_out"])
        _close(torch.stack(new_state[0]), S["h_out"])
        _close(torch.stack(new_state[1]), S["s_out"])
        np.testing.assert_allclose(loss.item(), S["loss"], rtol=1e-4)
        loss.backward()
        _grad_close(model.named_parameters(), _sub(S, "grad/"))
        state = new_state


def _joiner(sb, cls, J, **kw):
    P = _sub(J, "param/")
    V, E = P["embedding.weight"].shape
    Jd, De = P["enc_proj.weight"].shape
    m = cls(enc_out_dim=De, pred_emb_dim=E, join_dim=Jd, vocab_size=V, **kw).cuda()
    m.load_state_dict({k: torch.tensor(v) for k, v in P.items()}, strict=True)
    return m


@pytest.mark.parametrize("case", ["joiner", "joiner16"])
def test_joiners_match_reference_classes(cuda_device, G, case):
    """RNNTPredictorJoiner / RNNTCompactPredictorJoiner logits == the reference classes' (model.py:112-200)."""
    import statecatcher_b200 as sb
    J = _sub(G, case + "/")
    enc_out, prefix = torch.tensor(J["enc_out"]).cuda(), torch.tensor(J["prefix"]).cuda()
    pad = _joiner(sb, sb.RNNTPredictorJoiner, J)
    _close(pad(enc_out, prefix), J["logits_padded"])
    cj = _joiner(sb, sb.RNNTCompactPredictorJoiner, J)
    got = cj(enc_out, prefix, torch.tensor(J["in_lens"]).cuda(), torch.tensor(J["tgt_lens"]).cuda())
    assert tuple(got.shape) == J["logits_compact"].shape
    _close(got, J["logits_compact"])


@pytest.mark.parametrize("case", ["joiner", "joiner16"])
def test_compute_loss_rnnt_matches_reference_composition(cuda_device, G, case):
    """sb.compute_loss(mode='rnnt') with sb.RNNTLoss behind the reference's keyword call: loss,
    d enc_out and every joiner gradient against the reference's compute_loss (torchaudio under
    warp_rnnt's signature)."""
    import statecatcher_b200 as sb

    class Enc(torch.nn.Module):
        def forward(self, feats, masks, state):
            return feats, state

    class Args:
        debug = False
        compact_rnnt = False

    J = _sub(G, case + "/")
    pad = _joiner(sb, sb.RNNTPredictorJoiner, J)
    e = torch.tensor(J["enc_out"]).cuda().requires_grad_(True)
    loss, _, eo, _ = sb.compute_loss("rnnt", sb.RNNTLoss, Enc(), e, None, torch.tensor(J["tokens"]).cuda(),
                                     torch.tensor(J["in_lens"]).cuda(), torch.tensor(J["tgt_lens"]).cuda(), blank_id=0,
                                     use_rnnt_joiner=pad, input_state=None, args=Args())
    assert eo is e
    np.testing.assert_allclose(loss.item(), J["rnnt_loss"], rtol=1e-4)
    loss.backward()
    _grad_close([("enc_out", e)], {"enc_out": J["rnnt_grad_enc_out"]})
    _grad_close(pad.named_parameters(), _sub(J, "rnnt_grad/"))


def test_fused_head_matches_reference_composition(cuda_device, G):
    """RNNTFusedHead (joiner + log_softmax + loss in blocks of frames) on the 16-byte-aligned case,
    fp32, both with kept blocks and with recomputation."""
    import statecatcher_b200 as sb
    J = _sub(G, "joiner16/")
    for keep in (True, False):
        head = _joiner(sb, sb.RNNTFusedHead, J, chunk_frames=4, keep_blocks=keep)
        e = torch.tensor(J["enc_out"]).cuda().requires_grad_(True)
        loss = head(e, torch.tensor(J["tokens"]).cuda(), J["in_lens"].tolist(), J["tgt_lens"].tolist(), blank_id=0)
        assert abs(loss.item() - float(J["rnnt_loss"])) <= 2e-4 * abs(float(J["rnnt_loss"]))
        loss.backward()
        _grad_close([("enc_out", e)], {"enc_out": J["rnnt_grad_enc_out"]}, tol=1e-3)
        _grad_close(head.named_parameters(), _sub(J, "rnnt_grad/"), tol=1e-3)


@pytest.mark.parametrize("name", ["ties", "wide", "one_frame"])
@pytest.mark.parametrize("blank", [0, 2])
def test_greedy_decoder_matches_reference_decoder(cuda_device, G, name, blank):
    import statecatcher_b200 as sb
    D = _sub(G, f"decoder/{name}/")
    x = torch.tensor(D["x"].astype(np.float32)).cuda()
    got = sb.ctc_greedy_decoder(x, torch.tensor(D["lens"]).cuda(), blank=blank)
    assert [len(s) for s in got] == D[f"b{blank}/counts"].tolist()
    assert [int(t) for s in got for t in s] == D[f"b{blank}/flat"].tolist()
