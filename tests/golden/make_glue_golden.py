#!/usr/bin/env python
"""Golden fixtures for the code AROUND the encoder, made by running the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_glue_golden.py            ->  tests/golden/glue_cases.npz

What is executed is the reference's own code, imported from /root/reference:

* ``model.compute_loss`` (model.py:37-110) on ``model.ASRModel`` (model.py:282-398) for three
  consecutive segments with carried state, CTC mode, ``nn.CTCLoss(blank=0, zero_infinity=True)``
  exactly as train.py:142 builds it, a frame mask with a stream that runs out of audio, and
  (second case) the input projection of model.py:305-309.  ``model.py`` imports ``xlstm`` (absent
  here; used only by the xLSTM encoder branch) — it is stubbed in ``sys.modules``.  ``ASRModel``
  builds the Triton network (model.py:310), which needs a GPU; its ``encoder`` attribute is
  replaced by the reference's own ``lucyrnn.LucyRNN(kernel_impl="native")`` — the module the north
  star names — so everything else on the path (detach gate, mask multiply, positional encoder
  call, log_softmax/transpose, criterion call, 4-tuple return) is the reference's.
* ``model.RNNTPredictorJoiner`` / ``model.RNNTCompactPredictorJoiner`` (model.py:112-200): logits
  for seeded weights, padded and compact; and ``compute_loss(mode="rnnt")`` with a criterion that
  has warp_rnnt's keyword signature and evaluates ``torchaudio.functional.rnnt_loss`` (warp_rnnt
  itself is absent: the loss VALUE stays the substitute pin of rnnt_cases.npz; what this pins is the
  reference's blank-prefix construction, joiner call and log_softmax in front of it).
* ``decoder.ctc_greedy_decoder`` (decoder.py:3-30) on tie-heavy and ragged inputs.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, "/root/reference")

# ---- stub the one import of model.py that is not installed (xLSTM encoder branch only) ----
_x = types.ModuleType("xlstm")
_xl = types.ModuleType("xlstm.xlstm_large")
_xm = types.ModuleType("xlstm.xlstm_large.model")
_xm.xLSTMLargeConfig = type("xLSTMLargeConfig", (), {})
_xm.xLSTMLarge = type("xLSTMLarge", (nn.Module,), {})
sys.modules.update({"xlstm": _x, "xlstm.xlstm_large": _xl, "xlstm.xlstm_large.model": _xm})

import decoder as ref_decoder              # noqa: E402  (the reference, unmodified)
import lucyrnn as ref_lucyrnn              # noqa: E402
import model as ref_model                  # noqa: E402
from lucyrnn_conf import LucyRNNConfig      # noqa: E402
from oracle import lucy_oracle as LO        # noqa: E402


def asr_model(cfg, P, feat_dim, proj_dim, proj_sd=None):
    m = ref_model.ASRModel(None, cfg, cfg.vocab_size, feat_dim, proj_dim, debug=False)
    enc = ref_lucyrnn.LucyRNN(cfg)                      # native PyTorch path (lucyrnn.py:72-191)
    enc.load_state_dict(P, strict=True)
    m.encoder = enc
    if proj_sd is not None:
        m.proj.load_state_dict(proj_sd)
    return m


def ctc_pipeline(out, tag, proj_dim, seed):
    F, H, L, V, B, T, NSEG = 7, 12, 2, 9, 3, 21, 3
    in_dim = proj_dim if proj_dim > 0 else F
    kw = dict(input_dim=in_dim, hidden_dim=H, num_layers=L, vocab_size=V, fused_ops=True, layer_norm=False,
              is_training=True)
    cfg = LucyRNNConfig(kernel_impl="native", **kw)
    P = LO.random_params(cfg, seed, dtype=torch.float32)
    g = torch.Generator().manual_seed(seed + 1)
    proj_sd = None
    if proj_dim > 0:
        proj_sd = {"weight": 0.4 * torch.randn(proj_dim, F, generator=g), "bias": 0.1 * torch.randn(proj_dim, generator=g)}
        for k, v in proj_sd.items():
            out[f"{tag}/proj/{k}"] = v.numpy()
    m = asr_model(cfg, P, F, proj_dim, proj_sd)
    for k, v in kw.items():
        out[f"{tag}/cfg_{k}"] = np.asarray(v)
    for k, v in P.items():
        out[f"{tag}/param/{k}"] = v.numpy()
    crit = nn.CTCLoss(blank=0, zero_infinity=True)       # train.py:142
    state = None
    for seg in range(NSEG):
        feats = torch.randn(B, T, F, generator=g)
        mask = torch.ones(B, T, dtype=torch.bool)
        cut = T - 4 - 3 * seg                            # stream 1 runs out of audio inside every segment
        mask[1, cut:] = False
        in_lens = [T, cut, T]
        U = 5
        tokens = torch.randint(1, V, (B, U), generator=g)
        tgt_lens = [U, 2, U - 1]
        m.zero_grad()
        loss, new_state, enc_out, again = ref_model.compute_loss("ctc", crit, m, feats, mask, tokens, in_lens, tgt_lens,
                                                                 blank_id=0, input_state=state)
        assert again is new_state
        loss.backward()
        p = f"{tag}/seg{seg}/"
        out[p + "feats"], out[p + "mask"], out[p + "tokens"] = feats.numpy(), mask.numpy(), tokens.numpy()
        out[p + "in_lens"], out[p + "tgt_lens"] = np.asarray(in_lens), np.asarray(tgt_lens)
        out[p + "loss"], out[p + "enc_out"] = loss.detach().numpy(), enc_out.detach().numpy()
        out[p + "h_out"] = np.stack([t.detach().numpy() for t in new_state[0]])
        out[p + "s_out"] = np.stack([t.detach().numpy() for t in new_state[1]])
        for k, q in m.named_parameters():
            out[p + "grad/" + k] = (q.grad if q.grad is not None else torch.zeros_like(q)).numpy()
        state = new_state


class WarpRNNTStandIn:
    """warp_rnnt.RNNTLoss's keyword signature (model.py:97-105) evaluated by torchaudio."""

    def __call__(self, log_probs, labels, frames_lengths, labels_lengths, blank_id=0, compact=False, gather=True):
        import torchaudio.functional as TAF
        assert not compact
        nll = TAF.rnnt_loss(log_probs, labels.int(), torch.as_tensor(frames_lengths).int(),
                            torch.as_tensor(labels_lengths).int(), blank=blank_id, reduction="none",
                            fused_log_softmax=False)
        return nll.mean()


def joiner_cases(out, tag, dims, seed):
    g = torch.Generator().manual_seed(seed)
    B, T, U, De, E, J, V = dims
    pad = ref_model.RNNTPredictorJoiner(De, E, J, V, debug=False)
    sd = {k: 0.5 * torch.randn(v.shape, generator=g) for k, v in pad.state_dict().items()}
    pad.load_state_dict(sd)
    cj = ref_model.RNNTCompactPredictorJoiner(De, E, J, V, debug=False)
    cj.load_state_dict(sd)
    assert list(cj.state_dict()) == list(sd)
    enc_out = torch.randn(B, T, De, generator=g)
    tokens = torch.randint(1, V, (B, U), generator=g)
    in_lens, tgt_lens = torch.tensor([T, T - 2, 1]), torch.tensor([U, 2, 0])
    prefix = torch.cat([torch.zeros(B, 1, dtype=tokens.dtype), tokens], 1)
    for k, v in sd.items():
        out[tag + "/param/" + k] = v.numpy()
    out[tag + "/enc_out"], out[tag + "/tokens"], out[tag + "/prefix"] = enc_out.numpy(), tokens.numpy(), prefix.numpy()
    out[tag + "/in_lens"], out[tag + "/tgt_lens"] = in_lens.numpy(), tgt_lens.numpy()
    out[tag + "/logits_padded"] = pad(enc_out, prefix).detach().numpy()
    out[tag + "/logits_compact"] = cj(enc_out, prefix, in_lens, tgt_lens).detach().numpy()

    # compute_loss(mode="rnnt") around the padded joiner: the "model" is anything returning (enc_out, state)
    class Enc(nn.Module):
        def forward(self, feats, masks, state):
            return feats, state

    args = types.SimpleNamespace(debug=False, compact_rnnt=False)
    e = enc_out.clone().requires_grad_(True)
    loss, st, eo, _ = ref_model.compute_loss("rnnt", WarpRNNTStandIn(), Enc(), e, None, tokens, in_lens, tgt_lens,
                                             blank_id=0, use_rnnt_joiner=pad, input_state=None, args=args)
    pad.zero_grad()
    loss.backward()
    out[tag + "/rnnt_loss"] = loss.detach().numpy()
    out[tag + "/rnnt_grad_enc_out"] = e.grad.numpy()
    for k, q in pad.named_parameters():
        out[tag + "/rnnt_grad/" + k] = q.grad.numpy()


def decoder_cases(out, seed):
    g = torch.Generator().manual_seed(seed)
    specs = {"ties": (4, 40, 7), "wide": (2, 150, 1024), "one_frame": (3, 1, 5)}
    for name, (B, T, V) in specs.items():
        x = torch.randint(0, 3, (B, T, V), generator=g).float()     # coarse scores: exact ties, long repeats
        lens = torch.tensor([T] + [int(torch.randint(0, T + 1, (1,), generator=g)) for _ in range(B - 1)])
        for blank in (0, 2):
            dec = ref_decoder.ctc_greedy_decoder(x, lens, blank=blank)
            flat = np.asarray([t for seq in dec for t in seq], dtype=np.int64)
            out[f"decoder/{name}/b{blank}/flat"] = flat
            out[f"decoder/{name}/b{blank}/counts"] = np.asarray([len(s) for s in dec], dtype=np.int64)
        out[f"decoder/{name}/x"], out[f"decoder/{name}/lens"] = x.numpy().astype(np.int8), lens.numpy()


def main():
    out = {}
    ctc_pipeline(out, "ctc_plain", proj_dim=-1, seed=401)
    ctc_pipeline(out, "ctc_proj", proj_dim=6, seed=402)
    joiner_cases(out, "joiner", (3, 6, 4, 10, 5, 8, 10), 403)          # odd sizes: scalar kernel paths
    joiner_cases(out, "joiner16", (3, 11, 4, 24, 8, 16, 16), 405)      # 16-byte-aligned rows: vector kernel paths
    decoder_cases(out, 404)
    path = os.path.join(HERE, "glue_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
