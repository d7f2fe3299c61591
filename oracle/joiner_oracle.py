"""RNN-T predictor/joiner oracle — TEST INFRASTRUCTURE ONLY (never imported by the product).

Restates model.py:112-200 and the RNN-T branch of ``compute_loss`` (model.py:73-105) in numpy fp64:

  prefix  = [blank] + tokens                                         model.py:76-83
  pred    = Linear_pred(Embedding(prefix))         (B, U+1, J)       model.py:133-137
  enc     = Linear_enc(enc_out)                    (B, T,   J)       model.py:136
  joint   = tanh(enc[:, :, None] + pred[:, None])  (B, T, U+1, J)    model.py:140-141
  logits  = Linear_out(joint)                      (B, T, U+1, V)    model.py:142
  compact: only the T_b x (U_b+1) live nodes of each utterance, t-major, utterances
  concatenated -> (sum_b T_b*(U_b+1), V)                             model.py:174-196

and the hand-derived backward of ``mean_b nll_b`` through log_softmax, the joiner and both
projections (the reference gets it from autograd).  **Pinned** on tests/golden/glue_cases.npz,
which tests/golden/make_glue_golden.py produces by running the reference's own joiner classes and
``compute_loss``; the loss value under it is rnnt_oracle's (substitute pin: torchaudio, because
warp_rnnt is absent — see oracle/rnnt_oracle.py).

Parameters are a dict with the reference's state_dict keys: ``embedding.weight [V,E]``,
``enc_proj.weight [J,De]``/``.bias``, ``pred_proj.weight [J,E]``/``.bias``, ``joiner.weight [V,J]``/``.bias``.
"""
from __future__ import annotations

import numpy as np

from . import rnnt_oracle


def blank_prefix(tokens, blank=0):
    tokens = np.asarray(tokens, dtype=np.int64)
    return np.concatenate([np.full((tokens.shape[0], 1), blank, dtype=np.int64), tokens], axis=1)


def _f64(P):
    return {k: np.asarray(v, dtype=np.float64) for k, v in P.items()}


def projections(P, enc_out, prefix):
    P = _f64(P)
    emb = P["embedding.weight"][np.asarray(prefix, dtype=np.int64)]            # (B, U+1, E)
    pred = emb @ P["pred_proj.weight"].T + P["pred_proj.bias"]
    enc = np.asarray(enc_out, dtype=np.float64) @ P["enc_proj.weight"].T + P["enc_proj.bias"]
    return enc, pred, emb


def joiner_padded(P, enc_out, prefix):
    """(B,T,De), (B,U+1) -> logits (B,T,U+1,V)."""
    enc, pred, _ = projections(P, enc_out, prefix)
    joint = np.tanh(enc[:, :, None, :] + pred[:, None, :, :])
    Pd = _f64(P)
    return joint @ Pd["joiner.weight"].T + Pd["joiner.bias"]


def joiner_compact(P, enc_out, prefix, in_lens, tgt_lens):
    """-> logits (sum_b T_b*(U_b+1), V), rows ordered (b, t, u)."""
    enc, pred, _ = projections(P, enc_out, prefix)
    Pd = _f64(P)
    rows = []
    for b in range(enc.shape[0]):
        T, U1 = int(in_lens[b]), int(tgt_lens[b]) + 1
        j = np.tanh(enc[b, :T, None, :] + pred[b, None, :U1, :])
        rows.append(j.reshape(T * U1, -1))
    joint = np.concatenate(rows, 0) if rows else np.zeros((0, enc.shape[-1]))
    return joint @ Pd["joiner.weight"].T + Pd["joiner.bias"]


def rnnt_head_loss_and_grads(P, enc_out, tokens, in_lens, tgt_lens, blank=0):
    """compute_loss(mode='rnnt') with a mean-over-batch transducer loss.

    Returns (loss, d_enc_out, dict of parameter gradients)."""
    Pd = _f64(P)
    prefix = blank_prefix(tokens, blank)
    enc, pred, emb = projections(Pd, enc_out, prefix)
    joint = np.tanh(enc[:, :, None, :] + pred[:, None, :, :])
    logits = joint @ Pd["joiner.weight"].T + Pd["joiner.bias"]
    nll, dlogits = rnnt_oracle.rnnt_loss_and_grad_logits(logits, tokens, in_lens, tgt_lens, blank)
    B = logits.shape[0]
    dlogits = dlogits / B                                                      # mean over the batch
    g = {}
    g["joiner.weight"] = np.einsum("btuv,btuj->vj", dlogits, joint)
    g["joiner.bias"] = dlogits.sum((0, 1, 2))
    dpre = (dlogits @ Pd["joiner.weight"]) * (1.0 - joint * joint)             # (B,T,U+1,J)
    d_enc, d_pred = dpre.sum(2), dpre.sum(1)
    x = np.asarray(enc_out, dtype=np.float64)
    g["enc_proj.weight"] = np.einsum("btj,btd->jd", d_enc, x)
    g["enc_proj.bias"] = d_enc.sum((0, 1))
    g["pred_proj.weight"] = np.einsum("buj,bue->je", d_pred, emb)
    g["pred_proj.bias"] = d_pred.sum((0, 1))
    d_emb = d_pred @ Pd["pred_proj.weight"]
    ge = np.zeros_like(Pd["embedding.weight"])
    np.add.at(ge, prefix, d_emb)
    g["embedding.weight"] = ge
    return float(nll.mean()), d_enc @ Pd["enc_proj.weight"], g
