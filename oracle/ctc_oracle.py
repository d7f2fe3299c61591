"""CTC oracle — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy fp64 restatement of what the reference obtains from
``torch.nn.CTCLoss(blank=0, zero_infinity=True)`` applied to
``enc_out.log_softmax(-1).transpose(0,1)`` (model.py:70-71, train.py:142):
per-utterance negative log-likelihood over the 2U+1 extended-label lattice,
'mean' reduction = mean_b(nll_b / max(U_b,1)), and the gradient with respect to
the *logits* (log-softmax backward folded in).  Semantics pinned in SURVEY.md
Appendix B; parity is pinned by tests/golden (torch 2.11.0 CPU F.ctc_loss).
"""
from __future__ import annotations

import numpy as np

NEG_INF = -np.inf


def _lse(*xs):
    m = max(xs)
    if m == NEG_INF:
        return NEG_INF
    return m + np.log(sum(np.exp(x - m) for x in xs))


def extended_labels(y, blank=0):
    """l' = [blank, y1, blank, y2, ..., blank]  (length 2U+1)."""
    ext = np.full(2 * len(y) + 1, blank, dtype=np.int64)
    ext[1::2] = y
    return ext


def ctc_utterance(logp, y, blank=0):
    """logp [T,V] normalised log-probs, y [U] labels.  Returns nll, alpha[T,S], beta[T,S]."""
    T = logp.shape[0]
    ext = extended_labels(y, blank)
    S = len(ext)
    alpha = np.full((max(T, 1), S), NEG_INF)
    beta = np.full((max(T, 1), S), NEG_INF)
    if T == 0:
        return (0.0 if len(y) == 0 else np.inf), alpha, beta
    alpha[0, 0] = logp[0, ext[0]]
    if S > 1:
        alpha[0, 1] = logp[0, ext[1]]
    for t in range(1, T):
        for s in range(S):
            a = alpha[t - 1, s]
            b = alpha[t - 1, s - 1] if s >= 1 else NEG_INF
            c = alpha[t - 1, s - 2] if (s >= 2 and ext[s] != blank and ext[s] != ext[s - 2]) else NEG_INF
            alpha[t, s] = _lse(a, b, c) + logp[t, ext[s]]
    ll = _lse(alpha[T - 1, S - 1], alpha[T - 1, S - 2] if S > 1 else NEG_INF)
    beta[T - 1, S - 1] = logp[T - 1, ext[S - 1]]
    if S > 1:
        beta[T - 1, S - 2] = logp[T - 1, ext[S - 2]]
    for t in range(T - 2, -1, -1):
        for s in range(S):
            a = beta[t + 1, s]
            b = beta[t + 1, s + 1] if s + 1 < S else NEG_INF
            c = beta[t + 1, s + 2] if (s + 2 < S and ext[s] != blank and ext[s] != ext[s + 2]) else NEG_INF
            beta[t, s] = _lse(a, b, c) + logp[t, ext[s]]
    return -ll, alpha, beta


def ctc_loss_and_grad(logits, targets, in_lens, tgt_lens, blank=0, reduction="mean",
                      zero_infinity=True):
    """logits [B,T,V] (unnormalised), targets [B,Umax] ints, lengths as sequences.

    Returns (loss, nll[B], dlogits[B,T,V]) with dlogits = d loss / d logits:
    scale_b * (softmax - occupancy) for t < T_b and exactly 0 beyond, rows of an
    infeasible utterance all-zero (zero_infinity).  scale_b = 1/(B*max(U_b,1)) for
    'mean', 1 for 'sum'/'none' (for 'none' the grad is that of sum_b nll_b).
    """
    logits = np.asarray(logits, dtype=np.float64)
    B, T, V = logits.shape
    m = logits.max(-1, keepdims=True) if V else logits
    lse = m + np.log(np.exp(logits - m).sum(-1, keepdims=True))
    logp = logits - lse
    nll = np.zeros(B)
    grad = np.zeros_like(logits)
    for b in range(B):
        Tb, Ub = int(in_lens[b]), int(tgt_lens[b])
        y = np.asarray(targets[b][:Ub], dtype=np.int64) if Ub > 0 else np.zeros(0, np.int64)
        n, alpha, beta = ctc_utterance(logp[b, :Tb], y, blank)
        if not np.isfinite(n):
            nll[b] = 0.0 if zero_infinity else np.inf
            continue
        nll[b] = n
        ext = extended_labels(y, blank)
        scale = 1.0 / (B * max(Ub, 1)) if reduction == "mean" else 1.0
        for t in range(Tb):
            occ = np.zeros(V)
            ab = alpha[t] + beta[t]
            for s in range(len(ext)):
                if ab[s] > NEG_INF:
                    occ[ext[s]] += np.exp(ab[s] + n - logp[b, t, ext[s]])
            grad[b, t] = scale * (np.exp(logp[b, t]) - occ)
    if reduction == "mean":
        loss = float(np.mean(nll / np.maximum(np.asarray(tgt_lens, dtype=np.float64)[:B], 1.0)))
    elif reduction == "sum":
        loss = float(nll.sum())
    else:
        loss = nll.copy()
    return loss, nll, grad
