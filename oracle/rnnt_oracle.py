"""RNN-T oracle — TEST INFRASTRUCTURE ONLY.  **PARITY UNPINNED** (see oracle/__init__.py).

The reference's transducer loss is ``warp_rnnt.RNNTLoss`` (train.py:39, 144) called as
``criterion(log_probs=..., labels=..., frames_lengths=..., labels_lengths=...,
blank_id=..., compact=..., gather=True)`` (model.py:97-105).  The package is absent from
/root/reference, requirements.txt and this image, and the call matches no published
warp-rnnt API, so there is nothing of the reference's to pin against.  This file restates
the published algorithm (Graves 2012, "Sequence Transduction with RNNs", eqs. 16-20;
SURVEY.md Appendix C) in numpy fp64; tests cross-check it against
``torchaudio.functional.rnnt_loss`` where torchaudio is importable.

Lattice for one utterance: nodes (t,u), 0<=t<T, 0<=u<=U.
  alpha(0,0)=0
  alpha(t,u)=lse(alpha(t-1,u)+blank(t-1,u), alpha(t,u-1)+label(t,u-1))
  ll = alpha(T-1,U) + blank(T-1,U)
where blank(t,u)=log_probs[t,u,blank], label(t,u)=log_probs[t,u,y_{u+1}].
"""
from __future__ import annotations

import numpy as np

NEG_INF = -np.inf


def _lse2(a, b):
    m = max(a, b)
    if m == NEG_INF:
        return NEG_INF
    return m + np.log(np.exp(a - m) + np.exp(b - m))


def rnnt_utterance(lp_blank, lp_label):
    """lp_blank [T,U+1], lp_label [T,U] (label(t,u) for u<U).  Returns nll, alpha, beta."""
    T, U1 = lp_blank.shape
    U = U1 - 1
    alpha = np.full((T, U1), NEG_INF)
    beta = np.full((T, U1), NEG_INF)
    alpha[0, 0] = 0.0
    for t in range(T):
        for u in range(U1):
            if t == 0 and u == 0:
                continue
            a = alpha[t - 1, u] + lp_blank[t - 1, u] if t > 0 else NEG_INF
            b = alpha[t, u - 1] + lp_label[t, u - 1] if u > 0 else NEG_INF
            alpha[t, u] = _lse2(a, b)
    ll = alpha[T - 1, U] + lp_blank[T - 1, U]
    beta[T - 1, U] = lp_blank[T - 1, U]
    for t in range(T - 1, -1, -1):
        for u in range(U, -1, -1):
            if t == T - 1 and u == U:
                continue
            a = beta[t + 1, u] + lp_blank[t, u] if t + 1 < T else NEG_INF
            b = beta[t, u + 1] + lp_label[t, u] if u < U else NEG_INF
            beta[t, u] = _lse2(a, b)
    return -ll, alpha, beta


def rnnt_loss_and_grad(log_probs, labels, frame_lens, label_lens, blank=0):
    """log_probs [B,T,U+1,V] (already normalised), labels [B,U].

    Returns nll[B] and d(sum_b nll_b)/d log_probs — non-zero only at the blank and
    label entries of each live lattice node (the 'gather=True' contract of the
    reference call, model.py:104).  Utterances with T_b==0 give nll=0, zero grad.
    """
    lp = np.asarray(log_probs, dtype=np.float64)
    B, T, U1, V = lp.shape
    nll = np.zeros(B)
    grad = np.zeros_like(lp)
    for b in range(B):
        Tb, Ub = int(frame_lens[b]), int(label_lens[b])
        if Tb <= 0:
            continue
        y = np.asarray(labels[b][:Ub], dtype=np.int64)
        lpb = lp[b, :Tb, :Ub + 1, blank]
        lpl = np.stack([lp[b, :Tb, u, y[u]] for u in range(Ub)], 1) if Ub > 0 else np.zeros((Tb, 0))
        n, alpha, beta = rnnt_utterance(lpb, lpl)
        nll[b] = n
        ll = -n
        for t in range(Tb):
            for u in range(Ub + 1):
                if alpha[t, u] == NEG_INF:
                    continue
                nb = beta[t + 1, u] if t + 1 < Tb else (0.0 if u == Ub else NEG_INF)
                if nb > NEG_INF:
                    grad[b, t, u, blank] -= np.exp(alpha[t, u] + lpb[t, u] + nb - ll)
                if u < Ub and beta[t, u + 1] > NEG_INF:
                    grad[b, t, u, y[u]] -= np.exp(alpha[t, u] + lpl[t, u] + beta[t, u + 1] - ll)
    return nll, grad


def rnnt_loss_and_grad_logits(logits, labels, frame_lens, label_lens, blank=0):
    """Same, from unnormalised joint logits, with the log-softmax backward folded in:
    dlogits[t,u,:] = softmax*(sum of node grads) ... i.e. g - softmax*sum_v(g)."""
    x = np.asarray(logits, dtype=np.float64)
    m = x.max(-1, keepdims=True)
    lse = m + np.log(np.exp(x - m).sum(-1, keepdims=True))
    lp = x - lse
    nll, g = rnnt_loss_and_grad(lp, labels, frame_lens, label_lens, blank)
    dx = g - np.exp(lp) * g.sum(-1, keepdims=True)
    return nll, dx


# The known-answer vector the transducer-loss packages publish in their own test suites (HawkAaron/warp-transducer
# `test.py` "small test", reused by 1ytic/warp-rnnt — the package the reference imports, train.py:39 — in
# `pytorch_binding/warp_rnnt/test.py` with `log_softmax(acts)` as input, and by torchaudio's `get_basic_data`):
# B=1, T=2, U=2, V=5, blank 0, labels [1, 2]; cost and d cost / d acts as printed there (written here from those
# published tests — no copy of either package is in this image; torchaudio 2.11 reproduces every digit).
WARP_KAT_ACTS = [[[[0.1, 0.6, 0.1, 0.1, 0.1], [0.1, 0.1, 0.6, 0.1, 0.1], [0.1, 0.1, 0.2, 0.8, 0.1]],
                  [[0.1, 0.6, 0.1, 0.1, 0.1], [0.1, 0.1, 0.2, 0.1, 0.1], [0.7, 0.1, 0.2, 0.1, 0.1]]]]
WARP_KAT_LABELS = [[1, 2]]
WARP_KAT_COST = 4.495666
WARP_KAT_GRADS = [[[[-0.13116688, -0.3999269, 0.17703125, 0.17703125, 0.17703125],
                    [-0.18572757, 0.12247056, -0.18168412, 0.12247056, 0.12247056],
                    [-0.32091254, 0.06269141, 0.06928472, 0.12624499, 0.06269141]],
                   [[0.05456069, -0.21824276, 0.05456069, 0.05456069, 0.05456069],
                    [0.12073959, 0.12073959, -0.48295835, 0.12073959, 0.12073959],
                    [-0.6925882, 0.16871116, 0.18645467, 0.16871116, 0.16871116]]]]
