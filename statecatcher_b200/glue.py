"""Host glue around the hot path, mirroring the reference's model.py.

* ``detach_states`` / ``assert_all_detached``  — model.py:11-35 (bit-exact state handoff:
  tensors are detached, never copied or re-typed).
* ``compute_loss``  — model.py:37-110: detach carried state, run the model, apply the loss
  head, return the reference's 4-tuple ``(loss, state, enc_out, state)``.
* ``LucyASRModel``  — the LucyRNN branch of ``ASRModel`` (model.py:282-398): optional input
  projection, zero-masking of padded frames (model.py:376-377), positional encoder call
  (model.py:384/388).  The LSTM/xLSTM branches are other encoders and out of scope.
"""
from __future__ import annotations

from typing import Any, Optional

import torch
import torch.nn as nn

from .ctc import CTCLoss, ctc_loss_from_logits
from .lucyrnn import LucyRNN
from .lucyrnn_conf import LucyRNNConfig


def detach_states(states):
    if states is None:
        return None
    if isinstance(states, torch.Tensor):
        return states.detach()
    if isinstance(states, dict):
        return {k: detach_states(v) for k, v in states.items()}
    if isinstance(states, tuple):
        return tuple(detach_states(v) for v in states)
    if isinstance(states, list):
        return [detach_states(v) for v in states]
    return states


def assert_all_detached(x):
    if isinstance(x, torch.Tensor):
        assert not x.requires_grad, "Tensor still requires grad"
    elif isinstance(x, (list, tuple)):
        for v in x:
            assert_all_detached(v)
    elif isinstance(x, dict):
        for v in x.values():
            assert_all_detached(v)


class LucyASRModel(nn.Module):
    def __init__(self, cfg: LucyRNNConfig, frontend: Optional[nn.Module] = None, feat_dim: int = 80,
                 proj_dim: int = -1, compute_dtype=None):
        super().__init__()
        self.frontend = frontend
        self.cfg = cfg                      # train.py:484 reads model.cfg.stack_order
        self.encoder = LucyRNN(cfg, compute_dtype=compute_dtype)
        self.enc_out_dim = cfg.vocab_size
        if proj_dim > 0:
            self.proj = nn.Linear(feat_dim, proj_dim)

    def forward(self, feats, mask, states=None):
        if hasattr(self, "proj"):
            feats = self.proj(feats)
        if mask is not None:
            feats = feats * mask.unsqueeze(-1).float()
        if states is not None:
            return self.encoder(feats, states)
        return self.encoder(feats)


def compute_loss(mode: str, criterion, model: nn.Module, feats, masks, tokens, in_lens, tgt_lens,
                 blank_id: int, use_rnnt_joiner: Optional[nn.Module] = None,
                 input_state: Optional[Any] = None, args=None, compact=False):
    if input_state:                                          # truthiness gate, model.py:60
        input_state = detach_states(input_state)
        if args and getattr(args, "debug", False):
            assert_all_detached(input_state)
    enc_out, output_state = model(feats, masks, input_state)
    if mode == "ctc":
        if isinstance(criterion, CTCLoss):
            # fused head: log-softmax folded into the CTC kernels, (B,T,V) read in place
            loss = ctc_loss_from_logits(enc_out, tokens, in_lens, tgt_lens, criterion.blank,
                                        criterion.reduction, criterion.zero_infinity)
        else:
            loss = criterion(enc_out.log_softmax(-1).transpose(0, 1), tokens, in_lens, tgt_lens)
    elif mode == "rnnt":
        assert use_rnnt_joiner is not None, "Joiner module required for RNN-T mode"
        prefix = torch.cat([torch.full((tokens.size(0), 1), blank_id, dtype=tokens.dtype,
                                       device=tokens.device), tokens], dim=1)
        if args is not None and getattr(args, "compact_rnnt", False):
            logits = use_rnnt_joiner(enc_out, prefix, in_lens, tgt_lens)
        else:
            logits = use_rnnt_joiner(enc_out, prefix)
        log_probs = logits.float().log_softmax(dim=-1)
        loss = criterion(log_probs=log_probs, labels=tokens, frames_lengths=in_lens,
                         labels_lengths=tgt_lens, blank_id=blank_id, compact=compact, gather=True)
    else:
        raise ValueError(f"Unknown mode: {mode}")
    return loss, output_state, enc_out, output_state
