#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference, which does not exist on the
GPU box):   python tests/golden/make_golden.py

* lucy_*.npz  — /root/reference/lucyrnn.py ``LucyRNN`` (kernel_impl="native") driven the
  way model.py:60-71 / train.py:460-580 drive it: two consecutive segments, state
  detached and carried, ``nn.CTCLoss(blank=0, zero_infinity=True)`` on
  ``log_softmax(-1).transpose(0,1)``, ``loss.backward()`` per segment.
* ctc_cases.npz — torch 2.11.0 CPU ``F.ctc_loss`` on the edge cases of SURVEY.md App. B.
* rnnt_cases.npz — ``torchaudio.functional.rnnt_loss`` (the substitute pin for the absent
  warp_rnnt; parity with the reference itself is UNPINNED).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, "/root/reference")

import lucyrnn as ref_lucyrnn            # noqa: E402  (the reference, unmodified)
from lucyrnn_conf import LucyRNNConfig    # noqa: E402
from oracle import lucy_oracle as LO      # noqa: E402

CASES = {
    # name: (cfg kwargs, B, T, seed)
    "train_fused_noln": (dict(is_training=True, fused_ops=True, layer_norm=False), 3, 17, 11),
    "train_fused_ln": (dict(is_training=True, fused_ops=True, layer_norm=True), 3, 17, 12),
    "train_unfused_ln": (dict(is_training=True, fused_ops=False, layer_norm=True), 3, 17, 13),
    "train_unfused_noln": (dict(is_training=True, fused_ops=False, layer_norm=False), 3, 17, 14),
    "step_fused_noln": (dict(is_training=False, fused_ops=True, layer_norm=False), 3, 17, 15),
    "step_fused_ln": (dict(is_training=False, fused_ops=True, layer_norm=True), 3, 17, 16),
    "step_unfused_ln": (dict(is_training=False, fused_ops=False, layer_norm=True), 3, 17, 17),
    "train_fused_ln_prefixsum": (dict(is_training=True, fused_ops=True, layer_norm=True,
                                      decay_mode="prefix_sum", lambda_decay=0.05), 3, 17, 18),
    "train_fused_noln_stack3": (dict(is_training=True, fused_ops=True, layer_norm=False,
                                     stack_order=3), 3, 17, 19),
    "train_fused_noln_nostate": (dict(is_training=True, fused_ops=True, layer_norm=False,
                                      return_last_states=False), 2, 9, 20),
}
SMALL = dict(input_dim=5, hidden_dim=8, num_layers=2, vocab_size=6)
MEDIUM = {"medium_train_fused_noln": (dict(input_dim=80, hidden_dim=32, num_layers=3, vocab_size=33,
                                           is_training=True, fused_ops=True, layer_norm=False), 2, 64, 31),
          "medium_train_fused_ln": (dict(input_dim=80, hidden_dim=32, num_layers=3, vocab_size=33,
                                         is_training=True, fused_ops=True, layer_norm=True), 2, 64, 32)}


def run_case(name, kw, B, T, seed):
    full = dict(SMALL)
    full.update(kw)
    cfg = LucyRNNConfig(kernel_impl="native", **full)
    P = LO.random_params(cfg, seed, dtype=torch.float32)
    model = ref_lucyrnn.LucyRNN(cfg)
    missing = model.load_state_dict(P, strict=True)
    g = torch.Generator().manual_seed(seed + 1000)
    V = cfg.vocab_size
    Tout = (T - T % cfg.stack_order) // cfg.stack_order
    out = {"cfg_" + k: np.asarray(v) for k, v in full.items()}
    out["B"], out["T"] = np.asarray(B), np.asarray(T)
    for k, v in P.items():
        out["param/" + k] = v.numpy()
    crit = torch.nn.CTCLoss(blank=0, zero_infinity=True)
    state = None
    for seg in range(2):
        x = torch.randn(B, T, cfg.input_dim, generator=g)
        Umax = max(1, Tout // 3)
        tgt_lens = [int(torch.randint(0, Umax + 1, (1,), generator=g)) for _ in range(B)]
        tgt_lens[0] = Umax
        in_lens = [Tout] * B
        if B > 1:
            in_lens[1] = max(1, Tout - 2)
        tokens = torch.randint(1, V, (B, Umax), generator=g)
        if state:                                       # model.py:60-61
            state = tuple([t.detach() for t in part] for part in state)
            out[f"seg{seg}/h_in"] = np.stack([t.numpy() for t in state[0]])
            out[f"seg{seg}/s_in"] = np.stack([t.numpy() for t in state[1]])
        model.zero_grad()
        res = model(x, state) if state is not None else model(x)
        if cfg.return_last_states:
            logits, new_state = res
        else:
            logits, new_state = res, None
        loss = crit(logits.log_softmax(-1).transpose(0, 1), tokens, in_lens, tgt_lens)
        loss.backward()
        out[f"seg{seg}/x"] = x.numpy()
        out[f"seg{seg}/tokens"] = tokens.numpy()
        out[f"seg{seg}/in_lens"] = np.asarray(in_lens)
        out[f"seg{seg}/tgt_lens"] = np.asarray(tgt_lens)
        out[f"seg{seg}/logits"] = logits.detach().numpy()
        out[f"seg{seg}/loss"] = loss.detach().numpy()
        if new_state is not None:
            out[f"seg{seg}/h_out"] = np.stack([t.detach().numpy() for t in new_state[0]])
            out[f"seg{seg}/s_out"] = np.stack([t.detach().numpy() for t in new_state[1]])
        for k, p in model.named_parameters():
            out[f"seg{seg}/grad/" + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
        state = new_state
    np.savez_compressed(os.path.join(HERE, f"lucy_{name}.npz"), **out)
    print("wrote", name, {k: v for k, v in full.items() if k not in SMALL or SMALL[k] != v})


def ctc_cases():
    g = torch.Generator().manual_seed(77)
    out = {}
    # (T, V, list of (in_len, labels))
    specs = {
        "basic": (12, 7, [(12, [1, 2, 3]), (10, [4, 4, 5]), (12, [6]), (7, [1, 2, 1, 2])]),
        "repeats_tight": (6, 5, [(5, [1, 1, 2]), (6, [3, 3, 3]), (4, [1, 2, 3, 4])]),
        "infeasible": (5, 5, [(3, [1, 1]), (2, [1, 2, 3]), (5, [2, 3])]),      # first two -> inf -> zeroed
        "empty_target": (8, 6, [(8, []), (5, []), (8, [1, 2])]),
        "zero_frames": (6, 6, [(0, []), (0, [1, 2]), (6, [3])]),
        "all_empty": (4, 5, [(4, []), (3, [])]),
        "long": (60, 33, [(60, list(range(1, 21))), (45, [5] * 10), (60, [1, 2] * 12), (1, [7])]),
    }
    for name, (T, V, utts) in specs.items():
        B = len(utts)
        Umax = max(len(u[1]) for u in utts)
        tokens = torch.zeros(B, Umax, dtype=torch.long)
        for b, (_, lab) in enumerate(utts):
            if lab:
                tokens[b, :len(lab)] = torch.tensor(lab)
        in_lens = [u[0] for u in utts]
        tgt_lens = [len(u[1]) for u in utts]
        logits = (torch.randn(B, T, V, generator=g) * 2).requires_grad_(True)
        logp = logits.log_softmax(-1).transpose(0, 1)
        loss = torch.nn.functional.ctc_loss(logp, tokens, in_lens, tgt_lens, blank=0,
                                            reduction="mean", zero_infinity=True)
        loss.backward()
        nll = torch.nn.functional.ctc_loss(logp.detach(), tokens, in_lens, tgt_lens, blank=0,
                                           reduction="none", zero_infinity=True)
        out[name + "/logits"] = logits.detach().numpy()
        out[name + "/tokens"] = tokens.numpy()
        out[name + "/in_lens"] = np.asarray(in_lens)
        out[name + "/tgt_lens"] = np.asarray(tgt_lens)
        out[name + "/loss"] = loss.detach().numpy()
        out[name + "/nll"] = nll.numpy()
        out[name + "/grad"] = logits.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "ctc_cases.npz"), **out)
    print("wrote ctc cases", list(specs))


def rnnt_cases():
    import torchaudio
    g = torch.Generator().manual_seed(99)
    out = {}
    specs = {"basic": (3, 9, 4, 6, [(9, 4), (7, 2), (9, 0)]),
             "single": (1, 1, 1, 5, [(1, 1)]),
             "wide": (2, 12, 7, 11, [(12, 7), (5, 3)])}
    for name, (B, T, U, V, lens) in specs.items():
        logits = torch.randn(B, T, U + 1, V, generator=g).requires_grad_(True)
        targets = torch.randint(1, V, (B, U), generator=g).int()
        fl = torch.tensor([l[0] for l in lens]).int()
        tl = torch.tensor([l[1] for l in lens]).int()
        nll = torchaudio.functional.rnnt_loss(logits, targets, fl, tl, blank=0, reduction="none",
                                              fused_log_softmax=True)
        nll.sum().backward()
        out[name + "/logits"] = logits.detach().numpy()
        out[name + "/targets"] = targets.numpy()
        out[name + "/frame_lens"] = fl.numpy()
        out[name + "/label_lens"] = tl.numpy()
        out[name + "/nll"] = nll.detach().numpy()
        out[name + "/grad"] = logits.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "rnnt_cases.npz"), **out)
    print("wrote rnnt cases", list(specs))


if __name__ == "__main__":
    torch.set_num_threads(4)
    for name, (kw, B, T, seed) in CASES.items():
        run_case(name, kw, B, T, seed)
    for name, (kw, B, T, seed) in MEDIUM.items():
        run_case(name, kw, B, T, seed)
    ctc_cases()
    rnnt_cases()
