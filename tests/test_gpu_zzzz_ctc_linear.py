"""GPU: the EXPERIMENTAL linear-domain CTC lattice kernel (SC_CTC_WAVE=3, opt-in) against the default kernel and
torch's fp64 ctc_loss.  The kernel was written after round 1's GPU budget was spent and has never run, so this
file sorts after every other GPU test AND only runs when SC_RUN_EXPERIMENTAL=1 is set: a kernel that has never
run must not be able to hang or poison the default `pytest -m gpu` run.
    SC_RUN_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_zzzz_ctc_linear.py -q"""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("SC_RUN_EXPERIMENTAL") != "1",
                                 reason="experimental kernel (never run on a GPU yet): set SC_RUN_EXPERIMENTAL=1")]


@pytest.mark.parametrize("shape", [(5, 700, 64, 150), (3, 50, 11, 6), (2, 3000, 1024, 150)], ids=["cfg2like", "small", "cfg2row"])
def test_linear_domain_lattice_kernel_agrees(cuda_device, monkeypatch, shape):
    """SC_CTC_WAVE=3 and 4 (EXPERIMENTAL: linear-domain recursion with a per-node exponent, opt-in) against the default
    kernel and torch's fp64 ctc_loss: loss, gradient, ragged lengths, an empty transcript, a one-frame
    utterance, an infeasible one, repeated labels, and peaky logits (the input that breaks a column-scaled
    linear recursion)."""
    from statecatcher_b200 import ctc_loss_from_logits
    B, T, V, U = shape
    g = torch.Generator().manual_seed(B * 1000 + T)
    for scale in (2.0, 6.0):
        logits = torch.randn(B, T, V, generator=g) * scale
        tokens = torch.randint(1, V, (B, U), generator=g)
        tokens[0, 1] = tokens[0, 0]                               # a repeat: no skip transition there
        inl = [T] * B
        tgl = [U] * B
        inl[1], tgl[1] = max(T - 50, 1), max(U // 2, 1)
        if B >= 3:
            inl[2], tgl[2] = T, 0                                 # empty transcript
        if B >= 4:
            inl[3], tgl[3] = 1, 1                                 # one frame, one label
        if B >= 5:
            inl[4], tgl[4] = max(U // 2, 1), U                    # fewer frames than labels: infeasible -> 0 loss, 0 grad
        xd = logits.double().requires_grad_(True)
        ref = torch.nn.functional.ctc_loss(xd.log_softmax(-1).transpose(0, 1), tokens, inl, tgl, zero_infinity=True)
        ref.backward()
        out = {}
        for wave in ("2", "3", "4"):
            monkeypatch.setenv("SC_CTC_WAVE", wave)
            x = logits.cuda().requires_grad_(True)
            loss = ctc_loss_from_logits(x, tokens.cuda(), inl, tgl, zero_infinity=True)
            loss.backward()
            out[wave] = (loss.item(), x.grad.cpu())
        monkeypatch.delenv("SC_CTC_WAVE")
        for wave in ("3", "4"):                               # 3: block-barrier structure, 4: pair-per-thread wavefront
            np.testing.assert_allclose(out[wave][0], ref.item(), rtol=1e-4, err_msg=wave)
            np.testing.assert_allclose(out[wave][1].numpy(), xd.grad.numpy(), rtol=2e-3, atol=2e-7, err_msg=wave)
            assert (out[wave][1] - out["2"][1]).abs().max().item() < 5e-6, wave
