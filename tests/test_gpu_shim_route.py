"""GPU: the shim route end to end.  The criterion is built the way train.py:142 builds it —
`torch.nn.CTCLoss(blank=0, zero_infinity=True)` — after importing a shim module, the loss is computed by the
reference's call shape `criterion(enc_out.log_softmax(-1).transpose(0, 1), tokens, in_lens, tgt_lens)`
(model.py:70-71), and everything is compared with tests/golden/glue_cases.npz, which the reference's own
`compute_loss` + `ASRModel` produced.  Where the reference tree is mounted (authoring container) the
driver of the three carried segments is the reference's OWN model.compute_loss / ASRModel imported through
the shims; on the GPU box (no /root/reference) it is the package's mirror of the same functions."""
import importlib

import numpy as np
import pytest
import torch

from conftest import load_golden
from shim_env import shim_imports

pytestmark = pytest.mark.gpu


def _sub(G, prefix):
    return {k[len(prefix):]: v for k, v in G.items() if k.startswith(prefix)}


def test_train_py_call_shape_reaches_k3_and_matches_the_reference(cuda_device, monkeypatch):
    import statecatcher_b200 as sb
    from statecatcher_b200 import _lib
    monkeypatch.delenv("SC_SHIM_CTC", raising=False)
    C = _sub(load_golden("glue_cases"), "ctc_plain/")
    cfg_kw = {k[4:]: v.item() for k, v in C.items() if k.startswith("cfg_")}
    F = C["seg0/feats"].shape[-1]
    with shim_imports() as use_ref:
        if use_ref:
            model_py = importlib.import_module("model")              # /root/reference/model.py through the shims
            cfg = model_py.LucyRNNConfig(kernel_impl="triton", **cfg_kw)
            asr = model_py.ASRModel(None, cfg, cfg.vocab_size, F, -1, debug=False).cuda()
            compute_loss = model_py.compute_loss
        else:
            importlib.import_module("lucyrnn")                        # what model.py:7 does; installs the route
            cfg = sb.LucyRNNConfig(kernel_impl="triton", **cfg_kw)
            asr = sb.LucyASRModel(cfg, frontend=None, feat_dim=F, proj_dim=-1).cuda()
            compute_loss = sb.compute_loss
        assert isinstance(asr.encoder, sb.LucyRNN)
        asr.encoder.load_state_dict({k: torch.tensor(v) for k, v in _sub(C, "param/").items()}, strict=True)
        crit = torch.nn.CTCLoss(blank=0, zero_infinity=True)          # train.py:142, verbatim
        assert getattr(type(crit), "_statecatcher_b200_routed", False) and not isinstance(crit, sb.CTCLoss)
        state = None
        for seg in range(3):
            S = _sub(C, f"seg{seg}/")
            asr.zero_grad(set_to_none=True)
            calls = []
            orig_call = _lib.call
            monkeypatch.setattr(sb.ctc, "call", lambda name, *a: (calls.append(name), orig_call(name, *a))[1])
            loss, new_state, enc_out, _ = compute_loss(
                "ctc", crit, asr, torch.tensor(S["feats"]).cuda(), torch.tensor(S["mask"]).cuda(),
                torch.tensor(S["tokens"]).cuda(), S["in_lens"].tolist(), S["tgt_lens"].tolist(), blank_id=0, input_state=state)
            loss.backward()
            monkeypatch.setattr(sb.ctc, "call", orig_call)
            assert calls == ["sc_ctc_emissions", "sc_ctc_lattice", "sc_ctc_bwd"], calls     # K3, not ATen's ctc_loss_gpu
            want = S["enc_out"]
            np.testing.assert_allclose(enc_out.detach().cpu().numpy(), want, rtol=1e-4, atol=2e-5 * max(1.0, np.abs(want).max()))
            np.testing.assert_allclose(loss.item(), S["loss"], rtol=1e-4)
            np.testing.assert_allclose(torch.stack(new_state[0]).cpu().numpy(), S["h_out"], rtol=1e-4, atol=2e-5)
            for k, p in asr.named_parameters():
                w = S["grad/" + k]
                got = p.grad.cpu().numpy() if p.grad is not None else np.zeros_like(w)
                assert np.abs(got - w).max() <= 2e-4 * max(1e-3, np.abs(w).max()), (seg, k)
            state = new_state


def test_fp16_autocast_with_grad_scaler_flow(cuda_device):
    """train.py:515-526, 549-566 (`--use-scaler`): forward under `autocast(float16)`, `scaler.scale(loss).backward()`,
    `scaler.unscale_`, clip, `scaler.step`, `scaler.update`.  The module serves fp16 autocast with its bf16 path
    (fp32 master weights, fp32 accumulation and state); scaling by 2^16 is exact in bf16/fp32, so the unscaled
    gradients must equal the unscaled-run gradients of the same bf16 path to rounding, the optimizer must step,
    and an overflowing step (inf injected) must be SKIPPED with the scale halved."""
    import statecatcher_b200 as sb
    from torch.amp import GradScaler, autocast
    torch.manual_seed(0)
    cfg = sb.LucyRNNConfig(input_dim=80, hidden_dim=128, num_layers=2, vocab_size=64, fused_ops=True, layer_norm=False)
    model = sb.LucyRNN(cfg).cuda()
    torch.nn.init.normal_(model.output_proj.weight, std=0.05)
    crit = sb.CTCLoss(blank=0, zero_infinity=True)
    g = torch.Generator().manual_seed(1)
    B, T = 4, 64
    x = torch.randn(B, T, 80, generator=g).cuda()
    tok = torch.randint(1, 64, (B, 8), generator=g).cuda()
    inl, tgl = [T, T, 50, T], [8, 5, 3, 0]

    def run(scaler):
        model.zero_grad(set_to_none=True)
        with autocast(device_type="cuda", dtype=torch.float16):
            logits, _ = model(x)
            assert logits.dtype == torch.bfloat16                      # fp16 autocast -> the bf16 kernels
            loss = crit(logits.log_softmax(-1).transpose(0, 1), tok, inl, tgl)     # model.py:70-71
        if scaler is None:
            loss.backward()
        else:
            scaler.scale(loss).backward()
        return loss.detach()

    base_loss = run(None)
    base = {k: p.grad.clone() for k, p in model.named_parameters()}
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    scaler = GradScaler("cuda", init_scale=2.0 ** 16)
    loss = run(scaler)
    assert torch.equal(loss, base_loss)
    scaler.unscale_(opt)                                                # train.py:551
    for k, p in model.named_parameters():
        ref = base[k]
        assert torch.isfinite(p.grad).all(), k
        err = (p.grad - ref).abs().max().item()
        assert err <= 2e-2 * max(ref.abs().max().item(), 1e-6), (k, err)    # bf16 dlogits/dG are rounded at a different scale
    torch.nn.utils.clip_grad_norm_(model.parameters(), 50.0)            # train.py:553
    before = {k: p.detach().clone() for k, p in model.named_parameters()}
    scaler.step(opt)                                                     # train.py:563
    scaler.update()
    assert scaler.get_scale() == 2.0 ** 16
    moved = [k for k, p in model.named_parameters() if not torch.equal(p, before[k])]
    assert "output_proj.weight" in moved and "layers.0.W_fused.weight" in moved
    # overflow: the step is skipped and the scale backs off
    run(scaler)
    model.output_proj.weight.grad[0, 0] = float("inf")
    before = {k: p.detach().clone() for k, p in model.named_parameters()}
    scaler.unscale_(opt)
    scaler.step(opt)
    scaler.update()
    assert all(torch.equal(p, before[k]) for k, p in model.named_parameters())
    assert scaler.get_scale() == 2.0 ** 15
