"""CPU: host logic of optim.FusedAdam / optim.Lion with the C-ABI call replaced by a recorder — which
entry points are called, how many times, with which pointer tables and scalars.  (The arithmetic is
checked on the GPU in tests/test_gpu_optim.py and tests/test_gpu_zz_optim_ext.py.)"""
import ctypes

import pytest
import torch


@pytest.fixture
def recorder(monkeypatch):
    from statecatcher_b200 import optim, _lib
    calls = []
    monkeypatch.setattr(optim, "call", lambda name, *a: calls.append((name, a)))
    monkeypatch.setattr(optim, "stream", lambda: 0)
    monkeypatch.setattr(_lib, "require_cuda", lambda t, name: None)
    return calls


def _params(n=5):
    ps = [torch.zeros(s, requires_grad=True) for s in [(3, 4), (7,), (0,), (2, 2), (9,)][:n]]
    for p in ps:
        p.grad = torch.ones_like(p)
    return ps


def _table(arg):
    return [arg[i] or 0 for i in range(len(arg))]


@pytest.mark.parametrize("cls", ["adam", "lion"])
def test_per_tensor_calls(recorder, cls):
    from statecatcher_b200.optim import FusedAdam, Lion
    ps = _params()
    ps[3].grad = None                                      # a parameter without gradient is skipped
    opt = (FusedAdam(ps, max_grad_norm=50.0, multi_tensor=False) if cls == "adam"
           else Lion(ps, max_grad_norm=50.0, multi_tensor=False))
    opt.step()
    names = [c[0] for c in recorder]
    assert names == ["sc_sumsq_accum"] * 4 + ["sc_adam_step" if cls == "adam" else "sc_lion_step"] * 4
    step_calls = [c for c in recorder if c[0].endswith("_step")]
    assert [c[1][0] for c in step_calls] == [p.data_ptr() for p in ps if p.grad is not None]
    assert opt.grad_norm is not None and opt.grad_norm.dim() == 0
    if cls == "adam":
        assert all(c[1][10] == 1 for c in step_calls)      # 1-based step count
        recorder.clear()
        opt.step()
        assert all(c[1][10] == 2 for c in recorder if c[0] == "sc_adam_step")


@pytest.mark.parametrize("cls", ["adam", "lion"])
def test_multi_tensor_calls(recorder, cls):
    from statecatcher_b200.optim import FusedAdam, Lion
    ps = _params()
    ps[3].grad = None
    live = [p for p in ps if p.grad is not None]
    opt = (FusedAdam(ps, lr=2e-3, max_grad_norm=50.0, multi_tensor=True) if cls == "adam"
           else Lion(ps, lr=2e-3, max_grad_norm=50.0, multi_tensor=True))
    opt.step()
    assert [c[0] for c in recorder] == ["sc_sumsq_accum_multi", f"sc_{cls}_step_multi"]
    (_, a_sum), (_, a_step) = recorder
    assert _table(a_sum[0]) == [p.grad.data_ptr() for p in live] and list(a_sum[1]) == [p.numel() for p in live]
    assert a_sum[2] == 4
    assert _table(a_step[0]) == [p.data_ptr() for p in live]
    assert _table(a_step[1]) == [p.grad.data_ptr() for p in live]
    assert _table(a_step[2]) == [opt.state[p]["exp_avg"].data_ptr() for p in live]
    if cls == "adam":
        assert _table(a_step[3]) == [opt.state[p]["exp_avg_sq"].data_ptr() for p in live]
        assert list(a_step[4]) == [p.numel() for p in live] and a_step[5] == 4
        assert a_step[6] == pytest.approx(2e-3) and a_step[11] == 1 and a_step[14] == 1
        assert a_step[12] == a_sum[3] != 0 and a_step[13] == 50.0   # the norm accumulator feeds the clip
    else:
        assert list(a_step[3]) == [p.numel() for p in live] and a_step[4] == 4
        assert a_step[5] == pytest.approx(2e-3) and (a_step[6], a_step[7]) == (0.9, 0.99)
        assert a_step[9] == a_sum[3] != 0 and a_step[10] == 50.0
    assert isinstance(a_step[0], ctypes.Array)


def test_multi_tensor_adam_groups_by_step_count(recorder):
    """A parameter that joins later (its first gradient arrives at step 3) has its own bias correction."""
    from statecatcher_b200.optim import FusedAdam
    ps = _params(2)
    ps[1].grad = None
    opt = FusedAdam(ps, multi_tensor=True)
    opt.step(); opt.step()
    ps[1].grad = torch.ones_like(ps[1])
    recorder.clear()
    opt.step()
    steps = sorted((c[1][11], c[1][5]) for c in recorder if c[0] == "sc_adam_step_multi")
    assert steps == [(1, 1), (3, 1)]
    assert not any(c[0].startswith("sc_sumsq") for c in recorder)          # no clipping requested


def test_param_groups_keep_their_hyperparameters(recorder):
    from statecatcher_b200.optim import Lion
    a, b = _params(2)
    opt = Lion([{"params": [a], "lr": 1e-3}, {"params": [b], "weight_decay": 0.5}], lr=1e-4, multi_tensor=True)
    opt.step()
    (_, x), (_, y) = recorder
    assert x[5] == pytest.approx(1e-3) and x[8] == 0.0
    assert y[5] == pytest.approx(1e-4) and y[8] == 0.5


def test_rejects_what_the_kernels_do_not_take(recorder):
    from statecatcher_b200.optim import FusedAdam, Lion
    p = torch.zeros(4, dtype=torch.float64, requires_grad=True)
    p.grad = torch.zeros_like(p)
    for opt in (FusedAdam([p]), Lion([p])):
        with pytest.raises(TypeError):
            opt.step()
    q = torch.zeros(4, 4, requires_grad=True)
    q.grad = torch.zeros(4, 4).t().contiguous().t()           # non-contiguous gradient
    with pytest.raises(TypeError):
        Lion([q]).step()
    with pytest.raises(ValueError):
        Lion([q], betas=(0.9, 1.5))


@pytest.mark.parametrize("cls", ["adam", "lion"])
def test_gradscaler_drives_the_fused_optimizers(recorder, cls):
    """train.py:513-566 under --use-scaler: scale -> backward -> unscale_ -> clip -> scaler.step(optimizer).
    The scaler steps the optimizer when no gradient overflowed and skips it otherwise."""
    from statecatcher_b200.optim import FusedAdam, Lion
    p = torch.ones(6, requires_grad=True)
    opt = FusedAdam([p], max_grad_norm=50.0) if cls == "adam" else Lion([p], max_grad_norm=50.0)
    scaler = torch.amp.GradScaler("cpu", init_scale=1024.0)
    scaler.scale((p * 2).sum()).backward()
    assert torch.equal(p.grad, torch.full((6,), 2048.0))
    scaler.unscale_(opt)
    assert torch.equal(p.grad, torch.full((6,), 2.0))
    scaler.step(opt)
    scaler.update()
    assert [c[0] for c in recorder][-1] == ("sc_adam_step_multi" if cls == "adam" else "sc_lion_step_multi")   # multi-tensor is the default
    recorder.clear()
    opt.zero_grad()
    scaler.scale((p * float("inf")).sum()).backward()
    scaler.unscale_(opt)
    scaler.step(opt)                                         # overflow: the step is skipped
    scaler.update()
    assert not recorder and scaler.get_scale() < 1024.0
