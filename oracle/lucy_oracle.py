"""LucyRNN oracle — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

A functional (weights-in-a-dict) restatement, in plain torch on the CPU, of the
reference encoder ``/root/reference/lucyrnn.py``.  Two evaluation strategies
are provided for the same function:

* ``forward_looped``  walks time exactly the way the reference does (decay scan
  loop then one cell evaluation per timestep for the training path,
  lucyrnn.py:109-170; nested t/l loop for the step path, lucyrnn.py:172-184).
  It is the strategy timed as the CPU baseline, because it is the reference's
  own CPU algorithm (including the per-timestep re-projection).
* ``forward_closed``  is the algebraic form: time-parallel
  projections followed by two diagonal first-order linear scans.  It is what
  the CUDA kernels implement, so it doubles as their line-by-line spec.

Both run in any float dtype (fp64 = the tight truth for kernel unit tests) and
both are differentiable by autograd.  ``backward_closed`` is the hand-derived
reverse-time adjoint (Appendix A.3) in the same notation as the CUDA backward.

Weights use the reference's state_dict keys (lucyrnn.py:15-30, 85) so the same
dict loads into the reference module, the oracle and the B200 module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

LN_EPS = 1e-5  # torch.nn.LayerNorm default, used by lucyrnn.py:17-20
GATES = ("r", "z", "k", "v", "h", "decay")  # chunk order of W_fused, lucyrnn.py:49


@dataclass
class OracleConfig:
    """Field-for-field mirror of lucyrnn_conf.py:3-16."""
    input_dim: int
    hidden_dim: int
    num_layers: int
    vocab_size: int
    return_last_states: bool = True
    kernel_impl: str = "native"
    is_training: bool = True
    fused_ops: bool = False
    layer_norm: bool = True
    stack_order: int = 1
    decay_mode: str = "learned"
    lambda_decay: float = 0.001


# --------------------------------------------------------------------------- #
# parameters
# --------------------------------------------------------------------------- #
def param_shapes(cfg) -> Dict[str, Tuple[int, ...]]:
    """state_dict keys and shapes of the reference module (lucyrnn.py:9-31, 80-87)."""
    H = cfg.hidden_dim
    out: Dict[str, Tuple[int, ...]] = {}
    for l in range(cfg.num_layers):
        fin = cfg.input_dim * cfg.stack_order if l == 0 else H
        pre = f"layers.{l}."
        out[pre + "input_proj.weight"] = (H, fin)
        out[pre + "input_proj.bias"] = (H,)
        if cfg.layer_norm:
            for nm in ("in", "r", "z", "h"):
                out[pre + f"layernorm_{nm}.weight"] = (H,)
                out[pre + f"layernorm_{nm}.bias"] = (H,)
        if cfg.fused_ops:
            out[pre + "W_fused.weight"] = (6 * H, H)
            out[pre + "W_fused.bias"] = (6 * H,)
        else:
            for g in GATES:
                out[pre + f"W_{g}.weight"] = (H, H)
                out[pre + f"W_{g}.bias"] = (H,)
    out["output_proj.weight"] = (cfg.vocab_size, H)
    out["output_proj.bias"] = (cfg.vocab_size,)
    return out


def random_params(cfg, seed: int, dtype=torch.float64, scale: float = 0.5,
                  perturb_ln: bool = True) -> Dict[str, torch.Tensor]:
    """Seeded *generic* weights (not the reference init): every tensor non-trivial,
    including LN affine terms and output_proj, so no term of the function is hidden
    by a zero/identity weight."""
    g = torch.Generator().manual_seed(seed)
    P = {}
    for name, shp in param_shapes(cfg).items():
        if "layernorm" in name:
            if name.endswith("weight"):
                t = 1.0 + (0.2 * torch.randn(shp, generator=g, dtype=torch.float64) if perturb_ln else 0.0)
            else:
                t = 0.1 * torch.randn(shp, generator=g, dtype=torch.float64) if perturb_ln else torch.zeros(shp, dtype=torch.float64)
        elif name.endswith("weight"):
            fan_in = shp[1]
            t = torch.randn(shp, generator=g, dtype=torch.float64) * (scale / math.sqrt(fan_in) * 2.0)
        else:
            t = 0.1 * torch.randn(shp, generator=g, dtype=torch.float64)
        P[name] = t.to(dtype)
    return P


def reference_init_params(cfg, seed: int, out_std: float = 0.02, dtype=torch.float32):
    """The reference's init (orthogonal on >=2-D weights, LN w=1 b=0: lucyrnn.py:34-42;
    nn.Linear default bias init) except output_proj.weight ~ N(0,out_std) instead of the
    zero init of lucyrnn.py:86-87, which would make all logits equal."""
    g = torch.Generator().manual_seed(seed)
    P = {}
    for name, shp in param_shapes(cfg).items():
        if "layernorm" in name:
            t = torch.ones(shp) if name.endswith("weight") else torch.zeros(shp)
        elif name == "output_proj.weight":
            t = torch.randn(shp, generator=g) * out_std
        elif name == "output_proj.bias":
            t = torch.zeros(shp)
        elif name.endswith("weight"):
            t = torch.empty(shp)
            torch.nn.init.orthogonal_(t, generator=g)
        else:
            fan_in = param_shapes(cfg)[name.replace("bias", "weight")][1]
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shp, generator=g) * 2 - 1) * bound
        P[name] = t.to(dtype)
    return P


# --------------------------------------------------------------------------- #
# building blocks
# --------------------------------------------------------------------------- #
def _ln(x, P, key, cfg):
    """nn.LayerNorm(H) or Identity (lucyrnn.py:17-20)."""
    if not cfg.layer_norm:
        return x
    w, b = P[key + ".weight"], P[key + ".bias"]
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + LN_EPS) * w + b


def _lin(x, P, key):
    return x @ P[key + ".weight"].transpose(-1, -2) + P[key + ".bias"]


def _gate(u, P, pre, cfg, name):
    """One named gate pre-activation from u (fused chunk or separate Linear)."""
    if cfg.fused_ops:
        H = cfg.hidden_dim
        i = GATES.index(name)
        w = P[pre + "W_fused.weight"][i * H:(i + 1) * H]
        b = P[pre + "W_fused.bias"][i * H:(i + 1) * H]
        return u @ w.transpose(-1, -2) + b
    return _lin(u, P, pre + f"W_{name}")


def _stack_frames(x, cfg):
    """Frame stacking, lucyrnn.py:92-99."""
    k = cfg.stack_order
    if k > 1:
        B, T, F = x.shape
        T2 = T - (T % k)
        x = x[:, :T2, :].reshape(B, T2 // k, F * k)
    return x


def _init_state(cfg, B, dtype, state):
    """lucyrnn.py:101-107.  Returns fresh Python lists (the caller's lists are not mutated)."""
    if state is None:
        h = [torch.zeros(B, cfg.hidden_dim, dtype=dtype) for _ in range(cfg.num_layers)]
        s = [torch.zeros(B, cfg.hidden_dim, dtype=dtype) for _ in range(cfg.num_layers)]
    else:
        h = [t.to(dtype) for t in state[0]]
        s = [t.to(dtype) for t in state[1]]
    return h, s


def _cell(P, pre, cfg, x_t, h_prev, s_prev):
    """One LucyRNNCell evaluation, lucyrnn.py:44-70 (mask=None)."""
    u = _ln(_lin(x_t, P, pre + "input_proj"), P, pre + "layernorm_in", cfg)
    z = torch.sigmoid(_ln(_gate(u, P, pre, cfg, "z"), P, pre + "layernorm_z", cfg))
    k = _gate(u, P, pre, cfg, "k")
    v = _gate(u, P, pre, cfg, "v")
    d = torch.sigmoid(_gate(u, P, pre, cfg, "decay"))
    s = d * s_prev + k * v
    if cfg.fused_ops:
        a = _gate(u, P, pre, cfg, "h") + s                    # lucyrnn.py:54
    else:
        a = _lin(u + s, P, pre + "W_h")                        # lucyrnn.py:62
    c = torch.tanh(_ln(a, P, pre + "layernorm_h", cfg))
    h = (1 - z) * c + z * h_prev                               # lucyrnn.py:64
    return h, s


def _prefix_sum_scan(kv, lam):
    """decay_mode='prefix_sum', lucyrnn.py:126-142 (1e-7 guards included)."""
    B, T, H = kv.shape
    t = torch.arange(T, dtype=torch.float32).to(kv.dtype).view(1, T, 1)
    dec = torch.exp(-lam * t).expand(B, T, H)
    logw = torch.cumsum(torch.log(dec + 1e-7), dim=1)
    w = torch.exp(logw)
    return torch.cumsum(kv * w, dim=1) / (w + 1e-7)


# --------------------------------------------------------------------------- #
# strategy 1: the reference's own loop structure
# --------------------------------------------------------------------------- #
def forward_looped(P, cfg, x, state=None):
    """Same order of evaluation as lucyrnn.py:89-191."""
    if cfg.decay_mode not in ("learned", "prefix_sum"):
        raise ValueError(f"Unknown decay_mode: {cfg.decay_mode}")
    x = _stack_frames(x, cfg)
    B, T, _ = x.shape
    dt = x.dtype
    h, s = _init_state(cfg, B, dt, state)
    if cfg.is_training:
        inp = x
        for l in range(cfg.num_layers):
            pre = f"layers.{l}."
            u = _ln(_lin(inp, P, pre + "input_proj"), P, pre + "layernorm_in", cfg)
            kv = _gate(u, P, pre, cfg, "k") * _gate(u, P, pre, cfg, "v")
            if cfg.decay_mode == "prefix_sum":
                S_all = _prefix_sum_scan(kv, cfg.lambda_decay)
            else:
                d = torch.sigmoid(_gate(u, P, pre, cfg, "decay"))
                run = torch.zeros(B, cfg.hidden_dim, dtype=dt)
                steps = []
                for t in range(T):                               # lucyrnn.py:153-158
                    run = d[:, t] * run + kv[:, t]
                    steps.append(run)
                S_all = torch.stack(steps, dim=1)
            outs = []
            for t in range(T):                                   # lucyrnn.py:160-166
                h[l], _ = _cell(P, pre, cfg, inp[:, t], h[l], S_all[:, t])
                outs.append(h[l])
            inp = torch.stack(outs, dim=1)
        enc = inp
    else:
        frames = []
        for t in range(T):                                       # lucyrnn.py:172-184
            cur = x[:, t]
            for l in range(cfg.num_layers):
                h[l], s[l] = _cell(P, f"layers.{l}.", cfg, cur, h[l], s[l])
                cur = h[l]
            frames.append(cur)
        enc = torch.stack(frames, dim=1)
    logits = _lin(enc, P, "output_proj")
    return logits, (h, s)


def _cell_as_timed(P, pre, cfg, x_t, h_prev, s_prev):
    """_cell with the reference's operation GRANULARITY as well (lucyrnn.py:44-70): one 6H-wide
    fused projection per call and the dead r gate (with its LayerNorm) evaluated, as the reference
    does.  Same values as _cell; used where the oracle stands in for the reference's CPU COST."""
    u = _ln(_lin(x_t, P, pre + "input_proj"), P, pre + "layernorm_in", cfg)
    if not cfg.fused_ops:
        torch.sigmoid(_ln(_lin(u, P, pre + "W_r"), P, pre + "layernorm_r", cfg))      # dead gate, lucyrnn.py:56
        return _cell(P, pre, cfg, x_t, h_prev, s_prev)
    r, z, k, v, h_pre, dl = _lin(u, P, pre + "W_fused").chunk(6, dim=-1)
    torch.sigmoid(_ln(r, P, pre + "layernorm_r", cfg))                                  # dead gate, lucyrnn.py:50
    z = torch.sigmoid(_ln(z, P, pre + "layernorm_z", cfg))
    s = torch.sigmoid(dl) * s_prev + k * v
    c = torch.tanh(_ln(h_pre + s, P, pre + "layernorm_h", cfg))
    return (1 - z) * c + z * h_prev, s


def forward_looped_as_timed(P, cfg, x, state=None):
    """forward_looped with the reference's memory behaviour too: per-step selects ``t[:, i, :]`` and
    the in-place ``layer_output[:, t, :] = h`` of lucyrnn.py:153-166, whose backward zero-fills a full
    (B,T,H) tensor per timestep.  This is the strategy bench.py times as the CPU
    baseline when the reference tree itself is not mounted (kind "port")."""
    if cfg.decay_mode != "learned" or not cfg.is_training:
        return forward_looped(P, cfg, x, state)
    x = _stack_frames(x, cfg)
    B, T, _ = x.shape
    h, s = _init_state(cfg, B, x.dtype, state)
    inp = x.clone()                                                  # lucyrnn.py:110
    for l in range(cfg.num_layers):
        pre = f"layers.{l}."
        u = _ln(_lin(inp, P, pre + "input_proj"), P, pre + "layernorm_in", cfg)
        if cfg.fused_ops:
            _, _, k, v, _, dl = _lin(u, P, pre + "W_fused").chunk(6, dim=-1)
        else:
            k, v, dl = _lin(u, P, pre + "W_k"), _lin(u, P, pre + "W_v"), _lin(u, P, pre + "W_decay")
        kv, d = k * v, torch.sigmoid(dl)
        run = torch.zeros(B, cfg.hidden_dim, dtype=x.dtype)
        steps = []
        for t in range(T):                                           # lucyrnn.py:153-158
            run = d[:, t, :] * run + kv[:, t, :]
            steps.append(run.unsqueeze(1))
        S_all = torch.cat(steps, dim=1)
        out = torch.zeros(B, T, cfg.hidden_dim, dtype=x.dtype)
        for t in range(T):                                           # lucyrnn.py:160-166
            h[l], _ = _cell_as_timed(P, pre, cfg, inp[:, t, :], h[l], S_all[:, t, :])
            out[:, t, :] = h[l]
        inp = out
    return _lin(inp, P, "output_proj"), (h, s)


# --------------------------------------------------------------------------- #
# strategy 2: closed form (Appendix A) — the spec of the CUDA path
# --------------------------------------------------------------------------- #
def _linear_scan(a, b, x0):
    """y_t = a_t*y_{t-1} + b_t over dim 1, y_{-1}=x0.  Returns all y."""
    y = x0
    out = []
    # unbind, not a[:, t]: the backward of T separate selects zero-fills a full [B,T,H] tensor per
    # step (O(T^2) traffic — the very cost SURVEY.md 0.10 measures in the reference); unbind's
    # backward is one stack.  Same values either way.
    for a_t, b_t in zip(a.unbind(1), b.unbind(1)):
        y = a_t * y + b_t
        out.append(y)
    if not out:
        return a.new_zeros(a.shape)
    return torch.stack(out, dim=1)


def layer_closed(P, l, cfg, inp, h0, s0, want_aux=False):
    """One layer in closed form.  Returns (Hout[B,T,H], h_T, s_T, aux)."""
    pre = f"layers.{l}."
    u = _ln(_lin(inp, P, pre + "input_proj"), P, pre + "layernorm_in", cfg)
    zraw = _gate(u, P, pre, cfg, "z")
    k = _gate(u, P, pre, cfg, "k")
    v = _gate(u, P, pre, cfg, "v")
    d = torch.sigmoid(_gate(u, P, pre, cfg, "decay"))
    kv = k * v
    if cfg.is_training:
        if cfg.decay_mode == "prefix_sum":
            S = _prefix_sum_scan(kv, cfg.lambda_decay)
        else:
            S = _linear_scan(d, kv, torch.zeros_like(s0))       # A.1: carried s ignored
        sp = d * S + kv                                          # second application, lucyrnn.py:53
        s_T = s0                                                 # returned unchanged, lucyrnn.py:165
    else:
        S = _linear_scan(d, kv, s0)                              # A.2
        sp = S
        s_T = S[:, -1] if S.shape[1] > 0 else s0
    if cfg.fused_ops:
        a = _gate(u, P, pre, cfg, "h") + sp
    else:
        a = _lin(u + sp, P, pre + "W_h")
    c = torch.tanh(_ln(a, P, pre + "layernorm_h", cfg))
    zh = torch.sigmoid(_ln(zraw, P, pre + "layernorm_z", cfg))
    Hout = _linear_scan(zh, (1 - zh) * c, h0)
    h_T = Hout[:, -1] if Hout.shape[1] > 0 else h0
    aux = dict(u=u, z=zraw, k=k, v=v, d=d, S=S, sp=sp, a=a, c=c, zh=zh) if want_aux else None
    return Hout, h_T, s_T, aux


def forward_closed(P, cfg, x, state=None):
    if cfg.decay_mode not in ("learned", "prefix_sum"):
        raise ValueError(f"Unknown decay_mode: {cfg.decay_mode}")
    x = _stack_frames(x, cfg)
    B = x.shape[0]
    h, s = _init_state(cfg, B, x.dtype, state)
    inp = x
    for l in range(cfg.num_layers):
        inp, h[l], s[l], _ = layer_closed(P, l, cfg, inp, h[l], s[l])
    logits = _lin(inp, P, "output_proj")
    return logits, (h, s)


# --------------------------------------------------------------------------- #
# hand-derived backward of the scan stage (Appendix A.3) — spec of the CUDA bwd
# --------------------------------------------------------------------------- #
def scan_backward_closed(g, z, k, v, p, q, h0, s0, training: bool,
                         ln_z=None, ln_h=None):
    """Adjoint of the fused-ops scan stage.

    Inputs are the gate pre-activations z,k,v,p,q [B,T,H] (p = the 'h' chunk, q = decay
    logits), g = dL/dHout [B,T,H], h0/s0 the carried state.  ln_z / ln_h are optional
    (weight,bias) pairs.  Returns dz,dk,dv,dp,dq and (dw,db) pairs for the two LNs.
    No gradient is produced for h0/s0 (detached between segments, model.py:60-61).
    """
    B, T, H = g.shape
    d = torch.sigmoid(q)
    kv = k * v
    if training:
        S = _linear_scan(d, kv, torch.zeros_like(s0))
        sp = d * S + kv
        Sprev = torch.cat([torch.zeros_like(S[:, :1]), S[:, :-1]], 1)
    else:
        S = _linear_scan(d, kv, s0)
        sp = S
        Sprev = torch.cat([s0.unsqueeze(1), S[:, :-1]], 1)
    a = p + sp

    def ln_fwd(x, wb):
        if wb is None:
            return x, None
        mu = x.mean(-1, keepdim=True)
        rstd = torch.rsqrt(((x - mu) ** 2).mean(-1, keepdim=True) + LN_EPS)
        xh = (x - mu) * rstd
        return xh * wb[0] + wb[1], (xh, rstd)

    def ln_bwd(dy, wb, saved):
        if wb is None:
            return dy, None
        xh, rstd = saved
        dw = (dy * xh).sum((0, 1))
        db = dy.sum((0, 1))
        gx = dy * wb[0]
        dx = rstd * (gx - gx.mean(-1, keepdim=True) - xh * (gx * xh).mean(-1, keepdim=True))
        return dx, (dw, db)

    an, sav_h = ln_fwd(a, ln_h)
    zn, sav_z = ln_fwd(z, ln_z)
    c = torch.tanh(an)
    zh = torch.sigmoid(zn)
    Hout = _linear_scan(zh, (1 - zh) * c, h0)
    Hprev = torch.cat([h0.unsqueeze(1), Hout[:, :-1]], 1)

    gam = torch.zeros_like(g)
    run = torch.zeros(B, H, dtype=g.dtype)
    for t in range(T - 1, -1, -1):                     # gamma_t = g_t + zh_{t+1} gamma_{t+1}
        run = g[:, t] + (zh[:, t + 1] * run if t + 1 < T else 0)
        gam[:, t] = run
    dzh = gam * (Hprev - c)
    dan = gam * (1 - zh) * (1 - c * c)
    dzn = dzh * zh * (1 - zh)
    da, gln_h = ln_bwd(dan, ln_h, sav_h)
    dz, gln_z = ln_bwd(dzn, ln_z, sav_z)
    dp = da
    dsp = da
    sig = torch.zeros_like(g)
    run = torch.zeros(B, H, dtype=g.dtype)
    for t in range(T - 1, -1, -1):
        nxt = d[:, t + 1] * run if t + 1 < T else 0
        run = (d[:, t] * dsp[:, t] if training else dsp[:, t]) + nxt
        sig[:, t] = run
    if training:
        dkv = dsp + sig
        dd = S * dsp + Sprev * sig
    else:
        dkv = sig
        dd = Sprev * sig
    dk = dkv * v
    dv = dkv * k
    dq = dd * d * (1 - d)
    return dict(dz=dz, dk=dk, dv=dv, dp=dp, dq=dq, ln_h=gln_h, ln_z=gln_z, Hout=Hout)


# --------------------------------------------------------------------------- #
# host glue restated from model.py
# --------------------------------------------------------------------------- #
def detach_states(states):
    """model.py:11-25: rebuild the nested container with detached tensors."""
    if isinstance(states, torch.Tensor):
        return states.detach()
    if isinstance(states, dict):
        return {k: detach_states(v) for k, v in states.items()}
    if isinstance(states, tuple):
        return tuple(detach_states(v) for v in states)
    if isinstance(states, list):
        return [detach_states(v) for v in states]
    return states


def train_segments(P, cfg, xs: List[torch.Tensor], targets, in_lens, tgt_lens,
                   looped: bool = True):
    """K consecutive segments with carried, detached state and a CTC loss per segment
    (model.py:60-71, train.py:460-580).  P tensors must have requires_grad=True.
    Returns list of losses, list of logits, final state; grads accumulate into P[*].grad."""
    fwd = forward_looped if looped else forward_closed
    state = None
    losses, outs = [], []
    crit = torch.nn.CTCLoss(blank=0, zero_infinity=True)
    for i, x in enumerate(xs):
        if state:
            state = detach_states(state)
        logits, state = fwd(P, cfg, x, state)
        logp = logits.log_softmax(-1).transpose(0, 1)
        loss = crit(logp, targets[i], in_lens[i], tgt_lens[i])
        loss.backward()
        losses.append(loss.detach())
        outs.append(logits.detach())
    return losses, outs, state
