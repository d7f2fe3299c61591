"""LucyRNN — B200-native drop-in for the reference encoder.

Mirrors the public surface of /root/reference/lucyrnn.py (``LucyRNNCell`` lines 8-70,
``LucyRNN`` lines 72-191): same constructor, same ``forward(x, hidden_states=None,
masks=None) -> (logits, (h_list, s_list))``, same state_dict keys, same initialisation, same
two behaviours selected by ``config.is_training``:

* training path (lucyrnn.py:109-170): S scan from zero, decay applied a second time inside
  the cell, carried ``s`` ignored and returned unchanged, ``h`` carried;
* step path (lucyrnn.py:172-184): ``h`` and ``s`` both carried, single application.

Nothing here loops over time in Python.  Each layer is one autograd node that enqueues
hand-written sm_100a kernels through the C-ABI (``_lib``): projection GEMMs (K1), the fused
recurrent scan and its reverse-time adjoint (K2), and LayerNorm/row helpers for the
non-default flag combinations.  Precision: fp32 tensors run the fp32 path (rtol 1e-4 vs the
reference); under ``torch.autocast`` (or with ``compute_dtype=torch.bfloat16``) activations
are bf16 with fp32 accumulation, fp32 master weights and fp32 carried state.

Deliberate, documented differences from the reference:
* ``masks`` must be None (the argument is dead/broken upstream and
  ``ASRModel`` never passes it, model.py:384/388);
* the dead ``r`` gate (lucyrnn.py:50/56) is never computed; its weight-gradient rows are
  exact zeros (fused) / None (unfused), as autograd gives upstream;
* gradients do not flow into the carried state (it is detached upstream, model.py:60-61).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, ops
from .lucyrnn_conf import LucyRNNConfig

FOLD_PROJECTIONS = True   # bf16 fast path: G = x (Wf W_in)^T + (Wf b_in + bf)
_GATES_UNFUSED = ("z", "k", "v", "decay")  # gate blocks computed from u (W_h acts on u+s')


class LucyRNNCell(nn.Module):
    """Parameter container with the reference's names and init (lucyrnn.py:9-42).

    ``forward`` evaluates one step (lucyrnn.py:44-70, mask=None) through the CUDA step path.
    """

    def __init__(self, input_dim, hidden_dim, fused_ops=False, layer_norm=True):
        super().__init__()
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.fused_ops = fused_ops
        self.layer_norm = layer_norm
        self.input_proj = nn.Linear(input_dim, hidden_dim)
        self.layernorm_in = nn.LayerNorm(hidden_dim) if layer_norm else nn.Identity()
        self.layernorm_r = nn.LayerNorm(hidden_dim) if layer_norm else nn.Identity()
        self.layernorm_z = nn.LayerNorm(hidden_dim) if layer_norm else nn.Identity()
        self.layernorm_h = nn.LayerNorm(hidden_dim) if layer_norm else nn.Identity()
        if fused_ops:
            self.W_fused = nn.Linear(hidden_dim, 6 * hidden_dim)
        else:
            self.W_r = nn.Linear(hidden_dim, hidden_dim)
            self.W_z = nn.Linear(hidden_dim, hidden_dim)
            self.W_k = nn.Linear(hidden_dim, hidden_dim)
            self.W_v = nn.Linear(hidden_dim, hidden_dim)
            self.W_h = nn.Linear(hidden_dim, hidden_dim)
            self.W_decay = nn.Linear(hidden_dim, hidden_dim)
        self.init_weights()

    def init_weights(self):
        for name, param in self.named_parameters():
            if "weight" in name and param.dim() > 1:
                nn.init.orthogonal_(param)
        if self.layer_norm:
            for ln in (self.layernorm_in, self.layernorm_r, self.layernorm_z, self.layernorm_h):
                nn.init.constant_(ln.bias, 0)
                nn.init.constant_(ln.weight, 1.0)

    @_lib.on_tensor_device
    def forward(self, x, h_prev, s_prev, mask=None):
        if mask is not None:
            raise NotImplementedError("LucyRNNCell: mask must be None (dead argument upstream)")
        meta = _LayerMeta(self, train_mode=False, decay_mode=0, lam=0.0, dtype=_compute_dtype(x, None))
        xin = x.unsqueeze(1)
        if xin.dtype != meta.dtype:
            xin = ops.cast(xin.contiguous(), meta.dtype)
        out, hT, sT = _LucyLayerFn.apply(meta, xin.contiguous(), h_prev.float().contiguous(),
                                         s_prev.float().contiguous(), *meta.params(self))
        return hT, sT


def _compute_dtype(x: torch.Tensor, override) -> torch.dtype:
    if override is not None:
        return override
    if torch.is_autocast_enabled("cuda"):
        return torch.bfloat16        # fp16 autocast (train.py:515) is served by the bf16 path too
    if x.dtype == torch.bfloat16:
        return torch.bfloat16
    return torch.float32


class _LayerMeta:
    """Static description of one layer call (flags + parameter order)."""

    def __init__(self, cell: LucyRNNCell, train_mode: bool, decay_mode: int, lam: float, dtype):
        self.fused = cell.fused_ops
        self.ln = cell.layer_norm
        self.train_mode = train_mode
        self.decay_mode = decay_mode
        self.lam = lam
        self.dtype = dtype
        self.H = cell.hidden_dim
        names = ["input_proj.weight", "input_proj.bias"]
        if self.ln:
            names += ["layernorm_in.weight", "layernorm_in.bias", "layernorm_z.weight", "layernorm_z.bias",
                      "layernorm_h.weight", "layernorm_h.bias"]
        if self.fused:
            names += ["W_fused.weight", "W_fused.bias"]
        else:
            for g in _GATES_UNFUSED + ("h",):
                names += [f"W_{g}.weight", f"W_{g}.bias"]
        self.names = names
        self.index = {n: i for i, n in enumerate(names)}
        # the fully fused scan kernel serves the configuration model.py:232-245 wires
        self.fast = self.fused and not self.ln and decay_mode == 0
        # fold input_proj into W_fused (see _forward_folded) on the bf16 training path
        self.fold = self.fast and dtype == torch.bfloat16 and FOLD_PROJECTIONS

    def params(self, cell: LucyRNNCell):
        out = []
        for n in self.names:
            mod, attr = n.split(".")
            out.append(getattr(getattr(cell, mod), attr))
        return out


def _w(t: torch.Tensor, dtype) -> torch.Tensor:
    """Weight in the compute dtype (bf16 copy made by our cast kernel; fp32 passes through)."""
    t = t.detach()
    if t.dtype == dtype:
        return t if t.is_contiguous() else t.contiguous()
    return ops.cast(t.contiguous(), dtype)


class _LucyLayerFn(torch.autograd.Function):
    """One LucyRNN layer over a whole segment: x[B,T,in] -> Hout[B,T,H], h_T, s_T."""

    @staticmethod
    def forward(ctx, meta: _LayerMeta, x, h0, s0, *params):
        B, T, Fin = x.shape
        H, M, cd = meta.H, B * T, meta.dtype
        p = lambda n: params[meta.index[n]]          # noqa: E731
        x2 = x.reshape(M, Fin)
        if meta.fold:
            return _LucyLayerFn._forward_folded(ctx, meta, x2, h0, s0, p, B, T, Fin)
        W_in = _w(p("input_proj.weight"), cd)
        pre = ops.gemm_fwd(x2, W_in, p("input_proj.bias").detach())
        sv = {}
        if meta.ln:
            u, sv["mu_in"], sv["rs_in"] = ops.layernorm_fwd(pre, p("layernorm_in.weight").detach(),
                                                            p("layernorm_in.bias").detach())
        else:
            u = pre
        if meta.fused:
            Wg = _w(p("W_fused.weight")[H:], cd)                       # r rows skipped
            bg = p("W_fused.bias").detach()[H:]
            nblk = 5                                                   # z k v p q
        else:
            Wg = _w(torch.cat([p(f"W_{g}.weight").detach() for g in _GATES_UNFUSED], 0), cd)
            bg = torch.cat([p(f"W_{g}.bias").detach() for g in _GATES_UNFUSED], 0)
            nblk = 4                                                   # z k v q
        G = ops.gemm_fwd(u, Wg, bg)                                    # [M, nblk*H]
        blk = lambda i: G[:, i * H:(i + 1) * H]                        # noqa: E731
        z, k, v = blk(0), blk(1), blk(2)
        q = blk(4) if meta.fused else blk(3)
        if meta.fast:
            Hout, hT, sT, ck = ops.scan_fwd(G, B, T, H, h0, s0, meta.train_mode)
            ctx.save_for_backward(x2, pre, u, G, Hout, h0, s0, ck, W_in, Wg)
        else:
            addend = blk(3) if meta.fused else u
            A, S_all, sT = ops.sscan_fwd(k, v, q, addend, s0, B, T, H, meta.train_mode,
                                         meta.decay_mode, meta.lam)
            if meta.fused:
                A2, W_h = A, None
            else:
                W_h = _w(p("W_h.weight"), cd)
                A2 = ops.gemm_fwd(A, W_h, p("W_h.bias").detach())
            if meta.ln:
                An, sv["mu_h"], sv["rs_h"] = ops.layernorm_fwd(A2, p("layernorm_h.weight").detach(),
                                                               p("layernorm_h.bias").detach())
                Zn, sv["mu_z"], sv["rs_z"] = ops.layernorm_fwd(z, p("layernorm_z.weight").detach(),
                                                               p("layernorm_z.bias").detach())
            else:
                An, Zn = A2, z
            Hout, hT = ops.hscan_fwd(An, Zn, h0, B, T, H)
            ctx.save_for_backward(x2, pre, u, G, Hout, h0, s0, S_all, W_in, Wg, A, A2, An, Zn, W_h)
        ctx.meta, ctx.sv, ctx.shape = meta, sv, (B, T, Fin)
        ctx.lnp = [params[meta.index[n]].detach() for n in
                   ("layernorm_in.weight", "layernorm_z.weight", "layernorm_h.weight")] if meta.ln else None
        if sT is None:
            sT = h0.new_empty(0)
        Hout3 = Hout.view(B, T, H)
        ctx.mark_non_differentiable(hT, sT)
        return Hout3, hT, sT

    # ---- projection folding (bf16, fused_ops, no LayerNorm) -------------------------------
    # With layer_norm=False nothing non-linear sits between input_proj and W_fused
    # (lucyrnn.py:113-116 with Identity norms), so  G = (x W_in^T + b_in) Wf^T + bf
    #                                                 = x (Wf W_in)^T + (Wf b_in + bf).
    # The [5H,in] product Wc is rebuilt from the current weights every call (a 10-GFLOP
    # tensor-core GEMM) and the per-frame work drops from 2(in*H + 5H^2) to 2*5H*in FLOPs
    # forward and the same ratio backward; `u` is never materialised.  The backward maps the
    # gradient of Wc back onto W_fused / input_proj exactly (chain rule in weight space, the
    # fp32 dWc carried through the bf16 tensor cores as a two-term hi/lo expansion).
    # The bias rides along as one extra input column: with x_ext = [x | 1] and
    # Win_ext = [W_in | b_in | 0..] (padded to a multiple of 8 columns for TMA alignment),
    # Wc_ext = Wf . Win_ext holds Wc in its first `in` columns and Wf b_in in column `in`, and
    # the same extension makes the backward's weight-space GEMMs produce db_in and the
    # dbc (x) b_in outer-product term of dWf with no separate matrix-vector kernels.
    @staticmethod
    def _forward_folded(ctx, meta, x2, h0, s0, p, B, T, Fin):
        H, cd = meta.H, meta.dtype
        dev = x2.device
        PADC = 8
        Wfb = _w(p("W_fused.weight").detach()[H:], cd)                # [5H,H]  bf16
        Win_ext = torch.zeros(H, Fin + PADC, dtype=cd, device=dev)    # [W_in | b_in | 0]
        ops.cast(p("input_proj.weight").detach(), cd, out=Win_ext[:, :Fin])
        ops.cast(p("input_proj.bias").detach().view(H, 1), cd, out=Win_ext[:, Fin:Fin + 1])
        Wc_ext = ops.gemm_dgrad(Wfb, Win_ext, out_dtype=torch.float32)  # [5H, in+8] = Wf . Win_ext
        Wc = ops.cast(Wc_ext[:, :Fin], cd)                            # [5H,in] bf16, contiguous
        bc = Wc_ext[:, Fin] + p("W_fused.bias").detach()[H:]          # Wf b_in + bf  (fp32)
        G = ops.gemm_fwd(x2, Wc, bc)                                  # [M,5H]
        Hout, hT, sT, ck = ops.scan_fwd(G, B, T, H, h0, s0, meta.train_mode)
        ctx.save_for_backward(x2, G, Hout, h0, s0, ck, Wc, Wfb, Win_ext)
        ctx.meta, ctx.sv, ctx.shape = meta, {}, (B, T, Fin)
        if sT is None:
            sT = h0.new_empty(0)
        ctx.mark_non_differentiable(hT, sT)
        return Hout.view(B, T, H), hT, sT

    @staticmethod
    def _backward_folded(ctx, g2):
        meta = ctx.meta
        B, T, Fin = ctx.shape
        H = meta.H
        x2, G, Hout, h0, s0, ck, Wc, Wfb, Win_ext = ctx.saved_tensors
        dev = g2.device
        E = Win_ext.shape[1]                                          # in + pad
        grads = [None] * len(meta.names)
        gi = meta.index
        dG, dbg = ops.scan_bwd(G, Hout, h0, s0, ck, g2, B, T, H, meta.train_mode)
        dWc_ext = torch.zeros(5 * H, E, dtype=torch.float32, device=dev)   # [dWc | dbc | 0]
        ops.gemm_wgrad(dG, x2, out=dWc_ext[:, :Fin], accumulate=True)
        dWc_ext[:, Fin].copy_(dbg)
        dx = ops.gemm_dgrad(dG, Wc).view(B, T, Fin) if ctx.needs_input_grad[1] else None
        hl = ops.split_bf16(dWc_ext)                                  # [5H, 2E] = [hi | lo]
        Win2 = torch.cat([Win_ext, Win_ext], dim=1)                   # [H, 2E]
        dWf = torch.empty(6 * H, H, dtype=torch.float32, device=dev)
        dWf[:H].zero_()                                               # dead r gate: exact zeros
        ops.gemm_fwd(hl, Win2, None, out=dWf[H:])                     # dWc W_in^T + dbc (x) b_in
        both = ops.gemm_wgrad(Wfb, hl)                                # Wf^T [hi|lo] -> [H, 2E]
        dbf = torch.zeros(6 * H, dtype=torch.float32, device=dev)
        dbf[H:].copy_(dbg)
        grads[gi["W_fused.weight"]], grads[gi["W_fused.bias"]] = dWf, dbf
        grads[gi["input_proj.weight"]] = both[:, :Fin] + both[:, E:E + Fin]
        grads[gi["input_proj.bias"]] = both[:, Fin] + both[:, E + Fin]
        return (None, dx, None, None, *grads)

    @staticmethod
    def backward(ctx, dHout, _dhT, _dsT):
        meta, sv = ctx.meta, ctx.sv
        B, T, Fin = ctx.shape
        H, M, cd = meta.H, B * T, meta.dtype
        dev = dHout.device
        g2 = dHout.reshape(M, H)
        if g2.dtype != cd:
            g2 = ops.cast(g2.contiguous(), cd)
        elif not g2.is_contiguous():
            g2 = g2.contiguous()
        if meta.fold:
            return _LucyLayerFn._backward_folded(ctx, g2)
        grads = [None] * len(meta.names)
        gi = meta.index
        if meta.fast:
            x2, pre, u, G, Hout, h0, s0, ck, W_in, Wg = ctx.saved_tensors
            dG, dbg = ops.scan_bwd(G, Hout, h0, s0, ck, g2, B, T, H, meta.train_mode)
            du_extra = None
        else:
            x2, pre, u, G, Hout, h0, s0, S_all, W_in, Wg, A, A2, An, Zn, W_h = ctx.saved_tensors
            nblk = 5 if meta.fused else 4
            dG = torch.empty(M, nblk * H, dtype=cd, device=dev)
            dblk = lambda i: dG[:, i * H:(i + 1) * H]                  # noqa: E731
            blk = lambda i: G[:, i * H:(i + 1) * H]                    # noqa: E731
            qi = 4 if meta.fused else 3
            # where dA (grad of the tensor the S scan produced) must end up
            dA_final = dblk(3) if meta.fused else torch.empty(M, H, dtype=cd, device=dev)
            if meta.ln:
                dAn = torch.empty(M, H, dtype=cd, device=dev)
                dZn = torch.empty(M, H, dtype=cd, device=dev)
            else:
                dAn = dA_final if meta.fused else torch.empty(M, H, dtype=cd, device=dev)
                dZn = dblk(0)
            ops.hscan_bwd(An, Zn, Hout, h0, g2, dAn, dZn, B, T, H)
            # bias gradients of the gate projection = column sums of dG: formed by the kernels that write dG's blocks
            # (LayerNorm backward for z and, when fused, p; the S-scan backward for k, v, q) instead of by a pass
            # over the whole tensor; blocks nobody sums (no LayerNorm: z, p) fall back to sc_colsum below
            dbg = torch.zeros(nblk * H, dtype=torch.float32, device=dev)
            summed = [False] * nblk
            if meta.ln:
                w_in, w_z, w_h = ctx.lnp
                dA2_buf = dA_final if meta.fused else torch.empty(M, H, dtype=cd, device=dev)
                dA2, dw, db = ops.layernorm_bwd(dAn, A2, w_h, sv["mu_h"], sv["rs_h"], dx=dA2_buf,
                                                dxsum=dbg[3 * H:4 * H] if meta.fused else None)
                grads[gi["layernorm_h.weight"]], grads[gi["layernorm_h.bias"]] = dw, db
                _, dw, db = ops.layernorm_bwd(dZn, blk(0), w_z, sv["mu_z"], sv["rs_z"], dx=dblk(0), dxsum=dbg[:H])
                grads[gi["layernorm_z.weight"]], grads[gi["layernorm_z.bias"]] = dw, db
                summed[0] = True
                if meta.fused:
                    summed[3] = True
            else:
                dA2 = dAn
            if meta.fused:
                dA = dA2                                            # already sits in dG's p block
                du_extra = None
            else:
                grads[gi["W_h.weight"]] = ops.gemm_wgrad(dA2, A)
                grads[gi["W_h.bias"]] = ops.colsum(dA2)
                dA = ops.gemm_dgrad(dA2, W_h, out=dA_final)
                du_extra = dA                                       # addend was u (lucyrnn.py:62)
            if qi == 3:                                             # k, v, q adjacent: blocks 1..3
                ksum = dbg[H:4 * H]
                ops.sscan_bwd(blk(1), blk(2), blk(qi), S_all, s0, dA, dblk(1), dblk(2), dblk(qi),
                              B, T, H, meta.train_mode, meta.decay_mode, meta.lam, dsum=ksum)
            else:                                                   # fused: k, v in blocks 1, 2 and q in block 4
                ksum = torch.zeros(3 * H, dtype=torch.float32, device=dev)
                ops.sscan_bwd(blk(1), blk(2), blk(qi), S_all, s0, dA, dblk(1), dblk(2), dblk(qi),
                              B, T, H, meta.train_mode, meta.decay_mode, meta.lam, dsum=ksum)
                dbg[H:3 * H].copy_(ksum[:2 * H])
                dbg[4 * H:].copy_(ksum[2 * H:])
            summed[1] = summed[2] = summed[qi] = True
            for i, done in enumerate(summed):
                if not done:
                    ops.colsum(dblk(i), out=dbg[i * H:(i + 1) * H])
        # gate projection backward
        if meta.fused:
            dWf = torch.empty(6 * H, H, dtype=torch.float32, device=dev)
            dWf[:H].zero_()                                         # dead r gate: exact zeros
            ops.gemm_wgrad(dG, u, out=dWf[H:], accumulate=False)
            dbf = torch.zeros(6 * H, dtype=torch.float32, device=dev)
            dbf[H:].copy_(dbg)
            grads[gi["W_fused.weight"]], grads[gi["W_fused.bias"]] = dWf, dbf
        else:
            dWcat = ops.gemm_wgrad(dG, u)
            for i, g in enumerate(_GATES_UNFUSED):
                grads[gi[f"W_{g}.weight"]] = dWcat[i * H:(i + 1) * H]
                grads[gi[f"W_{g}.bias"]] = dbg[i * H:(i + 1) * H]
        du = ops.gemm_dgrad(dG, Wg)
        if du_extra is not None:
            du.add_(du_extra)
        if meta.ln:
            db_in = torch.zeros(H, dtype=torch.float32, device=dev)
            dpre, dw, db = ops.layernorm_bwd(du, pre, ctx.lnp[0], sv["mu_in"], sv["rs_in"], dxsum=db_in)
            grads[gi["layernorm_in.weight"]], grads[gi["layernorm_in.bias"]] = dw, db
        else:
            dpre = du
            db_in = ops.colsum(dpre)
        grads[gi["input_proj.weight"]] = ops.gemm_wgrad(dpre, x2)
        grads[gi["input_proj.bias"]] = db_in
        dx = None
        if ctx.needs_input_grad[1]:
            dx = ops.gemm_dgrad(dpre, W_in).view(B, T, Fin)
        return (None, dx, None, None, *grads)


class _LinearFn(torch.autograd.Function):
    """output_proj (lucyrnn.py:85, 186) through the K1 kernels."""

    @staticmethod
    def forward(ctx, x, weight, bias, cd):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        W = _w(weight, cd)
        y = ops.gemm_fwd(x2, W, bias.detach())
        ctx.save_for_backward(x2, W)
        ctx.shp = shp
        return y.view(*shp[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, W = ctx.saved_tensors
        d2 = dy.reshape(-1, dy.shape[-1])
        if d2.dtype != x2.dtype:
            d2 = ops.cast(d2.contiguous(), x2.dtype)
        elif not d2.is_contiguous():
            d2 = d2.contiguous()
        dx = ops.gemm_dgrad(d2, W).view(ctx.shp) if ctx.needs_input_grad[0] else None
        dW = ops.gemm_wgrad(d2, x2)
        db = ops.colsum(d2)
        return dx, dW, db, None


class LucyRNN(nn.Module):
    """Drop-in for /root/reference/lucyrnn.py:72-191 (see module docstring)."""

    def __init__(self, config: LucyRNNConfig, compute_dtype: Optional[torch.dtype] = None):
        super().__init__()
        self.config = config
        if self.config.kernel_impl not in ["native", "triton"]:
            raise ValueError("kernel_impl must be either 'native' or 'triton'")
        _lib.load()                              # fail loudly at construction if the .so is absent
        self.compute_dtype = compute_dtype       # None: fp32 unless autocast / bf16 input
        self.layers = nn.ModuleList()
        for i in range(config.num_layers):
            layer_input_dim = config.input_dim * config.stack_order if i == 0 else config.hidden_dim
            self.layers.append(LucyRNNCell(layer_input_dim, config.hidden_dim, config.fused_ops, config.layer_norm))
        self.output_proj = nn.Linear(config.hidden_dim, config.vocab_size)
        nn.init.zeros_(self.output_proj.weight)
        nn.init.zeros_(self.output_proj.bias)

    @_lib.on_tensor_device
    def forward(self, x, hidden_states=None, masks=None):
        cfg = self.config
        if masks is not None:
            raise NotImplementedError(
                "LucyRNN.forward: masks must be None — the reference's mask path is dead/broken "
                "(lucyrnn.py:164/176 + 66-68) and ASRModel zero-multiplies features instead (model.py:376-377)")
        if cfg.decay_mode not in ("learned", "prefix_sum"):
            raise ValueError(f"Unknown decay_mode: {cfg.decay_mode}")
        _lib.require_cuda(x, "LucyRNN input")
        batch_size, seq_len, feat_dim = x.size()
        if cfg.stack_order > 1:                                     # lucyrnn.py:92-99
            stack = cfg.stack_order
            trim_len = seq_len - (seq_len % stack)
            x = x[:, :trim_len, :].reshape(batch_size, trim_len // stack, feat_dim * stack)
            seq_len = x.size(1)
        H = cfg.hidden_dim
        if hidden_states is None:                                   # lucyrnn.py:101-105
            h = [torch.zeros(batch_size, H, device=x.device) for _ in range(cfg.num_layers)]
            s = [torch.zeros(batch_size, H, device=x.device) for _ in range(cfg.num_layers)]
        else:
            h, s = hidden_states                                    # caller's lists, updated in place
            # the reference fails with a broadcast error when a carried state does not match the
            # batch (e.g. a last partial batch, lucyrnn.py:64); here the scan would read h0/s0 out
            # of bounds, so check before anything is enqueued
            for nm, lst in (("h", h), ("s", s)):
                if len(lst) != cfg.num_layers:
                    raise RuntimeError(f"LucyRNN.forward: hidden_states {nm} has {len(lst)} entries, "
                                       f"the model has {cfg.num_layers} layers")
                for l, t in enumerate(lst):
                    if tuple(t.shape) != (batch_size, H):
                        raise RuntimeError(f"LucyRNN.forward: carried state {nm}[{l}] has shape {tuple(t.shape)}, "
                                           f"expected ({batch_size}, {H}) for this batch")
                    if t.device != x.device:
                        raise RuntimeError(f"LucyRNN.forward: carried state {nm}[{l}] is on {t.device}, input on {x.device}")
        cd = _compute_dtype(x, self.compute_dtype)
        inp = x.contiguous()
        if inp.dtype != cd:
            inp = ops.cast(inp, cd) if not inp.requires_grad else inp.to(cd)
        # decay_mode only alters the training path's scan (lucyrnn.py:126-142)
        dmode = 1 if (cfg.decay_mode == "prefix_sum" and cfg.is_training) else 0
        for l, layer in enumerate(self.layers):
            meta = _LayerMeta(layer, cfg.is_training, dmode, float(cfg.lambda_decay), cd)
            h0 = h[l].detach().float().contiguous()
            s0 = s[l].detach().float().contiguous()
            inp, hT, sT = _LucyLayerFn.apply(meta, inp, h0, s0, *meta.params(layer))
            h[l] = hT
            if not cfg.is_training:
                s[l] = sT                                           # training path: s passes through
        logits = _LinearFn.apply(inp, self.output_proj.weight, self.output_proj.bias, cd)
        if cfg.return_last_states:
            return logits, (h, s)
        return logits


class LucyRNNtriton(LucyRNN):
    """Import-compatibility alias: model.py:9/310 constructs ``LucyRNNtriton(cfg)``.

    The reference class of that name (lucyrnn_triton.py:77-155) is a different, forward-only
    7-gate network that BASELINE.json's north_star retires; this alias runs the LucyRNN of
    lucyrnn.py on the CUDA kernels, with no Triton inside.
    """
