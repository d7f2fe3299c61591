"""RNN-T head on the sm_100a wavefront kernels (K4).

Reference call sites: train.py:39/144 (`from warp_rnnt import RNNTLoss`, used as the
criterion) and model.py:73-105 (`criterion(log_probs=..., labels=..., frames_lengths=...,
labels_lengths=..., blank_id=..., compact=..., gather=True)`), joiners model.py:112-200.
warp_rnnt is absent from the reference tree, its requirements and this image, and that call
matches no published API, so PARITY IS UNPINNED: the loss implemented here
is the standard transducer negative log-likelihood (Graves 2012), checked against
oracle/rnnt_oracle.py and torchaudio.functional.rnnt_loss.

``RNNTLoss`` keeps the keyword signature of the reference call.  ``reduction`` defaults to
'mean' over the batch so the value can be ``.backward()``-ed as train.py:526-536 does.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import call, ptr, stream
from .ctc import _lens
from .lucyrnn import _LinearFn, _compute_dtype


class _RNNTFn(torch.autograd.Function):
    """log_probs is (B,T,U1,V) padded, or (rows,V) compact with row offsets (model.py:147-200)."""

    @staticmethod
    def forward(ctx, log_probs, labels, frame_lens, label_lens, blank, offsets, T, U1):
        V = log_probs.shape[-1]
        B = labels.shape[0]
        dev = log_probs.device
        U1p = (U1 + 3) & ~3
        ws = lambda: torch.empty(B, T + U1, U1p, dtype=torch.float32, device=dev)   # noqa: E731
        eb, el, alpha, beta = ws(), ws(), ws(), ws()
        nll = torch.empty(B, dtype=torch.float32, device=dev)
        ldl = labels.stride(0) if labels.numel() else max(U1 - 1, 1)
        call("sc_rnnt_fwd", ptr(log_probs), ptr(labels), ldl, ptr(frame_lens), ptr(label_lens),
             B, T, U1, V, blank, ptr(offsets), ptr(eb), ptr(el), ptr(alpha), ptr(beta), ptr(nll), stream())
        ctx.save_for_backward(labels, frame_lens, label_lens, eb, el, alpha, beta, nll, offsets)
        ctx.cfg = (B, T, U1, V, blank, ldl, tuple(log_probs.shape))
        return nll

    @staticmethod
    def backward(ctx, gnll):
        labels, frame_lens, label_lens, eb, el, alpha, beta, nll, offsets = ctx.saved_tensors
        B, T, U1, V, blank, ldl, shape = ctx.cfg
        grad = torch.empty(shape, dtype=torch.float32, device=nll.device)
        w = gnll.to(torch.float32).contiguous()
        rows = shape[0] if offsets is not None else 0
        call("sc_rnnt_bwd", ptr(labels), ldl, ptr(frame_lens), ptr(label_lens), B, T, U1, V, blank,
             ptr(offsets), rows, ptr(eb), ptr(el), ptr(alpha), ptr(beta), ptr(nll), ptr(w), ptr(grad), stream())
        return grad, None, None, None, None, None, None, None


def _reduce(nll, reduction):
    if reduction == "mean":
        return nll.mean()
    if reduction == "sum":
        return nll.sum()
    if reduction in (None, "none"):
        return nll
    raise ValueError(f"unknown reduction {reduction!r}")


@_lib.on_tensor_device
def rnnt_loss(log_probs, labels, frames_lengths, labels_lengths, blank: int = 0,
              reduction: str = "mean", compact: bool = False) -> torch.Tensor:
    """log_probs: (B,T,U+1,V) normalised (log_softmax of the joint), or with compact=True the
    packed (sum_b T_b*(U_b+1), V) layout the compact joiner produces; labels (B,U)."""
    _lib.require_cuda(log_probs, "rnnt_loss input")
    x = log_probs.float().contiguous()
    lab = labels.to(device=x.device, dtype=torch.int64).contiguous()
    B = lab.size(0)
    fl, _ = _lens(frames_lengths, x.device, B, "frames_lengths")
    ll, ll_max = _lens(labels_lengths, x.device, B, "labels_lengths")
    if compact:
        if x.dim() != 2:
            raise ValueError("compact rnnt_loss expects (rows, V) log-probs")
        rows = fl * (ll + 1)
        offsets = (torch.cumsum(rows, 0) - rows).contiguous()
        T, U1 = int(fl.max().item()), int(ll.max().item()) + 1       # host sync: sizes of the workspaces
        if int(rows.sum().item()) != x.size(0):
            raise ValueError("compact log_probs row count does not match sum(T_b*(U_b+1))")
        return _reduce(_RNNTFn.apply(x, lab, fl, ll, int(blank), offsets, T, U1), reduction)
    if x.dim() != 4:
        raise ValueError("rnnt_loss expects (B,T,U+1,V) log-probs")
    _, T, U1, V = x.shape
    if lab.dim() != 2 or x.size(0) != B or lab.size(1) < U1 - 1:
        raise ValueError("labels must be (B,U) with U >= log_probs.size(2)-1")
    if ll_max is not None and ll_max > U1 - 1:              # list lengths (what model.py passes): checked on the host, as
        raise ValueError(f"labels_lengths up to {ll_max} exceed the lattice width {U1 - 1}")   # warp_rnnt / torchaudio do
    return _reduce(_RNNTFn.apply(x, lab, fl, ll, int(blank), None, T, U1), reduction)


def RNNTLoss(log_probs, labels, frames_lengths, labels_lengths, blank_id: int = 0, compact: bool = False,
             gather: bool = True, reduction: str = "mean", **_unused):
    """Callable with the keyword signature model.py:97-105 uses on `warp_rnnt.RNNTLoss`."""
    return rnnt_loss(log_probs, labels, frames_lengths, labels_lengths, blank=blank_id, reduction=reduction,
                     compact=compact)


class RNNTPredictorJoiner(nn.Module):
    """model.py:112-145: Embedding -> Linear predictor, Linear encoder projection, broadcast
    add, tanh, Linear to vocab.  The three Linear layers run on the K1 GEMM kernels."""

    def __init__(self, enc_out_dim: int, pred_emb_dim: int, join_dim: int, vocab_size: int, debug: bool = False):
        super().__init__()
        self.embedding = nn.Embedding(vocab_size, pred_emb_dim)
        self.enc_proj = nn.Linear(enc_out_dim, join_dim)
        self.pred_proj = nn.Linear(pred_emb_dim, join_dim)
        self.joiner = nn.Linear(join_dim, vocab_size)
        self.debug = debug

    def _lin(self, x, layer):
        cd = _compute_dtype(x, None)
        if x.dtype != cd:
            x = x.to(cd)
        return _LinearFn.apply(x.contiguous(), layer.weight, layer.bias, cd)

    @_lib.on_tensor_device
    def forward(self, enc_out: torch.Tensor, prefix: torch.Tensor):
        pred_emb = self.embedding(prefix)                        # (B, U+1, E)
        enc = self._lin(enc_out, self.enc_proj)                  # (B, T, J)
        pred = self._lin(pred_emb, self.pred_proj)               # (B, U+1, J)
        joint = torch.tanh(enc.unsqueeze(2) + pred.unsqueeze(1))  # (B, T, U+1, J)
        return self._lin(joint, self.joiner)                     # (B, T, U+1, V)


class RNNTCompactPredictorJoiner(RNNTPredictorJoiner):
    """model.py:147-200: joint only over the live T_b x (U_b+1) nodes of every utterance, packed
    as (sum_b T_b*(U_b+1), J) rows -> (rows, V) logits (no padding work, no padded lattice)."""

    @_lib.on_tensor_device
    def forward(self, enc_out, prefix, in_lens, tgt_lens):
        B = enc_out.size(0)
        in_lens = [int(v) for v in (in_lens.tolist() if isinstance(in_lens, torch.Tensor) else in_lens)]
        tgt_lens = [int(v) for v in (tgt_lens.tolist() if isinstance(tgt_lens, torch.Tensor) else tgt_lens)]
        enc = self._lin(enc_out, self.enc_proj)                  # (B, T, J)
        pred = self._lin(self.embedding(prefix), self.pred_proj)  # (B, U+1, J)
        parts = []
        for b in range(B):
            T, U1 = in_lens[b], tgt_lens[b] + 1
            if T > 0:
                parts.append(torch.tanh(enc[b, :T].unsqueeze(1) + pred[b, :U1].unsqueeze(0)).reshape(T * U1, -1))
        joint = torch.cat(parts, 0) if parts else enc.new_zeros(0, enc.size(-1))
        return self._lin(joint, self.joiner)                     # (rows, V)


# ------------------------------------------------------------------ fused joint head ----
class _RNNTFusedFn(torch.autograd.Function):
    """joiner (model.py:136-144) + log_softmax + transducer loss without the (B,T,U+1,V) tensor.

    Forward: enc/pred projections once; then per block of ``chunk`` frames: joint = tanh(enc+pred)
    -> logits (K1 GEMM) -> per-node lse and blank/label log-probs into the skewed lattice arrays;
    the logits block is dropped.  Lattice recursion (K4) gives nll.  Backward: node gradients,
    then per block the joint and logits are recomputed, dlogits formed in one pass and pushed
    back through the joiner GEMMs; d_enc / d_pred reductions of the broadcast add by a custom
    kernel.  Peak memory is one block, not the whole lattice.
    """

    @staticmethod
    def forward(ctx, enc_out, pred_emb, labels, fl, ll, blank, chunk, cd, keep_blocks, We, be, Wp, bp, Wo, bo):
        B, T, De = enc_out.shape
        U1, J, V = pred_emb.shape[1], We.shape[0], Wo.shape[0]
        dev = enc_out.device
        wcast = lambda w: w.detach() if w.dtype == cd else ops.cast(w.detach().contiguous(), cd)   # noqa: E731
        Wec, Wpc, Woc = wcast(We), wcast(Wp), wcast(Wo)
        e2 = enc_out.reshape(B * T, De)
        p2 = pred_emb.reshape(B * U1, -1)
        if e2.dtype != cd:
            e2 = ops.cast(e2.contiguous(), cd)
        if p2.dtype != cd:
            p2 = ops.cast(p2.contiguous(), cd)
        encp = ops.gemm_fwd(e2, Wec, be.detach()).view(B, T, J)
        predp = ops.gemm_fwd(p2, Wpc, bp.detach()).view(B, U1, J)
        U1p = (U1 + 3) & ~3
        ws = lambda: torch.empty(B, T + U1, U1p, dtype=torch.float32, device=dev)   # noqa: E731
        eb, el, alpha, beta = ws(), ws(), ws(), ws()
        lse = torch.empty(B, max(T, 1), U1, dtype=torch.float32, device=dev)
        nll = torch.empty(B, dtype=torch.float32, device=dev)
        ldl = labels.stride(0) if labels.numel() else max(U1 - 1, 1)
        dtc = _lib.dt(encp)
        # Block buffers are allocated once and reused by every block (views for the tail block) —
        # unless the whole joint + logits lattice fits comfortably in free HBM (cfg4: 88 GB of the
        # B200's 180 GB), in which case every block keeps its own and the backward skips the
        # recomputation of the joiner (one elementwise pass and one GEMM per block).
        rows_max = B * min(chunk, max(T, 1)) * U1
        esz = torch.empty((), dtype=cd).element_size()
        need = B * T * U1 * (J + V) * esz
        if keep_blocks is not None:
            keep = keep_blocks
        else:
            # free = what the driver still has + what torch's caching allocator holds but is not using (after the
            # first step the previous step's blocks sit there: cudaMemGetInfo alone would say "no room")
            free = torch.cuda.mem_get_info(dev)[0] + torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)
            keep = T > 0 and need < 0.75 * free
        kept = []
        if not keep:
            joint_buf = torch.empty(rows_max, J, dtype=cd, device=dev)
            logits_buf = torch.empty(rows_max, V, dtype=cd, device=dev)
        for t0 in range(0, T, chunk):
            Tc = min(chunk, T - t0)
            if keep:
                joint = torch.empty(B * Tc * U1, J, dtype=cd, device=dev)
                logits = torch.empty(B * Tc * U1, V, dtype=cd, device=dev)
                kept.append((joint, logits))
            else:
                joint, logits = joint_buf[:B * Tc * U1], logits_buf[:B * Tc * U1]
            ech = encp[:, t0:t0 + Tc]
            call("sc_joint_fwd", ptr(ech), ech.stride(0), ech.stride(1), ptr(predp), predp.stride(0), predp.stride(1),
                 ptr(joint), B, Tc, U1, J, dtc, stream())
            ops.gemm_fwd(joint, Woc, bo.detach(), out=logits)
            call("sc_rnnt_lse_gather", ptr(logits), _lib.dt(logits), ptr(labels), ldl, ptr(fl), ptr(ll), B, T, t0, Tc,
                 U1, V, blank, ptr(lse), ptr(eb), ptr(el), stream())
        call("sc_rnnt_lattice", ptr(fl), ptr(ll), B, T, U1, ptr(eb), ptr(el), ptr(alpha), ptr(beta), ptr(nll), stream())
        ctx.kept = kept                                   # plain attribute: not inputs/outputs of the node
        ctx.save_for_backward(e2, p2, encp, predp, labels, fl, ll, eb, el, alpha, beta, lse, nll, Wec, Wpc, Woc, bo.detach())
        ctx.cfg = (B, T, U1, J, V, De, blank, chunk, cd, ldl, tuple(pred_emb.shape))
        return nll

    @staticmethod
    def backward(ctx, gnll):
        (e2, p2, encp, predp, labels, fl, ll, eb, el, alpha, beta, lse, nll, Wec, Wpc, Woc, bo) = ctx.saved_tensors
        B, T, U1, J, V, De, blank, chunk, cd, ldl, pshape = ctx.cfg
        dev = nll.device
        gb = torch.empty(B, max(T, 1), U1, dtype=torch.float32, device=dev)
        gl = torch.empty_like(gb)
        w = gnll.to(torch.float32).contiguous()
        call("sc_rnnt_node_grads", ptr(fl), ptr(ll), B, T, U1, ptr(eb), ptr(el), ptr(alpha), ptr(beta), ptr(nll), ptr(w),
             ptr(gb), ptr(gl), stream())
        d_encp = torch.empty(B, T, J, dtype=cd, device=dev)
        d_predp = torch.zeros(B, U1, J, dtype=torch.float32, device=dev)
        dWo = torch.zeros(V, J, dtype=torch.float32, device=dev)
        dbo = torch.zeros(V, dtype=torch.float32, device=dev)
        dtc = _lib.dt(encp)
        rows_max = B * min(chunk, max(T, 1)) * U1
        kept = ctx.kept
        joint_buf = logits_buf = None                    # recompute buffers, allocated on first use
        dlogits_buf = torch.empty(rows_max, V, dtype=cd, device=dev)
        dJ_buf = torch.empty(rows_max, J, dtype=cd, device=dev)
        for bi, t0 in enumerate(range(0, T, chunk)):
            Tc = min(chunk, T - t0)
            n = B * Tc * U1
            dlogits, dJ = dlogits_buf[:n], dJ_buf[:n]
            ech = encp[:, t0:t0 + Tc]
            if kept and kept[bi] is not None:
                joint, logits = kept[bi]
                kept[bi] = None                          # released as the backward passes it (a 2nd backward recomputes)
            else:
                if joint_buf is None:
                    joint_buf = torch.empty(rows_max, J, dtype=cd, device=dev)
                    logits_buf = torch.empty(rows_max, V, dtype=cd, device=dev)
                joint, logits = joint_buf[:n], logits_buf[:n]
                call("sc_joint_fwd", ptr(ech), ech.stride(0), ech.stride(1), ptr(predp), predp.stride(0), predp.stride(1),
                     ptr(joint), B, Tc, U1, J, dtc, stream())
                ops.gemm_fwd(joint, Woc, bo, out=logits)
            call("sc_rnnt_dlogits", ptr(logits), _lib.dt(logits), ptr(lse), ptr(gb), ptr(gl), ptr(labels), ldl, ptr(ll),
                 B, T, t0, Tc, U1, V, blank, ptr(dlogits), ptr(dbo), stream())     # dbo += column sums (fused)
            ops.gemm_wgrad(dlogits, joint, out=dWo, accumulate=True)
            ops.gemm_dgrad(dlogits, Woc, out=dJ)
            dch = d_encp[:, t0:t0 + Tc]
            call("sc_joint_bwd", ptr(dJ), ptr(ech), ech.stride(0), ech.stride(1), ptr(predp), predp.stride(0),
                 predp.stride(1), ptr(dch), dch.stride(0), dch.stride(1), ptr(d_predp), B, Tc, U1, J, dtc, stream())
        del dlogits_buf, dJ_buf
        de2 = d_encp.view(B * T, J)
        dp2 = d_predp.view(B * U1, J)
        if dp2.dtype != cd:
            dp2 = ops.cast(dp2, cd)
        dWe, dbe = ops.gemm_wgrad(de2, e2), ops.colsum(de2)
        dWp, dbp = ops.gemm_wgrad(dp2, p2), ops.colsum(dp2)
        d_enc = ops.gemm_dgrad(de2, Wec).view(B, T, De) if ctx.needs_input_grad[0] else None
        d_pred = ops.gemm_dgrad(dp2, Wpc).view(pshape) if ctx.needs_input_grad[1] else None
        return (d_enc, d_pred, None, None, None, None, None, None, None, dWe, dbe, dWp, dbp, dWo, dbo)


class RNNTFusedHead(nn.Module):
    """Joiner + transducer loss in one node: same parameters (and state_dict keys) as
    ``RNNTPredictorJoiner`` (model.py:112-145), same loss as ``RNNTLoss(joiner(...).log_softmax(-1))``,
    but the (B,T,U+1,V) logits exist only one block of ``chunk_frames`` frames at a time, so
    configs[3] (J=512, V=1024, T=3000, U<=150) runs at batch 64."""

    def __init__(self, enc_out_dim: int, pred_emb_dim: int, join_dim: int, vocab_size: int,
                 chunk_frames: int = 64, compute_dtype=None, keep_blocks: Optional[bool] = None):
        super().__init__()
        self.embedding = nn.Embedding(vocab_size, pred_emb_dim)
        self.enc_proj = nn.Linear(enc_out_dim, join_dim)
        self.pred_proj = nn.Linear(pred_emb_dim, join_dim)
        self.joiner = nn.Linear(join_dim, vocab_size)
        self.chunk_frames = chunk_frames
        self.compute_dtype = compute_dtype
        self.keep_blocks = keep_blocks        # None: keep the blocks' joint/logits when they fit in free HBM, else recompute

    @_lib.on_tensor_device
    def forward(self, enc_out, tokens, frames_lengths, labels_lengths, blank_id: int = 0, reduction: str = "mean"):
        _lib.require_cuda(enc_out, "RNNTFusedHead input")
        B = enc_out.size(0)
        tokens = tokens.to(device=enc_out.device, dtype=torch.int64).contiguous()
        prefix = torch.cat([torch.full((B, 1), blank_id, dtype=torch.int64, device=enc_out.device), tokens], dim=1)
        fl, _ = _lens(frames_lengths, enc_out.device, B, "frames_lengths")
        ll, ll_max = _lens(labels_lengths, enc_out.device, B, "labels_lengths")
        if ll_max is not None and ll_max > tokens.size(1):
            raise ValueError(f"labels_lengths up to {ll_max} exceed the {tokens.size(1)} tokens given")
        cd = _compute_dtype(enc_out, self.compute_dtype)
        pred_emb = self.embedding(prefix)
        nll = _RNNTFusedFn.apply(enc_out.contiguous(), pred_emb.contiguous(), tokens, fl, ll, int(blank_id),
                                 int(self.chunk_frames), cd, self.keep_blocks, self.enc_proj.weight, self.enc_proj.bias,
                                 self.pred_proj.weight, self.pred_proj.bias, self.joiner.weight, self.joiner.bias)
        return _reduce(nll, reduction)
