"""Frontend in front of the hot path: MFCC / mel-dB features, the frame
mask and ``in_lens``.

Mirrors ``make_frontend`` (model.py:250-279) and ``compute_frame_mask`` (train.py:296-306): the
modules take the waveform batch ``(B, S)`` and return ``(B, 80, T)`` exactly like the torchaudio
transforms the reference builds, so ``frontend(batch).transpose(1, 2).contiguous()``
(train.py:473-475) keeps working — the returned tensor is a transposed view of the ``(B, T, 80)``
buffer the kernel writes, which makes that ``.contiguous()`` free.  One fused CUDA kernel
(``csrc/sc_frontend.cu``); no torchaudio, no cuFFT, no CPU fallback.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
from torch import nn

from . import _lib
from ._lib import call, ptr, stream

N_FFT, HOP, N_MELS = 400, 160, 80
MEL_KWARGS = {"n_fft": 400, "win_length": 400, "hop_length": 160, "n_mels": 80, "center": False,
              "power": 2.0, "mel_scale": "htk"}


def _tables(sample_rate: int) -> torch.Tensor:
    lib = _lib.load()
    n = int(lib.sc_frontend_tables_len())
    host = torch.empty(n, dtype=torch.float32)
    _lib.check(lib.sc_frontend_tables(host.data_ptr(), n, int(sample_rate)), "sc_frontend_tables")
    return host


class _Frontend(nn.Module):
    mode = 0
    top_db = -1.0

    def __init__(self, sample_rate: int = 16000):
        super().__init__()
        self.sample_rate = int(sample_rate)
        self.register_buffer("tables", _tables(self.sample_rate), persistent=False)

    @staticmethod
    def num_frames(n_samples: int) -> int:
        return 0 if n_samples < N_FFT else 1 + (n_samples - N_FFT) // HOP

    @_lib.on_tensor_device
    def features(self, waveform: torch.Tensor, frame_mask: torch.Tensor = None) -> torch.Tensor:
        """(..., S) -> (..., T, 80) contiguous fp32 (the layout the encoder consumes).  With
        ``frame_mask`` (B, T) bool (MFCC only) masked frames come out as zeros: model.py:377's
        ``feats * mask.unsqueeze(-1).float()`` without another pass over the features."""
        _lib.require_cuda(waveform, "frontend input")
        if self.tables.device != waveform.device:
            raise RuntimeError("frontend tables and waveform are on different devices: move the module with .to(device)")
        lead = waveform.shape[:-1]
        S = waveform.shape[-1]
        w = waveform.reshape(-1, S).to(torch.float32)
        if w.stride(-1) != 1:
            w = w.contiguous()
        B, T = w.shape[0], self.num_frames(S)
        out = torch.empty(B, T, N_MELS, dtype=torch.float32, device=w.device)
        if B and T:
            gmax = torch.empty(1, dtype=torch.int32, device=w.device) if self.mode == 1 else None
            fm, ldm = None, 0
            if frame_mask is not None:
                fm = frame_mask.to(torch.bool).reshape(B, -1)
                if fm.shape[1] != T or self.mode != 0:
                    raise ValueError(f"frame_mask must be ({B}, {T}) and the frontend 'mfcc'")
                fm = fm if fm.stride(-1) == 1 else fm.contiguous()
                ldm = fm.stride(0)
            call("sc_frontend", ptr(w), w.stride(0), B, S, ptr(self.tables), self.mode, float(self.top_db),
                 ptr(fm), ldm, ptr(out), out.stride(0), ptr(gmax), stream())
        return out.reshape(*lead, T, N_MELS)

    @torch.no_grad()
    def forward(self, waveform: torch.Tensor) -> torch.Tensor:
        """(..., S) -> (..., 80, T), the torchaudio layout (a transposed view)."""
        return self.features(waveform).transpose(-1, -2)


class MFCC(_Frontend):
    """``torchaudio.transforms.MFCC(sample_rate, n_mfcc=80, dct_type=2, norm='ortho', log_mels=True,
    melkwargs=MEL_KWARGS)`` as built at model.py:262-270."""
    mode = 0


class MelDB(_Frontend):
    """``MelSpectrogram(sample_rate, **MEL_KWARGS)`` + ``AmplitudeToDB(top_db=80.0)`` (model.py:271-278).
    As in torchaudio, a 2-D or 3-D input gets ONE dB floor for the whole batch."""
    mode = 1
    top_db = 80.0


def make_frontend(ftype: str, sample_rate: int):
    """model.py:250-279: -> (frontend module, mel_kwargs)."""
    if ftype == "mfcc":
        return MFCC(sample_rate), dict(MEL_KWARGS)
    if ftype == "mel":
        return MelDB(sample_rate), dict(MEL_KWARGS)
    raise ValueError(f"Unsupported frontend: {ftype}")


def _mask_geometry(S: int, subsample: float) -> Tuple[int, int]:
    T = int(S / subsample)
    S_trim = S - (S % T)
    sub = int(subsample)
    if S_trim != T * sub:        # what sample_mask.view(B, T, int(subsample)) raises upstream
        raise RuntimeError(f"shape '[-1, {T}, {sub}]' is invalid for input with {S_trim} samples per stream")
    return T, sub


@_lib.on_tensor_device
def frame_mask_and_lens(sample_mask: torch.Tensor, n_feat_frames: int, stack_order: int = 1
                        ) -> Tuple[torch.Tensor, List[int]]:
    """train.py:486-490 in one kernel: ``sample_mask`` (B, S) bool -> (frame_mask (B, T) bool,
    in_lens list[int]) with ``subsample = S / n_feat_frames * stack_order``."""
    _lib.require_cuda(sample_mask, "sample_mask")
    B, S = sample_mask.shape
    subsample = S / n_feat_frames
    subsample *= float(stack_order)
    T, sub = _mask_geometry(S, subsample)
    m = sample_mask.to(torch.bool)
    if m.stride(-1) != 1:
        m = m.contiguous()
    fm = torch.empty(B, T, dtype=torch.bool, device=m.device)
    lens = torch.empty(B, dtype=torch.int64, device=m.device)
    call("sc_frame_mask", ptr(m), m.stride(0), B, S, T, sub, float(subsample), int(n_feat_frames), ptr(fm), ptr(lens),
         stream())
    return fm, lens.tolist()


def featurize(frontend: _Frontend, waveform: torch.Tensor, sample_mask: torch.Tensor
              ) -> Tuple[torch.Tensor, torch.Tensor, List[int]]:
    """train.py:473-490 + model.py:377 in two launches: (B, S) waveform and sample mask ->
    (feats (B, T, 80) with masked frames zeroed, frame_mask (B, T) bool, in_lens list[int])."""
    T = frontend.num_frames(waveform.shape[-1])
    fm, lens = frame_mask_and_lens(sample_mask, T)
    if fm.shape[1] != T:                 # upstream asserts this (train.py:492)
        raise AssertionError(f"Mismatch: feats={T} vs mask={fm.shape[1]}")
    return frontend.features(waveform, fm if frontend.mode == 0 else None), fm, lens


def compute_frame_mask(sample_mask: torch.Tensor, subsample: float) -> torch.Tensor:
    """train.py:296-306 with the same signature."""
    B, S = sample_mask.shape
    T, sub = _mask_geometry(S, subsample)
    m = sample_mask.to(torch.bool)
    if m.stride(-1) != 1:
        m = m.contiguous()
    _lib.require_cuda(m, "sample_mask")
    fm = torch.empty(B, T, dtype=torch.bool, device=m.device)
    lens = torch.empty(B, dtype=torch.int64, device=m.device)
    call("sc_frame_mask", ptr(m), m.stride(0), B, S, T, sub, float(subsample), T, ptr(fm), ptr(lens), stream())
    return fm
