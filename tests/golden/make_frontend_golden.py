#!/usr/bin/env python
"""Golden vectors for the frontend, made with the reference's own frontend
objects: torchaudio.transforms.MFCC / MelSpectrogram + AmplitudeToDB built exactly as
model.py:250-279 builds them, and compute_frame_mask / the in_lens line copied in behaviour from
train.py:296-306, 486-490 (train.py itself does not import here: jiwer/ffmpeg are missing).
Run in the build container (torchaudio 2.11.0 CPU); output: tests/golden/frontend_cases.npz."""
import os

import numpy as np
import torch
import torchaudio

HERE = os.path.dirname(os.path.abspath(__file__))
MEL = dict(n_fft=400, win_length=400, hop_length=160, n_mels=80, center=False, power=2.0, mel_scale="htk")


def frontends(sr):
    mfcc = torchaudio.transforms.MFCC(sample_rate=sr, n_mfcc=80, dct_type=2, norm="ortho", log_mels=True, melkwargs=MEL)
    mel = torch.nn.Sequential(torchaudio.transforms.MelSpectrogram(sample_rate=sr, **MEL),
                              torchaudio.transforms.AmplitudeToDB(top_db=80.0))
    return mfcc, mel


def reference_frame_mask(sample_mask, subsample):
    B, S = sample_mask.shape
    T = int(S / subsample)
    S_trim = S - (S % T)
    sample_mask = sample_mask[:, :S_trim]
    return sample_mask.view(B, T, int(subsample)).any(dim=2)


def main():
    g = torch.Generator().manual_seed(20261018)
    out = {}
    cases = [("noise_1s", 16000, 2, 16000), ("speechlike_2s", 16000, 3, 32000), ("short", 16000, 1, 400),
             ("sr8k", 8000, 2, 8000 + 37), ("ragged_tail", 16000, 2, 16000 + 159)]
    for name, sr, B, S in cases:
        t = torch.arange(S, dtype=torch.float64) / sr
        wav = torch.randn(B, S, generator=g, dtype=torch.float64) * 0.05
        if name != "noise_1s":      # harmonic content with a moving envelope: wide dynamic range per frame
            for b in range(B):
                f0 = 110.0 * (b + 1)
                for h in range(1, 12):
                    wav[b] += (0.5 / h) * torch.sin(2 * np.pi * f0 * h * t) * (0.5 + 0.5 * torch.sin(2 * np.pi * 1.5 * t + b))
        if name == "speechlike_2s":
            wav[1, S // 2:] = 0.0       # digital silence: exercises log(0 + 1e-6) and the dB floor
        wav = wav.float()
        mfcc, mel = frontends(sr)
        with torch.no_grad():
            out[f"{name}/wav"] = wav.numpy()
            out[f"{name}/sr"] = np.int64(sr)
            out[f"{name}/mfcc"] = mfcc(wav).transpose(1, 2).contiguous().numpy()
            out[f"{name}/mel_db"] = mel(wav).transpose(1, 2).contiguous().numpy()
    # frame mask / in_lens: 30 s and 1 s batches, ragged valid lengths
    for name, S, nfeat, valid in [("mask_1s", 16000, 98, [16000, 8000, 161, 0]),
                                  ("mask_30s", 480000, 2998, [480000, 479999, 240000, 12345])]:
        m = torch.zeros(len(valid), S, dtype=torch.bool)
        for b, v in enumerate(valid):
            m[b, :v] = True
        m[1, 5] = False                 # a hole inside valid audio
        subsample = m.size(1) / nfeat
        subsample *= float(1)
        fm = reference_frame_mask(m, subsample)
        in_lens = (m.sum(dim=1) / subsample).clamp(max=nfeat).long()
        out[f"{name}/valid"] = np.array(valid)
        out[f"{name}/S"] = np.int64(S)
        out[f"{name}/nfeat"] = np.int64(nfeat)
        out[f"{name}/frame_mask"] = fm.numpy()
        out[f"{name}/in_lens"] = in_lens.numpy()
    np.savez_compressed(os.path.join(HERE, "frontend_cases.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
