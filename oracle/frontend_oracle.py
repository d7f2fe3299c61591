"""Frontend oracle — TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product).

Restates in numpy fp64 what the reference's frontend does before the hot path:
  * make_frontend (model.py:250-279): torchaudio.transforms.MFCC(n_mfcc=80, dct_type=2,
    norm='ortho', log_mels=True, melkwargs={n_fft=400, win_length=400, hop_length=160,
    n_mels=80, center=False, power=2.0, mel_scale='htk'}) or
    MelSpectrogram(**melkwargs) + AmplitudeToDB(top_db=80.0);
  * train.py:473-475: feats = frontend(batch).transpose(1, 2)  -> (B, T, 80);
  * compute_frame_mask (train.py:296-306) and the in_lens line (train.py:490).
The arithmetic lives in torchaudio (third party, unpinned in requirements.txt; 2.11.0 installed
here), so this file follows torchaudio's published algorithm (functional.spectrogram,
melscale_fbanks, create_dct, amplitude_to_DB) and is PINNED against torchaudio itself through the
golden vectors made by tests/golden/make_frontend_golden.py.
"""
import math

import numpy as np

N_FFT, HOP, N_MELS, N_MFCC = 400, 160, 80, 80
N_FREQ = N_FFT // 2 + 1


def hann_window(n=N_FFT):
    """torch.hann_window(n) (periodic)."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def melscale_fbanks(sample_rate, n_freqs=N_FREQ, n_mels=N_MELS):
    """torchaudio.functional.melscale_fbanks(n_freqs, 0, sr/2, n_mels, sr, None, 'htk') -> [n_freqs, n_mels]."""
    f_min, f_max = 0.0, float(sample_rate // 2)
    all_freqs = np.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = np.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(0.0, np.minimum(down, up))


def create_dct(n_mfcc=N_MFCC, n_mels=N_MELS):
    """torchaudio.functional.create_dct(n_mfcc, n_mels, 'ortho') -> [n_mels, n_mfcc]."""
    n = np.arange(n_mels, dtype=np.float64)
    k = np.arange(n_mfcc, dtype=np.float64)[:, None]
    dct = np.cos(math.pi / n_mels * (n + 0.5) * k)
    dct[0] *= 1.0 / math.sqrt(2.0)
    dct *= math.sqrt(2.0 / n_mels)
    return dct.T


def num_frames(n_samples):
    return 0 if n_samples < N_FFT else 1 + (n_samples - N_FFT) // HOP


def mel_power(wav, sample_rate=16000):
    """(B, S) waveform -> (B, T, 80) mel power spectrogram (center=False, power=2)."""
    wav = np.asarray(wav, dtype=np.float64)
    B, S = wav.shape
    T = num_frames(S)
    idx = np.arange(T)[:, None] * HOP + np.arange(N_FFT)[None, :]
    frames = wav[:, idx] * hann_window()[None, None, :]            # (B, T, 400)
    spec = np.abs(np.fft.rfft(frames, axis=-1)) ** 2               # (B, T, 201)
    return spec @ melscale_fbanks(sample_rate)


def mfcc(wav, sample_rate=16000):
    """(B, S) -> (B, T, 80): the 'mfcc' frontend followed by train.py:475's transpose."""
    return np.log(mel_power(wav, sample_rate) + 1e-6) @ create_dct()


def mel_db(wav, sample_rate=16000, top_db=80.0):
    """(B, S) -> (B, T, 80): the 'mel' frontend (AmplitudeToDB with ONE cut-off for the batch)."""
    x = 10.0 * np.log10(np.maximum(mel_power(wav, sample_rate), 1e-10))
    if x.size:
        x = np.maximum(x, x.max() - top_db)
    return x


def frame_mask_and_lens(sample_mask, n_feat_frames, stack_order=1):
    """train.py:486-490 + compute_frame_mask (296-306): sample_mask (B, S) bool ->
    (frame_mask (B, T) bool, in_lens list[int]).  Follows the reference literally, including the
    fp32 division of the in_lens line and the view() that fails when S_trim != T*int(subsample)."""
    m = np.asarray(sample_mask).astype(bool)
    B, S = m.shape
    subsample = S / n_feat_frames
    subsample *= float(stack_order)
    T = int(S / subsample)
    S_trim = S - (S % T)
    sub_i = int(subsample)
    if S_trim != T * sub_i:
        raise RuntimeError(f"shape '[{B}, {T}, {sub_i}]' is invalid for input of size {B * S_trim}")
    frame_mask = m[:, :S_trim].reshape(B, T, sub_i).any(axis=2)
    q = m.sum(axis=1).astype(np.float32) / np.float32(subsample)
    in_lens = np.minimum(q, np.float32(n_feat_frames)).astype(np.int64)
    return frame_mask, [int(v) for v in in_lens]
