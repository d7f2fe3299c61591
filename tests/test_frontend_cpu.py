"""CPU: the frontend oracle (oracle/frontend_oracle.py) pinned against golden vectors made with
the reference's own torchaudio frontend objects (tests/golden/make_frontend_golden.py), and the
host-side table builder of the C-ABI (sc_frontend_tables: pure C, no GPU needed) checked against
the oracle — including a numpy emulation of the kernel's twice-folded real DFT."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import frontend_oracle as FO

WAV_CASES = ["noise_1s", "speechlike_2s", "short", "sr8k", "ragged_tail"]


@pytest.fixture(scope="module")
def G():
    return load_golden("frontend_cases")


@pytest.mark.parametrize("name", WAV_CASES)
def test_oracle_matches_torchaudio_golden(G, name):
    wav, sr = G[name + "/wav"], int(G[name + "/sr"])
    # torchaudio computes in fp32 (cuFFT/pocketfft + fp32 matmuls): its own noise is ~1e-4 on
    # MFCC values up to ~120 and ~6e-4 dB on the weakest mel bands
    np.testing.assert_allclose(FO.mfcc(wav, sr), G[name + "/mfcc"], rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(FO.mel_db(wav, sr), G[name + "/mel_db"], rtol=1e-5, atol=3e-3)
    assert G[name + "/mfcc"].shape[1] == FO.num_frames(wav.shape[1])


@pytest.mark.parametrize("name", ["mask_1s", "mask_30s"])
def test_oracle_frame_mask_and_lens_bit_exact(G, name):
    S, nfeat, valid = int(G[name + "/S"]), int(G[name + "/nfeat"]), G[name + "/valid"]
    m = np.zeros((len(valid), S), bool)
    for b, v in enumerate(valid):
        m[b, :v] = True
    m[1, 5] = False
    fm, lens = FO.frame_mask_and_lens(m, nfeat)
    assert np.array_equal(fm, G[name + "/frame_mask"])
    assert lens == G[name + "/in_lens"].tolist()


def test_oracle_frame_mask_rejects_what_the_reference_rejects():
    with pytest.raises(RuntimeError):
        FO.frame_mask_and_lens(np.ones((1, 1000), bool), 7, 3)   # T = 2, S_trim = 1000 != 2 * 428


def _unpack(tab):
    NB = 104
    o = 400 + 4 * NB * NB
    # bases are stored [13 stages][4 transforms][8 rows][104 bins] (one bulk copy per stage)
    t = dict(win=tab[:400], bas=tab[400:o].reshape(13, 4, 8, NB).transpose(1, 0, 2, 3).reshape(4, NB, NB))
    t["melw"] = tab[o:o + 80 * 32].reshape(80, 32); o += 80 * 32
    t["lo"] = tab[o:o + 80].view(np.int32); o += 80
    t["cnt"] = tab[o:o + 80].view(np.int32); o += 80
    t["dct"] = tab[o:o + 6400].reshape(80, 80)
    assert o + 6400 == tab.size
    return t


@pytest.mark.parametrize("sr", [8000, 16000, 22050, 48000])
def test_table_builder_matches_oracle(sr):
    from statecatcher_b200 import frontend as FE
    t = _unpack(FE._tables(sr).numpy())
    np.testing.assert_allclose(t["win"], FO.hann_window(), atol=1e-7)
    np.testing.assert_allclose(t["dct"], FO.create_dct(), atol=1e-7)
    fb = np.zeros((201, 80))
    for m in range(80):
        fb[t["lo"][m]:t["lo"][m] + t["cnt"][m], m] = t["melw"][m, :t["cnt"][m]]
    np.testing.assert_allclose(fb, FO.melscale_fbanks(sr), atol=1e-7)
    assert t["cnt"].max() <= 32


def test_table_builder_argument_errors():
    from statecatcher_b200 import _lib
    lib = _lib.load()
    buf = torch.empty(16)
    assert lib.sc_frontend_tables(buf.data_ptr(), 16, 16000) == -1      # buffer too small
    assert lib.sc_frontend_tables(None, 1 << 20, 16000) == -1
    assert lib.sc_frontend(None, 0, 0, 0, None, 0, 0.0, None, 0, None, 0, None, None) == -1
    assert lib.sc_frame_mask(None, 0, 1, 1, 1, 1, 1.0, 1, None, None, None) == -1


def test_folded_dft_tables_reproduce_rfft():
    """The kernel's algorithm in numpy: window, fold n<->400-n, fold n<->200-n, four ~100x100
    real transforms from the shipped tables == numpy's rfft of the windowed frame."""
    from statecatcher_b200 import frontend as FE
    t = _unpack(FE._tables(16000).numpy())
    rng = np.random.default_rng(0)
    for _ in range(3):
        x = rng.standard_normal(400)
        a = x * t["win"].astype(np.float64)
        V = np.zeros((4, 104))
        V[0, 0], V[1, 0] = a[0] + a[200], a[0] - a[200]
        for n in range(1, 100):
            e1, e2 = a[n] + a[400 - n], a[200 - n] + a[200 + n]
            o1, o2 = a[n] - a[400 - n], a[200 - n] - a[200 + n]
            V[:, n] = e1 + e2, e1 - e2, o1 - o2, o1 + o2
        V[0, 100], V[3, 100] = a[100] + a[300], a[100] - a[300]
        R = [V[q] @ t["bas"][q].astype(np.float64) for q in range(4)]
        re, im = np.zeros(201), np.zeros(201)
        for k in range(201):
            re[k], im[k] = R[k & 1][k >> 1], R[2 + (k & 1)][k >> 1]
        X = np.fft.rfft(a)
        np.testing.assert_allclose(re, X.real, atol=5e-6)
        np.testing.assert_allclose(im, X.imag, atol=5e-6)


def test_frontend_module_contract_without_gpu():
    import statecatcher_b200 as sb
    fe, kw = sb.make_frontend("mfcc", 16000)
    assert kw == {"n_fft": 400, "win_length": 400, "hop_length": 160, "n_mels": 80, "center": False,
                  "power": 2.0, "mel_scale": "htk"}
    assert isinstance(fe, sb.MFCC) and isinstance(sb.make_frontend("mel", 16000)[0], sb.MelDB)
    assert list(fe.state_dict()) == []                     # tables are not checkpointed
    with pytest.raises(ValueError):
        sb.make_frontend("fbank", 16000)
    with pytest.raises(RuntimeError):                      # no CPU fallback
        fe(torch.zeros(1, 16000))
    assert fe.num_frames(399) == 0 and fe.num_frames(400) == 1 and fe.num_frames(480000) == 2998
