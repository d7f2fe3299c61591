"""GPU parity: the fused frontend kernel (sc_frontend / sc_frame_mask) against the fp64 oracle
and against golden vectors produced by the reference's torchaudio frontend objects."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import frontend_oracle as FO

pytestmark = pytest.mark.gpu
WAV_CASES = ["noise_1s", "speechlike_2s", "short", "sr8k", "ragged_tail"]
# fp32 arithmetic on both sides (ours: folded DFT + FMA; torchaudio: FFT): MFCC values reach ~120,
# mel dB of bands 60+ dB below the frame maximum carry ~1e-3 dB of fp32 noise
MFCC_TOL = dict(rtol=1e-5, atol=2e-3)
DB_TOL = dict(rtol=1e-5, atol=5e-3)


@pytest.fixture(scope="module")
def G():
    return load_golden("frontend_cases")


@pytest.mark.parametrize("name", WAV_CASES)
def test_mfcc_and_mel_match_oracle_and_torchaudio_golden(cuda_device, G, name):
    import statecatcher_b200 as sb
    wav, sr = G[name + "/wav"], int(G[name + "/sr"])
    x = torch.tensor(wav).cuda()
    mfcc, _ = sb.make_frontend("mfcc", sr)
    mel, _ = sb.make_frontend("mel", sr)
    mfcc, mel = mfcc.cuda(), mel.cuda()
    got = mfcc(x)
    assert got.shape == (wav.shape[0], 80, FO.num_frames(wav.shape[1]))
    feats = got.transpose(1, 2)
    assert feats.is_contiguous()                           # train.py:475's .contiguous() is free
    np.testing.assert_allclose(feats.cpu().numpy(), FO.mfcc(wav, sr), **MFCC_TOL)
    np.testing.assert_allclose(feats.cpu().numpy(), G[name + "/mfcc"], **MFCC_TOL)
    got_db = mel(x).transpose(1, 2).cpu().numpy()
    np.testing.assert_allclose(got_db, FO.mel_db(wav, sr), **DB_TOL)
    np.testing.assert_allclose(got_db, G[name + "/mel_db"], **DB_TOL)


def test_too_short_and_empty_inputs(cuda_device):
    import statecatcher_b200 as sb
    fe = sb.MFCC(16000).cuda()
    assert fe(torch.zeros(2, 399, device="cuda")).shape == (2, 80, 0)
    assert fe(torch.zeros(0, 1600, device="cuda")).shape == (0, 80, 8)
    one = fe(torch.zeros(16000, device="cuda"))            # unbatched (S,) like torchaudio
    assert one.shape == (80, 98)
    np.testing.assert_allclose(one.cpu().numpy(), FO.mfcc(np.zeros((1, 16000)))[0].T, atol=1e-4)


def test_strided_rows_and_batch_independence(cuda_device):
    """Rows of a wider buffer (ldw > S); a stream's features do not depend on its batch mates."""
    import statecatcher_b200 as sb
    g = torch.Generator().manual_seed(5)
    big = torch.randn(3, 20000, generator=g).cuda()
    fe = sb.MFCC(16000).cuda()
    a = fe(big[:, :16000])
    for b in range(3):
        assert torch.equal(a[b], fe(big[b:b + 1, :16000].contiguous())[0])


def test_mel_db_floor_is_one_value_for_the_batch(cuda_device):
    import statecatcher_b200 as sb
    g = torch.Generator().manual_seed(6)
    wav = torch.randn(2, 8000, generator=g)
    wav[1] *= 1e-6                                           # 120 dB quieter: entirely on the floor
    got = sb.MelDB(16000).cuda()(wav.cuda()).transpose(1, 2).cpu().numpy()
    want = FO.mel_db(wav.numpy())
    np.testing.assert_allclose(got, want, **DB_TOL)
    assert np.allclose(got[1], got.max() - 80.0, atol=1e-3)


def test_full_size_segment_batch(cuda_device):
    """configs[1] shape of use: 30 s segments at 16 kHz (2998 frames), against the fp64 oracle."""
    import statecatcher_b200 as sb
    g = torch.Generator().manual_seed(7)
    B, S = 4, 480000
    t = torch.arange(S) / 16000.0
    wav = 0.05 * torch.randn(B, S, generator=g)
    for b in range(B):
        wav[b] += 0.4 * torch.sin(2 * np.pi * (200.0 + 150 * b) * t) * (0.5 + 0.5 * torch.sin(2 * np.pi * 0.7 * t))
    fe = sb.MFCC(16000).cuda()
    got = fe.features(wav.cuda())
    assert got.shape == (B, 2998, 80)
    np.testing.assert_allclose(got.cpu().numpy(), FO.mfcc(wav.numpy()), **MFCC_TOL)


@pytest.mark.parametrize("name", ["mask_1s", "mask_30s"])
def test_frame_mask_and_lens_bit_exact(cuda_device, G, name):
    import statecatcher_b200 as sb
    S, nfeat, valid = int(G[name + "/S"]), int(G[name + "/nfeat"]), G[name + "/valid"]
    m = torch.zeros(len(valid), S, dtype=torch.bool)
    for b, v in enumerate(valid):
        m[b, :int(v)] = True
    m[1, 5] = False
    fm, lens = sb.frame_mask_and_lens(m.cuda(), nfeat)
    assert fm.dtype == torch.bool and np.array_equal(fm.cpu().numpy(), G[name + "/frame_mask"])
    assert lens == G[name + "/in_lens"].tolist()
    ofm, olens = FO.frame_mask_and_lens(m.numpy(), nfeat)
    assert np.array_equal(fm.cpu().numpy(), ofm) and lens == olens
    fm2 = sb.compute_frame_mask(m.cuda(), S / nfeat)        # the reference's two-argument form
    assert torch.equal(fm2, fm)


def test_frame_mask_random_patterns_and_stack_order(cuda_device):
    import statecatcher_b200 as sb
    g = torch.Generator().manual_seed(8)
    m = torch.rand(5, 48000, generator=g) > 0.999           # sparse hits: most frames empty
    m[3] = False
    for nfeat, stack in [(300, 1), (300, 3), (150, 2)]:
        fm, lens = sb.frame_mask_and_lens(m.cuda(), nfeat, stack)
        ofm, olens = FO.frame_mask_and_lens(m.numpy(), nfeat, stack)
        assert np.array_equal(fm.cpu().numpy(), ofm) and lens == olens
    with pytest.raises(RuntimeError):
        sb.frame_mask_and_lens(torch.ones(1, 1000, dtype=torch.bool, device="cuda"), 7, 3)


def test_featurize_folds_the_frame_mask(cuda_device, G):
    """featurize = frontend + compute_frame_mask + in_lens + model.py:377's feats*mask: masked
    frames are zeros, live frames equal the unmasked features, lens as the reference computes."""
    import statecatcher_b200 as sb
    g = torch.Generator().manual_seed(11)
    B, S = 3, 16000
    wav = torch.randn(B, S, generator=g) * 0.1
    valid = [16000, 9000, 0]
    m = torch.zeros(B, S, dtype=torch.bool)
    for b, v in enumerate(valid):
        m[b, :v] = True
    fe = sb.MFCC(16000).cuda()
    feats, fm, lens = sb.featurize(fe, wav.cuda(), m.cuda())
    plain = fe.features(wav.cuda())
    ofm, olens = FO.frame_mask_and_lens(m.numpy(), plain.shape[1])
    assert np.array_equal(fm.cpu().numpy(), ofm) and lens == olens
    want = plain * fm.unsqueeze(-1).float()
    assert torch.equal(feats, want)
    assert (feats[2] == 0).all() and (feats[1, lens[1] + 1:] == 0).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_mask_rows_is_the_reference_multiply(cuda_device, dtype):
    """sc_mask_rows == feats * mask.unsqueeze(-1).float() bit for bit (incl. -0.0), and
    LucyASRModel.forward routes through it."""
    import statecatcher_b200 as sb
    from statecatcher_b200 import ops
    g = torch.Generator().manual_seed(12)
    x = torch.randn(4, 37, 80, generator=g).to(dtype).cuda()
    mask = (torch.rand(4, 37, generator=g) > 0.4).cuda()
    got = ops.mask_rows(x, mask)
    want = (x * mask.unsqueeze(-1).float()).to(dtype)
    assert got.dtype == dtype and torch.equal(got.view(torch.int16 if dtype == torch.bfloat16 else torch.int32),
                                              want.view(torch.int16 if dtype == torch.bfloat16 else torch.int32))
    cfg = sb.LucyRNNConfig(input_dim=80, hidden_dim=32, num_layers=1, vocab_size=11, fused_ops=True, layer_norm=False)
    model = sb.LucyASRModel(cfg).cuda()
    torch.nn.init.normal_(model.encoder.output_proj.weight, std=0.1)
    xf = x.float()
    a, _ = model(xf, mask)
    b_, _ = model.encoder(xf * mask.unsqueeze(-1).float())
    assert torch.equal(a, b_)
