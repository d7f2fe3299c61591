"""CPU: the arithmetic of the experimental linear-domain CTC recursion (alpha = m * 2^e, fp32 mantissa, int
exponent; sc_ctc.cu ctc_alpha_beta_lin_kernel) emulated in numpy with every operation rounded to fp32, against
an fp64 log-domain recursion AND the CTC oracle.  Pins the claim the kernel rests on: with the mantissa
renormalised every step no lattice node is lost and occupancies agree to ~1e-5; without it they do not."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "profiles"))
import ctc_linear_exp_study as study  # noqa: E402

from oracle import ctc_oracle  # noqa: E402


def _case(T, U, V, scale, seed):
    rng = np.random.default_rng(seed)
    logits = (rng.standard_normal((T, V)) * scale).astype(np.float32).astype(np.float64)
    y = rng.integers(1, V, U)
    y[2] = y[1]
    ext = study.ext_labels(y)
    lse = np.log(np.exp(logits - logits.max(1, keepdims=True)).sum(1)) + logits.max(1)
    lp = logits[:, ext] - lse[:, None]
    return logits, y, ext, lp


def _occupancy(a, b, lp2, S):
    ll = np.logaddexp2(a[-1, S - 1], a[-1, S - 2])
    with np.errstate(invalid="ignore", over="ignore"):
        return np.nan_to_num(np.exp2(a + b - lp2 - ll)), ll


def test_fp64_recursion_of_the_study_is_the_oracle():
    logits, y, ext, lp = _case(40, 6, 12, 2.0, 0)
    lse = np.log(np.exp(logits - logits.max(1, keepdims=True)).sum(1)) + logits.max(1)
    nll, alpha, beta = ctc_oracle.ctc_utterance(logits - lse[:, None], y, blank=0)     # oracle gathers from [T,V] itself
    a = study.fp64_log_domain(lp, ext)
    np.testing.assert_allclose(a[np.isfinite(a)], alpha[np.isfinite(a)], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(-np.logaddexp(a[-1, -1], a[-1, -2]), nll, rtol=1e-12)


def _run(T, U, V, scale, K):
    logits, y, ext, lp = _case(T, U, V, scale, 1)
    S = len(ext)
    ref_a = study.fp64_log_domain(lp, ext) / np.log(2.0)
    ref_b = study.fp64_log_domain(lp[::-1, ::-1], ext[::-1])[::-1, ::-1] / np.log(2.0)
    shift = lp.max(1, keepdims=True)
    p = np.exp((lp - shift).astype(np.float32)).astype(np.float32)
    a = study.linear_exp_domain(p, ext, K) + (np.cumsum(shift[:, 0]) / np.log(2.0))[:, None]
    b = (study.linear_exp_domain(p[::-1, ::-1], ext[::-1], K)[::-1, ::-1]
         + (np.cumsum(shift[::-1, 0])[::-1] / np.log(2.0))[:, None])
    g_ref, ll_ref = _occupancy(ref_a, ref_b, lp / np.log(2.0), S)
    g, ll = _occupancy(a, b, lp / np.log(2.0), S)
    return g, g_ref, ll, ll_ref


def test_per_step_renormalisation_keeps_every_node():
    for scale in (2.0, 6.0):
        g, g_ref, ll, ll_ref = _run(400, 40, 64, scale, K=1)
        assert abs(ll - ll_ref) <= 1e-7 * abs(ll_ref)
        big = g_ref > 1e-3
        assert (np.abs(g - g_ref)[big] / g_ref[big]).max() < 1e-4
        assert not ((g == 0) & (g_ref > 1e-6)).any()
        np.testing.assert_allclose(g.sum(1), 1.0, atol=1e-4)      # occupancies of a frame sum to one


def test_sparser_renormalisation_fails_on_peaky_input():
    """Why the exponent extraction stays inside the step: renormalising every 8th step loses lattice nodes."""
    g, g_ref, ll, ll_ref = _run(400, 40, 64, 6.0, K=8)
    big = g_ref > 1e-3
    lost = ((g == 0) & (g_ref > 1e-6)).any()
    assert lost or not np.isfinite(ll) or (np.abs(g - g_ref)[big] / g_ref[big]).max() > 1e-3


def test_the_kernels_own_arithmetic_header_on_the_host(tmp_path):
    """statecatcher_b200/csrc/sc_ctc_lin_math.h — the exponent-field arithmetic the experimental kernel compiles —
    built as plain C++ (tests/ctc_lin_host.cpp) and run over whole utterances: log2(alpha) of every node against
    the fp64 recursion, including masked emissions (probability zero) and a lattice that dies completely."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    assert gxx, "g++ is part of this image"
    exe = tmp_path / "ctc_lin_host"
    subprocess.run([gxx, "-O2", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "ctc_lin_host.cpp")], check=True)

    def run(lp2, ext):
        T, S = lp2.shape
        skip = np.zeros(S, np.uint8)
        skip[2:] = (ext[2:] != 0) & (ext[2:] != ext[:-2])
        src, dst = tmp_path / "in.bin", tmp_path / "out.bin"
        with open(src, "wb") as f:
            f.write(np.array([T, S], np.int32).tobytes() + skip.tobytes() + np.ascontiguousarray(lp2, np.float32).tobytes())
        subprocess.run([str(exe), str(src), str(dst)], check=True)
        return np.fromfile(dst, np.float64).reshape(T, S)

    for scale, T, U, V in ((2.0, 600, 40, 64), (6.0, 600, 40, 64), (2.0, 50, 24, 9)):
        logits, y, ext, lp = _case(T, U, V, scale, 3)
        shift = lp.max(1, keepdims=True)
        lp2 = ((lp - shift) / np.log(2.0)).astype(np.float32)
        got = run(lp2, ext) + (np.cumsum(shift[:, 0]) / np.log(2.0))[:, None]
        # reference on the SAME fp32 emissions, so that only the recursion's arithmetic is compared
        ref = (study.fp64_log_domain(lp2.astype(np.float64) * np.log(2.0), ext) / np.log(2.0)
               + (np.cumsum(shift[:, 0]) / np.log(2.0))[:, None])
        live = np.isfinite(ref)
        assert (np.isfinite(got) == live).all()                  # same nodes reachable, none lost
        # values are log2 of probabilities thousands of binades apart: compare the probabilities' ratio
        assert np.abs(got[live] - ref[live]).max() < 3e-4, (scale, np.abs(got[live] - ref[live]).max())
    # masked emissions: label 1 can never be emitted -> every node from the first label on is dead for good
    lp2 = np.zeros((6, 5), np.float32)
    lp2[:, 1] = -np.inf
    got = run(lp2, np.array([0, 1, 0, 2, 0]))
    assert np.isfinite(got[:, 0]).all() and not np.isfinite(got[:, 1:]).any()
    np.testing.assert_allclose(got[:, 0], 0.0, atol=1e-6)
