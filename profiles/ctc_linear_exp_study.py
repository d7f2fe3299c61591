#!/usr/bin/env python
"""Feasibility study on the CPU (numpy emulation of fp32 arithmetic): can the CTC alpha/beta recursion run in
the LINEAR domain with a per-node exponent (alpha = m * 2^e, fp32 mantissa, int exponent) and still give the
gradient within rtol 1e-4 of the fp64 log-domain result at the cfg2 shape (T=3000, U<=150, V=1024)?

Why: the shipped recursion's step is min/max -> ex2 x2 -> lg2 (three dependent MUFU round trips, 285 cycles per
step measured); a column-scaled linear recursion was rejected because nodes 2^-125 below the column maximum
underflow although the other direction's mass makes their occupancy matter.  A per-node exponent
removes that failure; this script measures what fp32 mantissas cost in accuracy.

    python profiles/ctc_linear_exp_study.py [T] [U] [V] [renorm_every]
"""
import sys

import numpy as np

f32 = np.float32


def ext_labels(y):
    ext = np.zeros(2 * len(y) + 1, dtype=np.int64)
    ext[1::2] = y
    return ext


def fp64_log_domain(lp, ext):
    """Reference: log-domain recursion in fp64 (natural log).  lp [T,S] = log p of the lattice labels."""
    T, S = lp.shape
    skip = np.zeros(S, bool)
    skip[2:] = (ext[2:] != 0) & (ext[2:] != ext[:-2])
    NEG = -np.inf
    alpha = np.full((T, S), NEG)
    alpha[0, :2] = lp[0, :2]
    for t in range(1, T):
        a = alpha[t - 1]
        b = np.concatenate([[NEG], a[:-1]])
        c = np.where(skip, np.concatenate([[NEG, NEG], a[:-2]]), NEG)
        m = np.maximum(np.maximum(a, b), c)
        ms = np.where(np.isfinite(m), m, 0.0)
        alpha[t] = ms + np.log(np.exp(a - ms) + np.exp(b - ms) + np.exp(c - ms)) + lp[t]
        alpha[t][~np.isfinite(m)] = NEG
    return alpha


def linear_exp_domain(p, ext, renorm_every=1):
    """alpha = m * 2^e: fp32 mantissa m, int32 exponent e, every operation rounded to fp32.
    p [T,S] fp32 linear emission probabilities of the lattice labels (rows pre-scaled so that max = 1, the
    shift is added back in the exponent domain by the caller).  Returns log2(alpha) as float64 for comparison."""
    T, S = p.shape
    skip = np.zeros(S, bool)
    skip[2:] = (ext[2:] != 0) & (ext[2:] != ext[:-2])
    m = np.zeros(S, f32)
    e = np.zeros(S, np.int32)
    m[:2] = p[0, :2]
    out = np.full((T, S), -np.inf)

    def store(t):
        with np.errstate(divide="ignore"):
            out[t] = np.where(m > 0, np.log2(m.astype(np.float64)) + e, -np.inf)

    def renorm():
        nonlocal m, e
        mm, de = np.frexp(m)                       # integer ops on the exponent field in the kernel
        m = mm.astype(f32)
        e = np.where(m > 0, e + de.astype(np.int32), 0).astype(np.int32)

    renorm()
    store(0)
    BIG = np.int32(-(1 << 30))
    for t in range(1, T):
        mb = np.concatenate([[f32(0)], m[:-1]]).astype(f32)
        eb = np.concatenate([[0], e[:-1]]).astype(np.int32)
        mc = np.where(skip, np.concatenate([[f32(0), f32(0)], m[:-2]]), f32(0)).astype(f32)
        ec = np.concatenate([[0, 0], e[:-2]]).astype(np.int32)
        ea = np.where(m > 0, e, BIG)
        ebb = np.where(mb > 0, eb, BIG)
        ecc = np.where(mc > 0, ec, BIG)
        emax = np.maximum(np.maximum(ea, ebb), ecc)
        live = emax > BIG

        def aligned(mm, ee):                        # m * 2^(e - emax): exponent-field subtract, flush below 2^-40
            d = np.where(live, ee.astype(np.int64) - emax, 0)
            d = np.clip(d, -60, 0).astype(np.int32)
            return np.where((mm > 0) & (d > -40), np.ldexp(mm, d), f32(0)).astype(f32)

        s = (aligned(m, ea) + aligned(mb, ebb)).astype(f32)
        s = (s + aligned(mc, ecc)).astype(f32)
        m = (s * p[t]).astype(f32)
        e = np.where(live, emax, 0).astype(np.int32)
        if t % renorm_every == 0:
            renorm()
        store(t)
    return out


def kernel_rows(p, lp2_frac_int, ext, with_emission, EB=16):
    """What ctc_alpha_beta_lin_kernel STORES: fp32 rows lg2(m) + float(e - eref), eref = largest live exponent of
    the previous column at the start of each EB-row emission block; mantissa/exponent handling as in the kernel
    (emission split into integer part -> exponent and 2^fraction in [1,2), terms below 2^-60 dropped)."""
    pf, ei = lp2_frac_int
    T, S = pf.shape
    skip = np.zeros(S, bool)
    skip[2:] = (ext[2:] != 0) & (ext[2:] != ext[:-2])
    DEAD = np.int64(-(1 << 28))
    m = np.zeros(S, f32)
    e = np.full(S, DEAD)
    rows = np.full((T, S), -np.inf, dtype=f32)

    def renorm(v, ebase):
        mm, de = np.frexp(v)
        live = v > 0
        return np.where(live, mm, 0).astype(f32), np.where(live, ebase + de, DEAD)

    def sc(mm, d):
        return np.where(d < -60, f32(0), mm * np.exp2(np.maximum(d, -60)).astype(f32)).astype(f32)

    m[:2], e[:2] = renorm(pf[0, :2], ei[0, :2])
    with np.errstate(divide="ignore"):
        rows[0, :2] = (np.log2(m[:2]).astype(f32) + e[:2].astype(f32)) if with_emission else f32(0)
    eref = 0
    for t in range(1, T):
        if t % EB == 0 and (e > DEAD).any():
            eref = e[e > DEAD].max()
        mb = np.concatenate([[f32(0)], m[:-1]]); eb = np.concatenate([[DEAD], e[:-1]])
        mc = np.where(skip, np.concatenate([[f32(0), f32(0)], m[:-2]]), f32(0))
        ec = np.where(skip, np.concatenate([[DEAD, DEAD], e[:-2]]), DEAD)
        emax = np.maximum(np.maximum(e, eb), ec)
        s = ((sc(m, e - emax) + sc(mb, eb - emax)).astype(f32) + sc(mc, ec - emax)).astype(f32)
        v = (s * pf[t]).astype(f32)
        m, e_new = renorm(v, emax + ei[t])
        with np.errstate(divide="ignore"):
            if with_emission:
                rows[t] = np.where(m > 0, np.log2(m).astype(f32) + (e_new - eref).astype(f32), -np.inf)
            else:
                rows[t] = np.where(s > 0, np.log2(s).astype(f32) + (emax - eref).astype(f32), -np.inf)
        e = e_new
    return rows


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    U = int(sys.argv[2]) if len(sys.argv) > 2 else 150
    V = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    K = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    rng = np.random.default_rng(0)
    for scale, what in ((2.0, "random logits x2 (bench input)"), (6.0, "peaky logits x6")):
        logits = (rng.standard_normal((T, V)) * scale).astype(np.float32).astype(np.float64)
        y = rng.integers(1, V, U)
        y[5] = y[4]                                 # one repeated label (no skip transition)
        ext = ext_labels(y)
        lse = np.log(np.exp(logits - logits.max(1, keepdims=True)).sum(1)) + logits.max(1)
        lp = logits[:, ext] - lse[:, None]          # [T,S] natural log, fp64
        ref_a = fp64_log_domain(lp, ext) / np.log(2.0)
        ref_b = fp64_log_domain(lp[::-1, ::-1], ext[::-1])[::-1, ::-1] / np.log(2.0)
        # fp32 inputs of the kernel: per-frame shift to the largest lattice emission, linear probabilities
        shift = lp.max(1, keepdims=True)
        p = np.exp((lp - shift).astype(f32)).astype(f32)
        csum = np.cumsum(shift[:, 0]) / np.log(2.0)
        a = linear_exp_domain(p, ext, K) + csum[:, None]
        b = (linear_exp_domain(p[::-1, ::-1], ext[::-1], K)[::-1, ::-1]
             + (np.cumsum(shift[::-1, 0])[::-1] / np.log(2.0))[:, None])
        S = len(ext)
        ll_ref = np.logaddexp2(ref_a[-1, S - 1], ref_a[-1, S - 2])
        ll = np.logaddexp2(a[-1, S - 1], a[-1, S - 2])
        # occupancy gamma[t,s] = alpha*beta / (p * Z); gradient = softmax - sum_s gamma -> compare gamma itself
        lp2 = lp / np.log(2.0)
        with np.errstate(invalid="ignore", over="ignore"):
            g_ref = np.exp2(ref_a + ref_b - lp2 - ll_ref)
            g = np.exp2(a + b - lp2 - ll)
        g_ref = np.nan_to_num(g_ref)
        g = np.nan_to_num(g)
        err = np.abs(g - g_ref)
        dead_wrong = int(((g == 0) & (g_ref > 1e-6)).sum())
        print(f"{what}: T={T} U={U} V={V} renorm every {K}")
        print(f"  log2-likelihood: fp64 {ll_ref:.6f}  linear+exp {ll:.6f}  rel err {abs(ll - ll_ref) / abs(ll_ref):.2e}")
        print(f"  occupancy: max abs err {err.max():.3e} (values up to {g_ref.max():.3f}), "
              f"max rel err where gamma > 1e-3: {(err / np.maximum(g_ref, 1e-30))[g_ref > 1e-3].max():.3e}, "
              f"row-sum err {np.abs(g.sum(1) - 1).max():.3e}, nodes lost (0 vs > 1e-6): {dead_wrong}")
        if K == 1:
            # the kernel's own storage format, then the gradient pass's per-frame normalisation in fp32
            lp2s = ((lp - shift) / np.log(2.0)).astype(f32)
            fl = np.floor(lp2s)
            split = (np.exp2((lp2s - fl).astype(f32)).astype(f32), fl.astype(np.int64))
            ra = kernel_rows(p, split, ext, True)
            rb = kernel_rows(p[::-1, ::-1], (split[0][::-1, ::-1], split[1][::-1, ::-1]), ext[::-1], False)[::-1, ::-1]
            ab = (ra + rb).astype(f32)                       # beta rows carry no emission: alpha + beta is the log-occupancy
            ab = np.where(np.isfinite(ab), ab, -np.inf)
            gk = np.exp2((ab - ab.max(1, keepdims=True)).astype(f32)).astype(f32)
            gk = gk / gk.sum(1, keepdims=True)
            errk = np.abs(gk - g_ref)
            print(f"  kernel storage format (fp32 rows relative to a per-16-row exponent, per-frame normalisation): "
                  f"max abs err {errk.max():.3e}, max rel err where gamma > 1e-3: "
                  f"{(errk / np.maximum(g_ref, 1e-30))[g_ref > 1e-3].max():.3e}, largest |stored value| "
                  f"{np.abs(ra[np.isfinite(ra)]).max():.1f} / {np.abs(rb[np.isfinite(rb)]).max():.1f}")


if __name__ == "__main__":
    main()
