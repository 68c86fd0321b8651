"""Numerical prototype (numpy, emulated fp32 + FMA) of the series-kernel recurrences.

Development aid, not part of the product or the tests: it answers "which recurrence / anchor period
keeps the L=2000 fp32 series within 1e-5 of the fp64 series on the E-set" before CUDA is written.
    python oracle/proto_series_fp32.py
"""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from oracle import so3_oracle as O  # noqa: E402

f32 = np.float32


def fma(a, b, c):
    return (a.astype(np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)


def mul(a, b):
    return (a * b).astype(f32)


def ex2_approx(x, rng):
    y = np.exp2(x.astype(np.float64))
    y = y * (1 + rng.uniform(-2.0 ** -22, 2.0 ** -22, size=x.shape))  # MUFU.EX2 ~2 ulp
    return y.astype(f32)


def sincos_exact_of_product(m, om):
    a = np.float64(m) * om.astype(np.float64)
    return np.sin(a).astype(f32), np.cos(a).astype(f32)


def series_chi(om, eps, L, K, rng, ex2=True):
    n = om.shape[0]
    cexp = (-(eps.astype(np.float64) ** 2) * np.log2(np.e)).astype(f32)
    sw, cw = np.sin(om.astype(np.float64)).astype(f32), np.cos(om.astype(np.float64)).astype(f32)
    chi = np.ones(n, f32)
    D = np.zeros(n, f32)
    F = np.full(n, 0.5, f32)  # l = 0: (0+1/2) * e_0 * chi_0 = 0.5
    Fp = np.zeros(n, f32)
    s = np.zeros(n, f32)
    c = np.ones(n, f32)
    for m in range(1, L):
        if K and (m % K == 0):
            s, c = sincos_exact_of_product(m, om)
        else:
            s, c = fma(s, cw, mul(c, sw)), fma(c, cw, -mul(s, sw).astype(np.float64))
        mf = f32(m)
        chi = fma(c, f32(2.0), chi)
        D = fma(s, mf, D)
        t = f32(m * (m + 1))
        x = mul(np.full(n, t, f32), cexp)
        e = ex2_approx(x, rng)
        p = mul(e, f32(m + 0.5))
        F = fma(p, chi, F)
        Fp = fma(p, D, Fp)
    f = 2.0 * F.astype(np.float64)
    g = -2.0 * Fp.astype(np.float64) / F.astype(np.float64)
    return f, g


def series_sc(om, eps, L, K, rng):
    """S/C form: sum w_l sin((l+1/2)w), sum w_l (l+1/2) cos((l+1/2) w); f = S/sin(w/2)."""
    n = om.shape[0]
    cexp = (-(eps.astype(np.float64) ** 2) * np.log2(np.e)).astype(f32)
    sw, cw = np.sin(om.astype(np.float64)).astype(f32), np.cos(om.astype(np.float64)).astype(f32)
    S = np.zeros(n, f32)
    C = np.zeros(n, f32)
    h0 = (om.astype(np.float64) / 2)
    s, c = np.sin(h0).astype(f32), np.cos(h0).astype(f32)
    sh, ch = s.copy(), c.copy()
    for l in range(0, L):
        if l > 0:
            if K and (l % K == 0):
                a = (l + 0.5) * om.astype(np.float64)
                s, c = np.sin(a).astype(f32), np.cos(a).astype(f32)
            else:
                s, c = fma(s, cw, mul(c, sw)), fma(c, cw, -mul(s, sw).astype(np.float64))
        t = f32(l * (l + 1))
        e = ex2_approx(mul(np.full(n, t, f32), cexp), rng)
        p = mul(e, f32(l + 0.5))
        S = fma(p, s, S)
        q = mul(p, f32(l + 0.5))
        C = fma(q, c, C)
    S64, C64 = S.astype(np.float64), C.astype(np.float64)
    f = 2 * S64 / sh
    g = (C / S).astype(np.float64) - 0.5 * (ch / sh).astype(np.float64)
    return f, g


def eset(n, rng, kmax=4.0):
    eps = np.exp(rng.uniform(np.log(6.4e-3), 0.0, n)).astype(f32)
    k = rng.uniform(0, kmax, n)
    om = np.minimum(eps * np.sqrt(2.0) * k, 3.0).astype(f32)
    return om, eps


def main():
    rng = np.random.default_rng(0)
    n = 4000
    om, eps = eset(n, rng)
    ft, gt = O.igso3_series(om.astype(np.float64), eps.astype(np.float64), 2000)
    for name, fn in (("chi", series_chi), ("sc", series_sc)):
        for K in (0, 64, 32, 16, 8):
            f, g = fn(om, eps, 2000, K, rng)
            ef = np.abs(f - ft) / ft
            eg = np.abs(g - gt) / np.maximum(np.abs(gt), 1e-300)
            egs = np.abs(g - gt) / np.maximum(np.abs(gt), 1.0 / np.maximum(om, 1e-6))  # vs 1/omega scale
            print(f"{name:4s} K={K:3d}  f: max {ef.max():.2e} p99 {np.quantile(ef, .99):.2e} med {np.median(ef):.1e} |"
                  f" g rel: max {eg.max():.2e} p99 {np.quantile(eg, .99):.2e} | g/(1/w): max {egs.max():.2e}")


if __name__ == "__main__" and len(sys.argv) == 1:
    main()


def series_naive_best(om, eps, L, kahan=False):
    """every term exact (fp64) then rounded to fp32; fp32 accumulation -> floor for any fp32 series."""
    n = om.shape[0]
    S = np.zeros(n, f32); comp = np.zeros(n, f32)
    om64, v = om.astype(np.float64), eps.astype(np.float64) ** 2
    for l in range(L):
        term = ((2 * l + 1) * np.exp(-l * (l + 1) * v) * np.sin((l + 0.5) * om64)).astype(f32)
        if kahan:
            y = (term - comp).astype(f32); t = (S + y).astype(f32); comp = ((t - S).astype(f32) - y).astype(f32); S = t
        else:
            S = (S + term).astype(f32)
    return S.astype(np.float64) / np.sin(om64 / 2)


def binned():
    rng = np.random.default_rng(1)
    n = 6000
    om, eps = eset(n, rng, 4.0)
    k = om / (np.sqrt(2.0) * eps)
    ft, gt = O.igso3_series(om.astype(np.float64), eps.astype(np.float64), 2000)
    fb = series_naive_best(om, eps, 2000)
    fk = series_naive_best(om, eps, 2000, kahan=True)
    fc, gc = series_chi(om, eps, 2000, 16, rng)
    fs, gs = series_sc(om, eps, 2000, 16, rng)
    print("k-bin      naive-best  kahan-best   chi16-f    sc16-f    chi16-g    sc16-g(abs*w)")
    for lo in np.arange(0, 4, 0.5):
        m = (k >= lo) & (k < lo + 0.5)
        e = lambda a: (np.abs(a - ft) / ft)[m].max()
        eg = lambda a: (np.abs(a - gt) / np.maximum(np.abs(gt), 1e-300))[m].max()
        print(f"[{lo:.1f},{lo+.5:.1f})  {e(fb):.2e}   {e(fk):.2e}   {e(fc):.2e}  {e(fs):.2e}  {eg(gc):.2e}  {(np.abs(gs-gt)*om)[m].max():.2e}")
    # by eps
    print("eps-bin")
    for lo, hi in ((6e-3, 2e-2), (2e-2, 6e-2), (6e-2, 0.2), (0.2, 0.5), (0.5, 1.01)):
        m = (eps >= lo) & (eps < hi) & (k < 3)
        e = lambda a: (np.abs(a - ft) / ft)[m].max()
        print(f"[{lo},{hi}) k<3: naive {e(fb):.2e} chi {e(fc):.2e} sc {e(fs):.2e}")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "binned":
    binned()


def series_w(om, eps, L, K, rng):
    """Suffix-weight (cosine-series) form, evaluated downward in m:
       P_m = sum_{l>=m} (l+1/2) e_l ;  f = 2 (P_0 + 2 sum_{m>=1} P_m cos(m w)) ; f' = -4 sum m P_m sin(m w)."""
    n = om.shape[0]
    cexp = (-(eps.astype(np.float64) ** 2) * np.log2(np.e)).astype(f32)
    sw, cw = np.sin(om.astype(np.float64)).astype(f32), np.cos(om.astype(np.float64)).astype(f32)
    P = np.zeros(n, f32); F = np.zeros(n, f32); Fp = np.zeros(n, f32)
    s, c = sincos_exact_of_product(L - 1, om)
    for m in range(L - 1, 0, -1):
        if m != L - 1:
            if K and (m % K == K - 1):
                s, c = sincos_exact_of_product(m, om)
            else:  # rotate by -w
                s, c = fma(s, cw, -mul(c, sw).astype(np.float64)), fma(c, cw, mul(s, sw))
        t = f32(m * (m + 1))
        e = ex2_approx(mul(np.full(n, t, f32), cexp), rng)
        p = mul(e, f32(m + 0.5))
        P = (P + p).astype(f32)
        F = fma(P, c, F)
        Fp = fma(mul(P, f32(m)), s, Fp)
    P0 = (P + f32(0.5)).astype(f32)
    Ftot = fma(F, f32(2.0), P0)
    f = 2.0 * Ftot.astype(np.float64)
    g = -2.0 * Fp.astype(np.float64) / Ftot.astype(np.float64)
    return f, g


def binned_w():
    rng = np.random.default_rng(1)
    n = 6000
    om, eps = eset(n, rng, 4.0)
    k = om / (np.sqrt(2.0) * eps)
    ft, gt = O.igso3_series(om.astype(np.float64), eps.astype(np.float64), 2000)
    res = {"chi32": series_chi(om, eps, 2000, 32, rng), "w32": series_w(om, eps, 2000, 32, rng), "w16": series_w(om, eps, 2000, 16, rng)}
    print("k-bin     " + "".join(f"{k_:>22s}" for k_ in res))
    for lo in np.arange(0, 4, 0.5):
        m = (k >= lo) & (k < lo + 0.5)
        row = ""
        for f, g in res.values():
            row += f"   f {(np.abs(f - ft) / ft)[m].max():.1e} g {(np.abs(g - gt) / np.abs(gt))[m].max():.1e}"
        print(f"[{lo:.1f},{lo+.5:.1f})" + row)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "w":
    binned_w()
