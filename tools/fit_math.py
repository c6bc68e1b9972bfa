#!/usr/bin/env python3
"""Fit the float32 polynomial kernels used by the portable transcendental functions.

The oracle (oracle/qo_math.h) and the CUDA kernels (quatro_b200/csrc/qb_math.cuh) each carry
their own copy of atan2f / acosf / sincosf written as a FIXED sequence of IEEE float32
add/mul/div/sqrt operations (no FMA contraction), so CPU and GPU produce bit-identical
Darboux-angle bins and covariance eigen-roots.  This script derives the coefficients
(least squares on Chebyshev nodes in float64, rounded to float32) and reports the error of the
float32 evaluation against the float64 truth.  Run: python tools/fit_math.py
"""
import numpy as np

f32 = np.float32


def cheb_nodes(a, b, n):
    k = np.arange(n)
    x = np.cos(np.pi * (2 * k + 1) / (2 * n))
    return 0.5 * (a + b) + 0.5 * (b - a) * x


def fit(fun, a, b, deg, n=4000):
    z = cheb_nodes(a, b, n)
    y = fun(z)
    V = np.vander(z, deg + 1, increasing=True)
    c, *_ = np.linalg.lstsq(V, y, rcond=None)
    return c.astype(np.float32)


def horner32(c, z):
    z = z.astype(f32)
    acc = np.full_like(z, c[-1], dtype=f32)
    for k in range(len(c) - 2, -1, -1):
        acc = (acc * z).astype(f32)
        acc = (acc + c[k]).astype(f32)
    return acc


def ulp_err(approx32, truth64):
    t = truth64
    a = approx32.astype(np.float64)
    ulp = np.spacing(np.abs(t).astype(f32)).astype(np.float64)
    return np.max(np.abs(a - t) / ulp)


def show(name, c):
    print(f"// {name}")
    print("  " + ", ".join(f"{float(v):.9e}f" for v in c))


def main():
    rng = np.random.default_rng(0)
    # ---- atan on |t| <= tan(pi/8): atan(t) = t + t*z*P(z), z = t*t
    T = np.tan(np.pi / 8) * 1.0001

    def fa(z):
        t = np.sqrt(np.maximum(z, 1e-300))
        return np.where(z < 1e-12, -1.0 / 3.0 + z / 5.0, (np.arctan(t) - t) / (t * z))

    ca = fit(fa, 0.0, T * T, 6)
    show("atan P(z), 7 coeffs", ca)
    t = rng.uniform(-T, T, 2_000_000).astype(f32)
    z = (t * t).astype(f32)
    p = horner32(ca, z)
    r = (t + (t * (z * p).astype(f32)).astype(f32)).astype(f32)
    print("//   atan core max ulp err:", ulp_err(r, np.arctan(t.astype(np.float64))))

    # ---- asin on [0, 0.5]: asin(x) = x + x*z*Q(z), z = x*x in [0, 0.25]
    def fs(z):
        x = np.sqrt(np.maximum(z, 1e-300))
        return np.where(z < 1e-12, 1.0 / 6.0 + 3.0 * z / 40.0, (np.arcsin(x) - x) / (x * z))

    cs = fit(fs, 0.0, 0.2501, 6)
    show("asin Q(z), 7 coeffs", cs)
    x = rng.uniform(0, 0.5, 2_000_000).astype(f32)
    z = (x * x).astype(f32)
    q = horner32(cs, z)
    r = (x + (x * (z * q).astype(f32)).astype(f32)).astype(f32)
    print("//   asin core max ulp err:", ulp_err(r, np.arcsin(x.astype(np.float64))))

    # ---- sin / cos on [0, 1.1]: sin(x) = x + x*z*S(z); cos(x) = 1 - z/2 + z*z*C(z)
    def fsin(z):
        x = np.sqrt(np.maximum(z, 1e-300))
        return np.where(z < 1e-12, -1.0 / 6.0 + z / 120.0, (np.sin(x) - x) / (x * z))

    def fcos(z):
        x = np.sqrt(np.maximum(z, 1e-300))
        return np.where(z < 1e-6, 1.0 / 24.0 - z / 720.0, (np.cos(x) - 1.0 + 0.5 * z) / (z * z))

    csn = fit(fsin, 0.0, 1.21, 5)
    ccs = fit(fcos, 0.0, 1.21, 5)
    show("sin S(z), 6 coeffs", csn)
    show("cos C(z), 6 coeffs", ccs)
    x = rng.uniform(0, 1.1, 2_000_000).astype(f32)
    z = (x * x).astype(f32)
    s = horner32(csn, z)
    rs = (x + (x * (z * s).astype(f32)).astype(f32)).astype(f32)
    c = horner32(ccs, z)
    rc = ((f32(1.0) - (f32(0.5) * z).astype(f32)).astype(f32) + ((z * z).astype(f32) * c).astype(f32)).astype(f32)
    print("//   sin max ulp err:", ulp_err(rs, np.sin(x.astype(np.float64))))
    print("//   cos max ulp err:", ulp_err(rc, np.cos(x.astype(np.float64))))


if __name__ == "__main__":
    main()
