"""
NumPy model of csrc/sq_nufft.cu: type-1 NUFFT of point particles onto the non-negative integer modes of a periodic
box, with centred modes, the exponential-of-semicircle kernel and per-axis deconvolution.  Used to derive the
(nf, w, beta) rule in Nufft::init and to check the formulas; run it to print the error of the rule against the direct
sum for a range of grid sizes.  CPU only; not imported by the package.
"""

import math

import numpy as np


def pick(n, eps):
    cands = sorted([2 ** k for k in range(5, 11)] + [3 * 2 ** k for k in range(4, 9)])
    nf = next((x for x in cands if x >= 2.0 * n), 0)
    if nf == 0 or nf > 160:
        nf = next(x for x in cands if x >= 1.5 * n)
    sigma = nf / n
    w = math.ceil((math.log(1 / eps) + math.log(5)) / (math.pi * math.sqrt(1 - 1 / sigma)))
    w = max(4, min(16, w))
    beta = 0.975 * math.pi * w * (1 - 1 / (2 * sigma))
    return nf, w, beta


def phihat(xi, beta, nquad=96):
    x, wt = np.polynomial.legendre.leggauss(nquad)
    th = 0.25 * np.pi * (x + 1)
    f = np.exp(beta * (np.cos(th) - 1)) * np.cos(th)
    return 2 * 0.25 * np.pi * (wt * f) @ np.cos(np.outer(np.sin(th), xi))


def kernel_values(x, w, beta):
    """l0 and the w kernel values of fine-grid coordinate(s) x."""
    l0 = np.ceil(x - w / 2)
    z = (l0[..., None] + np.arange(w) - x[..., None]) * (2 / w)
    return l0.astype(int), np.exp(beta * (np.sqrt(np.maximum(1 - z * z, 0)) - 1))


def nufft3d(s, n, eps=1e-9):
    """rho[b, a, c] = sum_j exp(2 pi i (a sx + b sy + c sz)), a, b, c in [0, n) (the reference's meshgrid order)."""
    nf, w, beta = pick(n, eps)
    k0 = n // 2
    g = s * nf
    l0, val = zip(*(kernel_values(g[:, d], w, beta) for d in range(3)))
    c = np.exp(2j * np.pi * k0 * s.sum(axis=1))
    grid = np.zeros((nf, nf, nf), complex)
    for ix in range(w):
        for iy in range(w):
            for iz in range(w):
                np.add.at(grid, ((l0[0] + ix) % nf, (l0[1] + iy) % nf, (l0[2] + iz) % nf),
                          c * val[0][:, ix] * val[1][:, iy] * val[2][:, iz])
    kk = np.arange(n) - k0
    G = np.fft.ifftn(grid) * nf ** 3
    G = G[np.ix_(kk % nf, kk % nf, kk % nf)]          # [a, b, c]
    dec = 1 / (0.5 * w * phihat(np.pi * kk * w / nf, beta))
    G = G * dec[:, None, None] * dec[None, :, None] * dec[None, None, :]
    return np.transpose(G, (1, 0, 2))


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    s = rng.random((400, 3))
    for n in (5, 12, 16, 21):
        for eps in (1e-9, 1e-13):
            got = nufft3d(s, n, eps)
            k = np.arange(n)
            ph = [np.exp(2j * np.pi * np.outer(k, s[:, d])) for d in range(3)]
            want = np.einsum("aj,bj,cj->bac", *ph)
            print(n, eps, pick(n, eps)[:2], np.linalg.norm(got - want) / np.linalg.norm(want))
