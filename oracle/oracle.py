"""
CPU oracle for the S(q) / g(r) hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import this module.  Nothing under
``mdcraft_b200/`` does; the product path fails loudly without its CUDA library.

It wraps ``oracle/oracle.c`` (the arithmetic, see that file's header for what
is pinned against the reference and what is "parity unpinned") and restates
the class-level bookkeeping of the reference in NumPy:

* ``structure_factor``  ~ ``StructureFactor.__init__/_prepare/_single_frame/
  _conclude``  (``/root/reference/src/mdcraft/analysis/structure.py:1822-2017``)
* ``rdf``               ~ ``RadialDistributionFunction._prepare/_single_frame/
  _conclude``  (``structure.py:849-1010``)
"""

from __future__ import annotations

import ctypes
import os
import subprocess
from itertools import combinations_with_replacement
from typing import Iterable, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle.c with the committed Makefile (gcc, OpenMP, no FMA contraction)."""
    src = os.path.join(_HERE, "oracle.c")
    if (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    ):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        c_dp = ctypes.POINTER(ctypes.c_double)
        c_fp = ctypes.POINTER(ctypes.c_float)
        c_lp = ctypes.POINTER(ctypes.c_int64)
        c_ip = ctypes.POINTER(ctypes.c_int)
        L.orc_delta_fourier_transform_sum.argtypes = [c_dp, ctypes.c_int64, c_dp, ctypes.c_int64, c_dp]
        L.orc_sq_single_frame.argtypes = [c_dp, ctypes.c_int64, c_dp, c_lp, ctypes.c_int, c_ip, ctypes.c_int, c_dp]
        L.orc_histogram_bin_edges.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int, c_dp]
        L.orc_histogram.argtypes = [c_dp, ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_double, c_lp]
        L.orc_pair_distance.argtypes = [c_fp, c_fp, c_fp]
        L.orc_pair_distance.restype = ctypes.c_double
        L.orc_radial_histogram.argtypes = [
            c_fp, ctypes.c_int64, c_fp, ctypes.c_int64, c_fp, ctypes.c_int,
            ctypes.c_double, ctypes.c_double, ctypes.c_int64, ctypes.c_int64, c_lp,
        ]
        L.orc_radial_histogram_triclinic.argtypes = L.orc_radial_histogram.argtypes
        L.orc_pair_distance_triclinic.argtypes = [c_fp, c_fp, c_fp]
        L.orc_pair_distance_triclinic.restype = ctypes.c_double
        L.orc_triclinic_vectors.argtypes = [c_fp, c_fp]
        L.orc_triclinic_vectors.restype = ctypes.c_int
        L.orc_num_threads.restype = ctypes.c_int
        L.orc_set_num_threads.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _lp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))


# ----------------------------------------------------------------------------
# S(q)
# ----------------------------------------------------------------------------


def delta_fourier_transform_sum(qs: np.ndarray, rs: np.ndarray) -> np.ndarray:
    """``structure.py:1240-1280``: ``F[i] = sum_j exp(1j q_i . r_j)`` (complex128)."""
    qs = np.ascontiguousarray(qs, dtype=np.float64)
    rs = np.ascontiguousarray(rs, dtype=np.float64)
    out = np.empty((qs.shape[0], 2), dtype=np.float64)
    lib().orc_delta_fourier_transform_sum(_dp(qs), qs.shape[0], _dp(rs), rs.shape[0], _dp(out))
    return out.view(np.complex128).reshape(-1)


def wavevector_grid(
    dimensions: Sequence[float],
    n_points: int = 32,
    q_max: Optional[float] = None,
    n_surfaces: Optional[int] = None,
    n_surface_points: int = 8,
):
    """
    ``structure.py:1822-1887``: first-octant reciprocal-lattice grid (+ optional
    off-lattice surface points, cubic boxes only) and the ``q_max`` filter.
    Returns ``(wavevectors (Nq,3), wavenumbers (Nq,))``.
    """
    dimensions = np.asarray(dimensions)
    if len(dimensions) != 3:
        raise ValueError("'dimensions' must have length 3.")
    if np.allclose(dimensions, dimensions[0]):
        grid = 2 * np.pi * np.arange(n_points) / dimensions[0]
        wavevectors = np.stack(np.meshgrid(grid, grid, grid), -1).reshape(-1, 3)
        if n_surfaces:
            n_theta, n_phi = _closest_factors(n_surface_points)
            theta = np.linspace(
                np.pi / (2 * n_theta + 4), np.pi / 2 - np.pi / (2 * n_theta + 4), n_theta
            )
            phi = np.linspace(
                np.pi / (2 * n_phi + 4), np.pi / 2 - np.pi / (2 * n_phi + 4), n_phi
            )
            extra = np.einsum(
                "o,tpd->otpd",
                grid[1 : n_surfaces + 1],
                np.stack(
                    (
                        np.sin(theta) * np.cos(phi)[:, None],
                        np.sin(theta) * np.sin(phi)[:, None],
                        np.tile(np.cos(theta)[None, :], (n_phi, 1)),
                    ),
                    axis=-1,
                ),
            ).reshape((n_surfaces * n_surface_points, 3))
            wavevectors = np.vstack((wavevectors, extra))
    else:
        wavevectors = np.stack(
            np.meshgrid(*[2 * np.pi * np.arange(n_points) / L for L in dimensions]),
            axis=-1,
        ).reshape(-1, 3)
    wavenumbers = np.linalg.norm(wavevectors, axis=1)
    if q_max is not None:
        keep = wavenumbers <= q_max
        wavevectors = wavevectors[keep]
        wavenumbers = wavenumbers[keep]
    return wavevectors, wavenumbers


def _closest_factors(value: int):
    """
    ``algorithm/utility.py:16-75`` get_closest_factors(value, 2, reverse=True):
    the reference's greedy assignment of prime factors (largest first), which
    is not always the closest pair (72 -> (12, 6)).
    """
    n_factors = 2
    rt = value ** (1 / n_factors)
    rt_int = int(np.round(rt))
    if np.isclose(rt, rt_int):
        return rt_int, rt_int
    primes, v, p = [], int(value), 2
    while p * p <= v:
        while v % p == 0:
            primes.append(p)
            v //= p
        p += 1
    if v > 1:
        primes.append(v)
    i = 0
    factors = [1] * n_factors
    for j, f in enumerate(primes[::-1]):
        while True:
            if i < n_factors:
                m = factors[i] * f
                if m <= rt_int or (j < n_factors and factors[i] == 1):
                    factors[i] = m
                    break
                i += 1
            else:
                factors[int(np.argmin(factors))] *= f
                break
    hi, lo = sorted(factors)[::-1]
    return hi, lo


def sq_pairs(n_groups: int, mode: Optional[str]):
    """``structure.py:1917-1921``."""
    if mode == "partial":
        return tuple(combinations_with_replacement(range(n_groups), 2))
    if mode == "pair":
        return ((0, n_groups - 1),)
    return ((None, None),)


def sq_accumulate(
    frames: Iterable[np.ndarray],
    group_sizes: Sequence[int],
    wavevectors: np.ndarray,
    pairs,
) -> tuple:
    """
    Per-frame accumulation (``structure.py:1940-1970``).  ``frames`` yields the
    concatenated float32 ``(N, 3)`` positions of all groups (group order).
    Returns ``(ssf_sum (n_pairs, Nq), n_frames)`` — un-normalised.
    """
    wavevectors = np.ascontiguousarray(wavevectors, dtype=np.float64)
    offsets = np.concatenate(([0], np.cumsum(group_sizes))).astype(np.int64)
    pr = np.array(
        [(-1, -1) if p[0] is None else p for p in pairs], dtype=np.int32
    ).reshape(-1, 2)
    ssf = np.zeros((len(pairs), wavevectors.shape[0]))
    n_frames = 0
    for pos in frames:
        pos64 = np.ascontiguousarray(pos, dtype=np.float64)  # structure.py:1943 (f32 -> f64)
        lib().orc_sq_single_frame(
            _dp(wavevectors), wavevectors.shape[0], _dp(pos64), _lp(offsets),
            len(group_sizes), _ip(pr), len(pairs), _dp(ssf),
        )
        n_frames += 1
    return ssf, n_frames


def sq_conclude(ssf_sum, n_frames, n_total, wavenumbers, unique=True, sort=True):
    """``structure.py:1993-2017`` verbatim semantics (the O(N_unique x Nq) isclose loop)."""
    ssf = ssf_sum / (n_frames * n_total)
    wn = np.unique(wavenumbers.round(11)) if unique else wavenumbers
    if unique:
        ssf = np.hstack(
            [ssf[:, np.isclose(q, wavenumbers)].mean(axis=1, keepdims=True) for q in wn]
        )
    if sort:
        order = np.argsort(wn)
        wn = wn[order]
        ssf = ssf[:, order]
    return wn, ssf


def structure_factor(
    frames: Iterable[np.ndarray],
    group_sizes: Sequence[int],
    dimensions: Sequence[float],
    *,
    mode: Optional[str] = None,
    n_points: int = 32,
    q_max: Optional[float] = None,
    n_surfaces: Optional[int] = None,
    n_surface_points: int = 8,
    wavevectors: Optional[np.ndarray] = None,
    sort: bool = True,
    unique: bool = True,
) -> dict:
    """Whole-class restatement; returns the reference's ``results`` keys."""
    if wavevectors is not None:
        wv = np.asarray(wavevectors, dtype=np.float64)
        wn = np.linalg.norm(wv, axis=1)
        if q_max is not None:
            keep = wn <= q_max
            wv, wn = wv[keep], wn[keep]
    else:
        wv, wn = wavevector_grid(dimensions, n_points, q_max, n_surfaces, n_surface_points)
    pairs = sq_pairs(len(group_sizes), mode)
    ssf_sum, n_frames = sq_accumulate(frames, group_sizes, wv, pairs)
    wn_out, ssf = sq_conclude(ssf_sum, n_frames, int(np.sum(group_sizes)), wn, unique, sort)
    return {"pairs": pairs, "ssf": ssf, "wavenumbers": wn_out, "wavevectors": wv, "ssf_sum": ssf_sum, "n_frames": n_frames}


def scsf(frames: Iterable[np.ndarray], n_chains: int, n_monomers: int, dimensions, *, n_points: int = 32,
         q_max: Optional[float] = None, wavevectors: Optional[np.ndarray] = None, sort: bool = True,
         unique: bool = True) -> dict:
    """``SingleChainStructureFactor`` for one group (``polymer.py:1382-1445``, form="exp")."""
    if wavevectors is not None:
        wv = np.asarray(wavevectors, dtype=np.float64)
        wn = np.linalg.norm(wv, axis=1)
        if q_max is not None:
            keep = wn <= q_max
            wv, wn = wv[keep], wn[keep]
    else:
        wv, wn = wavevector_grid(dimensions, n_points, q_max)
    acc = np.zeros((1, wv.shape[0]))
    n_frames = 0
    for pos in frames:
        chains = np.asarray(pos, dtype=np.float64).reshape(n_chains, n_monomers, 3)
        for chain in chains:
            rho = delta_fourier_transform_sum(wv, chain)
            acc[0] += (rho * rho.conj()).real
        n_frames += 1
    acc /= n_chains * n_monomers * n_frames
    wn_out = np.unique(wn.round(11))
    if unique:
        acc = np.hstack([acc[:, np.isclose(q, wn)].mean(axis=1, keepdims=True) for q in wn_out])
    else:
        wn_out = wn
    if sort and unique:
        order = np.argsort(wn_out)
        wn_out, acc = wn_out[order], acc[:, order]
    return {"scsf": acc, "wavenumbers": wn_out}


def isf(
    frames: Sequence[np.ndarray],
    group_sizes: Sequence[int],
    dimensions: Sequence[float],
    *,
    mode: Optional[str] = None,
    n_points: int = 32,
    q_max: Optional[float] = None,
    wavevectors: Optional[np.ndarray] = None,
    n_lags: Optional[int] = None,
    incoherent: bool = False,
    sort: bool = True,
    unique: bool = True,
) -> dict:
    """
    ``IntermediateScatteringFunction`` restated (``structure.py:2603-2828``, form="exp"): a sliding window of the
    last ``n_lags`` frames; ``cisf[lag] += Re(rho(t - lag) conj rho(t))``, ``iisf[lag] += Re sum exp(iq.(r(t) - r(t - lag)))``.
    """
    frames = [np.asarray(f, dtype=np.float64) for f in frames]
    n_frames = len(frames)
    if wavevectors is not None:
        wv = np.asarray(wavevectors, dtype=np.float64)
        wn = np.linalg.norm(wv, axis=1)
        if q_max is not None:
            keep = wn <= q_max
            wv, wn = wv[keep], wn[keep]
    else:
        wv, wn = wavevector_grid(dimensions, n_points, q_max)
    pairs = sq_pairs(len(group_sizes), mode)
    n_lags = int(min(n_lags or np.inf, n_frames))
    offsets = np.concatenate(([0], np.cumsum(group_sizes))).astype(int)
    groups = [slice(0, offsets[-1])] if mode is None else [slice(offsets[i], offsets[i + 1]) for i in range(len(group_sizes))]
    rho = [[delta_fourier_transform_sum(wv, f[g]) for g in groups] for f in frames]
    cisf = np.zeros((n_lags, len(pairs), wv.shape[0]))
    iisf = np.zeros((n_lags, len(groups), wv.shape[0])) if incoherent else None
    for t in range(n_frames):
        for lag in range(min(n_lags, t + 1)):
            s = t - lag
            for i, (j, k) in enumerate(pairs):
                j = 0 if j is None else j
                k = 0 if k is None else k
                if j == k:
                    cisf[lag, i] += (rho[s][j] * rho[t][j].conj()).real
                    if incoherent:
                        iisf[lag, j] += delta_fourier_transform_sum(wv, frames[t][groups[j]] - frames[s][groups[j]]).real
                else:
                    cisf[lag, i] += (rho[s][j] * rho[t][k].conj()).real + (rho[s][k] * rho[t][j].conj()).real
    n_total = int(np.sum(group_sizes))
    norm = n_total * np.arange(n_frames, n_frames - n_lags, -1)[:, None, None]
    cisf /= norm
    if incoherent:
        iisf /= norm
    wn_out = np.unique(wn.round(11)) if unique else wn
    if unique:
        cisf = np.stack([cisf[:, :, np.isclose(q, wn)].mean(axis=2) for q in wn_out], axis=-1)
        if incoherent:
            iisf = np.stack([iisf[:, :, np.isclose(q, wn)].mean(axis=2) for q in wn_out], axis=-1)
    if sort:
        order = np.argsort(wn_out)
        wn_out = wn_out[order]
        cisf = cisf[:, :, order]
        if incoherent:
            iisf = iisf[:, :, order]
    return {"pairs": pairs, "cisf": cisf, "iisf": iisf, "wavenumbers": wn_out, "n_lags": n_lags}


# ----------------------------------------------------------------------------
# g(r)
# ----------------------------------------------------------------------------


def histogram_bin_edges(range_, n_bins: int) -> np.ndarray:
    """``accelerated.py:13-43``."""
    edges = np.empty(n_bins + 1)
    lib().orc_histogram_bin_edges(float(min(range_)), float(max(range_)), int(n_bins), _dp(edges))
    return edges


def histogram(x: np.ndarray, n_bins: int, bin_edges: np.ndarray) -> np.ndarray:
    """``accelerated.py:46-81`` (LLVM-lowered arithmetic, see oracle.c)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    h = np.zeros(n_bins, dtype=np.int64)
    lib().orc_histogram(_dp(x), x.shape[0], int(n_bins), float(bin_edges[0]), float(bin_edges[-1]), _lp(h))
    return h


def pair_distance(a, b, box) -> float:
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    box = np.ascontiguousarray(box, dtype=np.float32)
    if _is_triclinic(box):
        return float(lib().orc_pair_distance_triclinic(_fp(a), _fp(b), _fp(box)))
    return float(lib().orc_pair_distance(_fp(a), _fp(b), _fp(box)))


def _is_triclinic(box) -> bool:
    return len(box) >= 6 and not np.all(np.asarray(box[3:6]) == 90)


def triclinic_vectors(dimensions) -> np.ndarray:
    """``MDAnalysis.lib.mdamath.triclinic_vectors(dimensions, dtype=float32)`` as restated in ``oracle.c``:
    the lower-triangular box matrix ``(3, 3)`` float32 (zeros for an invalid box)."""
    box = np.ascontiguousarray(dimensions, dtype=np.float32)
    m = np.zeros(6, dtype=np.float32)
    lib().orc_triclinic_vectors(_fp(box), _fp(m))
    return np.array([[m[0], 0, 0], [m[1], m[2], 0], [m[3], m[4], m[5]]], dtype=np.float32)


def radial_histogram(
    positions_a, positions_b, n_bins, range_, dimensions, *, exclusion=None
) -> np.ndarray:
    """``structure.py:34-131`` with a streamed brute-force ``capped_distance``."""
    a = np.ascontiguousarray(np.atleast_2d(positions_a), dtype=np.float32)
    b = np.ascontiguousarray(np.atleast_2d(positions_b), dtype=np.float32)
    box = np.ascontiguousarray(dimensions, dtype=np.float32)
    h = np.zeros(n_bins, dtype=np.int64)
    ex0, ex1 = (0, 0) if exclusion is None else (int(exclusion[0]), int(exclusion[1]))
    fn = lib().orc_radial_histogram_triclinic if _is_triclinic(box) else lib().orc_radial_histogram
    fn(
        _fp(a), a.shape[0], _fp(b), b.shape[0], _fp(box), int(n_bins),
        float(range_[0]), float(range_[1]), ex0, ex1, _lp(h),
    )
    return h


def capped_distance(reference, configuration, max_cutoff, min_cutoff=None, box=None):
    """
    Materialising restatement of ``MDAnalysis.lib.distances.capped_distance``
    (brute-force method) for SMALL inputs — used only as the stub behind the
    reference's real ``radial_histogram`` when generating golden vectors.
    Vectorised NumPy with the same op order as ``oracle.c:orc_pair_r2``.
    """
    a = np.atleast_2d(np.asarray(reference, dtype=np.float32))
    b = np.atleast_2d(np.asarray(configuration, dtype=np.float32))
    box = np.asarray(box, dtype=np.float32)
    if _is_triclinic(box):
        d = _triclinic_distance_array(a, b, box)
        mask = d <= max_cutoff
        if min_cutoff is not None:
            mask &= d > min_cutoff
        return np.argwhere(mask).astype(np.intp), d[mask]
    inv = (1.0 / box[:3].astype(np.float64)).astype(np.float32)
    r2 = np.zeros((a.shape[0], b.shape[0]))
    for k in range(3):
        dx = (b[None, :, k] - a[:, None, k]).astype(np.float64)  # float32 subtract, widened
        if box[k] > np.finfo(np.float32).eps:
            s = np.float64(inv[k]) * dx
            # C round(): half away from zero
            rs = np.where(s >= 0, np.floor(s + 0.5), np.ceil(s - 0.5))
            dx = np.float64(box[k]) * (s - rs)
        sq = dx * dx
        r2 = sq if k == 0 else r2 + sq
    d = np.sqrt(r2)
    mask = d <= max_cutoff
    if min_cutoff is not None:
        mask &= d > min_cutoff
    pairs = np.argwhere(mask)
    return pairs.astype(np.intp), d[mask]


def _triclinic_wrap(pos: np.ndarray, m: np.ndarray) -> np.ndarray:
    """``_triclinic_pbc`` (see ``oracle.c``) in NumPy: float32 coordinates moved into the primary cell."""
    ax, bx, by, cx, cy, cz = (np.float64(v) for v in (m[0, 0], m[1, 0], m[1, 1], m[2, 0], m[2, 1], m[2, 2]))
    bi0, bi4, bi8 = 1.0 / ax, 1.0 / by, 1.0 / cz
    a_y, b_z = bx * bi4, cy * bi8
    a_z = (cx - cy * a_y) * bi8
    x, y, z = (pos[:, k].astype(np.float64) for k in range(3))
    s = np.where((z < 0.0) | (z >= cz), np.floor(z * bi8), 0.0)
    z, y, x = z - s * cz, y - s * cy, x - s * cx
    lb = z * b_z
    s = np.where((y < lb) | (y >= lb + by), np.floor((y - lb) * bi4), 0.0)
    y, x = y - s * by, x - s * bx
    lb = y * a_y + z * a_z
    s = np.where((x < lb) | (x >= lb + ax), np.floor((x - lb) * bi0), 0.0)
    x = x - s * ax
    return np.stack((x, y, z), axis=1).astype(np.float32)


def _triclinic_distance_array(a: np.ndarray, b: np.ndarray, box: np.ndarray) -> np.ndarray:
    """``_calc_distance_array_triclinic`` (see ``oracle.c``) in NumPy, for small inputs."""
    m = triclinic_vectors(box)
    a, b = _triclinic_wrap(a, m), _triclinic_wrap(b, m)
    d = [(b[None, :, k] - a[:, None, k]).astype(np.float64) for k in range(3)]   # float32 subtract, widened
    best = np.full(d[0].shape, np.float64(np.finfo(np.float32).max))
    f64 = np.float64
    for ix in (-1, 0, 1):
        rx = d[0] + f64(m[0, 0] * np.float32(ix))
        for iy in (-1, 0, 1):
            ry0 = rx + f64(m[1, 0] * np.float32(iy))
            ry1 = d[1] + f64(m[1, 1] * np.float32(iy))
            for iz in (-1, 0, 1):
                rz0 = ry0 + f64(m[2, 0] * np.float32(iz))
                rz1 = ry1 + f64(m[2, 1] * np.float32(iz))
                rz2 = d[2] + f64(m[2, 2] * np.float32(iz))
                dsq = rz0 * rz0 + rz1 * rz1 + rz2 * rz2
                best = np.where(dsq < best, dsq, best)
    return np.sqrt(best)


def box_volume(dims) -> float:
    """``ts.volume`` (MDAnalysis ``box_volume``): product of the lengths, times the triclinic factor."""
    d = np.asarray(dims, dtype=np.float32)
    lx, ly, lz = (float(v) for v in d[:3])
    if not _is_triclinic(d):
        return lx * ly * lz
    a, b, g = np.deg2rad(d[3:6].astype(np.float64))
    ca, cb, cg = np.cos(a), np.cos(b), np.cos(g)
    return lx * ly * lz * float(np.sqrt(1 - ca * ca - cb * cb - cg * cg + 2 * ca * cb * cg))


def rdf(
    frames_a: Iterable[np.ndarray],
    frames_b: Optional[Iterable[np.ndarray]],
    dims_per_frame: Iterable[np.ndarray],
    n_bins: int,
    range_,
    *,
    exclusion=None,
    norm: Optional[str] = "rdf",
    drop_axis: Optional[int] = None,
) -> dict:
    """Whole-class restatement of the serial RDF path (``structure.py:849-1010``)."""
    edges = histogram_bin_edges(np.asarray(range_, dtype=float), n_bins)
    bins = (edges[:-1] + edges[1:]) / 2
    counts = np.zeros(n_bins, dtype=np.int64)
    vol = 0.0
    n_frames = 0
    fb = frames_a if frames_b is None else frames_b
    n_a = n_b = None
    for pa, pb, dims in zip(frames_a, fb, dims_per_frame):
        pa = np.array(pa, dtype=np.float32)
        pb = pa if frames_b is None else np.array(pb, dtype=np.float32)
        dims = np.array(dims, dtype=np.float32)
        n_a, n_b = pa.shape[0], pb.shape[0]
        if drop_axis is None:
            vol += box_volume(dims)
        else:
            pa[:, drop_axis] = 0
            pb[:, drop_axis] = 0
            dims[drop_axis] = dims[:3].max()
            vol += np.delete(dims[:3], drop_axis).prod()
        counts += radial_histogram(pa, pb, n_bins, range_, dims, exclusion=exclusion)
        n_frames += 1
    nrm = n_frames
    if norm is not None:
        if drop_axis is None:
            nrm = nrm * 4 * np.pi * np.diff(edges**3) / 3
        else:
            nrm = nrm * np.pi * np.diff(edges**2)
        if norm == "rdf":
            n2 = n_b - (exclusion[1] if exclusion else 0)
            nrm = nrm * (n_a * n2 * n_frames / vol)
    return {
        "bin_edges": edges, "bins": bins, "counts": counts, "rdf": counts / nrm,
        "n_frames": n_frames, "volume": vol,
    }


# ----------------------------------------------------------------------------
# frame ingest
# ----------------------------------------------------------------------------


def read_lammps_dump(path: str, dt: float = 1.0):
    """
    CPU restatement of ``LAMMPSDumpTrajectoryReader`` (reader.py:353-562) for the layouts the GPU reader takes:
    orthogonal or restricted triclinic (xy xz yz) box, unscaled coordinates, optional ids.  Pure Python (small files only); pinned by
    tests/golden/dumps.npz, which the reference's own reader produced.

    Returns ``positions f32 (F, N, 3)``, ``dimensions f32 (F, 6)``, ``steps i64 (F,)``, ``times f64 (F,)``,
    ``offsets`` (byte offset of every frame, reader.py:138-210).
    """
    with open(path, "rb") as fh:
        raw = fh.read()
    lines = raw.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    starts = np.concatenate(([0], np.cumsum([len(x) + 1 for x in lines])))
    positions, dims, steps, offsets = [], [], [], []
    i = 0
    while i + 9 <= len(lines):
        n = int(lines[i + 3])
        if i + 9 + n > len(lines):
            break
        offsets.append(int(starts[i]))
        steps.append(int(lines[i + 1]))
        bounds = [[float(v) for v in lines[i + 5 + ax].split()] for ax in range(3)]                # reader.py:415-417
        if b"xy xz yz" in lines[i + 4]:                                                              # :372-388
            (xlo, xhi, xy), (ylo, yhi, xz), (zlo, zhi, yz) = bounds
            xlo -= min(0, xy, xz, xy + xz)
            xhi -= max(0, xy, xz, xy + xz)
            ylo -= min(0, yz)
            yhi -= max(0, yz)
            vec = np.array([[xhi - xlo, 0, 0], [xy, yhi - ylo, 0], [xz, yz, zhi - zlo]], dtype=float)
            par = np.empty(6)                                                 # algorithm/topology.py:743-757
            par[:3] = np.linalg.norm(vec, axis=1)
            par[3] = np.degrees(np.arccos(np.dot(vec[1], vec[2]) / (par[1] * par[2])))
            par[4] = np.degrees(np.arccos(np.dot(vec[0], vec[2]) / (par[0] * par[2])))
            par[5] = np.degrees(np.arccos(np.dot(vec[0], vec[1]) / (par[0] * par[1])))
            dims.append(par.astype(np.float32))
        else:
            dims.append(np.array([b[1] - b[0] for b in bounds] + [90, 90, 90]).astype(np.float32))   # :418
        cols = {a.decode(): k for k, a in enumerate(lines[i + 8].split()[2:])}                       # :425
        axes = []
        for ax in "xyz":                                                                            # :438-447
            for c in ("u", "su", "", "s"):
                if ax + c in cols:
                    axes.append(cols[ax + c])
                    break
            else:
                axes.append(-1)
        pos = np.zeros((n, 3), dtype=np.float32)
        ids = np.empty(n, dtype=int)
        for a in range(n):
            values = lines[i + 9 + a].split()
            if "id" in cols:
                ids[a] = int(values[cols["id"]])
            # ts.positions[i] = tuple(str): NumPy casts str -> float64 -> float32                    # :503-505
            pos[a] = [0 if k == -1 else np.float32(float(values[k])) for k in axes]
        if "id" in cols:
            pos = pos[np.argsort(ids)]                                                               # :550-552
        positions.append(pos)
        i += 9 + n
    steps = np.array(steps, dtype=np.int64)
    return np.stack(positions), np.stack(dims), steps, steps * dt, np.array(offsets)


# ----------------------------------------------------------------------------
# density profiles
# ----------------------------------------------------------------------------


def density_profile_counts(frames: Iterable[np.ndarray], group_sizes, axes, n_bins, dimensions):
    """
    CPU restatement of ``DensityProfile._single_frame`` without recentring (profile.py:1128-1181): positions widened
    to float64, ``wrap`` (algorithm/topology.py:620-624), ``numba_histogram`` per axis and group
    (accelerated.py:46-81, through the C restatement pinned by histogram_pins.npz).  ``dimensions``: the three box
    lengths as the class holds them (float32 from the universe, float64 when given explicitly).

    Returns one int64 array ``(n_frames, n_groups, n_bins[k])`` per axis.
    """
    dims = np.asarray(dimensions)
    edges = np.concatenate(([0], np.cumsum(group_sizes)))
    out = [[] for _ in axes]
    for pos in frames:
        p = np.array(pos, dtype=np.float64)
        wrap = (p < 0) | (p > dims)
        p[wrap] -= (np.floor(p / dims) * dims)[wrap]
        for k, (ia, n) in enumerate(zip(axes, n_bins)):
            bin_edges = histogram_bin_edges((0.0, float(dims[ia])), int(n))
            out[k].append(np.stack([histogram(np.ascontiguousarray(p[a:b, ia]), int(n), bin_edges)
                                    for a, b in zip(edges[:-1], edges[1:])]))
    return [np.stack(o) for o in out]


def write_amber_netcdf(path: str, positions, times, cell_lengths=None, cell_angles=None, version: int = 2,
                       velocities=None, forces=None) -> None:
    """Test helper: an AMBER-convention NetCDF-3 trajectory (what NetCDFFile.write_header lays out,
    openmm/file.py:309-531) written with SciPy; version 2 = 64-bit offsets, the format the reference uses."""
    from scipy.io import netcdf_file

    positions = np.asarray(positions, dtype=np.float32)
    F, N = positions.shape[:2]
    f = netcdf_file(path, "w", version=version)
    f.Conventions, f.ConventionVersion, f.program = "AMBER", "1.0", "oracle"
    f.createDimension("frame", None)
    f.createDimension("spatial", 3)
    f.createDimension("atom", N)
    v_time = f.createVariable("time", "f", ("frame",))
    v_time.units = "picosecond"
    v_pos = f.createVariable("coordinates", "f", ("frame", "atom", "spatial"))
    v_pos.units = "angstrom"
    if cell_lengths is not None:
        f.createDimension("cell_spatial", 3)
        f.createDimension("cell_angular", 3)
        v_len = f.createVariable("cell_lengths", "d", ("frame", "cell_spatial"))
        v_ang = f.createVariable("cell_angles", "d", ("frame", "cell_angular"))
    if velocities is not None:
        v_vel = f.createVariable("velocities", "f", ("frame", "atom", "spatial"))
    if forces is not None:
        v_frc = f.createVariable("forces", "f", ("frame", "atom", "spatial"))
    for i in range(F):
        v_time[i] = times[i]
        v_pos[i] = positions[i]
        if cell_lengths is not None:
            v_len[i] = cell_lengths[i]
            v_ang[i] = cell_angles[i]
        if velocities is not None:
            v_vel[i] = velocities[i]
        if forces is not None:
            v_frc[i] = forces[i]
    f.close()


def read_amber_netcdf(path: str):
    """
    Independent CPU reader for the NetCDF ingest: SciPy's NetCDF-3 implementation.
    PARITY UNPINNED against the reference's NetCDFFile (openmm/file.py needs netCDF4 and OpenMM, both absent here);
    the file format itself is the contract.  Returns positions f32 (F, N, 3), times f32 (F,), cell_lengths and
    cell_angles f64 (F, 3) or None.
    """
    from scipy.io import netcdf_file

    f = netcdf_file(path, "r", mmap=False)
    pos = np.array(f.variables["coordinates"][:], dtype=np.float32)
    times = np.array(f.variables["time"][:], dtype=np.float32)
    box = "cell_lengths" in f.variables
    lengths = np.array(f.variables["cell_lengths"][:], dtype=np.float64) if box else None
    angles = np.array(f.variables["cell_angles"][:], dtype=np.float64) if box else None
    f.close()
    return pos, times, lengths, angles
