"""
ctypes binding of ``libmdcraft_b200.so`` (``include/mdcraft_b200.h``).

This is the only door between the Python host code and the CUDA kernels.  There
is no CPU fallback: if the library is missing, or no sm_100 device is present,
every entry point raises :class:`MdcError`.
"""

from __future__ import annotations

import ctypes
import os
import re
from ctypes import POINTER, byref, c_char_p, c_double, c_float, c_int, c_int64, c_void_p
from typing import Optional, Sequence

import numpy as np

from . import build as _build

MDC_OK = 0
MDC_ERR_INVALID = -1
MDC_ERR_CUDA = -2
MDC_ERR_NOMEM = -3
MDC_ERR_UNSUPPORTED = -4

PRECISION_FP32 = 0
PRECISION_FP64 = 1
PRECISIONS = {"fp32": PRECISION_FP32, "fp64": PRECISION_FP64}

SQ_ALGO_AUTO = 0
SQ_ALGO_DIRECT = 1
SQ_ALGO_FACTORED = 3
SQ_ALGO_NUFFT = 4
SQ_ALGOS = {"auto": SQ_ALGO_AUTO, "direct": SQ_ALGO_DIRECT, "factored": SQ_ALGO_FACTORED, "nufft": SQ_ALGO_NUFFT}

PROBES = {
    "ffma": 0, "mufu_sincos_pairs": 1, "dfma": 2, "f2f_f64_f32": 3, "atoms_spread": 4, "imad": 5,
    "ffma2": 6, "alu": 7, "i2f": 8, "atoms_hist": 9, "atoms_hist_base": 10, "ffma2_scalar": 11, "ffma2_loop": 12, "ffma2_loop_grouped": 13,
}


class MdcError(RuntimeError):
    """An error reported by libmdcraft_b200 (``code`` is the C return value)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libmdcraft_b200 error {code}: {message}")
        self.code = code


_dp = POINTER(c_double)
_fp = POINTER(c_float)
_lp = POINTER(c_int64)
_ip = POINTER(c_int)

#: every symbol declared in include/mdcraft_b200.h: name -> (restype, argtypes)
SIGNATURES = {
    "mdc_version": (c_int, []),
    "mdc_last_error": (c_char_p, []),
    "mdc_create": (c_int, [c_int, POINTER(c_void_p)]),
    "mdc_create_with_priority": (c_int, [c_int, c_int, POINTER(c_void_p)]),
    "mdc_destroy": (c_int, [c_void_p]),
    "mdc_device_info": (c_int, [c_void_p, _ip, _ip, _ip, _ip, _lp]),
    "mdc_synchronize": (c_int, [c_void_p]),
    "mdc_mark": (c_int, [c_void_p, c_int]),
    "mdc_elapsed_ms": (c_int, [c_void_p, _fp]),
    "mdc_sq_plan_create": (
        c_int,
        [c_void_p, c_int, _dp, c_double, _dp, c_int64, c_int, _lp, c_int, _ip, c_int, c_int, POINTER(c_void_p)],
    ),
    "mdc_sq_plan_destroy": (c_int, [c_void_p]),
    "mdc_sq_num_wavevectors": (c_int64, [c_void_p]),
    "mdc_sq_wavevectors": (c_int, [c_void_p, _dp]),
    "mdc_sq_plan_info": (c_int, [c_void_p, _ip]),
    "mdc_sq_push": (c_int, [c_void_p, c_void_p, c_int, c_int]),
    "mdc_sq_push_f64": (c_int, [c_void_p, c_void_p, c_int, c_int]),
    "mdc_sq_read": (c_int, [c_void_p, _dp, _lp]),
    "mdc_sq_set_shells": (c_int, [c_void_p, _lp, c_int64, _lp, _lp, c_int64]),
    "mdc_sq_read_shells": (c_int, [c_void_p, c_double, _dp, _lp]),
    "mdc_sq_reset": (c_int, [c_void_p]),
    "mdc_sq_accum_ptr": (c_int, [c_void_p, POINTER(c_void_p), _lp]),
    "mdc_sq_set_frames": (c_int, [c_void_p, c_int64]),
    "mdc_delta_fourier_transform_sum": (c_int, [c_void_p, _dp, c_int64, _dp, c_int64, c_int, _dp]),
    "mdc_scsf_plan_create": (
        c_int,
        [c_void_p, c_int, _dp, c_double, _dp, c_int64, c_int64, c_int64, c_int, c_int, POINTER(c_void_p)],
    ),
    "mdc_isf_plan_create": (
        c_int,
        [c_void_p, c_int, _dp, c_double, _dp, c_int64, c_int, _lp, c_int, _ip, c_int, c_int, c_int, c_int,
         POINTER(c_void_p)],
    ),
    "mdc_isf_push": (c_int, [c_void_p, c_void_p, c_int, c_int]),
    "mdc_isf_read": (c_int, [c_void_p, _dp, _dp, _lp]),
    "mdc_rdf_plan_create": (
        c_int,
        [c_void_p, c_int, c_double, c_double, c_int64, c_int64, c_int, c_int64, c_int64, POINTER(c_void_p)],
    ),
    "mdc_rdf_plan_destroy": (c_int, [c_void_p]),
    "mdc_rdf_push": (c_int, [c_void_p, c_void_p, c_void_p, _fp, c_int, c_int]),
    "mdc_rdf_read": (c_int, [c_void_p, _lp, _lp]),
    "mdc_rdf_set_blocks_per_sm": (c_int, [c_void_p, c_int]),
    "mdc_rdf_reset": (c_int, [c_void_p]),
    "mdc_rdf_accum_ptr": (c_int, [c_void_p, POINTER(c_void_p), _lp]),
    "mdc_radial_histogram": (
        c_int,
        [c_void_p, _fp, c_int64, _fp, c_int64, _fp, c_int, c_double, c_double, c_int64, c_int64, _lp],
    ),
    "mdc_radial_transform": (c_int, [c_void_p, c_int, _dp, _dp, _dp, c_int64, _dp, c_int64, _dp]),
    "mdc_dump_open": (c_int, [c_char_p, c_char_p, c_int, POINTER(c_void_p)]),
    "mdc_dump_close": (c_int, [c_void_p]),
    "mdc_dump_info": (c_int, [c_void_p, _lp]),
    "mdc_dump_headers": (c_int, [c_void_p, _lp, _fp, _dp, _lp]),
    "mdc_dump_block_bytes": (c_int64, [c_void_p, c_int64, c_int]),
    "mdc_dump_read": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int]),
    "mdc_profile_plan_create": (c_int, [c_void_p, c_int, _lp, c_int, _ip, _ip, _dp, POINTER(c_void_p)]),
    "mdc_profile_plan_destroy": (c_int, [c_void_p]),
    "mdc_profile_num_counts": (c_int64, [c_void_p]),
    "mdc_profile_push": (c_int, [c_void_p, c_void_p, c_int, c_int, _lp]),
    "mdc_profile_push_f64": (c_int, [c_void_p, c_void_p, c_int, c_int, _lp]),
    "mdc_profile_read": (c_int, [c_void_p, _lp, _lp]),
    "mdc_profile_reset": (c_int, [c_void_p]),
    "mdc_profile_accum_ptr": (c_int, [c_void_p, POINTER(c_void_p), _lp]),
    "mdc_ncdf_open": (c_int, [c_char_p, POINTER(c_void_p)]),
    "mdc_ncdf_close": (c_int, [c_void_p]),
    "mdc_ncdf_info": (c_int, [c_void_p, _lp]),
    "mdc_ncdf_headers": (c_int, [c_void_p, _dp, _dp, _dp]),
    "mdc_ncdf_read": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int]),
    "mdc_ncdf_has": (c_int, [c_void_p, c_int]),
    "mdc_ncdf_read_var": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p, c_int]),
    "mdc_sq_last_push_stats": (c_int, [c_void_p, _fp, _ip]),
    "mdc_rdf_last_push_stats": (c_int, [c_void_p, _fp, _ip]),
    "mdc_rdf_plan_info": (c_int, [c_void_p, _ip]),
    "mdc_probe_rate": (c_int, [c_void_p, c_int, _dp]),
    "mdc_probe_sincos": (c_int, [c_void_p, c_int, _dp]),
}

_lib: Optional[ctypes.CDLL] = None


def header_path() -> str:
    return os.path.join(os.path.dirname(_build.HERE), "include", "mdcraft_b200.h")


def declared_symbols() -> list:
    """Function names declared in ``include/mdcraft_b200.h``."""
    with open(header_path()) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mdc_[a-z0-9_]+)\s*\(", text)))


def lib() -> ctypes.CDLL:
    """Loads (building it in-tree if the sources are newer) the CUDA library."""
    global _lib
    if _lib is None:
        path = os.environ.get("MDCRAFT_B200_LIB") or _build.LIB_PATH   # override: kernel experiments only
        if path == _build.LIB_PATH and _build.is_stale():
            try:
                path = _build.build()
            except Exception as err:  # no nvcc on the box: use what travelled with the snapshot
                if not os.path.exists(path):
                    raise MdcError(
                        MDC_ERR_UNSUPPORTED,
                        f"{path} is missing and cannot be built ({err}); there is no CPU fallback",
                    ) from err
        L = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(code: int) -> None:
    if code != MDC_OK:
        raise MdcError(code, (lib().mdc_last_error() or b"").decode("utf-8", "replace"))


def _as(a: np.ndarray, ptr):
    return a.ctypes.data_as(ptr)


def _buffer(x, dtype=np.float32):
    """
    ``(address, on_device, keepalive)`` for a NumPy array (host) or any object
    with ``__cuda_array_interface__`` / a CUDA ``torch.Tensor`` (device).
    """
    if hasattr(x, "__cuda_array_interface__") and not isinstance(x, np.ndarray):
        cai = x.__cuda_array_interface__
        if np.dtype(cai["typestr"]) != np.dtype(dtype):
            raise TypeError(f"device buffer must be {np.dtype(dtype)}")
        if cai.get("strides") is not None:
            raise ValueError("device buffer must be C-contiguous")
        return c_void_p(cai["data"][0]), 1, x
    a = np.ascontiguousarray(x, dtype=dtype)
    return c_void_p(a.ctypes.data), 0, a


def _dtype_of(x):
    """NumPy dtype of a host array or of a ``__cuda_array_interface__`` buffer (``None`` if unknown)."""
    if isinstance(x, np.ndarray):
        return x.dtype
    cai = getattr(x, "__cuda_array_interface__", None)
    return np.dtype(cai["typestr"]) if cai else None


class Context:
    """One CUDA device (``mdc_create``)."""

    def __init__(self, device: int = 0, high_priority: bool = False):
        self._h = c_void_p()
        self.device = int(device)
        _check(lib().mdc_create_with_priority(self.device, int(bool(high_priority)), byref(self._h)))

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib().mdc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> dict:
        sm, ma, mi, khz = c_int(), c_int(), c_int(), c_int()
        mem = c_int64()
        _check(lib().mdc_device_info(self._h, byref(sm), byref(ma), byref(mi), byref(khz), byref(mem)))
        return {
            "sm_count": sm.value, "cc": (ma.value, mi.value), "sm_clock_khz": khz.value,
            "global_mem_bytes": mem.value,
        }

    def synchronize(self) -> None:
        _check(lib().mdc_synchronize(self._h))

    def mark(self, which: int) -> None:
        """Records the start (0) / stop (1) event on the compute stream shared by this context's plans."""
        _check(lib().mdc_mark(self._h, int(which)))

    def elapsed_ms(self) -> float:
        ms = c_float()
        _check(lib().mdc_elapsed_ms(self._h, byref(ms)))
        return ms.value

    def probe_rate(self, which) -> float:
        out = c_double()
        _check(lib().mdc_probe_rate(self._h, PROBES[which] if isinstance(which, str) else int(which), byref(out)))
        return out.value

    def probe_sincos(self, variant: int = 0) -> dict:
        out = (c_double * 4)()
        _check(lib().mdc_probe_sincos(self._h, int(variant), out))
        return {"mean_cos_err": out[0], "mean_sin_err": out[1], "rms_err": out[2], "max_err": out[3]}

    # -- one-shot functional seams ------------------------------------------
    def delta_fourier_transform_sum(self, qs, rs, precision: str = "fp32") -> np.ndarray:
        """``structure.py:1240-1280``: ``F[i] = sum_j exp(1j q_i . r_j)`` (complex128)."""
        qs = np.ascontiguousarray(qs, dtype=np.float64).reshape(-1, 3)
        rs = np.ascontiguousarray(rs, dtype=np.float64).reshape(-1, 3)
        out = np.empty((qs.shape[0], 2), dtype=np.float64)
        _check(lib().mdc_delta_fourier_transform_sum(
            self._h, _as(qs, _dp), qs.shape[0], _as(rs, _dp), rs.shape[0], PRECISIONS[precision], _as(out, _dp)))
        return out.view(np.complex128).reshape(-1)

    def radial_histogram(self, pos_a, pos_b, n_bins, range_, dimensions, exclusion=None) -> np.ndarray:
        """``structure.py:34-131``: ordered-pair distance histogram (int64)."""
        a = np.ascontiguousarray(np.atleast_2d(pos_a), dtype=np.float32)
        same = pos_b is None or pos_b is pos_a
        b = a if same else np.ascontiguousarray(np.atleast_2d(pos_b), dtype=np.float32)
        box = np.ascontiguousarray(dimensions, dtype=np.float32)
        if box.shape[0] == 3:
            box = np.concatenate((box, np.full(3, 90, np.float32)))
        ex0, ex1 = (0, 0) if exclusion is None else (int(exclusion[0]), int(exclusion[1]))
        out = np.zeros(int(n_bins), dtype=np.int64)
        _check(lib().mdc_radial_histogram(
            self._h, _as(a, _fp), a.shape[0], _as(a if same else b, _fp), b.shape[0], _as(box, _fp), int(n_bins),
            float(range_[0]), float(range_[1]), ex0, ex1, _as(out, _lp)))
        return out


    def radial_transform(self, kind: str, r, f, weights, q) -> np.ndarray:
        """``structure.py:134-215``: ``kind`` = ``"fourier3d"`` or ``"hankel2d"``; FP64 on the device."""
        r = np.ascontiguousarray(r, dtype=np.float64)
        f = np.ascontiguousarray(f, dtype=np.float64)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1)
        if not (r.shape == f.shape == w.shape) or r.ndim != 1:
            raise ValueError("r, f and weights must be 1-D arrays of equal length")
        out = np.empty(q.shape[0], dtype=np.float64)
        _check(lib().mdc_radial_transform(
            self._h, {"fourier3d": 0, "hankel2d": 1}[kind], _as(r, _dp), _as(f, _dp), _as(w, _dp), r.shape[0],
            _as(q, _dp), q.shape[0], _as(out, _dp)))
        return out


_contexts: dict = {}


def context(device: int = 0, lane: int = 0) -> Context:
    """
    One library context per CUDA device and LANE per process.  A context owns the streams its plans run on, so plans
    of different lanes execute concurrently on the device: the analysis classes put g(r) on lane 0 and the S(q)
    family on lane 1 (the g(r) pair kernel is issue bound, the NUFFT L2 / shared-memory bound: run side by side they
    hide most of the S(q) time).  Readers and the one-shot seams use lane 0.
    """
    ctx = _contexts.get((device, lane))
    if ctx is None:
        ctx = _contexts[(device, lane)] = Context(device, high_priority=lane != 0)
    return ctx


class _DeviceArray:
    """Exposes a library-owned accumulator through ``__cuda_array_interface__`` (for NCCL via torch)."""

    def __init__(self, ptr: int, n: int, typestr: str, owner):
        self.__cuda_array_interface__ = {
            "shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None,
        }
        self._owner = owner


class SqPlan:
    """``mdc_sq_plan``: accumulates S(q) pair sums over pushed frame blocks."""

    def __init__(
        self,
        ctx: Context,
        group_sizes: Sequence[int],
        pairs: Sequence[Sequence[Optional[int]]],
        *,
        n_points: int = 0,
        L0: Optional[Sequence[float]] = None,
        q_max: Optional[float] = None,
        wavevectors: Optional[np.ndarray] = None,
        precision: str = "fp32",
        algo: str = "auto",
    ):
        self.ctx = ctx
        self._h = c_void_p()
        gs = np.ascontiguousarray(group_sizes, dtype=np.int64)
        pr = np.ascontiguousarray(
            [(-1, -1) if p[0] is None else (int(p[0]), int(p[1])) for p in pairs], dtype=np.int32
        ).reshape(-1, 2)
        L = np.zeros(3) if L0 is None else np.ascontiguousarray(L0, dtype=np.float64)
        if wavevectors is not None:
            wv = np.ascontiguousarray(wavevectors, dtype=np.float64).reshape(-1, 3)
            n_exp, wv_p = wv.shape[0], _as(wv, _dp)
        else:
            wv, n_exp, wv_p = None, 0, None
        _check(lib().mdc_sq_plan_create(
            ctx._h, int(n_points), _as(L, _dp), -1.0 if q_max is None else float(q_max), wv_p, n_exp,
            gs.shape[0], _as(gs, _lp), pr.shape[0], _as(pr, _ip), PRECISIONS[precision], SQ_ALGOS[algo],
            byref(self._h)))
        self.n_particles = int(gs.sum())
        self.n_pairs = pr.shape[0]
        self.n_wavevectors = int(lib().mdc_sq_num_wavevectors(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib().mdc_sq_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> dict:
        """What the plan decided: lattice kernel family and, for the NUFFT, its fine grid and kernel width."""
        a = (c_int * 8)()
        _check(lib().mdc_sq_plan_info(self._h, a))
        names = {v: k for k, v in SQ_ALGOS.items() if k != "auto"}
        return {"algo": names.get(a[0], "explicit"), "nf": a[1], "w": a[2], "n_points": a[3], "n_sets": a[4],
                "chunks": a[5], "frame_block": a[6], "nufft_batch": a[7]}

    def wavevectors(self) -> np.ndarray:
        out = np.empty((self.n_wavevectors, 3), dtype=np.float64)
        _check(lib().mdc_sq_wavevectors(self._h, _as(out, _dp)))
        return out

    def push(self, pos, n_frames: Optional[int] = None) -> None:
        """``pos``: float32 (or float64: centres of mass) ``(F, N, 3)`` (or ``(N, 3)``), host array or device buffer."""
        f64 = _dtype_of(pos) == np.float64
        ptr, on_dev, keep = _buffer(pos, np.float64 if f64 else np.float32)
        shape = keep.shape if hasattr(keep, "shape") else None
        if n_frames is None:
            n_frames = 1 if len(shape) == 2 else int(shape[0])
        n_el = int(np.prod(shape))
        if n_el != n_frames * self.n_particles * 3:
            raise ValueError(f"positions have {n_el} elements, expected {n_frames} x {self.n_particles} x 3")
        _check((lib().mdc_sq_push_f64 if f64 else lib().mdc_sq_push)(self._h, ptr, int(n_frames), on_dev))

    def read(self, out: Optional[np.ndarray] = None):
        """Un-normalised sums ``(n_pairs, Nq)`` and the number of frames accumulated.  ``out``: optional
        C-contiguous float64 array of that shape to read into (e.g. a view of pinned memory)."""
        if out is None:
            out = np.empty((self.n_pairs, self.n_wavevectors), dtype=np.float64)
        elif out.dtype != np.float64 or not out.flags.c_contiguous or out.size != self.n_pairs * self.n_wavevectors:
            raise ValueError("out must be a C-contiguous float64 array with n_pairs x Nq elements")
        nf = c_int64()
        _check(lib().mdc_sq_read(self._h, _as(out, _dp), byref(nf)))
        return out, nf.value

    def set_shells(self, order: np.ndarray, lo: np.ndarray, hi: np.ndarray) -> None:
        """|q|-shell windows for :meth:`read_shells`: shell ``s`` = wavevectors ``order[lo[s]:hi[s]]``."""
        order = np.ascontiguousarray(order, dtype=np.int64)
        lo = np.ascontiguousarray(lo, dtype=np.int64)
        hi = np.ascontiguousarray(hi, dtype=np.int64)
        _check(lib().mdc_sq_set_shells(self._h, _as(order, _lp), order.shape[0], _as(lo, _lp), _as(hi, _lp), lo.shape[0]))
        self.n_shells = int(lo.shape[0])

    def read_shells(self, denom: float, out: Optional[np.ndarray] = None):
        """Shell means of ``sums / denom``, ``(n_pairs, n_shells)``, and the number of frames accumulated."""
        if out is None:
            out = np.empty((self.n_pairs, self.n_shells), dtype=np.float64)
        elif out.dtype != np.float64 or not out.flags.c_contiguous or out.size != self.n_pairs * self.n_shells:
            raise ValueError("out must be a C-contiguous float64 array with n_pairs x n_shells elements")
        nf = c_int64()
        _check(lib().mdc_sq_read_shells(self._h, float(denom), _as(out, _dp), byref(nf)))
        return out, nf.value

    def reset(self) -> None:
        _check(lib().mdc_sq_reset(self._h))

    def set_frames(self, n: int) -> None:
        _check(lib().mdc_sq_set_frames(self._h, int(n)))

    def accumulator(self) -> _DeviceArray:
        p, n = c_void_p(), c_int64()
        _check(lib().mdc_sq_accum_ptr(self._h, byref(p), byref(n)))
        return _DeviceArray(p.value, n.value, "<f8", self)

    def last_push_stats(self):
        ms, nl = c_float(), c_int()
        _check(lib().mdc_sq_last_push_stats(self._h, byref(ms), byref(nl)))
        return ms.value, nl.value


class ProfilePlan:
    """``mdc_profile_plan``: per-axis, per-group 1-D histograms of wrapped positions (DensityProfile)."""

    def __init__(self, ctx: Context, group_sizes: Sequence[int], axes: Sequence[int], n_bins: Sequence[int],
                 dimensions: Sequence[float]):
        self.ctx = ctx
        self._h = c_void_p()
        gs = np.ascontiguousarray(group_sizes, dtype=np.int64)
        ax = np.ascontiguousarray(axes, dtype=np.int32)
        nb = np.ascontiguousarray(n_bins, dtype=np.int32)
        dims = np.ascontiguousarray(dimensions, dtype=np.float64)
        if ax.shape != nb.shape or dims.shape != (3,):
            raise ValueError("axes and n_bins must have equal length and dimensions three entries")
        _check(lib().mdc_profile_plan_create(ctx._h, gs.shape[0], _as(gs, _lp), ax.shape[0], _as(ax, _ip), _as(nb, _ip),
                                             _as(dims, _dp), byref(self._h)))
        self.n_groups, self.n_particles = gs.shape[0], int(gs.sum())
        self.n_bins = [int(n) for n in nb]
        self.total = int(lib().mdc_profile_num_counts(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib().mdc_profile_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _split(self, flat: np.ndarray) -> list:
        """Count table ``(..., total)`` -> one ``(..., n_groups, n_bins[k])`` array per axis."""
        out, off = [], 0
        for n in self.n_bins:
            out.append(flat[..., off:off + self.n_groups * n].reshape(flat.shape[:-1] + (self.n_groups, n)))
            off += self.n_groups * n
        return out

    def push(self, pos, n_frames: Optional[int] = None, per_frame: bool = False):
        """``pos``: float32 (or float64, for recentred positions) ``(F, N, 3)``.  ``per_frame``: also return every
        frame's own counts, per axis ``(F, n_groups, n_bins)``."""
        f64 = getattr(pos, "dtype", None) == np.float64
        ptr, on_dev, keep = _buffer(pos, np.float64 if f64 else np.float32)
        shape = keep.shape
        if n_frames is None:
            n_frames = 1 if len(shape) == 2 else int(shape[0])
        if int(np.prod(shape)) != n_frames * self.n_particles * 3:
            raise ValueError(f"positions have {int(np.prod(shape))} elements, expected {n_frames} x {self.n_particles} x 3")
        out = np.empty((n_frames, self.total), dtype=np.int64) if per_frame else None
        fn = lib().mdc_profile_push_f64 if f64 else lib().mdc_profile_push
        _check(fn(self._h, ptr, int(n_frames), on_dev, None if out is None else _as(out, _lp)))
        return None if out is None else self._split(out)

    def read(self):
        """Summed counts per axis ``(n_groups, n_bins)`` and the number of frames pushed."""
        out = np.empty(self.total, dtype=np.int64)
        nf = c_int64()
        _check(lib().mdc_profile_read(self._h, _as(out, _lp), byref(nf)))
        return self._split(out), nf.value

    def reset(self) -> None:
        _check(lib().mdc_profile_reset(self._h))

    def accumulator(self) -> _DeviceArray:
        p, n = c_void_p(), c_int64()
        _check(lib().mdc_profile_accum_ptr(self._h, byref(p), byref(n)))
        return _DeviceArray(p.value, n.value, "<u8", self)


class DumpFile:
    """``mdc_dump``: a memory-mapped LAMMPS text dump whose atom lines are parsed on the GPU."""

    CONVENTIONS = ("u", "su", "", "s")

    def __init__(self, path: str, conventions: Optional[Sequence[Optional[str]]] = None, n_threads: int = 0):
        self._h = c_void_p()
        spec = None
        if conventions is not None:
            spec = ",".join("-" if c is None else c for c in conventions).encode()
        _check(lib().mdc_dump_open(os.fsencode(path), spec, int(n_threads), byref(self._h)))
        info = (c_int64 * 8)()
        _check(lib().mdc_dump_info(self._h, info))
        self.n_frames, self.n_atoms, self.has_id = int(info[0]), int(info[1]), bool(info[2])
        self.conventions = [None if info[3 + k] < 0 else self.CONVENTIONS[info[3 + k]] for k in range(3)]
        self.n_columns = int(info[6])
        self.steps = np.empty(self.n_frames, dtype=np.int64)
        self.box = np.empty((self.n_frames, 6), dtype=np.float32)
        self.origin = np.empty((self.n_frames, 3), dtype=np.float64)
        self.offsets = np.empty(self.n_frames + 1, dtype=np.int64)
        _check(lib().mdc_dump_headers(self._h, _as(self.steps, _lp), _as(self.box, _fp), _as(self.origin, _dp),
                                      _as(self.offsets, _lp)))

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib().mdc_dump_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def block_bytes(self, f0: int, n_frames: int) -> int:
        return int(lib().mdc_dump_block_bytes(self._h, int(f0), int(n_frames)))

    def last_host_parsed(self) -> int:
        """Values of the last ``read`` that the device's exact fast path handed to the host parser."""
        info = (c_int64 * 8)()
        _check(lib().mdc_dump_info(self._h, info))
        return int(info[7])

    def read(self, ctx: "Context", f0: int, n_frames: int, out=None):
        """Positions of frames ``[f0, f0 + n_frames)`` as float32 ``(n_frames, n_atoms, 3)`` in atom-id order.
        ``out``: a host array or a device buffer (``__cuda_array_interface__``) to fill; default: a new host array."""
        if out is None:
            out = np.empty((int(n_frames), self.n_atoms, 3), dtype=np.float32)
        ptr, on_dev, keep = _buffer(out)
        if not on_dev and keep is not out:
            raise ValueError("out must be a C-contiguous float32 array")
        n_el = int(np.prod(keep.shape)) if hasattr(keep, "shape") else None
        if n_el is not None and n_el != int(n_frames) * self.n_atoms * 3:
            raise ValueError(f"out has {n_el} elements, expected {n_frames} x {self.n_atoms} x 3")
        _check(lib().mdc_dump_read(self._h, ctx._h, int(f0), int(n_frames), ptr, on_dev))
        return out


class NcdfFile:
    """``mdc_ncdf``: a memory-mapped AMBER NetCDF-3 trajectory whose coordinates are byte-swapped on the GPU."""

    VARIABLES = {"coordinates": 0, "velocities": 1, "forces": 2}   # MDC_NCDF_*

    def __init__(self, path: str):
        self._h = c_void_p()
        _check(lib().mdc_ncdf_open(os.fsencode(path), byref(self._h)))
        info = (c_int64 * 4)()
        _check(lib().mdc_ncdf_info(self._h, info))
        self.n_frames, self.n_atoms = int(info[0]), int(info[1])
        self.has_box, self.has_time = bool(info[2]), bool(info[3])
        self.has_velocities = bool(lib().mdc_ncdf_has(self._h, self.VARIABLES["velocities"]))
        self.has_forces = bool(lib().mdc_ncdf_has(self._h, self.VARIABLES["forces"]))
        self.times = np.empty(self.n_frames, dtype=np.float64) if self.has_time else None
        self.cell_lengths = np.empty((self.n_frames, 3), dtype=np.float64) if self.has_box else None
        self.cell_angles = np.empty((self.n_frames, 3), dtype=np.float64) if self.has_box else None
        _check(lib().mdc_ncdf_headers(
            self._h, None if self.times is None else _as(self.times, _dp),
            None if self.cell_lengths is None else _as(self.cell_lengths, _dp),
            None if self.cell_angles is None else _as(self.cell_angles, _dp)))

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib().mdc_ncdf_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def read(self, ctx: "Context", f0: int, n_frames: int, out=None, variable: str = "coordinates"):
        """Coordinates (or ``variable="velocities"`` / ``"forces"``) of frames ``[f0, f0 + n_frames)`` as float32
        ``(n_frames, n_atoms, 3)``; ``out``: a host array or a device buffer to fill."""
        if out is None:
            out = np.empty((int(n_frames), self.n_atoms, 3), dtype=np.float32)
        ptr, on_dev, keep = _buffer(out)
        if not on_dev and keep is not out:
            raise ValueError("out must be a C-contiguous float32 array")
        if int(np.prod(keep.shape)) != int(n_frames) * self.n_atoms * 3:
            raise ValueError(f"out must hold {n_frames} x {self.n_atoms} x 3 elements")
        _check(lib().mdc_ncdf_read_var(self._h, ctx._h, self.VARIABLES[variable], int(f0), int(n_frames), ptr, on_dev))
        return out


class ScsfPlan(SqPlan):
    """``mdc_scsf_plan_create``: sum over chains of ``|rho_chain|^2`` in one accumulator row."""

    def __init__(self, ctx: Context, n_chains: int, n_monomers: int, *, n_points: int = 0,
                 L0: Optional[Sequence[float]] = None, q_max: Optional[float] = None,
                 wavevectors: Optional[np.ndarray] = None, precision: str = "fp32", algo: str = "auto"):
        self.ctx = ctx
        self._h = c_void_p()
        L = np.zeros(3) if L0 is None else np.ascontiguousarray(L0, dtype=np.float64)
        if wavevectors is not None:
            wv = np.ascontiguousarray(wavevectors, dtype=np.float64).reshape(-1, 3)
            n_exp, wv_p = wv.shape[0], _as(wv, _dp)
        else:
            n_exp, wv_p = 0, None
        _check(lib().mdc_scsf_plan_create(
            ctx._h, int(n_points), _as(L, _dp), -1.0 if q_max is None else float(q_max), wv_p, n_exp,
            int(n_chains), int(n_monomers), PRECISIONS[precision], SQ_ALGOS[algo], byref(self._h)))
        self.n_particles = int(n_chains) * int(n_monomers)
        self.n_pairs = 1   # one accumulator row
        self.n_wavevectors = int(lib().mdc_sq_num_wavevectors(self._h))


class IsfPlan(SqPlan):
    """``mdc_isf_plan_create``: coherent / incoherent intermediate scattering functions over a ring of lags."""

    def __init__(
        self,
        ctx: Context,
        group_sizes: Sequence[int],
        pairs: Sequence[Sequence[Optional[int]]],
        n_lags: int,
        *,
        incoherent: bool = False,
        n_points: int = 0,
        L0: Optional[Sequence[float]] = None,
        q_max: Optional[float] = None,
        wavevectors: Optional[np.ndarray] = None,
        precision: str = "fp32",
        algo: str = "auto",
    ):
        self.ctx = ctx
        self._h = c_void_p()
        gs = np.ascontiguousarray(group_sizes, dtype=np.int64)
        pr = np.ascontiguousarray(
            [(-1, -1) if p[0] is None else (int(p[0]), int(p[1])) for p in pairs], dtype=np.int32
        ).reshape(-1, 2)
        L = np.zeros(3) if L0 is None else np.ascontiguousarray(L0, dtype=np.float64)
        if wavevectors is not None:
            wv = np.ascontiguousarray(wavevectors, dtype=np.float64).reshape(-1, 3)
            n_exp, wv_p = wv.shape[0], _as(wv, _dp)
        else:
            n_exp, wv_p = 0, None
        _check(lib().mdc_isf_plan_create(
            ctx._h, int(n_points), _as(L, _dp), -1.0 if q_max is None else float(q_max), wv_p, n_exp,
            gs.shape[0], _as(gs, _lp), pr.shape[0], _as(pr, _ip), int(n_lags), 1 if incoherent else 0,
            PRECISIONS[precision], SQ_ALGOS[algo], byref(self._h)))
        self.n_particles = int(gs.sum())
        self.n_pairs = pr.shape[0]
        self.n_lags = int(n_lags)
        self.incoherent = bool(incoherent)
        self.n_rows = 1 if pr[0, 0] < 0 else gs.shape[0]
        self.n_wavevectors = int(lib().mdc_sq_num_wavevectors(self._h))

    def push(self, pos, n_frames: Optional[int] = None) -> None:
        """Frames must arrive in time order.  float32, or float64 (centres of mass: ``mdc_sq_push_f64``, which
        serves ISF plans too)."""
        if _dtype_of(pos) == np.float64:
            return super().push(pos, n_frames)
        ptr, on_dev, keep = _buffer(pos)
        if n_frames is None:
            n_frames = 1 if len(keep.shape) == 2 else int(keep.shape[0])
        if int(np.prod(keep.shape)) != n_frames * self.n_particles * 3:
            raise ValueError("positions have the wrong number of elements")
        _check(lib().mdc_isf_push(self._h, ptr, int(n_frames), on_dev))

    def read(self):
        """Un-normalised ``cisf (n_lags, n_pairs, Nq)``, ``iisf (n_lags, n_rows, Nq)`` or ``None``, frames."""
        cisf = np.empty((self.n_lags, self.n_pairs, self.n_wavevectors), dtype=np.float64)
        iisf = np.empty((self.n_lags, self.n_rows, self.n_wavevectors), dtype=np.float64) if self.incoherent else None
        nf = c_int64()
        _check(lib().mdc_isf_read(self._h, _as(cisf, _dp), None if iisf is None else _as(iisf, _dp), byref(nf)))
        return cisf, iisf, nf.value


class RdfPlan:
    """``mdc_rdf_plan``: accumulates the ordered-pair distance histogram over pushed frame blocks."""

    def __init__(self, ctx: Context, n_bins: int, range_, n_a: int, n_b: Optional[int] = None, exclusion=None):
        self.ctx = ctx
        self._h = c_void_p()
        self.same = n_b is None
        ex0, ex1 = (0, 0) if exclusion is None else (int(exclusion[0]), int(exclusion[1]))
        _check(lib().mdc_rdf_plan_create(
            ctx._h, int(n_bins), float(range_[0]), float(range_[1]), int(n_a), int(n_a if n_b is None else n_b),
            1 if self.same else 0, ex0, ex1, byref(self._h)))
        self.n_bins, self.n_a, self.n_b = int(n_bins), int(n_a), int(n_a if n_b is None else n_b)

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib().mdc_rdf_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def push(self, pos_a, pos_b, box, n_frames: Optional[int] = None) -> None:
        """``pos_*``: float32 ``(F, n, 3)`` host or device; ``box``: float32 ``(F, 6)`` or ``(6,)`` host."""
        pa, dev_a, ka = _buffer(pos_a)
        if n_frames is None:
            n_frames = 1 if len(ka.shape) == 2 else int(ka.shape[0])
        if int(np.prod(ka.shape)) != n_frames * self.n_a * 3:
            raise ValueError("pos_a has the wrong number of elements")
        if self.same:
            pb, kb = None, None
        else:
            pb, dev_b, kb = _buffer(pos_b)
            if dev_b != dev_a:
                raise ValueError("pos_a and pos_b must both be host or both be device buffers")
            if int(np.prod(kb.shape)) != n_frames * self.n_b * 3:
                raise ValueError("pos_b has the wrong number of elements")
        bx = np.ascontiguousarray(box, dtype=np.float32)
        if bx.ndim == 1:
            bx = np.ascontiguousarray(np.broadcast_to(bx, (n_frames, bx.shape[0])))
        if bx.shape != (n_frames, 6):
            raise ValueError("box must be (n_frames, 6)")
        _check(lib().mdc_rdf_push(self._h, pa, pb, _as(bx, _fp), int(n_frames), dev_a))

    def read(self):
        out = np.empty(self.n_bins, dtype=np.int64)
        nf = c_int64()
        _check(lib().mdc_rdf_read(self._h, _as(out, _lp), byref(nf)))
        return out, nf.value

    def set_blocks_per_sm(self, blocks_per_sm: int) -> None:
        """Caps the persistent pair kernel at ``blocks_per_sm`` blocks per SM (0 = all that fit), leaving room on every
        SM for the kernels of another context: see ``mdc_rdf_set_blocks_per_sm``."""
        _check(lib().mdc_rdf_set_blocks_per_sm(self._h, int(blocks_per_sm)))

    def reset(self) -> None:
        _check(lib().mdc_rdf_reset(self._h))

    def accumulator(self) -> _DeviceArray:
        p, n = c_void_p(), c_int64()
        _check(lib().mdc_rdf_accum_ptr(self._h, byref(p), byref(n)))
        return _DeviceArray(p.value, n.value, "<i8", self)

    def last_push_stats(self):
        ms, nl = c_float(), c_int()
        _check(lib().mdc_rdf_last_push_stats(self._h, byref(ms), byref(nl)))
        return ms.value, nl.value

    def info(self) -> dict:
        """Kernel path of the last frame block pushed (``mdc_rdf_plan_info``)."""
        a = (c_int * 8)()
        _check(lib().mdc_rdf_plan_info(self._h, a))
        return {"filter": bool(a[0]), "sorted": bool(a[1]), "ext_words": a[2], "copies_interleaved": a[3] == 32,
                "relative_rows": bool(a[4]), "triclinic": bool(a[5]), "queue": a[6], "frames": a[7]}
