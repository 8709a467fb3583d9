"""
Numba restatement of the reference's S(q) density kernel — TEST / BASELINE INFRASTRUCTURE ONLY.

``bench.py --impl reference`` times it beside the C/OpenMP port (``oracle/oracle.c``) so that the factor between the
port and the reference's own technology (Numba, ``fastmath``, ``prange`` over wavevectors) is on record.  The
reference's functions are ``delta_fourier_transform_sum`` + ``numba_delta_fourier_transform`` + ``numba_dot``
(``/root/reference/src/mdcraft/analysis/structure.py:1204-1280``, ``algorithm/accelerated.py:84-113``); its class
re-jits them with ``parallel=True`` when asked to (``structure.py:1901-1907``).  Nothing under ``mdcraft_b200/``
imports this module.
"""

from __future__ import annotations

import numpy as np


def available() -> bool:
    try:
        import numba  # noqa: F401

        return True
    except Exception:
        return False


_kernel = None


def _build():
    global _kernel
    if _kernel is None:
        import numba

        @numba.njit(fastmath=True, parallel=True)
        def density(qs, rs):
            out = np.empty(qs.shape[0], dtype=np.complex128)
            for i in numba.prange(qs.shape[0]):
                acc = 0.0j
                for j in range(rs.shape[0]):
                    phase = 0.0
                    for k in range(3):
                        phase += qs[i, k] * rs[j, k]
                    acc += np.exp(1j * phase)
                out[i] = acc
            return out

        _kernel = density
    return _kernel


def delta_fourier_transform_sum(qs: np.ndarray, rs: np.ndarray) -> np.ndarray:
    """``F[i] = sum_j exp(1j q_i . r_j)`` with Numba's thread pool (``numba.get_num_threads()`` threads)."""
    return _build()(np.ascontiguousarray(qs, dtype=np.float64), np.ascontiguousarray(rs, dtype=np.float64))


def num_threads() -> int:
    import numba

    return int(numba.get_num_threads())


def set_num_threads(n: int) -> None:
    import numba

    numba.set_num_threads(max(1, min(int(n), numba.config.NUMBA_NUM_THREADS)))
