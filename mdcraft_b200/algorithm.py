"""
Host-side helpers the S(q) / g(r) classes call on the way to the GPU.

They mirror the few ``mdcraft.algorithm`` functions that sit on the hot path's
*boundary* (SURVEY.md §8 a13 and row 4 of §2) and stay on the CPU:

* :func:`strip_unit` / :func:`is_unitless`  (``algorithm/unit.py:171-383``) —
  only the plain-number, pint and OpenMM pass-through needed for ``dimensions``,
  ``q_max`` and ``temperature``;
* :func:`get_closest_factors`  (``algorithm/utility.py:16-75``) — used for the
  ``n_surfaces`` off-lattice wavevectors;
* :func:`center_of_mass`  (``algorithm/molecule.py:16-441``) — the
  ``center_of_mass(group, grouping)`` call form used when
  ``groupings != "atoms"``;
* :func:`histogram_bin_edges`  (``algorithm/accelerated.py:13-43``).
"""

from __future__ import annotations

from numbers import Number
from typing import Any, Optional, Tuple

import numpy as np


def strip_unit(value: Any, unit_: Any = None) -> Tuple[Any, Any]:
    """
    Returns ``(magnitude, unit)``.  Plain numbers / arrays / lists pass through
    untouched with the requested unit, like the reference does
    (``unit.py:349-352``); ``pint`` and ``openmm.unit`` quantities are converted
    to ``unit_`` when one is given.
    """
    if value is None or isinstance(value, (Number, np.ndarray, list, tuple)):
        return value, unit_
    if hasattr(value, "value_in_unit") and hasattr(value, "unit"):  # openmm.unit.Quantity
        if unit_ is None or isinstance(unit_, str):
            return value._value, (unit_ if unit_ is not None else value.unit)
        return value.value_in_unit(unit_), unit_
    if hasattr(value, "magnitude") and hasattr(value, "units"):  # pint.Quantity
        if unit_ is None:
            return value.magnitude, value.units
        return value.to(unit_).magnitude, unit_
    if isinstance(value, str):
        try:
            return float(value), unit_
        except ValueError:
            try:
                import pint

                q = pint.UnitRegistry()(value)
            except ImportError as err:
                raise ValueError(
                    f"cannot parse {value!r}: quantities given as strings need the optional 'pint' package"
                ) from err
            return strip_unit(q, unit_)
    return value, unit_


def is_unitless(value: Any) -> bool:
    """``unit.py:355-383``."""
    return strip_unit(value)[1] is None


def _prime_factors(n: int) -> list:
    out, p = [], 2
    while p * p <= n:
        while n % p == 0:
            out.append(p)
            n //= p
        p += 1 if p == 2 else 2
    if n > 1:
        out.append(n)
    return out


def get_closest_factors(value: int, n_factors: int, reverse: bool = False) -> np.ndarray:
    """
    The reference's factorisation of ``value`` into ``n_factors`` integers that
    are "as close to each other as possible" (``utility.py:16-75``): the prime
    factors, largest first, are multiplied greedily into slots that stay at or
    below the rounded ``n``-th root.  Kept verbatim in behaviour, including the
    cases where the greedy choice is not the closest one (72 -> (12, 6)).
    """
    root = value ** (1 / n_factors)
    root_int = int(np.round(root))
    if np.isclose(root, root_int):
        return np.full(n_factors, root_int, dtype=int)
    slots = np.ones(n_factors, dtype=int)
    slot = 0
    for rank, prime in enumerate(reversed(_prime_factors(int(value)))):
        placed = False
        while slot < n_factors and not placed:
            candidate = slots[slot] * prime
            if candidate <= root_int or (rank < n_factors and slots[slot] == 1):
                slots[slot] = candidate
                placed = True
            else:
                slot += 1
        if not placed:
            slots[np.argmin(slots)] *= prime
    slots = np.sort(slots)
    return slots[::-1] if reverse else slots


def center_of_mass(group, grouping: Optional[str] = None) -> np.ndarray:
    """
    Centres of mass of ``group`` (``grouping=None``), of its residues or of its
    segments — the ``center_of_mass(group, grouping)`` call form of
    ``molecule.py:16-441`` (no image flags, masses and positions taken from the
    group).  Equal-sized sub-groups use the reference's single einsum
    (``molecule.py:428``); ragged ones fall back to one weighted mean each.
    """
    if grouping not in {None, "residues", "segments"}:
        raise ValueError(
            f"Invalid grouping: '{grouping}'. Valid options are None, 'residues', and 'segments'."
        )
    positions = group.positions
    masses = group.masses
    if grouping is None:
        return np.einsum("a,ad->d", masses, positions) / masses.sum()
    subgroups = getattr(group, grouping)
    sizes = [g.atoms.n_atoms for g in subgroups]
    if all(s == sizes[0] for s in sizes):
        n = getattr(group, f"n_{grouping}")
        m = masses.reshape((n, -1))
        return np.einsum("...a,...ad->...d", m, positions.reshape((n, -1, 3))) / m.sum(axis=-1, keepdims=True)
    return np.array([g.atoms.center_of_mass() for g in subgroups])


def histogram_bin_edges(range_, n_bins: int) -> np.ndarray:
    """
    ``accelerated.py:13-43``.  Numba compiles ``min + i * ((max - min) / n_bins)``
    with ``fastmath=True``; on an x86-64 host with FMA (probed: LLVM IR and
    bit-for-bit output) that is ``fma((max - min) * i, 1 / n_bins, min)`` with the
    last edge forced to ``max``.  The fused multiply-add is evaluated exactly
    here (rational arithmetic on ``n_bins + 1`` values, then one rounding).
    """
    from fractions import Fraction

    array = np.asarray(range_, dtype=float)
    lo, hi = float(array.min()), float(array.max())
    width, inv = hi - lo, 1.0 / n_bins
    edges = np.empty(n_bins + 1)
    if lo == 0.0:
        edges[:] = (width * np.arange(n_bins + 1)) * inv
    else:
        f_inv, f_lo = Fraction(inv), Fraction(lo)
        for i in range(n_bins + 1):
            edges[i] = float(Fraction(width * i) * f_inv + f_lo)
    edges[-1] = hi
    return edges
