"""
Density profiles
================

Drop-in for ``mdcraft.analysis.profile.DensityProfile``
(``/root/reference/src/mdcraft/analysis/profile.py:518-1500``), the 1-D cousin of the
g(r) histogram: number and charge density profiles along
the box axes.

Per frame the reference widens the float32 positions to float64, wraps them into
the box (``algorithm/topology.py:620-624``) and calls ``numba_histogram`` once per
axis and group (``profile.py:1128-1181``).  Here ``_single_frame`` only copies
positions into a pinned staging block; ``csrc/profile.cu`` wraps and bins on the
GPU in the reference's FP64 operation order, so the counts (hence the densities)
are identical.  The post-processing — surface charge densities and potential
profiles from the charge density (``profile.py:38-515, 1290-1500``) — is host
NumPy / SciPy on arrays of ``n_bins`` elements, as in the reference.

``recenter`` keeps its frame-to-frame state (image flags) and its centre-of-mass
shift on the host, in the reference's NumPy expressions (``topology.py:425-434``,
``profile.py:1137-1165``; a device reduction could not reproduce NumPy's summation
order bit for bit), and hands float64 positions to the same kernel.  ``parallel`` is
accepted and ignored.
"""

from __future__ import annotations

import logging
import warnings
from numbers import Real
from typing import Optional

import numpy as np

from .. import _capi
from ..algorithm import center_of_mass, strip_unit
from .base import DynamicAnalysisBase, Hash
from .structure import LANE_PROFILE, _ANGSTROM, _FrameBlock, _context, _ureg, histogram_bin_edges

__all__ = ["DensityProfile", "calculate_surface_charge_density", "calculate_potential_profile"]

#: e / (eps0 angstrom) in volts (real units) and 4 pi (reduced units), ``profile.py:29-34``
_ELECTROSTATIC_CONVERSION_FACTORS = (1.602176634e-19 / (8.8541878128e-12 * 1e-10), 4 * np.pi)
_GROUPINGS = {"atoms", "residues", "segments"}


def _is_unitless(value) -> bool:
    return not (hasattr(value, "units") or hasattr(value, "unit"))


def _check_profile_shapes(bins, rho_q) -> None:
    if bins.ndim != 1:
        raise ValueError("'bins' must be a one-dimensional array.")
    if rho_q.ndim not in {1, 2}:
        raise ValueError("'charge_density_profile' must be a one- or two-dimensional array.")
    if bins.shape[0] != rho_q.shape[-1]:
        raise ValueError("'bins' and 'charge_density_profile' must have the samelength in the last dimension.")


def _per_frame(value, name: str, what: str, rho_q) -> None:
    if value is not None and not isinstance(value, Real) and value.shape[0] != rho_q.shape[0]:
        raise ValueError(f"The number of {what} in '{name}' must be equal to the number of frames in "
                         "'charge_density_profile'.")


def _bins_and_length(bins, L):
    """Shifts uniformly spaced bins so that they start at dz / 2 (in place, with the reference's warning) and
    derives the system length when it is not given (``profile.py:129-145``)."""
    dz = bins[1] - bins[0]
    uniform = np.allclose(np.diff(bins), dz)
    if uniform and not np.isclose(bins[0], dz / 2):
        shift = bins[0] - dz / 2
        warnings.warn(f"'bins' currently does not start at zero, so it will be shifted left by {shift:.6g}.")
        bins -= shift
    if L is None:
        if not uniform:
            raise ValueError("'L' must be specified when 'bins' is not uniformly spaced.")
        L = bins[-1] - bins[0] + dz
    return dz, L


def calculate_surface_charge_density(bins, charge_density_profile, dielectric: float = None, *, L: float = None,
                                     dV=None, reduced: bool = False):
    """``profile.py:38-162``: sigma_q = (-int z rho_q(z) dz + eps dV / ecf) / L."""
    from scipy import integrate

    _check_profile_shapes(bins, charge_density_profile)
    if dV is not None:
        if dielectric is None:
            raise ValueError("'dielectric' must be specified when 'dV' is.")
        _per_frame(dV, "dVs", "potential differences", charge_density_profile)
    _, L = _bins_and_length(bins, L)
    y = -integrate.trapezoid(bins * charge_density_profile, bins)
    if dV is not None:
        y += dielectric * dV / _ELECTROSTATIC_CONVERSION_FACTORS[reduced]
    return y / L


def calculate_potential_profile(bins, charge_density_profile, dielectric: float, *, L: float = None, sigma_q=None,
                                dV=None, threshold: float = 1e-5, V0=0, method: str = "integral", pbc: bool = False,
                                reduced: bool = False):
    """``profile.py:165-515``: Poisson's equation in one dimension, by double integration or by a finite-difference
    matrix solve."""
    from scipy import integrate, sparse
    from scipy.sparse import linalg as sparse_linalg

    ecf = _ELECTROSTATIC_CONVERSION_FACTORS[reduced]
    rho_q = charge_density_profile
    _check_profile_shapes(bins, rho_q)
    _per_frame(sigma_q, "sigmas_q", "surface charge densities", rho_q)
    _per_frame(dV, "dVs", "potential differences", rho_q)
    _per_frame(V0, "V0s", "potentials", rho_q)
    dz, L = _bins_and_length(bins, L)
    if sigma_q is None and dV is not None:
        sigma_q = (dielectric * dV / ecf - integrate.trapezoid(bins * rho_q, bins)) / L

    if method == "integral":
        first = integrate.cumulative_trapezoid(rho_q, bins, initial=0)
        if sigma_q is None:
            warnings.warn("No surface charge density information. The value will be extracted from the integrated "
                          "charge density profile, which may be inaccurate due to numerical errors.")
            flat = np.abs(np.gradient(first if first.ndim == 1 else first.mean(axis=0))) < threshold
            cuts = np.where(np.diff(flat))[0] + 1
            if len(cuts) == 0:
                logging.warning("No bulk plateau region found in the integrated charge density profile. The average "
                                "value over the entire profile will be used.")
                sigma_q = first.mean(axis=-1, keepdims=True)
            else:
                mid = first.shape[-1] // 2
                lo, hi = cuts[cuts <= mid][-1], cuts[cuts >= mid][0]
                sigma_q = first[lo:hi].mean() if first.ndim == 1 else first[:, lo:hi].mean(axis=-1, keepdims=True)
        return V0 - ecf * integrate.cumulative_trapezoid(first - sigma_q, bins, initial=0) / dielectric
    if method == "matrix":
        if sigma_q is None:
            raise ValueError("No surface charge density information. Either 'sigma_q' or 'dV' must be provided when "
                             "method='matrix'.")
        if not np.allclose(np.diff(bins), dz):
            raise ValueError("'bins' must be uniformly spaced.")
        n = len(bins)
        A = sparse.diags((1, -2, 1), (-1, 0, 1), shape=(n, n), format="csc")
        b = rho_q.copy().T
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=sparse.SparseEfficiencyWarning)
            if pbc:
                A[0, -1] = A[-1, 0] = 1
                b *= -ecf * dz**2 / dielectric
                psi = np.empty_like(b)
                psi[1:] = sparse_linalg.spsolve(A[1:, 1:], b[1:])
                psi[0] = psi[-1]
                return psi
            A[0, :3] = -1.5, 2, -0.5
            A[-1, 0] = 1
            A[-1, -2:] = 0
            b[0] = ecf * dz * sigma_q / dielectric
            b[1:-1] *= -ecf * dz**2 / dielectric
            b[-1] = V0
            return sparse_linalg.spsolve(A, b)
    return None   # the reference falls through for an unknown method


class DensityProfile(DynamicAnalysisBase):
    """
    Number and charge density profiles along the box axes (``profile.py:518``).

    Same positional / keyword arguments and ``results`` keys as the reference: ``bins`` and ``bin_edges`` (one entry
    per axis), ``number_densities[ax]`` of shape ``(n_groups, n_bins)`` (``average=True``) or
    ``(n_groups, n_frames, n_bins)``, ``charge_densities[ax]`` when charges are known, ``times`` when
    ``average=False``, ``units``.  Extra keyword arguments: ``device``, ``frame_block``.
    """

    def __init__(self, groups, groupings="atoms", axes="xyz", n_bins=201, *, charges=None, dimensions=None,
                 recenter=None, dim_scales=1, average: bool = True, reduced: bool = False, parallel: bool = False,
                 verbose: bool = True, device: int = 0, frame_block: int = 16, **kwargs) -> None:
        self._groups = list(groups) if isinstance(groups, (list, tuple)) else [groups]
        self.universe = self._groups[0].universe
        super().__init__(self.universe.trajectory, parallel, verbose, **kwargs)
        self._n_groups = len(self._groups)

        def bad_grouping(gr):
            return ValueError(f"Invalid grouping '{gr}'. Valid values: '" + "', '".join(_GROUPINGS) + "'.")

        if isinstance(groupings, str):
            if groupings not in _GROUPINGS:
                raise bad_grouping(groupings)
            self._groupings = self._n_groups * [groupings]
        else:
            if self._n_groups != len(groupings):
                raise ValueError("The shape of 'groupings' is incompatible with that of 'groups'.")
            for gr in groupings:
                if gr not in _GROUPINGS:
                    raise bad_grouping(gr)
            self._groupings = groupings

        axis_list = (axes,) if isinstance(axes, int) else axes
        self._axis_indices = np.array([ord(a.lower()) - 120 if isinstance(a, str) else a for a in axis_list], dtype=int)
        self.axes = tuple(chr(120 + i) for i in self._axis_indices)
        if not all(ax in "xyz" for ax in self.axes):
            raise ValueError("Invalid axis passed in 'axes'.")
        self._n_axes = len(self.axes)

        if isinstance(n_bins, int):
            self._n_bins = n_bins * np.ones(self._n_axes, dtype=int)
        elif isinstance(n_bins, str):
            raise ValueError("'n_bins' must be an integer or an iterable object.")
        else:
            if len(n_bins) != self._n_axes:
                raise ValueError("The shape of 'n_bins' is incompatible with the number of axes to calculate density "
                                 "profiles along.")
            if not all(isinstance(n, int) for n in n_bins):
                raise ValueError("All bin counts in 'n_bins' must be integers.")
            self._n_bins = n_bins

        if charges is not None:
            if len(charges) != self._n_groups:
                raise ValueError("The shape of 'charges' is incompatible with that of 'groups'.")
            if reduced and not _is_unitless(charges):
                raise TypeError("'charges' cannot have units when 'reduced=True'.")
            self._charges = np.asarray(strip_unit(charges, "e")[0])
        elif hasattr(self.universe.atoms, "charges"):
            self._charges = np.empty(self._n_groups)
            for i, (ag, gr) in enumerate(zip(self._groups, self._groupings)):
                qs = getattr(ag, gr).charges
                if not np.allclose(qs, qs[0]):
                    self._charges = None
                    warnings.warn(f"Not all {gr} in `groups[{i}]` share the same charge. The charge density profile "
                                  "will not be calculated.")
                    break
                self._charges[i] = qs[0]
        else:
            self._charges = None

        if dimensions is not None:
            if len(dimensions) != 3:
                raise ValueError("'dimensions' must have length 3.")
            if reduced and not _is_unitless(dimensions):
                raise TypeError("'dimensions' cannot have units when 'reduced=True'.")
            self._dimensions = np.asarray(strip_unit(dimensions, "Å")[0])
        elif self.universe.dimensions is not None:
            self._dimensions = self.universe.dimensions[:3].copy()   # float32, as MDAnalysis hands it out
        else:
            raise ValueError("No system dimensions found or provided.")
        if isinstance(dim_scales, Real) or (len(dim_scales) == 3 and all(isinstance(f, Real) for f in dim_scales)):
            self._dimensions *= dim_scales
        else:
            raise ValueError("'dim_scales' must be a floating-point number or an array with shape (3,).")

        self._Ns = np.array([getattr(ag, f"n_{gr}") for ag, gr in zip(self._groups, self._groupings)], dtype=int)
        self._N = int(self._Ns.sum())
        edges = np.concatenate(([0], np.cumsum(self._Ns)))
        self._slices = [slice(int(a), int(b)) for a, b in zip(edges[:-1], edges[1:])]
        self._recenter = None if recenter is None else self._parse_recenter(recenter)

        self._dielectrics = np.full(self._n_axes, np.nan)
        self._dVs = np.full(self._n_axes, np.nan)
        self._average, self._reduced, self._verbose = average, reduced, verbose
        self._device, self._frame_block = int(device), max(1, int(frame_block))
        self._plan = None

    def _parse_recenter(self, recenter):
        """``profile.py:954-1005``: ``(indices or slice into the concatenated entities, target centre of mass)``."""
        if isinstance(recenter, tuple) and len(recenter) == 2:
            grp, com = recenter
            if len(com) != 3:
                raise ValueError("The target center of mass in 'recenter' must have length 3.")
            com = np.asarray(com)
        elif isinstance(recenter, (int, list)) or hasattr(recenter, "universe"):
            grp, com = recenter, self._dimensions / 2
        else:
            raise ValueError("Invalid value passed to 'recenter'. The argument must either be an atom group, its index "
                             "in 'groups', multiple groups/indices, or a tuple containing the aforementioned "
                             "information and a target center of mass, in that order.")

        def one(g):
            if isinstance(g, int):
                if not 0 <= g < self._n_groups:
                    raise ValueError("Invalid group index passed to 'recenter'.")
                return self._slices[g]
            for k, mine in enumerate(self._groups):
                if mine is g or bool(mine == g):
                    return self._slices[k]
            raise ValueError("The atom group passed to 'recenter' is not in 'groups'.")

        if isinstance(grp, int) or hasattr(grp, "universe"):
            return one(grp), com
        try:
            return np.r_[tuple(one(g) for g in grp)], com
        except ValueError:
            raise ValueError("Invalid atom group or index passed to 'recenter'.")

    def _entity_positions(self) -> np.ndarray:
        pos = np.empty((self._N, 3))
        for ag, gr, s in zip(self._groups, self._groupings, self._slices):
            pos[s] = ag.positions if gr == "atoms" else center_of_mass(ag, gr)
        return pos

    def _recentred(self) -> np.ndarray:
        """Unwrap across frames (topology.py:425-434, in place form) and shift the chosen entities' centre of mass
        onto its target (profile.py:1148-1165); float64, exactly the reference's NumPy expressions."""
        pos = self._entity_positions()
        dpos = pos - self._positions_old
        mask = np.abs(dpos) >= self._thresholds
        self._images[mask] -= np.sign(dpos[mask]).astype(int)
        self._positions_old[:] = pos[:]
        pos += self._images * self._dimensions
        sel, target = self._recenter
        masses = self._masses[sel]
        com = np.einsum("...a,...ad->...d", masses, pos[sel]) / masses.sum(axis=-1, keepdims=True)
        pos -= np.fromiter((0 if np.isnan(tx) else gx - tx for gx, tx in zip(com, target)), dtype=float, count=3)
        return pos

    # -- AnalysisBase protocol ------------------------------------------------
    def _prepare(self) -> None:
        self.results.bins, self.results.bin_edges = Hash(), Hash()
        for ia, ax, n in zip(self._axis_indices, self.axes, self._n_bins):
            self.results.bin_edges[ax] = histogram_bin_edges(np.asarray((0.0, self._dimensions[ia])), n)
            self.results.bins[ax] = (self.results.bin_edges[ax][:-1] + self.results.bin_edges[ax][1:]) / 2
        if self._plan is not None:
            self._plan.close()
        self._plan = _capi.ProfilePlan(_context(self._device, LANE_PROFILE), self._Ns, self._axis_indices,
                                       [int(n) for n in self._n_bins], np.asarray(self._dimensions, dtype=np.float64))
        self._stage = _FrameBlock(self._frame_block, self._N)
        if self._recenter is not None:
            self._masses = np.empty(self._N)
            for ag, gr, s in zip(self._groups, self._groupings, self._slices):
                self._masses[s] = getattr(ag, gr).masses
            self._sliced_trajectory[0]                       # profile.py:1038: start from the first analysed frame
            self._positions_old = self._entity_positions()
            self._images = np.zeros((self._N, 3), dtype=int)
            self._thresholds = self._dimensions / 2
            self._stage64 = np.empty((self._frame_block, self._N, 3))
        shape = [self._n_groups] if self._average else [self._n_groups, self.n_frames]
        self.results.number_densities = Hash({ax: np.zeros((*shape, n)) for ax, n in zip(self.axes, self._n_bins)})
        self._first_staged = 0
        density = _ANGSTROM**-3 if _ureg is not None else "1 / angstrom ** 3"
        self.results.units = Hash({"bins": _ANGSTROM, "number_densities": density})
        if self._charges is not None:
            self.results.charge_densities = Hash()
            self.results.units["charge_densities"] = (
                _ureg.elementary_charge / _ureg.angstrom**3 if _ureg is not None else "elementary_charge / angstrom ** 3"
            )
        if not self._average:
            self.results.times = np.fromiter((ts.time for ts in self._sliced_trajectory), dtype=float,
                                             count=self.n_frames)
            self.results.units["times"] = _ureg.picosecond if _ureg is not None else "picosecond"

    def _single_frame(self) -> None:
        if self._recenter is not None:
            self._stage64[self._stage.count] = self._recentred()
        else:
            slot = self._stage.slot()
            for ag, gr, s in zip(self._groups, self._groupings, self._slices):
                slot[s] = ag.positions if gr == "atoms" else center_of_mass(ag, gr)
        self._stage.count += 1
        if self._stage.full():
            self._flush()

    def _flush(self) -> None:
        n = self._stage.count
        if n == 0:
            return
        block = self._stage.view() if self._recenter is None else self._stage64[:n]
        per_frame = self._plan.push(block, n_frames=n, per_frame=not self._average)
        if per_frame is not None:
            lo = self._first_staged
            for ax, counts in zip(self.axes, per_frame):      # counts (n, n_groups, n_bins)
                self.results.number_densities[ax][:, lo:lo + n] = counts.transpose(1, 0, 2)
        self._first_staged += n
        self._stage.count = 0

    def _conclude(self) -> None:
        self._flush()
        if self._average:
            summed, _ = self._plan.read()
            for ax, counts in zip(self.axes, summed):
                self.results.number_densities[ax] += counts
        # profile.py:1236-1248, same expressions (and therefore the same dtype promotion)
        volume = np.prod(self._dimensions)
        for ax, n in zip(self.axes, self._n_bins):
            denom = volume / n
            if self._average:
                denom *= self.n_frames
            self.results.number_densities[ax] /= denom
            if self._charges is not None:
                self.results.charge_densities[ax] = np.einsum("g,g...b->...b", self._charges,
                                                               self.results.number_densities[ax])
        self._stage = None

    # -- post-processing (host) ----------------------------------------------------
    def _validate_input(self, value, unit_: str, name: str, n_axes: int):
        if value is None:
            return n_axes * [value]
        if self._reduced and not _is_unitless(value):
            raise TypeError(f"'{name}' cannot have units when 'reduced=True'.")
        value = strip_unit(value, unit_)[0]
        if isinstance(value, Real):
            return n_axes * [value]
        if len(value) != n_axes:
            raise ValueError(f"The length of '{name}' must match the number of axes.")
        return value

    def _select_axes(self, axes, dielectrics):
        if not hasattr(self.results, "charge_densities"):
            raise RuntimeError("Either call DensityProfile.run() before DensityProfile.calculate_potential_profiles() "
                               "or provide charge information when initializing the DensityProfile object.")
        if axes is None:
            axes, axis_indices = self.axes, self._axis_indices
        else:
            if isinstance(axes, Real) or any(not isinstance(ax, str) for ax in axes):
                raise ValueError("'axes' must only contain strings.")
            axes = tuple(axes)
            axis_indices = [ord(ax.lower()) - 120 for ax in axes]
        try:
            relative = [self.axes.index(ax) for ax in axes]
        except ValueError:
            raise ValueError("Invalid axis passed in 'axes'.")
        if dielectrics is not None:
            self._dielectrics[relative] = self._validate_input(dielectrics, "", "dielectrics", len(axes))
        elif np.any(np.isnan(self._dielectrics)):
            raise ValueError("No dielectric constants found or provided.")
        return axes, relative, self._dimensions[axis_indices]

    def calculate_surface_charge_densities(self, axes=None, dielectrics=None, *, dVs=None) -> None:
        """``profile.py:1290-1389``."""
        axes, relative, dimensions = self._select_axes(axes, dielectrics)
        if dVs is not None:
            self._dVs[relative] = self._validate_input(dVs, "V", "dVs", len(axes))
        if not hasattr(self.results, "surface_charge_densities"):
            shape = [self._n_axes] if self._average else [self._n_axes, self.n_frames]
            self.results.surface_charge_densities = np.full(shape, np.nan)
            self.results.units["surface_charge_densities"] = (
                _ureg.elementary_charge / _ureg.angstrom**2 if _ureg is not None else "elementary_charge / angstrom ** 2"
            )
        for i, (ax, dielectric, L, dV) in enumerate(zip(axes, self._dielectrics, dimensions, self._dVs)):
            self.results.surface_charge_densities[i] = calculate_surface_charge_density(
                self.results.bins[ax], self.results.charge_densities[ax], dielectric, L=L,
                dV=None if np.isnan(dV) else dV, reduced=self._reduced)

    def calculate_potential_profiles(self, axes=None, dielectrics=None, *, sigmas_q=None, dVs=None, thresholds=1e-5,
                                     V0s=0, methods="integral", pbcs=False) -> None:
        """``profile.py:1391-1500``."""
        axes, relative, dimensions = self._select_axes(axes, dielectrics)
        n_axes = len(axes)
        if not hasattr(self.results, "potentials"):
            self.results.potentials = Hash()
            self.results.units["potentials"] = _ureg.volt if _ureg is not None else "volt"
        if sigmas_q is None:
            if hasattr(self.results, "surface_charge_densities") and np.all(np.isfinite(self._dVs[relative])):
                sigmas_q = -self.results.surface_charge_densities
            else:
                sigmas_q = self._validate_input(sigmas_q, "e/Å^2", "sigmas_q", n_axes)
                if dVs is not None:
                    self._dVs[relative] = self._validate_input(dVs, "V", "dVs", n_axes)
        else:
            sigmas_q = self._validate_input(sigmas_q, "e/Å^2", "sigmas_q", n_axes)
        V0s = self._validate_input(V0s, "V", "V0s", n_axes)
        thresholds = self._validate_input(thresholds, "", "thresholds", n_axes)
        METHODS = {"integral", "matrix"}
        if isinstance(methods, str) and methods in METHODS:
            methods = n_axes * [methods]
        elif len(methods) != n_axes:
            raise ValueError("The length of 'methods' must match the number of axes.")
        else:
            for method in methods:
                if method not in METHODS:
                    raise ValueError(f"Invalid method '{method}'. Valid values: '" + "', '".join(METHODS) + "'.")
        if isinstance(pbcs, bool):
            pbcs = n_axes * [pbcs]
        elif len(pbcs) != n_axes:
            raise ValueError("The length of 'pbcs' must match the number of axes.")
        elif not all(isinstance(pbc, bool) for pbc in pbcs):
            raise ValueError("All values in 'pbcs' must be booleans.")
        for ax, dielectric, L, sigma_q, dV, threshold, V0, method, pbc in zip(
                axes, self._dielectrics, dimensions, sigmas_q, self._dVs, thresholds, V0s, methods, pbcs):
            self.results.potentials[ax] = calculate_potential_profile(
                self.results.bins[ax], self.results.charge_densities[ax], dielectric, L=L, sigma_q=sigma_q,
                dV=None if np.isnan(dV) else dV, threshold=threshold, V0=V0, method=method, pbc=pbc,
                reduced=self._reduced)
