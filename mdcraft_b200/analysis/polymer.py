"""
Polymeric analysis — single-chain structure factor
==================================================

B200-native drop-in for ``mdcraft.analysis.polymer.SingleChainStructureFactor``
(``/root/reference/src/mdcraft/analysis/polymer.py:1054-1445``).  It re-uses the S(q) kernels:
every chain is one group of the density kernels and the per-chain ``|rho|^2`` are summed into a single device
accumulator (``mdc_scsf_plan_create``).  The rest of ``polymer.py`` (gyradius, end-to-end vectors, relaxation
times) is outside the hot path and not provided.
"""

from __future__ import annotations

from typing import Union

import numpy as np

from .. import _capi
from ..algorithm import center_of_mass, get_closest_factors, strip_unit
from .base import Hash, NumbaAnalysisBase
from .structure import LANE_SQ, _ANGSTROM, _FrameBlock, _context, _ureg, combine_equal_wavenumbers


class SingleChainStructureFactor(NumbaAnalysisBase):
    r"""
    Single-chain structure factor :math:`S_\mathrm{sc}(q)` of monodisperse polymers: the average over chains of
    :math:`|\sum_{j \in \mathrm{chain}} e^{i q \cdot r_j}|^2 / N_\mathrm{monomers}`.

    ``results``: ``scsf (n_groups, N_unique or Nq)``, ``wavenumbers``, ``units``.

    ``n_chains`` / ``n_monomers`` may be integers or one value per group; when omitted they are taken from the
    topology (segments = chains), as in ``_PolymerAnalysisBase`` (``polymer.py:152-179``).  Unlike the reference,
    which only normalises correctly for a single group (``polymer.py:1428`` broadcasts the per-group factor along
    the wavevector axis, and ``polymer.py:171`` builds ``n_monomers`` with the wrong length), every group is
    normalised by its own ``n_chains * n_monomers * n_frames``; for one group the results are identical.
    """

    def __init__(
        self,
        groups,
        groupings: Union[str, tuple] = "atoms",
        n_chains: Union[int, tuple] = None,
        n_monomers: Union[int, tuple] = None,
        *,
        form: str = "exp",
        dimensions=None,
        n_points: int = 32,
        n_surfaces: int = None,
        n_surface_points: int = 8,
        q_max=None,
        wavevectors: np.ndarray = None,
        sort: bool = True,
        unique: bool = True,
        unwrap: bool = False,
        parallel: bool = False,
        verbose: bool = True,
        precision: str = "fp32",
        algorithm: str = "auto",
        device: int = 0,
        frame_block: int = 8,
        distributed: bool = False,
        **kwargs,
    ) -> None:
        self._groups = list(groups) if isinstance(groups, (list, tuple)) else [groups]
        self.universe = self._groups[0].universe
        super().__init__(self.universe.trajectory, verbose, **kwargs)

        GROUPINGS = {"atoms", "residues"}
        self._n_groups = len(self._groups)
        if isinstance(groupings, str):
            if groupings not in GROUPINGS:
                raise ValueError(f"Invalid grouping '{groupings}'. Valid values: '" + "', '".join(GROUPINGS) + "'.")
            self._groupings = self._n_groups * [groupings]
        else:
            if self._n_groups != len(groupings):
                raise ValueError("The shape of 'groupings' is incompatible with that of 'groups'.")
            for gr in groupings:
                if gr not in GROUPINGS:
                    raise ValueError(f"Invalid grouping '{gr}'. Valid values: '" + "', '".join(GROUPINGS) + "'.")
            self._groupings = groupings

        if n_chains is None or n_monomers is None:
            self._internal = True
            self._n_chains = np.empty(self._n_groups, dtype=int)
            self._n_monomers = np.empty_like(self._n_chains)
            for i, (ag, gr) in enumerate(zip(self._groups, self._groupings)):
                self._n_chains[i] = ag.segments.n_segments
                entities = ag.n_atoms if gr == "atoms" else ag.n_residues
                self._n_monomers[i] = entities // self._n_chains[i]
        else:
            self._internal = False

            def per_group(value, name):
                if isinstance(value, (int, np.integer)):
                    return np.full(self._n_groups, int(value), dtype=int)
                if len(value) != self._n_groups:
                    raise ValueError(f"The shape of '{name}' is incompatible with that of 'groups'.")
                return np.asarray(value, dtype=int)

            self._n_chains = per_group(n_chains, "n_chains")
            self._n_monomers = per_group(n_monomers, "n_monomers")
        if form not in {"exp", "trig"}:
            raise ValueError("Invalid form. Valid values: 'exp', 'trig'.")

        # wavevectors: same construction as StructureFactor (polymer.py:1302-1365 = structure.py:1822-1887)
        self._explicit = None
        self._lattice = None
        if wavevectors is not None:
            self._explicit = np.ascontiguousarray(wavevectors, dtype=np.float64).reshape(-1, 3)
        else:
            if dimensions is None:
                dimensions = self.universe.dimensions
                dimensions = None if dimensions is None else dimensions[:3]
            dimensions = strip_unit(dimensions, "Å")[0]
            if dimensions is None:
                raise ValueError(
                    "System dimensions were not found, but are required when 'wavevectors' is not specified."
                )
            elif len(dimensions) != 3:
                raise ValueError("'dimensions' must have length 3.")
            dimensions = np.asarray(dimensions)
            if np.allclose(dimensions, dimensions[0]):
                L0 = np.full(3, float(dimensions[0]))
                if n_surfaces:
                    grid = 2 * np.pi * np.arange(n_points) / dimensions[0]
                    n_theta, n_phi = get_closest_factors(n_surface_points, 2, reverse=True)
                    theta = np.linspace(np.pi / (2 * n_theta + 4), np.pi / 2 - np.pi / (2 * n_theta + 4), n_theta)
                    phi = np.linspace(np.pi / (2 * n_phi + 4), np.pi / 2 - np.pi / (2 * n_phi + 4), n_phi)
                    directions = np.stack(
                        (
                            np.sin(theta) * np.cos(phi)[:, None],
                            np.sin(theta) * np.sin(phi)[:, None],
                            np.tile(np.cos(theta)[None, :], (n_phi, 1)),
                        ),
                        axis=-1,
                    )
                    self._explicit = np.einsum("o,tpd->otpd", grid[1 : n_surfaces + 1], directions).reshape(
                        (n_surfaces * n_surface_points, 3)
                    )
            else:
                L0 = np.array([float(L) for L in dimensions])
            self._lattice = (int(n_points), L0)
        self._q_max = None if q_max is None else float(strip_unit(q_max, "Å^-1")[0])
        if self._explicit is not None and self._q_max is not None:
            self._explicit = self._explicit[np.linalg.norm(self._explicit, axis=1) <= self._q_max]

        self._form = form
        self._sort = sort
        self._unique = unique
        self._unwrap = unwrap   # stored, as in the reference; the lattice phases are periodic in the box anyway
        self._verbose = verbose
        self._precision = precision
        self._algorithm = algorithm
        self._device = int(device)
        self._frame_block = max(1, int(frame_block))
        self._distributed = bool(distributed)
        self._plans = []
        n_points_, L0_ = self._lattice if self._lattice is not None else (0, None)
        explicit = self._explicit if self._explicit is not None and len(self._explicit) else None
        for M, N in zip(self._n_chains, self._n_monomers):
            self._plans.append(_capi.ScsfPlan(
                _context(self._device, LANE_SQ), int(M), int(N), n_points=n_points_, L0=L0_, q_max=self._q_max,
                wavevectors=explicit, precision=precision, algo=algorithm,
            ))
        self._wavevectors = self._plans[0].wavevectors()
        self._wavenumbers = np.linalg.norm(self._wavevectors, axis=1)

    def _prepare(self) -> None:
        for plan in self._plans:
            plan.reset()
        self._stages = [_FrameBlock(self._frame_block, int(M * N)) for M, N in zip(self._n_chains, self._n_monomers)]
        self._frames_seen = 0
        self.results.scsf = np.zeros((self._n_groups, len(self._wavenumbers)))
        self.results.wavenumbers = np.unique(self._wavenumbers.round(11))
        self.results.units = Hash({"results.wavenumbers": _ANGSTROM**-1 if _ureg is not None else "1 / angstrom"})

    def _single_frame(self) -> None:
        for ag, gr, M, N, stage in zip(self._groups, self._groupings, self._n_chains, self._n_monomers, self._stages):
            if gr == "atoms":
                positions = ag.positions
            elif self._internal:
                positions = center_of_mass(ag, gr)
            else:   # polymer.py:1407-1410: equal-sized residues given by the chain / monomer counts
                pos = ag.positions.reshape(M, N, -1, 3).astype(np.float64)
                m = ag.masses.reshape(M, N, -1)
                positions = np.einsum("...a,...ad->...d", m, pos) / m.sum(axis=-1, keepdims=True)
            stage.slot()[:] = np.asarray(positions).reshape(M * N, 3)
            stage.count += 1
        self._frames_seen += 1
        if self._stages[0].full():
            self._flush()

    def _flush(self) -> None:
        for plan, stage in zip(self._plans, self._stages):
            if stage.count:
                plan.push(stage.view(), n_frames=stage.count)
                stage.count = 0

    def _conclude(self) -> None:
        self._flush()
        n_frames = self.n_frames
        if self._distributed:
            from .. import distributed as dist

            n_frames = int(round(dist.all_reduce_sum(np.array([float(self._frames_seen)]), device=self._device)[0]))
            for plan in self._plans:
                dist.all_reduce_device(plan.accumulator(), device=self._device)
            self.n_frames_total = n_frames
        scsf = np.vstack([plan.read()[0] for plan in self._plans])
        scsf /= (self._n_chains * self._n_monomers * n_frames)[:, None]
        if self._unique:
            scsf = combine_equal_wavenumbers(scsf, self._wavenumbers, self.results.wavenumbers)
        self.results.scsf = scsf
        if self._sort:
            order = np.argsort(self.results.wavenumbers)
            self.results.wavenumbers = self.results.wavenumbers[order]
            self.results.scsf = self.results.scsf[:, order]
        self._stages = None
