"""
AMBER NetCDF trajectories
=========================

Read side of ``mdcraft.openmm.file.NetCDFFile``
(``/root/reference/src/mdcraft/openmm/file.py:22-306``).  The reference goes through
the ``netCDF4`` library; here the NetCDF-3 header is parsed from a memory map and the
big-endian coordinates of a block of frames are byte-swapped on the GPU
(``csrc/ncdf.cu``), so that ``get_positions`` can also leave them in HBM
(``device=True``) for ``SqPlan.push`` / ``RdfPlan.push``.

Writing (``write_header`` / ``write_file`` / ``write_model``) belongs to the OpenMM
reporters and is not part of this package.  ``units=True`` needs ``openmm.unit``, as in
the reference; pass ``units=False`` where OpenMM is not installed.
"""

from __future__ import annotations

import warnings
from typing import Optional, Union

import numpy as np

from .. import _capi
from ..synthetic import MemoryTrajectory

__all__ = ["NetCDFFile", "NetCDFTrajectory"]


def _unit():
    try:
        from openmm import unit
    except ImportError as err:  # pragma: no cover - OpenMM is optional
        raise ImportError("units=True needs openmm.unit; pass units=False.") from err
    return unit


def _frame_list(frames, n: int):
    """Frame selection as the reference indexes a netCDF4 variable: None, an int, a list or a slice."""
    if frames is None:
        return np.arange(n), False
    if isinstance(frames, (int, np.integer)):
        return np.array([frames if frames >= 0 else n + frames]), True
    return np.arange(n)[frames], False


class NetCDFFile:
    """
    Reader for AMBER NetCDF trajectory files (``file.py:22``; ``mode`` must start with ``"r"``).

    Extra keyword argument: ``device`` (CUDA ordinal).
    """

    def __init__(self, file: str, mode: str = "r", restart: bool = False, *, device: int = 0, **kwargs):
        if not mode.startswith("r") or mode.startswith("r+"):
            raise NotImplementedError("Only reading (mode='r') is supported.")
        if not file.endswith((".nc", ".ncdf")):   # file.py:48-49
            file += ".nc"
        self._device = int(device)
        self._ctx = _capi.context(self._device)
        self._nc = _capi.NcdfFile(file)
        self._restart = False
        self._frame = self.get_num_frames()

    def close(self) -> None:
        self._nc.close()

    def get_num_frames(self) -> int:
        """``file.py:109-124``."""
        return self._nc.n_frames

    def get_num_atoms(self) -> int:
        """``file.py:126-141``."""
        return self._nc.n_atoms

    def get_dimensions(self, frames=None, units: bool = True):
        """``file.py:64-107``: ``(cell_lengths, cell_angles)``, float64 ``(n, 3)`` each (``(3,)`` for one frame)."""
        if not self._nc.has_box:
            raise KeyError("cell_lengths")
        idx, scalar = _frame_list(frames, self._nc.n_frames)
        lengths, angles = self._nc.cell_lengths[idx], self._nc.cell_angles[idx]
        if scalar:
            lengths, angles = lengths[0], angles[0]
        if units:
            u = _unit()
            return lengths * u.angstrom, angles * u.degree
        return lengths, angles

    def get_times(self, frames=None, units: bool = True):
        """``file.py:143-178``: float32, as stored."""
        idx, scalar = _frame_list(frames, self._nc.n_frames)
        times = self._nc.times[idx].astype(np.float32)
        times = times[0] if scalar else times
        return times * _unit().picosecond if units else times

    def get_positions(self, frames=None, units: bool = True, device: bool = False):
        """``file.py:180-215``: float32 ``(n, n_atoms, 3)`` (``(n_atoms, 3)`` for one frame).  ``device=True``: a CUDA
        ``torch`` tensor that never visits the host (``units`` must be False)."""
        return self._per_atom("coordinates", frames, units, device, lambda u: u.angstrom)

    def get_velocities(self, frames=None, units: bool = True, device: bool = False):
        """``file.py:217-261``: float32 velocities in angstrom / ps, ``None`` (with the reference's warning) if the
        file holds none."""
        if not self._nc.has_velocities:
            warnings.warn("The NetCDF file does not contain information about the atom velocities.")
            return None
        return self._per_atom("velocities", frames, units, device, lambda u: u.angstrom / u.picosecond)

    def get_forces(self, frames=None, units: bool = True, device: bool = False):
        """``file.py:263-306``: float32 forces in kcal / mol / angstrom, ``None`` (with the reference's warning) if
        the file holds none."""
        if not self._nc.has_forces:
            warnings.warn("The NetCDF file does not contain information about the forces acting on the atoms.")
            return None
        return self._per_atom("forces", frames, units, device, lambda u: u.kilocalorie_per_mole / u.angstrom)

    def _per_atom(self, variable: str, frames, units: bool, device: bool, unit_of):
        idx, scalar = _frame_list(frames, self._nc.n_frames)
        runs = np.split(idx, np.flatnonzero(np.diff(idx) != 1) + 1) if idx.size else []
        n_atoms = self._nc.n_atoms
        if device:
            import torch

            if units:
                raise ValueError("device=True returns a raw tensor: pass units=False.")
            out = torch.empty((idx.size, n_atoms, 3), dtype=torch.float32, device=f"cuda:{self._device}")
        else:
            out = np.empty((idx.size, n_atoms, 3), dtype=np.float32)
        at = 0
        for run in runs:   # consecutive frames go through one strided copy + one kernel
            self._nc.read(self._ctx, int(run[0]), run.size, out=out[at:at + run.size], variable=variable)
            at += run.size
        if scalar:
            out = out[0]
        return out * unit_of(_unit()) if units else out


class NetCDFTrajectory(MemoryTrajectory):
    """
    An AMBER NetCDF file behind the trajectory protocol of the analysis classes (``n_frames``, ``ts``, slicing,
    ``frames_block`` / ``frames_block_device``), so that ``synthetic.Universe(NetCDFTrajectory(path))`` drives
    ``StructureFactor`` / ``RadialDistributionFunction`` / ``DensityProfile`` with frame blocks that stay in HBM.
    """

    def __init__(self, filename: str, *, device: int = 0, frame_block: int = 64, dt: Optional[float] = None):
        self._file = NetCDFFile(filename, device=device)
        nc = self._file._nc
        self._device, self._frame_block = int(device), max(1, int(frame_block))
        self._cache_lo, self._cache = 0, None
        self._n_atoms = nc.n_atoms
        dims = None
        if nc.has_box:
            dims = np.concatenate((nc.cell_lengths, nc.cell_angles), axis=1).astype(np.float32)
        self._times = nc.times
        if dt is None:
            dt = float(self._times[1] - self._times[0]) if self._times is not None and len(self._times) > 1 else 1.0
        super().__init__(self._read_one, dims, n_frames=nc.n_frames, dt=dt)

    def _make_timestep(self, i: int, pos):
        ts = super()._make_timestep(i, pos)
        if self._times is not None:
            ts.time = float(self._times[i])
        return ts

    def _read_one(self, i: int) -> np.ndarray:
        lo = self._cache_lo
        if self._cache is None or not (lo <= i < lo + self._cache.shape[0]):
            n = min(self._frame_block, self._file.get_num_frames() - i)
            self._cache = self._file.get_positions(slice(i, i + n), units=False)
            self._cache_lo = lo = i
        return self._cache[i - lo]

    @property
    def n_atoms(self) -> int:
        return self._n_atoms

    def frames_block(self, lo: int, hi: int) -> np.ndarray:
        return self._file.get_positions(slice(int(lo), int(hi)), units=False)

    def frames_block_device(self, lo: int, hi: int):
        return self._file.get_positions(slice(int(lo), int(hi)), units=False, device=True)

    def close(self) -> None:
        self._file.close()
