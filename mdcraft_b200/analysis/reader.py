"""
Trajectory readers
==================

Drop-in for ``mdcraft.analysis.reader.LAMMPSDumpTrajectoryReader``
(``/root/reference/src/mdcraft/analysis/reader.py:24-572``), the step before the
S(q) / g(r) path.

The reference indexes the file line by line in Python (``reader.py:138-210``) and
then converts every atom line in a Python loop (``reader.py:495-548``).  Here the
file is memory mapped, the frame index is built by host threads counting newlines,
and a block of frames is parsed ON THE GPU: the raw text goes to HBM once, one
thread converts one atom line, and positions come out as float32 ``(F, N, 3)`` in
atom-id order, each value equal to ``float32(float(token))`` bit for bit, which is
what the reference's ``ts.positions[i] = tuple(str)`` stores.

:meth:`LAMMPSDumpTrajectoryReader.frames_block_device` leaves the block in HBM so
that it can be handed straight to the S(q) / g(r) plans (``_capi.SqPlan.push`` /
``_capi.RdfPlan.push``) without a host round trip.

Supported dump layouts (anything else raises): orthogonal and restricted
triclinic boxes (``ITEM: BOX BOUNDS xy xz yz``: the dimensions are formed as the
reference does, ``reader.py:372-388``); unscaled coordinates, wrapped (``x y z``)
or unwrapped (``xu yu zu``); with or without ``id``; a constant number of atoms.
``unwrap=True`` (image flags), scaled coordinates (``xs``, ``xsu``), general
triclinic boxes (``abc origin``), ``style grid`` dumps and ``extras`` are not: for scaled or unwrapped coordinates the reference adds the box origin to
EVERY atom once per atom (``reader.py:524-526`` sits inside the per-atom loop), so
there is no well-defined reference result to match unless the origin is zero.
"""

from __future__ import annotations

from typing import Optional, Sequence, Union

import numpy as np

from .. import _capi
from ..synthetic import MemoryTrajectory, Timestep

__all__ = ["LAMMPSDumpTrajectoryReader"]


class LAMMPSDumpTrajectoryReader(MemoryTrajectory):
    """
    LAMMPS dump trajectory reader (``reader.py:24-572``).

    Same positional / keyword arguments as the reference.  ``parallel`` and
    ``n_threads`` size the host thread pool that builds the frame index (the
    reference uses a process pool for the same job).  Extra keyword arguments:
    ``device`` (CUDA ordinal), ``frame_block`` (frames parsed per GPU pass, default:
    as many as fit 256 MiB of text, at most 64) and ``dt`` (time step length, as
    MDAnalysis readers take it).

    The object implements the trajectory protocol the analysis classes consume
    (``n_frames``, ``n_atoms``, ``ts``, ``trajectory[i]``, ``trajectory[a:b:c]``,
    iteration, ``check_slice_indices``, ``dt``), so
    ``synthetic.Universe(LAMMPSDumpTrajectoryReader(path))`` drives
    ``StructureFactor`` / ``RadialDistributionFunction`` where MDAnalysis is absent.
    """

    format = "LAMMPSDUMP"
    _CONVENTIONS = ["u", "su", "", "s"]

    def __init__(
        self,
        filename: str,
        conventions: Union[str, Sequence[Optional[str]], None] = None,
        unwrap: bool = False,
        *,
        extras=None,
        parallel: bool = False,
        n_threads: Optional[int] = None,
        device: int = 0,
        frame_block: Optional[int] = None,
        dt: float = 1.0,
        **kwargs,
    ) -> None:
        # reader.py:218-245, same messages
        if conventions is not None:
            if isinstance(conventions, str):
                if conventions not in self._CONVENTIONS:
                    raise ValueError(
                        f"Invalid convention '{conventions}'. Valid values: '" + "', '".join(self._CONVENTIONS) + "'."
                    )
                conventions = 3 * [conventions]
            elif len(conventions) != 3:
                raise ValueError("'conventions' must have length 3 when it is an array.")
            else:
                for c in conventions:
                    if c is not None and c not in self._CONVENTIONS:
                        raise ValueError(
                            f"Invalid convention '{c}'. Valid values: '" + "', '".join(self._CONVENTIONS) + "'."
                        )
                conventions = list(conventions)
        if unwrap:
            raise NotImplementedError("unwrap=True (image flags) is not supported by the GPU reader.")
        if extras is not None:
            raise NotImplementedError("'extras' are not supported by the GPU reader.")
        self.filename = str(filename)
        self._device = int(device)
        self._ctx = _capi.context(self._device)
        threads = 0 if (parallel and n_threads is None) else (int(n_threads) if n_threads else (0 if parallel else 1))
        try:
            self._file = _capi.DumpFile(self.filename, conventions, threads)
        except _capi.MdcError as err:
            if "coordinate information with the specified convention" in str(err):
                raise RuntimeError(str(err).split(": ", 1)[1]) from err
            raise
        self._conventions = list(self._file.conventions)
        if any(c is not None and "s" in c for c in self._conventions):
            self._file.close()
            raise NotImplementedError("Scaled coordinates (xs / xsu) are not supported by the GPU reader.")
        self._unwrap = False
        self._extras = None
        self._offsets = self._file.offsets[:-1].tolist()
        self._steps = self._file.steps
        n = self._file.n_atoms
        if frame_block is None:
            per_frame = max(1, self._file.block_bytes(0, 1))
            frame_block = int(max(1, min(64, (256 << 20) // per_frame)))
        self._frame_block = int(frame_block)
        self._cache_lo, self._cache = 0, None
        self._n_atoms = n
        super().__init__(self._read_one, self._file.box, n_frames=self._file.n_frames, dt=dt)

    # -- frames ------------------------------------------------------------
    def _make_timestep(self, i: int, pos) -> Timestep:
        # reader.py:372-373: data["step"], data["time"] = step * dt (MDAnalysis' Timestep.time returns data["time"])
        step = int(self._steps[i])
        return Timestep(pos, self._dims_of(i), frame=i, time=step * self.dt, data={"step": step, "time": step * self.dt})

    def _read_one(self, i: int) -> np.ndarray:
        lo = self._cache_lo
        if self._cache is None or not (lo <= i < lo + self._cache.shape[0]):
            n = min(self._frame_block, self._file.n_frames - i)
            self._cache = self._file.read(self._ctx, i, n)
            self._cache_lo = lo = i
        return self._cache[i - lo]

    def _read_frame(self, frame: int) -> Timestep:
        """``reader.py:335-351``."""
        return self[int(frame)]

    @property
    def n_atoms(self) -> int:
        return self._n_atoms

    def frames_block(self, lo: int, hi: int) -> np.ndarray:
        """Positions of frames ``[lo, hi)`` as one host ``(F, N, 3)`` float32 array (one GPU pass)."""
        return self._file.read(self._ctx, int(lo), int(hi) - int(lo))

    def frames_block_device(self, lo: int, hi: int):
        """Frames ``[lo, hi)`` parsed into a CUDA ``torch`` tensor ``(F, N, 3)`` that never leaves HBM."""
        import torch

        out = torch.empty((int(hi) - int(lo), self._n_atoms, 3), dtype=torch.float32, device=f"cuda:{self._device}")
        self._file.read(self._ctx, int(lo), int(hi) - int(lo), out=out)
        return out

    def host_parsed_values(self) -> int:
        """How many values of the last block needed the host parser (exotic number syntax or > 2^53 mantissas)."""
        return self._file.last_host_parsed()

    def close(self) -> None:
        """``reader.py:565-572``."""
        if getattr(self, "_file", None) is not None:
            self._file.close()
