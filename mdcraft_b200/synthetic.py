"""
In-memory trajectories and synthetic fluids
===========================================

Host-side plumbing only.  MDAnalysis is not installed in the build image, so
tests and benchmarks drive the analysis classes on the duck-typed
:class:`Universe` / :class:`AtomGroup` / :class:`MemoryTrajectory` below.  They
expose exactly the attributes the reference consumes from MDAnalysis on the
S(q) / g(r) path (``/root/reference/src/mdcraft/analysis/structure.py:805-811,
871-882, 1778-1780, 1815, 1826, 1889-1893, 1943-1944``):

* ``AtomGroup.universe``, ``.positions`` (fresh float32 ``(n, 3)`` copy),
  ``.n_atoms``, ``.n_residues``, ``.n_segments``, ``==``
* ``Universe.trajectory``, ``.dimensions``, ``.atoms``
* ``ts.dimensions`` (float32 ``(6,)``), ``ts.volume``, ``ts.frame``, ``ts.time``
* ``trajectory[i]``, ``trajectory[slice]``, ``len()``, ``.n_frames``, ``.dt``,
  ``.check_slice_indices``

The ``make_*`` generators produce the synthetic fluids of SURVEY.md §8(d)
(configs C1-C5) from fixed seeds, one frame at a time so that a
1,000-frame x 100,000-particle trajectory never has to live in host memory.
"""

from __future__ import annotations

from typing import Callable, Optional, Sequence, Union

import numpy as np

__all__ = [
    "Timestep",
    "MemoryTrajectory",
    "Universe",
    "AtomGroup",
    "cubic_box_length",
    "lj_like_frame",
    "uniform_frame",
    "make_universe",
    "CONFIGS",
    "make_config",
]


class Timestep:
    """One frame: positions, box and bookkeeping (MDAnalysis ``Timestep`` subset)."""

    __slots__ = ("positions", "_dimensions", "frame", "time", "data")

    def __init__(self, positions, dimensions, frame=0, time=0.0, data=None):
        self.positions = positions
        self._dimensions = dimensions
        self.frame = frame
        self.time = time
        self.data = {} if data is None else data

    @property
    def dimensions(self):
        # MDAnalysis hands out a fresh float32 copy on every access; the
        # reference mutates it in place when ``drop_axis`` is set
        # (structure.py:891).
        return None if self._dimensions is None else self._dimensions.copy()

    @property
    def volume(self) -> float:
        # [RECALL] MDAnalysis: box_volume(dimensions) — for an orthorhombic
        # cell the product of the three float32 lengths evaluated in double.
        d = self._dimensions
        if d is None:
            return 0.0
        lx, ly, lz = (float(v) for v in d[:3])
        if np.all(d[3:] == 90.0):
            return lx * ly * lz
        a, b, g = np.deg2rad(d[3:].astype(np.float64))
        ca, cb, cg = np.cos(a), np.cos(b), np.cos(g)
        return lx * ly * lz * np.sqrt(
            1 - ca * ca - cb * cb - cg * cg + 2 * ca * cb * cg
        )

    @property
    def n_atoms(self) -> int:
        return self.positions.shape[0]


class MemoryTrajectory:
    """
    A trajectory backed by an array ``(F, N, 3)`` or by a per-frame generator
    ``frame_fn(i) -> (N, 3) float32``.  ``dimensions`` is ``(6,)`` (NVT) or
    ``(F, 6)`` (NPT).
    """

    def __init__(
        self,
        positions: Union[np.ndarray, Callable[[int], np.ndarray]],
        dimensions: Optional[np.ndarray],
        n_frames: Optional[int] = None,
        dt: float = 1.0,
        _indices: Optional[np.ndarray] = None,
        _parent: Optional["MemoryTrajectory"] = None,
    ):
        if callable(positions):
            if n_frames is None:
                raise ValueError("n_frames is required with a frame generator.")
            self._frame_fn = positions
            self._array = None
            self._n_total = int(n_frames)
        else:
            arr = np.asarray(positions, dtype=np.float32)
            if arr.ndim == 2:
                arr = arr[None]
            self._array = arr
            self._frame_fn = None
            self._n_total = arr.shape[0]
        if dimensions is not None:
            dimensions = np.asarray(dimensions, dtype=np.float32)
            if dimensions.ndim == 1 and dimensions.shape[0] == 3:
                dimensions = np.concatenate(
                    (dimensions, np.full(3, 90.0, dtype=np.float32))
                )
        self._dimensions = dimensions
        self.dt = dt
        self._indices = (
            np.arange(self._n_total) if _indices is None else np.asarray(_indices)
        )
        # Slices share the "current frame" cursor with their parent, like
        # MDAnalysis' FrameIteratorSliced does with its reader.
        self._root = self if _parent is None else _parent._root
        if _parent is None:
            self._ts = None
            self._read(0)

    # -- reading ---------------------------------------------------------
    def _dims_of(self, i: int):
        d = self._dimensions
        if d is None:
            return None
        return d if d.ndim == 1 else d[i]

    def _read(self, i: int) -> Timestep:
        i = int(i)
        if self._root._array is not None:
            pos = self._root._array[i]
        else:
            pos = np.ascontiguousarray(self._root._frame_fn(i), dtype=np.float32)
        ts = self._root._make_timestep(i, pos)
        self._root._ts = ts
        return ts

    def _make_timestep(self, i: int, pos) -> Timestep:
        return Timestep(pos, self._dims_of(i), frame=i, time=i * self.dt)

    @property
    def ts(self) -> Timestep:
        return self._root._ts

    @property
    def n_frames(self) -> int:
        return len(self._indices)

    @property
    def n_atoms(self) -> int:
        return self.ts.n_atoms

    def __len__(self) -> int:
        return len(self._indices)

    def __iter__(self):
        for i in self._indices:
            yield self._read(i)

    def __getitem__(self, item):
        if isinstance(item, (int, np.integer)):
            n = len(self._indices)
            if item < -n or item >= n:
                raise IndexError(f"Index {item} exceeds length of trajectory ({n}).")
            return self._read(self._indices[item])
        if isinstance(item, slice):
            idx = self._indices[item]
        else:
            item = np.asarray(item)
            idx = self._indices[np.flatnonzero(item) if item.dtype == bool else item]
        return MemoryTrajectory(
            self._root._array if self._root._array is not None else self._root._frame_fn,
            self._root._dimensions,
            n_frames=self._root._n_total,
            dt=self._root.dt,
            _indices=idx,
            _parent=self,
        )

    def check_slice_indices(self, start, stop, step):
        """Same contract as ``ProtoReader.check_slice_indices``."""
        for v, name in ((start, "start"), (stop, "stop"), (step, "step")):
            if not (v is None or isinstance(v, (int, np.integer))):
                raise TypeError(f"Slice indices are not integers: {name}={v!r}")
        if step == 0:
            raise ValueError("Step size is zero")
        return slice(start, stop, step).indices(len(self))

    # -- optional bulk access used by the GPU frame-block pipeline ---------
    def frames_block(self, lo: int, hi: int) -> np.ndarray:
        """Positions of frames ``indices[lo:hi]`` as one ``(F, N, 3)`` array."""
        idx = self._indices[lo:hi]
        if self._root._array is not None:
            return self._root._array[idx]
        return np.stack([self._root._frame_fn(int(i)) for i in idx]).astype(
            np.float32, copy=False
        )


class Universe:
    """Minimal stand-in for ``MDAnalysis.Universe``."""

    def __init__(
        self,
        trajectory: MemoryTrajectory,
        resindices: Optional[np.ndarray] = None,
        segindices: Optional[np.ndarray] = None,
        masses: Optional[np.ndarray] = None,
    ):
        self.trajectory = trajectory
        n = trajectory.n_atoms
        self._resindices = (
            np.arange(n) if resindices is None else np.asarray(resindices)
        )
        self._segindices = (
            np.zeros(n, dtype=int) if segindices is None else np.asarray(segindices)
        )
        self._masses = np.ones(n) if masses is None else np.asarray(masses, float)
        self.atoms = AtomGroup(self, np.arange(n))

    @property
    def dimensions(self):
        return self.trajectory.ts.dimensions

    def select(self, lo: int, hi: int) -> "AtomGroup":
        return AtomGroup(self, np.arange(lo, hi))


class _GroupList(list):
    """``ResidueGroup`` / ``SegmentGroup`` stand-in: a list of sub-groups that knows its length by both names."""

    @property
    def n_residues(self) -> int:
        return len(self)

    @property
    def n_segments(self) -> int:
        return len(self)

    @property
    def masses(self) -> np.ndarray:
        """``ResidueGroup.masses`` / ``SegmentGroup.masses``: total mass of every member."""
        return np.array([g.masses.sum() for g in self])


class AtomGroup:
    """Minimal stand-in for ``MDAnalysis.AtomGroup`` (index list into a Universe)."""

    def __init__(self, universe: Universe, indices: np.ndarray):
        self.universe = universe
        self.indices = np.asarray(indices, dtype=np.intp)
        i = self.indices
        self._contiguous = bool(
            i.size == 0 or (i[-1] - i[0] + 1 == i.size and np.all(np.diff(i) == 1))
        )

    @property
    def positions(self) -> np.ndarray:
        pos = self.universe.trajectory.ts.positions
        if self._contiguous and self.indices.size:
            return np.array(pos[self.indices[0] : self.indices[-1] + 1], dtype=np.float32)
        return np.asarray(pos[self.indices], dtype=np.float32)

    @property
    def atoms(self) -> "AtomGroup":
        """MDAnalysis groups (and Residue / Segment objects) expose their atoms as ``.atoms``."""
        return self

    @property
    def masses(self) -> np.ndarray:
        return self.universe._masses[self.indices]

    @property
    def n_atoms(self) -> int:
        return int(self.indices.size)

    @property
    def n_residues(self) -> int:
        return int(np.unique(self.universe._resindices[self.indices]).size)

    @property
    def n_segments(self) -> int:
        return int(np.unique(self.universe._segindices[self.indices]).size)

    def _split(self, labels: np.ndarray):
        """Sub-groups of this group sharing a residue / segment index, in index order."""
        mine = labels[self.indices]
        out = _GroupList()
        for lab in np.unique(mine):
            out.append(AtomGroup(self.universe, self.indices[mine == lab]))   # ``.atoms`` of a Residue / Segment
        return out

    @property
    def residues(self):
        return self._split(self.universe._resindices)

    @property
    def segments(self):
        return self._split(self.universe._segindices)

    def center_of_mass(self) -> np.ndarray:
        m = self.masses
        return (m[:, None] * self.positions.astype(np.float64)).sum(axis=0) / m.sum()

    def __len__(self) -> int:
        return self.n_atoms

    def __eq__(self, other) -> bool:
        return (
            isinstance(other, AtomGroup)
            and other.universe is self.universe
            and np.array_equal(other.indices, self.indices)
        )

    def __hash__(self):
        return id(self)


# ----------------------------------------------------------------------------
# Synthetic fluids
# ----------------------------------------------------------------------------


def cubic_box_length(n: int, rho: float) -> np.float32:
    """Edge of the cubic box holding ``n`` particles at reduced density ``rho``."""
    return np.float32((n / rho) ** (1.0 / 3.0))


_LATTICES: dict = {}


def _simple_cubic(n: int, L: float) -> np.ndarray:
    """The first ``n`` sites of the smallest simple-cubic lattice with at least ``n`` sites in a box of edge ``L``."""
    key = (int(n), float(L))
    lat = _LATTICES.get(key)
    if lat is None:
        m = int(round(n ** (1.0 / 3.0)))
        while m**3 < n:
            m += 1
        g = (np.arange(m) + 0.5) * (float(L) / m)
        lat = np.ascontiguousarray(np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)[:n])
        if len(_LATTICES) < 8:
            _LATTICES[key] = lat
    return lat


def lj_like_frame(n: int, L: float, seed: int, jitter: float = 0.15) -> np.ndarray:
    """Simple-cubic lattice plus Gaussian jitter, wrapped into ``[0, L)``."""
    rng = np.random.default_rng(seed)
    pos = _simple_cubic(n, L)
    pos = pos + rng.normal(scale=jitter, size=pos.shape)
    pos = np.mod(pos, float(L)).astype(np.float32)
    pos[pos >= np.float32(L)] = 0.0
    return pos


def uniform_frame(n: int, L: float, seed: int) -> np.ndarray:
    """Ideal-gas-like (soft-core) configuration: uniform in ``[0, L)``."""
    rng = np.random.default_rng(seed)
    pos = (rng.random((n, 3)) * float(L)).astype(np.float32)
    pos[pos >= np.float32(L)] = 0.0
    return pos


def make_universe(
    n: int,
    n_frames: int,
    rho: float,
    kind: str,
    seed0: int,
    materialize: bool = False,
    box_scale: Optional[Sequence[float]] = None,
) -> Universe:
    """
    Build a synthetic NVT universe.  ``kind`` is ``"lj"`` (lattice + jitter) or
    ``"uniform"``.  Frame ``i`` is drawn from ``default_rng(seed0 + i)``.
    ``box_scale`` turns the cube into an orthorhombic cell (lengths
    ``L * box_scale``) by an affine stretch of the same points.
    """
    L = cubic_box_length(n, rho)
    scale = np.ones(3, np.float32) if box_scale is None else np.asarray(box_scale, np.float32)
    fn0 = lj_like_frame if kind == "lj" else uniform_frame

    def frame_fn(i: int) -> np.ndarray:
        p = fn0(n, L, seed0 + i)
        if box_scale is not None:
            p = p * scale
            lim = (L * scale).astype(np.float32)
            p = np.where(p >= lim, np.float32(0), p).astype(np.float32)
        return p

    dims = np.array([*(L * scale), 90, 90, 90], dtype=np.float32)
    if materialize:
        arr = np.stack([frame_fn(i) for i in range(n_frames)])
        traj = MemoryTrajectory(arr, dims)
    else:
        traj = MemoryTrajectory(frame_fn, dims, n_frames=n_frames)
    return Universe(traj)


#: name -> (n_particles, n_frames, rho*, kind, seed0)   
CONFIGS = {
    "C1": (1_000, 100, 0.8, "lj", 1000),
    "C2": (100_000, 1_000, 2.5, "uniform", 2000),
    "C3": (100_000, 1_000, 0.8, "lj", 3000),
    "C4": (200_000, 1_000, 0.8, "lj", 4000),
    "C5": (1_000_000, 10_000, 0.8, "uniform", 5000),
}


def make_config(name: str, n_frames: Optional[int] = None, n: Optional[int] = None, **kw) -> Universe:
    """Universe for one of the named configs, optionally with fewer frames / particles."""
    n0, f0, rho, kind, seed0 = CONFIGS[name]
    return make_universe(n or n0, n_frames or f0, rho, kind, seed0, **kw)
