"""
Loads the REAL reference modules from ``/root/reference`` in the build
container — TEST INFRASTRUCTURE ONLY (golden-vector generation and oracle
pinning).  ``/root/reference`` does not exist on the GPU box, so nothing that
runs there may call :func:`load`; use :func:`available` to gate.

Recipe: MDAnalysis and pint are not installed, so stub modules
are registered in ``sys.modules`` and the heavy package ``__init__`` files are
bypassed by pre-creating empty package objects whose ``__path__`` points into
the reference tree.  ``mdcraft.analysis.structure`` then imports unmodified and
its real ``StructureFactor`` / ``RadialDistributionFunction`` classes run on the
in-memory Universe from ``mdcraft_b200.synthetic``.  The only non-reference
arithmetic behind them is the ``capped_distance`` stub (``oracle.capped_distance``).
"""

from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

REF_SRC = "/root/reference/src"
_loaded = None


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "mdcraft", "analysis"))


class _Unit:
    """Just enough of a pint unit for ``results.units`` (never used numerically)."""

    def __init__(self, name="dimensionless"):
        self.name = name

    def __pow__(self, p):
        return _Unit(f"({self.name})**{p}")

    def __mul__(self, o):
        return _Unit(f"{self.name}*{getattr(o, 'name', o)}")

    __rmul__ = __mul__

    def __truediv__(self, o):
        return _Unit(f"{self.name}/{getattr(o, 'name', o)}")

    def __rtruediv__(self, o):
        return _Unit(f"{getattr(o, 'name', o)}/{self.name}")

    def m_as(self, unit):
        # only profile.py:30 asks for a magnitude: e / (eps0 * angstrom) in volts (CODATA 2018)
        return 1.602176634e-19 / (8.8541878128e-12 * 1e-10)

    def __repr__(self):
        return f"<unit {self.name}>"


class _UnitRegistry:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Unit(name)


class _AnalysisBase:
    """
    The run protocol of ``MDAnalysis.analysis.base.AnalysisBase`` [RECALL],
    consistent with how the reference re-implements it (base.py:386-389, 517).
    """

    def __init__(self, trajectory, verbose=False, **kwargs):
        from mdcraft_b200.analysis.base import Results

        self._trajectory = trajectory
        self._verbose = verbose
        self.results = Results()

    def _setup_frames(self, trajectory, start=None, stop=None, step=None, frames=None):
        import numpy as np

        self._trajectory = trajectory
        if frames is not None:
            self._sliced_trajectory = trajectory[frames]
        else:
            start, stop, step = trajectory.check_slice_indices(start, stop, step)
            self._sliced_trajectory = trajectory[start:stop:step]
        self.start, self.stop, self.step = start, stop, step
        self.n_frames = len(self._sliced_trajectory)
        self.frames = np.zeros(self.n_frames, dtype=int)
        self.times = np.zeros(self.n_frames)

    def _prepare(self):
        pass

    def _conclude(self):
        pass

    def run(self, start=None, stop=None, step=None, frames=None, verbose=None, **kwargs):
        self._setup_frames(self._trajectory, start=start, stop=stop, step=step, frames=frames)
        self._prepare()
        for i, ts in enumerate(self._sliced_trajectory):
            self._frame_index = i
            self._ts = ts
            self.frames[i] = ts.frame
            self.times[i] = ts.time
            self._single_frame()
        self._conclude()
        return self


def load():
    """Returns ``(structure_module, accelerated_module)`` of the real reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("/root/reference is not present (GPU box?)")
    sys.dont_write_bytecode = True  # the reference tree is read-only

    from mdcraft_b200 import synthetic
    from . import oracle

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if "MDAnalysis" not in sys.modules:
        mda = mod("MDAnalysis", AtomGroup=synthetic.AtomGroup, Universe=synthetic.Universe)
        mda.__path__ = []
        lib = mod("MDAnalysis.lib")
        lib.__path__ = []
        dist = mod(
            "MDAnalysis.lib.distances",
            capped_distance=oracle.capped_distance,
            minimize_vectors=lambda v, box: v,
        )
        ana = mod("MDAnalysis.analysis")
        ana.__path__ = []
        base = mod("MDAnalysis.analysis.base", AnalysisBase=_AnalysisBase)
        coords = mod("MDAnalysis.coordinates")
        coords.__path__ = []
        cbase = mod("MDAnalysis.coordinates.base", ReaderBase=object)
        mda.lib, lib.distances = lib, dist
        mda.analysis, ana.base = ana, base
        mda.coordinates, coords.base = coords, cbase
    if "pint" not in sys.modules:
        mod("pint", Quantity=type("Quantity", (), {}), UnitRegistry=_UnitRegistry)

    pkg = mod(
        "mdcraft", FOUND_OPENMM=False, Q_=sys.modules["pint"].Quantity,
        ureg=_UnitRegistry(), VERSION="reference",
    )
    pkg.__path__ = [os.path.join(REF_SRC, "mdcraft")]
    for sub in ("algorithm", "analysis"):
        m = mod(f"mdcraft.{sub}")
        m.__path__ = [os.path.join(REF_SRC, "mdcraft", sub)]
        setattr(pkg, sub, m)

    accelerated = importlib.import_module("mdcraft.algorithm.accelerated")
    sys.modules["mdcraft.algorithm"].accelerated = accelerated
    structure = importlib.import_module("mdcraft.analysis.structure")
    _loaded = (structure, accelerated)
    return _loaded


def load_polymer():
    """The real ``mdcraft.analysis.polymer`` (for ``SingleChainStructureFactor`` golden vectors)."""
    load()
    if "MDAnalysis.lib.log" not in sys.modules:
        log = types.ModuleType("MDAnalysis.lib.log")
        log.ProgressBar = lambda it, *a, **k: it
        sys.modules["MDAnalysis.lib.log"] = log
        sys.modules["MDAnalysis.lib"].log = log
    if "mdcraft.fit" not in sys.modules:
        fit = types.ModuleType("mdcraft.fit")
        fit.__path__ = [os.path.join(REF_SRC, "mdcraft", "fit")]
        sys.modules["mdcraft.fit"] = fit
        sys.modules["mdcraft"].fit = fit
    return importlib.import_module("mdcraft.analysis.polymer")


class _RefTimestep:
    """MDAnalysis ``Timestep`` subset the reference reader writes into (reader.py:353-562): float32 positions /
    velocities / forces, a float32 ``dimensions`` setter, ``data`` and ``dt``."""

    def __init__(self, n_atoms, dt=1.0, **kwargs):
        self.n_atoms = n_atoms
        self.positions = np.zeros((n_atoms, 3), dtype=np.float32)
        self.velocities = np.zeros((n_atoms, 3), dtype=np.float32)
        self.forces = np.zeros((n_atoms, 3), dtype=np.float32)
        self.has_velocities = self.has_forces = False
        self.data = {}
        self.dt = dt
        self.frame = -1
        self._dimensions = np.zeros(6, dtype=np.float32)

    @property
    def dimensions(self):
        return self._dimensions.copy()

    @dimensions.setter
    def dimensions(self, box):
        self._dimensions = np.asarray(box, dtype=np.float32)   # [RECALL] MDAnalysis stores the box as float32

    @property
    def time(self):
        return self.data.get("time", self.frame * self.dt)


class _RefReaderBase:
    """``MDAnalysis.coordinates.base.ReaderBase`` subset: what ``LAMMPSDumpTrajectoryReader.__init__`` relies on."""

    _Timestep = _RefTimestep

    def __init__(self, filename, convert_units=True, **kwargs):
        self.filename = filename
        self.convert_units = convert_units
        self._ts_kwargs = {"dt": kwargs.get("dt", 1.0)}

    def __len__(self):
        return self.n_frames

    def __getitem__(self, i):
        return self._read_frame(int(i))


def load_reader():
    """The real ``mdcraft.analysis.reader`` (LAMMPSDumpTrajectoryReader) on a stub ``ReaderBase``."""
    load()
    sys.modules["MDAnalysis.coordinates.base"].ReaderBase = _RefReaderBase
    dist = sys.modules["MDAnalysis.lib.distances"]
    if not hasattr(dist, "transform_StoR"):
        dist.transform_StoR = lambda coords, box: coords * np.asarray(box[:3])   # orthogonal boxes only; unused here
    if "MDAnalysis.lib.util" not in sys.modules:
        util = types.ModuleType("MDAnalysis.lib.util")
        util.anyopen = lambda f, *a, **k: open(f, *a, **k)
        util.store_init_arguments = lambda fn: fn
        sys.modules["MDAnalysis.lib.util"] = util
        sys.modules["MDAnalysis.lib"].util = util
    return importlib.import_module("mdcraft.analysis.reader")


def load_profile():
    """The real ``mdcraft.analysis.profile`` (DensityProfile)."""
    load()
    return importlib.import_module("mdcraft.analysis.profile")
