"""
Analysis base classes
=====================

Drop-in surface of ``/root/reference/src/mdcraft/analysis/base.py`` for the
S(q) / g(r) path: ``Hash`` (:68), ``SerialAnalysisBase`` (:106, ``run`` +
``save``), ``NumbaAnalysisBase`` (:209, ``run(..., n_threads=)``) and
``DynamicAnalysisBase`` (:521, the serial/parallel switch).

The reference's process-pool frame farm (``ParallelAnalysisBase.run``,
base.py:308-518) is *not* reproduced: on this path frames are processed in
blocks on the GPU and, across GPUs, one contiguous frame block per rank with a
single all-reduce (see :mod:`mdcraft_b200.distributed`).  ``parallel``,
``n_jobs``, ``module`` ... are accepted and ignored so existing call sites keep
working.

When MDAnalysis is importable the classes derive from its real
``AnalysisBase``; otherwise from a protocol-compatible stand-in
(``_setup_frames`` -> ``_prepare`` -> per-frame ``_single_frame`` ->
``_conclude``).
"""

from __future__ import annotations

import logging
import time
from typing import TextIO, Union

import numpy as np

logger = logging.getLogger("mdcraft_b200.analysis")

try:  # pragma: no cover - MDAnalysis is absent in the build image
    from MDAnalysis.analysis.base import AnalysisBase as _MDAAnalysisBase
    from MDAnalysis.analysis.base import Results as _MDAResults

    FOUND_MDANALYSIS = True
except ImportError:
    _MDAAnalysisBase = None
    _MDAResults = None
    FOUND_MDANALYSIS = False


class Hash(dict):
    """``dict`` with attribute access (base.py:68-103)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


if FOUND_MDANALYSIS:  # pragma: no cover
    Results = _MDAResults
    AnalysisBase = _MDAAnalysisBase
else:

    class Results(dict):
        """Attribute-accessible result container (``MDAnalysis.analysis.base.Results``)."""

        def __getattr__(self, key):
            try:
                return self[key]
            except KeyError as err:
                raise AttributeError(f"'Results' object has no attribute '{key}'") from err

        def __setattr__(self, key, value):
            self[key] = value

        def __delattr__(self, key):
            try:
                del self[key]
            except KeyError as err:
                raise AttributeError(f"'Results' object has no attribute '{key}'") from err

    class AnalysisBase:
        """Protocol-compatible stand-in for ``MDAnalysis.analysis.base.AnalysisBase``."""

        def __init__(self, trajectory, verbose: bool = False, **kwargs):
            self._trajectory = trajectory
            self._verbose = verbose
            self.results = Results()

        def _setup_frames(self, trajectory, start=None, stop=None, step=None, frames=None):
            self._trajectory = trajectory
            if frames is not None:
                if not all(opt is None for opt in (start, stop, step)):
                    raise ValueError("start/stop/step cannot be combined with frames")
                self._sliced_trajectory = trajectory[frames]
            else:
                start, stop, step = trajectory.check_slice_indices(start, stop, step)
                # range(), not slice, semantics: check_slice_indices normalises a backward run that ends on frame 0 to
                # stop = -1, which a Python slice would read as "the last frame"
                self._sliced_trajectory = trajectory[np.arange(start, stop, step)]
            self.start, self.stop, self.step = start, stop, step
            self.n_frames = len(self._sliced_trajectory)
            self.frames = np.zeros(self.n_frames, dtype=int)
            self.times = np.zeros(self.n_frames)

        def _prepare(self):
            pass

        def _single_frame(self):
            raise NotImplementedError("Only implemented in child classes")

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


class _RunLog:
    """
    ``verbose=True``: one ``logging.info`` line when a run starts and one when it ends (the reference logs the
    start and the wall time of its process farm the same way, base.py:378-384, 514-515): device, frames per block,
    the plan's algorithm, frames/s.  Silent when ``verbose`` is false.
    """

    def __init__(self, analysis, verbose=None, n_frames=None):
        self.analysis = analysis
        self.on = bool(getattr(analysis, "_verbose", False) if verbose is None else verbose)
        self.t0 = time.perf_counter()
        if self.on:
            n = getattr(analysis, "n_frames", None) if n_frames is None else n_frames
            logger.info("%s: analysing %s frames on cuda:%s (frames per block: %s)", type(analysis).__name__,
                        "the selected" if n is None else n, getattr(analysis, "_device", "?"),
                        getattr(analysis, "_frame_block", "?"))

    @staticmethod
    def count(trajectory, start, stop, step, frames):
        """Number of frames a ``run(start, stop, step, frames)`` will visit (``None`` if it cannot be told)."""
        try:
            if frames is not None:
                f = np.asarray(frames)
                return int(f.sum()) if f.dtype == bool else int(f.size)
            return len(range(*trajectory.check_slice_indices(start, stop, step)))
        except Exception:
            return None

    def done(self) -> None:
        if not self.on:
            return
        a = self.analysis
        dt = max(time.perf_counter() - self.t0, 1e-12)
        plan = getattr(a, "_plan", None)
        detail = ""
        if plan is not None and hasattr(plan, "info"):
            try:
                info = plan.info()
                if "algo" in info:          # S(q) plans
                    detail = f", algorithm {info['algo']}" + (f" (nf = {info['nf']}, w = {info['w']})" if info["algo"] == "nufft" else "")
                else:                       # g(r) plans: the kernel path of the last frame block
                    parts = ["filter kernel" if info["filter"] else "FP64 kernel for every pair"]
                    if info["triclinic"]:
                        parts.append("triclinic cell")
                    if info["sorted"]:
                        parts.append("Hilbert-ordered frames")
                    if info["relative_rows"]:
                        parts.append("relative rows")
                    if info["ext_words"]:
                        parts.append("one histogram copy per bank" if info["copies_interleaved"] else "extended histograms")
                    detail = ", " + ", ".join(parts)
            except Exception:  # pragma: no cover - diagnostics only
                detail = ""
        logger.info("%s: %d frames in %.3f s (%.1f frames/s) on cuda:%s%s", type(a).__name__, a.n_frames, dt,
                    a.n_frames / dt, getattr(a, "_device", "?"), detail)


class SerialAnalysisBase(AnalysisBase):
    """base.py:106-206."""

    def __init__(self, trajectory, verbose: bool = False, **kwargs):
        super().__init__(trajectory, verbose, **kwargs)

    def run(self, start=None, stop=None, step=None, frames=None, verbose=None, **kwargs):
        log = _RunLog(self, verbose, _RunLog.count(self._trajectory, start, stop, step, frames))
        out = super().run(start, stop, step, frames, verbose, **kwargs)
        log.done()
        return out

    def save(self, file: Union[str, TextIO], archive: bool = True, compress: bool = True, **kwargs) -> None:
        """Saves ``results`` in NumPy format (base.py:167-206)."""
        if archive and compress:
            np.savez_compressed(file, **self.results, **kwargs)
        elif archive:
            np.savez(file, **self.results, **kwargs)
        else:
            for data in self.results:
                np.save(f"{file}_{data}", self.results[data], **kwargs)


class NumbaAnalysisBase(SerialAnalysisBase):
    """base.py:209-278.  ``n_threads`` only ever sized Numba's pool; ignored on the GPU."""

    def run(self, start=None, stop=None, step=None, frames=None, verbose=None, *, n_threads=None, **kwargs):
        return super().run(start=start, stop=stop, step=step, frames=frames, verbose=verbose, **kwargs)


class DynamicAnalysisBase(SerialAnalysisBase):
    """
    base.py:521-602.  The reference switches between the serial loop and the
    process farm on ``parallel``; here both run the same GPU frame-block path,
    and the farm's keyword arguments (``n_jobs``, ``module``, ``block``,
    ``method``) are accepted and dropped.
    """

    _FARM_KWARGS = ("n_jobs", "module", "block", "method")

    def __init__(self, trajectory, parallel: bool = False, verbose: bool = False, **kwargs):
        super().__init__(trajectory, verbose, **kwargs)
        self._parallel = parallel

    def run(self, start=None, stop=None, step=None, frames=None, verbose=None, **kwargs):
        for k in self._FARM_KWARGS:
            kwargs.pop(k, None)
        return super().run(start, stop, step, frames, verbose, **kwargs)


def run_together(analyses, start=None, stop=None, step=None, frames=None):
    """
    Runs several analyses in ONE pass over their frames: every frame is read once and handed to each analysis'
    ``_single_frame`` in turn, exactly as their own ``run(start, stop, step, frames)`` would have done (same
    ``frames`` / ``times`` bookkeeping, same results).  Because the analysis families queue their kernels on
    different stream lanes (``_capi.context``), g(r) and S(q) of the same frame block execute side by side on the GPU.

    The analyses normally share one trajectory.  Analyses of DIFFERENT trajectories (e.g. g(r) of one system next to
    S(q) of another) are stepped in lockstep, frame i of every trajectory in the same turn; the selected frame ranges
    must then have equal lengths.  Returns the analyses.
    """
    analyses = list(analyses)
    if not analyses:
        return analyses
    trajectories = []
    for a in analyses:
        if not any(a._trajectory is t for t in trajectories):
            trajectories.append(a._trajectory)
        a._setup_frames(a._trajectory, start=start, stop=stop, step=step, frames=frames)
    if len({a.n_frames for a in analyses}) != 1:
        raise ValueError("run_together needs the same number of selected frames in every trajectory.")
    timers = [_RunLog(a) for a in analyses]
    for a in analyses:
        a._prepare()
    # g(r) next to analyses on other stream lanes: its persistent pair kernel leaves a quarter of every SM free
    pair_plans = [a._plan for a in analyses if hasattr(getattr(a, "_plan", None), "set_blocks_per_sm")]
    if pair_plans and len(pair_plans) < len(analyses):
        for plan in pair_plans:
            plan.set_blocks_per_sm(3)
    # trajectories that hand out frame blocks resident in HBM (the GPU readers): every block is parsed once and
    # consumed by each analysis on the device, when all of them can take it (see structure._DeviceBlockRun)
    sources = [a._device_block_ranges() if hasattr(a, "_device_block_ranges") and getattr(a, "_drop_axis", None) is None
               and type(a).__name__ in ("RadialDistributionFunction", "StructureFactor") else None for a in analyses]
    if len(trajectories) == 1 and all(src is not None for src in sources):
        from ..synthetic import Timestep

        src = sources[0][0]
        idx = np.asarray(analyses[0]._sliced_trajectory._indices, dtype=np.int64)
        block_len = min(a._frame_block for a in analyses)
        i = 0
        while i < idx.size:
            j = i + 1
            while j < idx.size and j - i < block_len and idx[j] == idx[j - 1] + 1:
                j += 1
            lo, hi = int(idx[i]), int(idx[j - 1]) + 1
            block = src.frames_block_device(lo, hi)
            dims = np.stack([np.asarray(src._dims_of(f), dtype=np.float32) for f in range(lo, hi)])
            stamps = [src._make_timestep(f, None) for f in range(lo, hi)]
            for a in analyses:
                for k, ts in enumerate(stamps):
                    a.frames[i + k], a.times[i + k] = ts.frame, ts.time
                a._consume_device_block(block, dims, [Timestep(None, d) for d in dims])
            i = j
        if idx.size:
            last = src[int(idx[-1])]
            for a in analyses:
                a._ts, a._frame_index = last, idx.size - 1
    else:
        groups = [[a for a in analyses if a._trajectory is t] for t in trajectories]
        for i, stamps in enumerate(zip(*(g[0]._sliced_trajectory for g in groups))):
            for ts, group in zip(stamps, groups):
                for a in group:
                    a._frame_index, a._ts = i, ts
                    a.frames[i], a.times[i] = ts.frame, ts.time
                    a._single_frame()
    for a, t in zip(analyses, timers):
        a._conclude()
        t.done()
    return analyses
