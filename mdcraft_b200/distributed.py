"""
Frame sharding across the GPUs of one node
==========================================

The reference farms frames out to a process pool and sums the per-frame
results on the host (``ParallelAnalysisBase.run`` ``base.py:308-518``,
``_single_frame_parallel`` ``structure.py:922-979``, reduction
``structure.py:985-988``).  Here every rank (one process per GPU,
``torchrun``-style) analyses ONE contiguous block of the sliced trajectory and
the on-device accumulators — the ``int64`` histogram and the ``float64`` S(q)
sums — are summed once, in ``_conclude``, with a single all-reduce: NCCL over
NVLink / NVSwitch directly on the library-owned device buffers when the process
group's backend is ``nccl``, a host all-reduce (``gloo``) otherwise.  Integer
sums keep g(r) counts bit-exact whatever the rank count.

``torch.distributed`` is plumbing only; nothing here touches the kernels.
"""

from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def _dist():
    import torch.distributed as dist

    return dist


def is_active() -> bool:
    try:
        dist = _dist()
    except ImportError:  # pragma: no cover
        return False
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def rank_and_world() -> Tuple[int, int]:
    if not is_active():
        return 0, 1
    dist = _dist()
    return dist.get_rank(), dist.get_world_size()


def frame_block(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """
    Contiguous block ``[lo, hi)`` of ``n_frames`` frames owned by ``rank``:
    ``ceil(n_frames / world)`` frames per rank, the last ranks may get fewer
    (or none).
    """
    if world < 1 or not 0 <= rank < world:
        raise ValueError("invalid rank / world size")
    per = -(-int(n_frames) // world)
    lo = min(rank * per, n_frames)
    return lo, min(lo + per, n_frames)


def shard_slice(trajectory_length: int, start=None, stop=None, step=None, rank: Optional[int] = None,
                world: Optional[int] = None) -> Tuple[int, Optional[int], int]:
    """
    ``(start, stop, step)`` for this rank such that the ranks together visit
    exactly the frames of ``range(*slice(start, stop, step).indices(n))``, each
    rank a contiguous run of them.  ``stop`` is ``None`` when a backward run ends
    on frame 0 (a literal ``-1`` would be read as "the last frame").
    """
    if rank is None or world is None:
        rank, world = rank_and_world()
    s0, s1, st = slice(start, stop, step).indices(int(trajectory_length))
    frames = range(s0, s1, st)
    lo, hi = frame_block(len(frames), rank, world)
    if lo >= hi:
        return s0, s0, st   # empty
    first = frames[lo]
    last = frames[hi - 1]
    if st > 0:
        return first, last + 1, st
    return first, (last - 1 if last > 0 else None), st


def run_sharded(analysis, start=None, stop=None, step=None, **kwargs):
    """
    ``analysis.run(...)`` on this rank's frame block.  The analysis must have
    been constructed with ``distributed=True`` so that ``_conclude`` all-reduces;
    afterwards every rank holds the full-trajectory ``results``.
    """
    n = len(analysis._trajectory)
    s0, s1, st = shard_slice(n, start, stop, step)
    return analysis.run(start=s0, stop=s1, step=st, **kwargs)


def run_together_sharded(analyses, start=None, stop=None, step=None):
    """
    :func:`mdcraft_b200.analysis.base.run_together` on this rank's frame block: several analyses (constructed with
    ``distributed=True``) fed from one pass over the rank's frames, each all-reducing in its ``_conclude``.
    """
    from .analysis.base import run_together

    analyses = list(analyses)
    n = len(analyses[0]._trajectory)
    s0, s1, st = shard_slice(n, start, stop, step)
    return run_together(analyses, start=s0, stop=s1, step=st)


def all_reduce_sum(values: np.ndarray, device: int = 0) -> np.ndarray:
    """Sum of a small host array over all ranks (identity when not distributed)."""
    values = np.ascontiguousarray(values)
    if not is_active():
        return values
    import torch

    dist = _dist()
    t = torch.from_numpy(values.copy())
    if dist.get_backend() == "nccl":
        t = t.cuda(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def all_reduce_device(accumulator, device: int = 0) -> None:
    """
    In-place sum over ranks of a library-owned device accumulator (anything
    exposing ``__cuda_array_interface__``).  With the ``nccl`` backend the
    all-reduce runs on the device buffer itself (no host round trip).
    """
    if not is_active():
        return
    import torch

    dist = _dist()
    t = torch.as_tensor(accumulator, device=torch.device("cuda", device))
    if dist.get_backend() == "nccl":
        torch.cuda.synchronize(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize(device)
    else:
        h = t.cpu()
        dist.all_reduce(h, op=dist.ReduceOp.SUM)
        t.copy_(h)
        torch.cuda.synchronize(device)


def all_reduce_host(array: np.ndarray) -> np.ndarray:
    """Sum of a host array over all ranks through the process group's CPU path (gloo)."""
    if not is_active():
        return array
    import torch

    dist = _dist()
    t = torch.from_numpy(np.ascontiguousarray(array).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()
