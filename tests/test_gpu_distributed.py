"""
GPU: the frame-sharded class path (``distributed=True`` + ``run_sharded`` / ``run_together_sharded``) on REAL kernel
output.  Two ranks share cuda:0 over a ``gloo`` process group (NCCL refuses two ranks on one device; with ``gloo``
``distributed.all_reduce_device`` moves the library-owned device accumulators through the host), each analyses its
contiguous frame block, and every rank must end up with the results of the serial run over all frames
(reference semantics: ``structure.py:985-988`` sums the per-frame results of the process farm, ``base.py:308-518``).
``bench.py --gpus N`` does the same over NCCL and checks it (``allreduce_check``).
"""

import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

N, F, L = 3000, 7, 15.0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _universes():
    from mdcraft_b200 import synthetic

    rng = np.random.default_rng(5)
    boxes = np.tile(np.array([L, L, L, 90, 90, 90], np.float32), (F, 1))
    boxes[:, :3] *= 1.0 + 0.01 * np.arange(F)[:, None]            # NPT: the volume sum is all-reduced too
    pos = (rng.random((F, N, 3)) * boxes[:, None, :3]).astype(np.float32)
    other = (rng.random((F, N, 3)) * L).astype(np.float32)
    u = synthetic.Universe(synthetic.MemoryTrajectory(pos, boxes))
    v = synthetic.Universe(synthetic.MemoryTrajectory(other, np.array([L, L, L, 90, 90, 90], np.float32)))
    return u, v


def _analyses(u, v, distributed):
    from mdcraft_b200.analysis.structure import RadialDistributionFunction, StructureFactor

    kw = dict(verbose=False, distributed=distributed)
    return (RadialDistributionFunction(u.select(0, 1200), u.select(1200, N), n_bins=64, range_=(0.0, L / 2), frame_block=2, **kw),
            RadialDistributionFunction(u.atoms, n_bins=50, range_=(0.0, 4.0), exclusion=(1, 1), frame_block=3, **kw),
            StructureFactor([v.select(0, 1000), v.select(1000, N)], mode="partial", n_points=10, frame_block=2, **kw))


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mdcraft_b200 import distributed

        u, v = _universes()
        cross, excl, sq = _analyses(u, v, True)
        distributed.run_sharded(cross, start=1)                              # 6 frames: 3 + 3
        distributed.run_together_sharded([excl, sq], step=2)                 # frames 0, 2, 4, 6: 2 + 2, two trajectories
        assert cross.n_frames_total == 6 and excl.n_frames_total == 4 and sq.n_frames_total == 4
        assert cross.n_frames == 3 and excl.n_frames == 2
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), cross_counts=cross.results.counts, cross_rdf=cross.results.rdf,
                 excl_counts=excl.results.counts, excl_rdf=excl.results.rdf, ssf=sq.results.ssf,
                 frames=np.asarray(cross.frames))
    finally:
        dist.destroy_process_group()


def test_sharded_classes_all_reduce_kernel_output(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    u, v = _universes()
    cross, excl, sq = _analyses(u, v, False)
    cross.run(start=1)
    excl.run(step=2)
    sq.run(step=2)
    seen = []
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npz")
        np.testing.assert_array_equal(got["cross_counts"], cross.results.counts)      # integer sums: bit-exact
        np.testing.assert_array_equal(got["excl_counts"], excl.results.counts)
        np.testing.assert_allclose(got["cross_rdf"], cross.results.rdf, rtol=1e-13)   # volume sum in another order
        np.testing.assert_allclose(got["excl_rdf"], excl.results.rdf, rtol=1e-13)
        np.testing.assert_allclose(got["ssf"], sq.results.ssf, rtol=1e-12, atol=1e-14)
        seen += list(got["frames"])
    assert sorted(seen) == list(range(1, F))                                           # each frame visited by exactly one rank
