"""CPU: the N > 1 host path over a world_size-2 gloo process group (no GPU involved)."""

import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mdcraft_b200 import distributed
        from oracle import oracle
        from mdcraft_b200 import synthetic

        assert distributed.is_active() and distributed.rank_and_world() == (rank, world)
        # each rank analyses its contiguous frame block (here with the CPU oracle standing in for
        # the kernels) and the integer histograms / volume / frame counts are summed once
        u = synthetic.make_config("C1", n_frames=5, n=300)
        L = float(u.dimensions[0])
        s0, s1, st = distributed.shard_slice(len(u.trajectory), None, None, None)
        frames = [u.trajectory[i].positions.copy() for i in range(s0, s1, st)]
        part = oracle.rdf(frames, None, [u.dimensions] * len(frames), 40, (0.0, L / 2), norm=None)
        counts = distributed.all_reduce_host(part["counts"].astype(np.int64))
        totals = distributed.all_reduce_sum(np.array([len(frames), part["volume"]], dtype=np.float64))
        np.save(os.path.join(out_dir, f"counts_{rank}.npy"), counts)
        np.save(os.path.join(out_dir, f"totals_{rank}.npy"), totals)
    finally:
        dist.destroy_process_group()


def test_frame_sharded_histogram_sums_to_the_serial_result(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from mdcraft_b200 import synthetic
    from oracle import oracle

    u = synthetic.make_config("C1", n_frames=5, n=300)
    L = float(u.dimensions[0])
    frames = [u.trajectory[i].positions.copy() for i in range(5)]
    serial = oracle.rdf(frames, None, [u.dimensions] * 5, 40, (0.0, L / 2), norm=None)
    for r in range(world):
        np.testing.assert_array_equal(np.load(tmp_path / f"counts_{r}.npy"), serial["counts"])
        t = np.load(tmp_path / f"totals_{r}.npy")
        assert t[0] == 5
        np.testing.assert_allclose(t[1], serial["volume"], rtol=1e-14)


def _make_sharded_analysis(universe):
    """A host-only analysis on the in-repo AnalysisBase protocol whose _conclude all-reduces like the real classes do."""
    from mdcraft_b200 import distributed
    from mdcraft_b200.analysis.base import SerialAnalysisBase

    class Impl(SerialAnalysisBase):
        def __init__(self, universe):
            super().__init__(universe.trajectory, False)
            self.universe = universe

        def _prepare(self):
            self.seen, self.acc = [], np.zeros(4)

        def _single_frame(self):
            self.seen.append(self._ts.frame)
            self.acc += np.histogram(self.universe.atoms.positions[:, 0], bins=4, range=(0.0, 1.0))[0]

        def _conclude(self):
            totals = distributed.all_reduce_sum(np.concatenate(([float(len(self.seen))], self.acc)))
            self.results.n_frames_total = int(totals[0])
            self.results.hist = totals[1:]

    return Impl(universe)


def _universes():
    from mdcraft_b200 import synthetic

    rng = np.random.default_rng(11)
    dims = np.array([1, 1, 1, 90, 90, 90], np.float32)
    return [synthetic.Universe(synthetic.MemoryTrajectory(rng.random((9, n, 3)).astype(np.float32), dims)) for n in (13, 6)]


def _sharded_worker(rank, world, port, out_dir):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mdcraft_b200 import distributed

        ua, ub = _universes()
        a, b = _make_sharded_analysis(ua), _make_sharded_analysis(ub)
        distributed.run_together_sharded([a, b], start=1)            # frames 1..8: 4 + 4, two trajectories in lockstep
        c = _make_sharded_analysis(ua)
        distributed.run_sharded(c, step=-2)                           # a backward run that ends on frame 0
        np.savez(os.path.join(out_dir, f"sharded_{rank}.npz"), a_hist=a.results.hist, b_hist=b.results.hist, a_seen=a.seen,
                 b_seen=b.seen, c_seen=c.seen, c_hist=c.results.hist, n=[a.results.n_frames_total, c.results.n_frames_total])
    finally:
        dist.destroy_process_group()


def test_run_together_sharded_visits_every_frame_once_and_reduces(tmp_path):
    world = 2
    mp.spawn(_sharded_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ua, ub = _universes()
    a, b, c = _make_sharded_analysis(ua).run(start=1), _make_sharded_analysis(ub).run(start=1), _make_sharded_analysis(ua).run(step=-2)
    seen_a, seen_c = [], []
    for r in range(world):
        got = np.load(tmp_path / f"sharded_{r}.npz")
        np.testing.assert_array_equal(got["a_hist"], a.results.hist)
        np.testing.assert_array_equal(got["b_hist"], b.results.hist)
        np.testing.assert_array_equal(got["c_hist"], c.results.hist)
        np.testing.assert_array_equal(got["a_seen"], got["b_seen"])
        assert list(got["n"]) == [8, 5]
        seen_a += list(got["a_seen"])
        seen_c += list(got["c_seen"])
    assert seen_a == list(range(1, 9)) and seen_c == [8, 6, 4, 2, 0]
