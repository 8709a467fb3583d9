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
