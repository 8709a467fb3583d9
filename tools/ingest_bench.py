"""GPU: LAMMPS dump ingest at N = 1e5 atoms: index, GPU parse, and the CPU restatement of the reference's reader beside it.
usage: python tools/ingest_bench.py [N] [frames]"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdcraft_b200 import _capi, synthetic  # noqa: E402
from mdcraft_b200.analysis.reader import LAMMPSDumpTrajectoryReader  # noqa: E402
from oracle import oracle  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
F = int(sys.argv[2]) if len(sys.argv) > 2 else 8
L = float(synthetic.cubic_box_length(N, 0.8))
tmp = tempfile.mkdtemp()
path = os.path.join(tmp, "traj.lammpsdump")
t0 = time.perf_counter()
rng = np.random.default_rng(0)
with open(path, "w") as fh:
    for f in range(F):
        pos = synthetic.lj_like_frame(N, L, 3000 + f)
        order = rng.permutation(N)
        fh.write(f"ITEM: TIMESTEP\n{f * 1000}\nITEM: NUMBER OF ATOMS\n{N}\nITEM: BOX BOUNDS pp pp pp\n0 {L!r}\n0 {L!r}\n0 {L!r}\n")
        fh.write("ITEM: ATOMS id type x y z\n")
        fh.write("\n".join("%d 1 %g %g %g" % (i + 1, *pos[i]) for i in order) + "\n")
size = os.path.getsize(path)
print(f"wrote {size / 1e6:.1f} MB, {F} frames x {N} atoms in {time.perf_counter() - t0:.1f} s", flush=True)

ctx = _capi.context(0)                      # CUDA context creation is not part of opening a file
torch.cuda.synchronize()
for threads in (1, 0):
    t0 = time.perf_counter()
    d = _capi.DumpFile(path, None, threads)
    t_index = time.perf_counter() - t0
    print(f"mdc_dump_open ({'1 thread' if threads else 'all host threads'}): {t_index * 1e3:.1f} ms for {size / 1e6:.0f} MB")
    d.close()
t0 = time.perf_counter()
r = LAMMPSDumpTrajectoryReader(path, parallel=True, frame_block=F)
t_open = time.perf_counter() - t0
dev = torch.empty((F, N, 3), dtype=torch.float32, device="cuda")
r._file.read(ctx, 0, F, out=dev)   # warm: page cache, device scratch
torch.cuda.synchronize()
res = {}
for fb in (1, 2, 4, F):
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        for lo in range(0, F, fb):
            r._file.read(ctx, lo, min(fb, F - lo), out=dev[lo:lo + fb])
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / (reps * F)
    res[fb] = dt
    print(f"GPU parse, {fb} frame(s) per pass: {dt * 1e3:.3f} ms/frame = {1 / dt:.0f} frames/s = {size / F / dt / 1e9:.2f} GB/s of text")
t0 = time.perf_counter()
host = r.frames_block(0, F)
t_host = (time.perf_counter() - t0) / F
print(f"GPU parse into host memory: {t_host * 1e3:.3f} ms/frame")
# CPU restatement of the reference's per-atom Python loop, one frame (the reference reader is slower still: it
# also builds a tuple and calls np.fromiter per atom)
one = os.path.join(tmp, "one.lammpsdump")
with open(path, "rb") as fh:
    blob = fh.read(int(r._file.offsets[1]))
with open(one, "wb") as fh:
    fh.write(blob)
t0 = time.perf_counter()
want = oracle.read_lammps_dump(one)[0]
t_cpu = time.perf_counter() - t0
np.testing.assert_array_equal(host[0], want[0])
print(f"CPU restatement of the reference reader: {t_cpu:.2f} s/frame; index + headers of {F} frames: {t_open * 1e3:.1f} ms")
# end to end from the FILE through the drop-in classes: dump text -> g(r) + S(q) (full q_max = 20 / sigma grid)
from mdcraft_b200.analysis import structure  # noqa: E402
from mdcraft_b200.analysis.base import run_together  # noqa: E402

u = synthetic.Universe(r)


def analyse(reps):
    t0 = time.perf_counter()
    g = structure.RadialDistributionFunction(u.atoms, n_bins=200, range_=(0.0, L / 2), exclusion=(1, 1), verbose=False,
                                             frame_block=8)
    q = structure.StructureFactor(u.atoms, n_points=161, q_max=20.0, verbose=False, frame_block=8)
    t1 = time.perf_counter()
    frames = np.tile(np.arange(F), reps)
    run_together([g, q], frames=frames)      # one pass over the file, g(r) and S(q) side by side on the device
    return time.perf_counter() - t1, t1 - t0, g, q


def marginal():
    """Seconds per frame from the difference between 7 passes and 1 pass over the file (plan creation, grid
    bookkeeping and _conclude's |q| shell averaging are per-analysis costs, not per-frame ones)."""
    analyse(1)                                  # warm: plans, NUFFT buffers
    t_1 = min(analyse(1)[0] for _ in range(3))
    runs = [analyse(7) for _ in range(3)]
    t_7 = min(r_[0] for r_ in runs)
    return (t_7 - t_1) / (6 * F), t_1, runs[0][1], runs[0][2], runs[0][3]


m_dev, t_dev, t_ctor, g1, q1 = marginal()
saved = structure._DeviceBlockRun._device_block_ranges
structure._DeviceBlockRun._device_block_ranges = lambda self: None     # force the per-frame protocol (host staging)
m_host, t_host, _, g2, q2 = marginal()
structure._DeviceBlockRun._device_block_ranges = saved
np.testing.assert_array_equal(g1.results.counts, g2.results.counts)
np.testing.assert_allclose(q1.results.ssf, q2.results.ssf, rtol=1e-9, atol=1e-12)
print(f"file -> g(r) + S(q) through the classes, marginal cost per frame: {m_dev * 1e3:.2f} ms = {1 / m_dev:.1f} frames/s with "
      f"device-resident blocks, {m_host * 1e3:.2f} ms = {1 / m_host:.1f} frames/s through the per-frame protocol; "
      f"per-analysis fixed cost: constructors {t_ctor * 1e3:.0f} ms, run() of {F} frames {t_dev * 1e3:.0f} ms")
best = min(res.values())
print(json.dumps({"ingest_frames_per_s": 1 / best, "text_gb_per_s": size / F / best / 1e9, "bytes_per_frame": size / F,
                  "cpu_frames_per_s": 1 / t_cpu, "open_ms": t_open * 1e3, "n_atoms": N, "frames": F,
                  "file_to_results_fps_device_blocks": 1 / m_dev, "file_to_results_fps_per_frame": 1 / m_host}))
