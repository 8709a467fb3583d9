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

t0 = time.perf_counter()
r = LAMMPSDumpTrajectoryReader(path, parallel=True, frame_block=F)
t_open = time.perf_counter() - t0
ctx = _capi.context(0)
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
best = min(res.values())
print(json.dumps({"ingest_frames_per_s": 1 / best, "text_gb_per_s": size / F / best / 1e9, "bytes_per_frame": size / F,
                  "cpu_frames_per_s": 1 / t_cpu, "open_ms": t_open * 1e3, "n_atoms": N, "frames": F}))
