"""GPU: frames/s of run_together([RadialDistributionFunction, StructureFactor]) on host trajectories for several frame_block pairs."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdcraft_b200 import synthetic
from mdcraft_b200.analysis.base import run_together
from mdcraft_b200.analysis.structure import RadialDistributionFunction, StructureFactor

N, F = 100_000, 192
Lg, Ls = float(synthetic.cubic_box_length(N, 2.5)), float(synthetic.cubic_box_length(N, 0.8))
ug = synthetic.Universe(synthetic.MemoryTrajectory(np.stack([synthetic.uniform_frame(N, Lg, 2000 + i) for i in range(F)]), np.array([Lg] * 3 + [90] * 3, np.float32)))
us = synthetic.Universe(synthetic.MemoryTrajectory(np.stack([synthetic.lj_like_frame(N, Ls, 3000 + i) for i in range(F)]), np.array([Ls] * 3 + [90] * 3, np.float32)))
import torch
for bg, bs in ((16, 8), (16, 16), (32, 16), (32, 32), (48, 48), (8, 8), (24, 12)):
    r = RadialDistributionFunction(ug.atoms, n_bins=200, range_=(0.0, Lg / 2), exclusion=(1, 1), verbose=False, frame_block=bg)
    s = StructureFactor(us.atoms, n_points=161, q_max=20.0, verbose=False, frame_block=bs)
    run_together([r, s], stop=48)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run_together([r, s])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"frame_block g(r) {bg:3d} S(q) {bs:3d}: {F / dt:7.1f} frames/s", flush=True)
    del r, s
