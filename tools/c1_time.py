import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np
from mdcraft_b200 import synthetic
from mdcraft_b200.analysis.structure import RadialDistributionFunction, StructureFactor
n, F, rho, kind, seed0 = synthetic.CONFIGS["C1"]
L = synthetic.cubic_box_length(n, rho)
frames = np.stack([synthetic.lj_like_frame(n, L, seed0 + f) for f in range(F)])
dims = np.array([L, L, L, 90, 90, 90], np.float32)
u = synthetic.Universe(synthetic.MemoryTrajectory(frames, dims))
for rep in range(3):
    t = time.perf_counter(); r = RadialDistributionFunction(u.atoms, n_bins=200, range_=(0, L / 2), verbose=False).run(); t1 = time.perf_counter() - t
    t = time.perf_counter(); s = StructureFactor([u.atoms], n_points=32, verbose=False).run(); t2 = time.perf_counter() - t
    print(f"C1 (N={n}, F={F}): g(r) {t1*1e3:.1f} ms total = {F/t1:.0f} frames/s; S(q) {t2*1e3:.1f} ms total = {F/t2:.0f} frames/s", flush=True)
