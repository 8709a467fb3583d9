"""GPU: NUFFT S(q) against the factored FP32 kernel and the FP64 kernel at full size, with timings.
usage: python tools/nufft_check.py [N] [n_points] [q_max]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdcraft_b200 import _capi, synthetic  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
n_points = int(sys.argv[2]) if len(sys.argv) > 2 else 161
q_max = float(sys.argv[3]) if len(sys.argv) > 3 else 20.0
if q_max <= 0:
    q_max = None
F = 2
ctx = _capi.Context(0)
L = synthetic.cubic_box_length(N, 0.8)
pos = np.stack([synthetic.lj_like_frame(N, L, 3000 + f) for f in range(F)]).astype(np.float32)
dev = torch.from_numpy(pos).cuda()


def run(algo, precision="fp32", reps=3, groups=None, pairs=((None, None),)):
    plan = _capi.SqPlan(ctx, groups or [N], pairs, n_points=n_points, L0=[L] * 3, q_max=q_max, algo=algo,
                        precision=precision)
    plan.push(dev, n_frames=F)
    out, _ = plan.read()
    plan.reset()
    ctx.synchronize()
    ctx.mark(0)
    for _ in range(reps):
        plan.push(dev, n_frames=F)
    ctx.mark(1)
    ms = ctx.elapsed_ms() / (reps * F)
    kms, nl = plan.last_push_stats()
    plan.close()
    return out / F, ms, nl


t0 = time.time()
a, ms_n, nl = run("nufft")
print(f"N={N} n_points={n_points} q_max={q_max} Nq={a.shape[1]}")
print(f"nufft    : {ms_n:8.3f} ms/frame  ({1e3 / ms_n:8.1f} frames/s), launches/push {nl}")
b, ms_f, nl = run("factored")
print(f"factored : {ms_f:8.3f} ms/frame  ({1e3 / ms_f:8.1f} frames/s), launches/push {nl}")
S = b / N
d = np.abs(a - b) / N
print("nufft vs factored: max |dS| = %.3e, max rel (S>0.01) = %.3e" % (d.max(), (d / np.abs(S))[np.abs(S) > 0.01].max()))
a64, ms_n64, _ = run("nufft", "fp64", reps=2)
print(f"nufft fp64: {ms_n64:8.3f} ms/frame")
d = np.abs(a - a64) / N
print("nufft fp32-mode vs fp64-mode: max |dS| = %.3e, max rel (S>0.01) = %.3e" % (d.max(), (d / np.abs(S))[np.abs(S) > 0.01].max()))
# FP64 reference arithmetic on a slice of wavevectors (explicit kernel)
plan = _capi.SqPlan(ctx, [N], [(None, None)], n_points=n_points, L0=[L] * 3, q_max=q_max, algo="nufft")
wv = plan.wavevectors()
plan.close()
sel = np.random.default_rng(0).choice(wv.shape[0], size=min(4096, wv.shape[0]), replace=False)
ref = _capi.SqPlan(ctx, [N], [(None, None)], wavevectors=wv[sel], precision="fp64")
ref.push(dev, n_frames=F)
r, _ = ref.read()
ctx.synchronize()
ctx.mark(0)
ref.push(dev, n_frames=F)
ctx.mark(1)
print(f"FP64 reference kernel: {sel.size * N * F / (ctx.elapsed_ms() * 1e-3) / 1e9:.2f} G terms/s")
ref.close()
r = r[0] / F / N
for name, arr in (("nufft", a), ("nufft64", a64), ("factored", b)):
    got = arr[0, sel] / N
    m = r > 0.01
    print(f"{name:9s} vs FP64 kernel on {sel.size} wavevectors: max |dS| = {np.abs(got - r).max():.3e}, "
          f"max rel (S>0.01) = {np.abs(got[m] / r[m] - 1).max():.3e}, rms rel = {np.sqrt(np.mean((got[m] / r[m] - 1) ** 2)):.3e}")
# partial mode, two groups
h = N // 2
pa, ms_pa, _ = run("nufft", groups=[h, N - h], pairs=((0, 0), (0, 1), (1, 1)))
print(f"nufft partial (3 pairs): {ms_pa:8.3f} ms/frame; sum rule max |sum - total| / N = {np.abs(pa.sum(axis=0) - a[0]).max() / N:.3e}")
print("wall", time.time() - t0)
