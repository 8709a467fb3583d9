#!/usr/bin/env python
"""
Secondary throughput figures (GPU box) for the BASELINE.json configurations that bench.py does not time, and for the
widened rows: writes gpurun_out/extra_bench.json.  Device-resident inputs, CUDA-event timing on the library's compute
stream, 2 warm-up pushes.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mdcraft_b200 import _capi, synthetic  # noqa: E402

ctx = _capi.Context(0)
out = {}


def frames(name, n=None, count=2):
    n0, _, rho, kind, seed0 = synthetic.CONFIGS[name]
    n = n or n0
    L = synthetic.cubic_box_length(n, rho)
    fn = synthetic.lj_like_frame if kind == "lj" else synthetic.uniform_frame
    pos = np.stack([fn(n, L, seed0 + i) for i in range(count)])
    return torch.from_numpy(pos).cuda(), float(L), n


def timed(push, n_frames, reps=3):
    for _ in range(2):
        push()
    ctx.synchronize()
    ctx.mark(0)
    for _ in range(reps):
        push()
    ctx.mark(1)
    return n_frames * reps / (ctx.elapsed_ms() * 1e-3)


# C3(i): default grid, every lattice algorithm and FP64 mode ("auto" = the plan's cost model; the choice is recorded)
pos, L, n = frames("C3", count=4)
for key, kw in (("c3_default_grid_auto", {}), ("c3_default_grid_nufft", {"algo": "nufft"}),
                ("c3_default_grid_factored", {"algo": "factored"}), ("c3_default_grid_direct", {"algo": "direct"}),
                ("c3_default_grid_fp64_mode", {"precision": "fp64"})):
    plan = _capi.SqPlan(ctx, [n], [(None, None)], n_points=32, L0=[L] * 3, **kw)
    out[key + "_fps"] = timed(lambda: plan.push(pos, n_frames=4), 4)
    out[key + "_algo"] = plan.info()["algo"]
    plan.close()
# C3(ii): the q_max = 20 / sigma grid in FP64 mode (auto: NUFFT at 1e-13)
plan = _capi.SqPlan(ctx, [n], [(None, None)], n_points=161, L0=[L] * 3, q_max=20.0, precision="fp64")
out["c3_full_grid_fp64_mode_fps"] = timed(lambda: plan.push(pos, n_frames=4), 4)
out["c3_full_grid_fp64_mode_algo"] = plan.info()["algo"]
plan.close()

# C4: binary electrolyte, partial S(q) (default grid) and the three g(r)
pos, L, n = frames("C4")
half = n // 2
box = np.array([L, L, L, 90, 90, 90], np.float32)
plan = _capi.SqPlan(ctx, [half, half], [(0, 0), (0, 1), (1, 1)], n_points=32, L0=[L] * 3)
out["c4_partial_sq_default_grid_fps"] = timed(lambda: plan.push(pos, n_frames=2), 2)
out["c4_partial_sq_default_grid_algo"] = plan.info()["algo"]
plan.close()
plan = _capi.SqPlan(ctx, [half, half], [(0, 0), (0, 1), (1, 1)], n_points=201, L0=[L] * 3, q_max=20.0)
out["c4_partial_sq_full_grid_fps"] = timed(lambda: plan.push(pos, n_frames=2), 2)
out["c4_partial_sq_full_grid_nq"] = plan.n_wavevectors
out["c4_partial_sq_full_grid_algo"] = plan.info()["algo"]
plan.close()
pa, pb = pos[:, :half].contiguous(), pos[:, half:].contiguous()
pp = _capi.RdfPlan(ctx, 200, (0.0, L / 2), half, None, (1, 1))
mm = _capi.RdfPlan(ctx, 200, (0.0, L / 2), half, None, (1, 1))
pm = _capi.RdfPlan(ctx, 200, (0.0, L / 2), half, half)


def c4_rdf():
    pp.push(pa, None, box, n_frames=2)
    mm.push(pb, None, box, n_frames=2)
    pm.push(pa, pb, box, n_frames=2)


out["c4_three_rdf_fps"] = timed(c4_rdf, 2)
for p in (pp, mm, pm):
    p.close()

# C5: one million particles
pos, L, n = frames("C5", count=1)
box = np.array([L, L, L, 90, 90, 90], np.float32)
plan = _capi.SqPlan(ctx, [n], [(None, None)], n_points=32, L0=[L] * 3)
out["c5_sq_default_grid_fps"] = timed(lambda: plan.push(pos, n_frames=1), 1)
out["c5_sq_default_grid_algo"] = plan.info()["algo"]
plan.close()
rdf = _capi.RdfPlan(ctx, 200, (0.0, L / 2), n, None, (1, 1))
out["c5_rdf_fps"] = timed(lambda: rdf.push(pos, None, box, n_frames=1), 1, reps=2)
out["c5_rdf_pairs_per_s"] = out["c5_rdf_fps"] * 0.5 * n * (n - 1.0)
rdf.close()

# widened rows: ISF (16 lags, with the self part) and single-chain S(q)
pos, L, n = frames("C3", n=20_000, count=8)
isf = _capi.IsfPlan(ctx, [n], [(None, None)], 16, incoherent=True, n_points=32, L0=[L] * 3)
out["isf_n20000_lags16_incoherent_fps"] = timed(lambda: isf.push(pos, n_frames=8), 8, reps=2)
out["isf_n20000_lags16_incoherent_algo"] = isf.info()["algo"]
isf.close()
M, N = 1000, 100
pos = torch.from_numpy(np.random.default_rng(0).random((2, M * N, 3)).astype(np.float32) * 40).cuda()
sc = _capi.ScsfPlan(ctx, M, N, n_points=32, L0=[40.0] * 3)
out["scsf_1000x100_fps"] = timed(lambda: sc.push(pos, n_frames=2), 2)
out["scsf_1000x100_algo"] = sc.info()["algo"]
sc.close()

os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/extra_bench.json", "w") as fh:
    json.dump(out, fh, indent=1)
print(json.dumps(out, indent=1))
