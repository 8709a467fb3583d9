"""Random S(q) configurations: every lattice kernel family (NUFFT, factored, direct; FP32 mode) against the FP64 kernel on
the SAME wavevectors (explicit list, numba_dot's operation order: the one the parity tests hold against the CPU oracle).
Particle counts from 50 to 40,000, cubic and orthorhombic cells, n_points 3 - 140 (every fine-grid family of the NUFFT),
q_max cuts, total / partial / pair modes with one to three groups, coordinates inside and far outside the cell.
Error model of DESIGN.md §3.10: |dS| <= 1e-5 S + 2 m kappa (sqrt(S_aa) + sqrt(S_bb)), kappa = 5e-7 (direct), 8e-8 (factored),
1e-8 (NUFFT), m = 5.

    python tools/sq_fuzz.py [cases] [seed]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdcraft_b200 import _capi  # noqa: E402

KAPPA = {"nufft": 1e-8, "factored": 8e-8, "direct": 5e-7}


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    ctx = _capi.context(0)
    bad = 0
    t0 = time.time()
    for case in range(cases):
        n_groups = int(rng.choice([1, 1, 2, 3]))
        sizes = [int(rng.choice([50, 400, 1500, 6000, 20000])) for _ in range(n_groups)]
        if sum(sizes) > 40000:
            sizes = [s // 2 for s in sizes]
        n = sum(sizes)
        mode = "total" if n_groups == 1 else str(rng.choice(["partial", "pair"]))
        if mode == "total":
            pairs = [(None, None)]
        elif mode == "pair":
            pairs = [(0, n_groups - 1)]
        else:
            pairs = [(i, j) for i in range(n_groups) for j in range(i, n_groups)]
        L = float(rng.uniform(6.0, 40.0))
        box = [L, L, L] if rng.random() < 0.5 else [L, L * float(rng.uniform(0.7, 1.4)), L * float(rng.uniform(0.7, 1.4))]
        n_points = int(rng.choice([3, 8, 17, 32, 33, 48, 64, 81, 100, 128, 140]))
        q_max = None if rng.random() < 0.5 else float(rng.uniform(0.3, 1.0) * 2 * np.pi * n_points / max(box))
        F = int(rng.integers(1, 4))
        spread = float(rng.choice([1.0, 1.0, 4.0]))
        pos = ((rng.random((F, n, 3)) * spread - (spread - 1) / 2) * box).astype(np.float32)
        tag = f"case {case}: sizes={sizes} mode={mode} box=({box[0]:.1f}, {box[1]:.1f}, {box[2]:.1f}) n_points={n_points} q_max={q_max} F={F} spread={spread}"
        try:
            ref = None
            line = []
            for algo in ("nufft", "factored", "direct"):
                plan = _capi.SqPlan(ctx, sizes, pairs, n_points=n_points, L0=box, q_max=q_max, algo=algo)
                if ref is None:
                    wv = plan.wavevectors()
                    rp = _capi.SqPlan(ctx, sizes, pairs, wavevectors=wv, precision="fp64")
                    rp.push(pos, n_frames=F)
                    ref = rp.read()[0] / (n * F)
                    rp.close()
                    # per-pair scale of the error model: sqrt of the self terms
                    self_rows = {}
                    if mode == "partial":
                        for k, (i, j) in enumerate(pairs):
                            if i == j:
                                self_rows[i] = np.sqrt(np.abs(ref[k]))
                assert plan.n_wavevectors == len(wv)
                used = plan.info()["algo"]
                plan.push(pos, n_frames=F)
                got = plan.read()[0] / (n * F)
                plan.close()
                worst = 0.0
                for k, (i, j) in enumerate(pairs):
                    if mode == "partial":
                        scale = self_rows[i] + self_rows[j]
                    else:
                        scale = 2.0 * np.sqrt(np.abs(ref[k])) + 1.0   # pair mode: the self terms are not on the plan
                    tol = 1e-5 * np.abs(ref[k]) + 2 * 5 * KAPPA[algo] * scale + 1e-12
                    worst = max(worst, float(np.max(np.abs(got[k] - ref[k]) / tol)))
                line.append(f"{used} {worst:.3f}")
                if not worst <= 1.0:
                    bad += 1
                    line[-1] += " <-- OUT OF TOLERANCE"
            print(tag, "->", "; ".join(line), f"(Nq = {len(wv)})", flush=True)
        except _capi.MdcError as err:
            print(tag, "-> refused:", str(err)[:100], flush=True)
    print(f"{cases} cases, {bad} out of tolerance, {time.time() - t0:.0f} s")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
