"""S(q) alone at the headline configuration (C3: N = 1e5 LJ-like, full lattice grid to q_max = 20/sigma, NUFFT):
kernel ms per frame.  MDC_NUFFT_PROFILE=1 prints the device time of every stage of the chain.

    python tools/sq_time.py [frames per push] [pushes]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdcraft_b200 import _capi, synthetic  # noqa: E402


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    n0, _, rho, kind, seed0 = synthetic.CONFIGS["C3"]
    L = float(synthetic.cubic_box_length(n0, rho))
    pos = np.stack([synthetic.lj_like_frame(n0, L, seed0 + i) for i in range(F)])
    import torch
    dev = torch.from_numpy(pos).cuda()
    ctx = _capi.context(0)
    plan = _capi.SqPlan(ctx, [n0], [(None, None)], n_points=161, L0=[L] * 3, q_max=20.0)
    best = 1e30
    for _ in range(reps):
        plan.push(dev, n_frames=F)
        ctx.synchronize()
        ms, nl = plan.last_push_stats()
        best = min(best, ms)
    ssf, _ = plan.read()
    print(f"S(q) C3 full grid ({plan.info()['algo']}): {best / F:.4f} ms per frame = {1e3 * F / best:.1f} frames/s "
          f"({nl} launches per push, checksum {float(ssf.sum()):.6e})")


if __name__ == "__main__":
    main()
