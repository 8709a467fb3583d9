"""g(r) alone at the headline configuration (C2: N = 1e5 uniform, 200 bins to L/2, exclusion (1, 1)): kernel ms per frame.

    python tools/rdf_time.py [frames per push] [pushes]

Environment switches of the library (MDC_RDF_REL, MDC_RDF_SORT, MDC_RDF_EXT ...) are read at first use: set them before.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdcraft_b200 import _capi, synthetic  # noqa: E402


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    n0, _, rho, kind, seed0 = synthetic.CONFIGS["C2"]
    L = synthetic.cubic_box_length(n0, rho)
    pos = np.stack([synthetic.uniform_frame(n0, L, seed0 + i) for i in range(F)])
    boxes = np.tile(np.array([L, L, L, 90, 90, 90], np.float32), (F, 1))
    ctx = _capi.context(0)
    excl = int(os.environ.get("RDF_TIME_EXCL", "1"))   # exclusion blocks of this many particles (3: "molecules")
    plan = _capi.RdfPlan(ctx, 200, (0.0, L / 2), n0, None, (excl, excl))
    bps = os.environ.get("RDF_TIME_BPS")
    if bps:
        plan.set_blocks_per_sm(int(bps))
    best = 1e30
    for _ in range(reps):
        plan.push(pos, None, boxes, n_frames=F)
        c, _ = plan.read()
        ms, nl = plan.last_push_stats()
        best = min(best, ms)
    print(f"g(r) C2{'' if excl == 1 else f' with exclusion ({excl}, {excl})'}: {best / F:.4f} ms per frame = {1e3 * F / best:.1f} frames/s ({nl} launches per push, counts sum {int(c.sum())})")


if __name__ == "__main__":
    main()
