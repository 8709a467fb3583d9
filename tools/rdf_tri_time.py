"""g(r) in a triclinic cell (60/90/75, N = 1e5, 200 bins to 0.49 x the smallest width): kernel ms per frame.
MDC_RDF_TRI_FILTER=0 in the environment sends the frames to the FP64 kernel (minimum_image_triclinic for every pair)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdcraft_b200 import _capi  # noqa: E402
from oracle import oracle  # noqa: E402  (cell matrix only)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    F = 2
    s = (n / 0.8) ** (1 / 3)
    box = np.array([s * 1.05, s, s * 1.1, 60, 90, 75], np.float32)
    m = oracle.triclinic_vectors(box).astype(np.float64)
    vol = abs(np.linalg.det(m))
    w = min(vol / np.linalg.norm(np.cross(m[(k + 1) % 3], m[(k + 2) % 3])) for k in range(3))
    rng = np.random.default_rng(3)
    pos = (rng.random((F, n, 3)) @ m).astype(np.float32)
    boxes = np.tile(box, (F, 1))
    ctx = _capi.context(0)
    plan = _capi.RdfPlan(ctx, 200, (0.0, 0.49 * w), n, None, (1, 1))
    best = 1e30
    for _ in range(3):
        plan.push(pos, None, boxes, n_frames=F)
        c, _ = plan.read()
        ms, nl = plan.last_push_stats()
        best = min(best, ms)
    print(f"triclinic g(r), N = {n}: {best / F:.3f} ms per frame = {1e3 * F / best:.1f} frames/s ({nl} launches, counts sum {int(c.sum())})")


if __name__ == "__main__":
    main()
