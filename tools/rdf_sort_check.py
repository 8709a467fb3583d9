"""g(r): sorted + culled path against the plain all-pairs path (kernel ms per push).

    python tools/rdf_sort_check.py [--half] [N ...]

--half: cutoffs L/2, 0.4 L, 0.3 L, exclusion (1, 1) as in the headline configuration.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdcraft_b200 import _capi  # noqa: E402


def main():
    ctx = _capi.context(0)
    args = sys.argv[1:]
    half = "--half" in args
    sizes = [int(a) for a in args if not a.startswith("--")] or [32768, 100_000, 250_000]
    rng = np.random.default_rng(5)
    for n in sizes:
        L = (n / 0.1) ** (1 / 3)                      # number density 0.1 / A^3
        F = 2
        boxes = np.tile(np.array([L, L, L, 90, 90, 90], np.float32), (F, 1))
        pos = (rng.random((F, n, 3)) * L).astype(np.float32)
        for rc in ((L / 2, 0.4 * L, 0.3 * L) if half else (3.0, 6.0, 12.0)):
            row = []
            for sort in ("0", "1"):
                os.environ["MDC_RDF_SORT"] = sort
                plan = _capi.RdfPlan(ctx, 200, (0.0, rc), n, None, (1, 1) if half else None)
                best = 1e30
                for _ in range(3):
                    plan.push(pos, None, boxes, n_frames=F)
                    c, _ = plan.read()
                    ms, nl = plan.last_push_stats()
                    best = min(best, ms)
                plan.close()
                row.append((best, int(c.sum()), nl, c))
            assert np.array_equal(row[0][3], row[1][3])
            print(f"N={n} L={L:.1f} rc={rc:.2f}: plain {row[0][0]:.3f} ms ({row[0][2]} launches)  sorted {row[1][0]:.3f} ms "
                  f"({row[1][2]} launches)  x{row[0][0] / row[1][0]:.2f}", flush=True)


if __name__ == "__main__":
    main()
