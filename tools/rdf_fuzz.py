"""Random g(r) configurations through the filter kernel against the all-FP64 kernel of the same library (which the parity
tests hold against the CPU oracle): particle counts around the tile and sort thresholds, one or two groups, cubic /
orthorhombic / triclinic cells, windows from zero or above, few or many bins, exclusion blocks, wrapped and unwrapped
coordinates, lattice points on bin edges, several frames with changing boxes.

    python tools/rdf_fuzz.py [cases] [seed]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdcraft_b200 import _capi  # noqa: E402
from oracle import oracle  # noqa: E402  (cell matrix only)


def counts(ctx, pa, pb, boxes, n_bins, window, excl, exact):
    os.environ["MDC_RDF_EXACT"] = "1" if exact else "0"
    plan = _capi.RdfPlan(ctx, n_bins, window, pa.shape[1], None if pb is None else pb.shape[1], excl)
    os.environ["MDC_RDF_EXACT"] = "0"
    plan.push(pa, pb, boxes, n_frames=pa.shape[0])
    c, _ = plan.read()
    plan.close()
    return c


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    ctx = _capi.context(0)
    bad = 0
    t0 = time.time()
    for case in range(cases):
        kind = rng.choice(["cubic", "ortho", "tri"])
        F = int(rng.integers(1, 4))
        na = int(rng.choice([700, 3000, 4096, 4500, 6000, 9000, 14000, 20000]))
        two = rng.random() < 0.35
        nb = int(rng.choice([900, 4096, 5000, 8000])) if two else None
        L = float(rng.uniform(8.0, 60.0))
        boxes, cells = [], []
        for f in range(F):
            s = 1.0 + 0.03 * f
            if kind == "cubic":
                b = [L * s, L * s, L * s, 90, 90, 90]
            elif kind == "ortho":
                b = [L * s, L * 1.13, L * 0.91 * s, 90, 90, 90]
            else:
                b = [L * s, L * 1.07, L * 0.95, float(rng.uniform(60, 120)), float(rng.uniform(65, 115)), float(rng.uniform(60, 120))]
            b = np.array(b, np.float32)
            try:
                m = oracle.triclinic_vectors(b).astype(np.float64)
            except Exception:
                m = None
            if m is None or not np.all(np.isfinite(m)) or m[2, 2] <= 0.2 * L:
                b = np.array([L * s, L, L, 90, 90, 90], np.float32)
                m = np.diag(b[:3].astype(np.float64))
            boxes.append(b)
            cells.append(m)
        boxes = np.stack(boxes)
        widths = []
        for m in cells:
            vol = abs(np.linalg.det(m))
            widths.append(min(vol / np.linalg.norm(np.cross(m[(k + 1) % 3], m[(k + 2) % 3])) for k in range(3)))
        w = min(widths)
        spread = rng.choice([1.0, 1.0, 3.0])       # wrapped or far outside the cell

        def make(n):
            out = []
            for m in cells:
                fr = rng.random((n, 3)) * spread - (spread - 1.0) / 2
                if rng.random() < 0.3:
                    k = n // 6
                    fr[:k] = np.round(rng.random((k, 3)) * 8) / 8          # lattice points along the cell vectors
                out.append((fr @ m).astype(np.float32))
            return np.stack(out)

        pa = make(na)
        pb = make(nb) if two else None
        r1 = float(w * rng.choice([0.5, 0.499, 0.45, 0.3, 0.12, 0.05, 0.62]))
        r0 = float(r1 * rng.choice([0.0, 0.0, 0.0, 0.2, 0.6]))
        n_bins = int(rng.choice([1, 7, 64, 200, 200, 333, 1000, 4000]))
        excl = None
        if rng.random() < 0.5:
            excl = (1, 1) if not two else (int(rng.integers(1, 4)), int(rng.integers(1, 4)))
            if not two and rng.random() < 0.3:
                e = int(rng.integers(2, 5))
                excl = (e, e)
        tag = f"case {case}: {kind} F={F} na={na} nb={nb} L={L:.1f} window=({r0:.3f}, {r1:.3f}) = {r1 / w:.3f} w, bins={n_bins} excl={excl} spread={spread}"
        try:
            fast = counts(ctx, pa, pb, boxes, n_bins, (r0, r1), excl, False)
            exact = counts(ctx, pa, pb, boxes, n_bins, (r0, r1), excl, True)
        except _capi.MdcError as err:
            print(tag, "-> refused:", str(err)[:80])
            continue
        ok = np.array_equal(fast, exact)
        bad += not ok
        print(tag, "->", "ok" if ok else f"DIFFERENT in {int((fast != exact).sum())} bins, sum {int(fast.sum())} vs {int(exact.sum())}", flush=True)
    print(f"{cases} cases, {bad} different, {time.time() - t0:.0f} s")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
