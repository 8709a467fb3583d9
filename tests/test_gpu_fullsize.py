"""
GPU: the BASELINE.json configurations at FULL size (C2-C5), checked through size-independent
properties and through oracle slices that the CPU finishes in seconds:

* g(r): the filter kernel against the all-FP64 kernel (bit-exact), a 512-particle slab of i against all j against
  the CPU oracle (bit-exact), evenness / self-pair / exclusion identities, the ideal-gas in-range fraction;
* S(q): FP32 (factored and direct) against the FP64-mode kernel on the whole grid under the error model, the
  FP64-mode kernel against the CPU oracle on a random subset of wavevectors, the partial sum rule.
"""

import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _frame(name, n=None, frame=0):
    from mdcraft_b200 import synthetic

    n0, _, rho, kind, seed0 = synthetic.CONFIGS[name]
    n = n or n0
    L = synthetic.cubic_box_length(n, rho)
    fn = synthetic.lj_like_frame if kind == "lj" else synthetic.uniform_frame
    return fn(n, L, seed0 + frame), np.array([L, L, L, 90, 90, 90], np.float32)


def _rdf_counts(ctx, pos_a, pos_b, box, n_bins, rng, exclusion=None, exact=False):
    from mdcraft_b200 import _capi

    old = os.environ.get("MDC_RDF_EXACT")
    os.environ["MDC_RDF_EXACT"] = "1" if exact else "0"
    try:
        plan = _capi.RdfPlan(ctx, n_bins, rng, len(pos_a), None if pos_b is None else len(pos_b), exclusion)
    finally:
        if old is None:
            del os.environ["MDC_RDF_EXACT"]
        else:
            os.environ["MDC_RDF_EXACT"] = old
    plan.push(pos_a[None], None if pos_b is None else pos_b[None], box, n_frames=1)
    counts, nf = plan.read()
    assert nf == 1
    plan.close()
    return counts


def test_c2_rdf_full_size(ctx):
    """C2: g(r) on 100,000 uniform particles, 200 bins to L/2."""
    from oracle import oracle

    pos, box = _frame("C2")
    n, L = len(pos), float(box[0])
    rng = (0.0, L / 2)
    fast = _rdf_counts(ctx, pos, None, box, 200, rng)
    # identities: N self pairs in bin 0, every other contribution comes in (i, j) / (j, i) twins
    assert fast[0] >= n and ((fast - np.eye(1, 200, 0, dtype=np.int64)[0] * n) % 2 == 0).all()
    excl = _rdf_counts(ctx, pos, None, box, 200, rng, exclusion=(1, 1))
    np.testing.assert_array_equal(excl, fast - np.eye(1, 200, 0, dtype=np.int64)[0] * n)
    # ideal gas: the fraction of ordered pairs within L/2 is pi/6, to ~1/sqrt(pairs)
    assert abs(excl.sum() / (n * (n - 1.0)) - np.pi / 6) < 2e-4
    # the path the headline runs: filter kernel, Hilbert-ordered frames, relative rows, one histogram copy per bank
    from mdcraft_b200 import _capi

    plan = _capi.RdfPlan(ctx, 200, rng, n, None, (1, 1))
    plan.push(pos[None], None, box, n_frames=1)
    info = plan.info()
    plan.close()
    assert info["filter"] and info["sorted"] and info["relative_rows"] and info["copies_interleaved"] and not info["triclinic"], info
    # a slab of i against all j, bit-exact against the CPU restatement of the reference
    slab = _rdf_counts(ctx, pos[:512], pos, box, 200, rng)
    np.testing.assert_array_equal(slab, oracle.radial_histogram(pos[:512], pos, 200, rng, box))
    # the FP32 filter changes nothing: all-FP64 kernel, same counts (20 % of the particles keeps this quick)
    sub = pos[:20_000]
    np.testing.assert_array_equal(_rdf_counts(ctx, sub, None, box, 200, rng),
                                  _rdf_counts(ctx, sub, None, box, 200, rng, exact=True))


@pytest.mark.parametrize("n_bins,lo,hi", [(200, 0.0, 0.5), (1000, 0.0, 0.45), (64, 0.11, 0.3), (2000, 0.0, 0.5)])
def test_rdf_filter_margin_with_unwrapped_coordinates(ctx, n_bins, lo, hi):
    """The FP32 filter's margin is derived from the RANGE of the coordinates (float32 subtraction error of the
    reference, rdf_frame_params): coordinates well outside the box (unwrapped, negative) widen it.  The filtered kernel
    must agree with the all-FP64 kernel on every count, and a slab with the CPU oracle, for narrow and wide bins."""
    from oracle import oracle

    rng = np.random.default_rng(n_bins)
    n, L = 40_000, 30.0
    box = np.array([L, L * 1.1, L * 0.93, 90, 90, 90], np.float32)
    pos = ((rng.random((n, 3)) * 2.6 - 0.7) * box[:3]).astype(np.float32)
    window = (lo * L * 0.93, hi * L * 0.93)
    fast = _rdf_counts(ctx, pos, None, box, n_bins, window, exclusion=(1, 1))
    exact = _rdf_counts(ctx, pos, None, box, n_bins, window, exclusion=(1, 1), exact=True)
    np.testing.assert_array_equal(fast, exact)
    slab = _rdf_counts(ctx, pos[:256], pos, box, n_bins, window)
    np.testing.assert_array_equal(slab, oracle.radial_histogram(pos[:256], pos, n_bins, window, box))
    # wrapped positions of the same configuration: a smaller range, a tighter margin, other counts near the edges
    wrapped = np.mod(pos.astype(np.float64), box[:3].astype(np.float64)).astype(np.float32)
    wrapped[wrapped >= box[:3]] = 0.0
    np.testing.assert_array_equal(_rdf_counts(ctx, wrapped, None, box, n_bins, window, exclusion=(1, 1)),
                                  _rdf_counts(ctx, wrapped, None, box, n_bins, window, exclusion=(1, 1), exact=True))
    slab = _rdf_counts(ctx, wrapped[:256], wrapped, box, n_bins, window)
    np.testing.assert_array_equal(slab, oracle.radial_histogram(wrapped[:256], wrapped, n_bins, window, box))


@pytest.mark.parametrize("n,n_bins,frac,tri", [(6000, 12_000, 0.45, False), (5000, 12_000, 0.006, False), (4500, 12_000, 0.45, True)])
def test_rdf_overflow_of_the_undecided_row_queue(ctx, n, n_bins, frac, tri):
    """Bins so narrow that the filter's margin is a large part of a bin (or exceeds half of it: the frame is then not
    usable and EVERY pair is undecided): far more undecided rows per tile pair than the shared-memory queue holds, so
    they take the per-block overflow table in global memory (park_row) and are resolved from there.  Counts must still
    be the exact kernel's and the oracle's; orthorhombic and triclinic cells."""
    from oracle import oracle

    rng = np.random.default_rng(n)
    if tri:
        box = np.array([22.0, 20.0, 24.0, 70, 100, 80], np.float32)
        m = oracle.triclinic_vectors(box).astype(np.float64)
        pos = ((rng.random((n, 3)) * 3 - 1) @ m).astype(np.float32)
        vol = abs(np.linalg.det(m))
        w = min(vol / np.linalg.norm(np.cross(m[(k + 1) % 3], m[(k + 2) % 3])) for k in range(3))
        window = (0.0, frac * w)
    else:
        L = 25.0
        box = np.array([L, L * 1.05, L * 0.97, 90, 90, 90], np.float32)
        pos = (rng.random((n, 3)) * box[:3]).astype(np.float32)
        window = (0.0, frac * L)
    fast = _rdf_counts(ctx, pos, None, box, n_bins, window, exclusion=(1, 1))
    np.testing.assert_array_equal(fast, _rdf_counts(ctx, pos, None, box, n_bins, window, exclusion=(1, 1), exact=True))
    np.testing.assert_array_equal(fast, oracle.radial_histogram(pos, pos, n_bins, window, box, exclusion=(1, 1)))


def test_rdf_random_configurations_filter_vs_exact_kernel(ctx, monkeypatch):
    """tools/rdf_fuzz.py with a fixed seed: 40 random configurations (sizes around the tile and sort thresholds, one or
    two groups, cubic / orthorhombic / triclinic cells, windows from zero or above and below or beyond half the cell
    width, 1 - 4000 bins, exclusion blocks, unwrapped coordinates, lattice points on bin edges, changing boxes) through
    the filter kernel and through the all-FP64 kernel: every histogram equal.  (3,500 more: profiles/rdf_fuzz_r02.txt.)"""
    import importlib.util

    spec = importlib.util.spec_from_file_location("rdf_fuzz", os.path.join(os.path.dirname(__file__), "..", "tools", "rdf_fuzz.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setattr(sys, "argv", ["rdf_fuzz.py", "40", "11"])
    assert mod.main() == 0


def test_sq_random_configurations_every_kernel_family_vs_fp64(ctx, monkeypatch):
    """tools/sq_fuzz.py with a fixed seed: 20 random S(q) configurations (50 - 40,000 particles, cubic and orthorhombic
    cells, n_points 3 - 140, q_max cuts, total / partial / pair modes, coordinates far outside the cell): the NUFFT, the
    factored and the direct kernel against the FP64 kernel on the same wavevectors, within the error model of DESIGN.md
    3.10.  (540 more: profiles/sq_fuzz_r02.txt.)"""
    import importlib.util

    spec = importlib.util.spec_from_file_location("sq_fuzz", os.path.join(os.path.dirname(__file__), "..", "tools", "sq_fuzz.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setattr(sys, "argv", ["sq_fuzz.py", "20", "7"])
    assert mod.main() == 0


def test_c4_binary_electrolyte_partials(ctx):
    """C4: 100k cations + 100k anions: ++, +-, -- g(r) and the partial S(q) sum rule."""
    from mdcraft_b200 import _capi

    pos, box = _frame("C4")
    n, L = len(pos), float(box[0])
    half = n // 2
    rng = (0.0, L / 2)
    pp = _rdf_counts(ctx, pos[:half], None, box, 200, rng, exclusion=(1, 1))
    mm = _rdf_counts(ctx, pos[half:], None, box, 200, rng, exclusion=(1, 1))
    pm = _rdf_counts(ctx, pos[:half], pos[half:], box, 200, rng)
    mp = _rdf_counts(ctx, pos[half:], pos[:half], box, 200, rng)
    allp = _rdf_counts(ctx, pos, None, box, 200, rng, exclusion=(1, 1))
    np.testing.assert_array_equal(pm, mp)                       # float32 subtraction is antisymmetric
    np.testing.assert_array_equal(pp + mm + pm + mp, allp)      # the four blocks tile the full pair matrix

    plan = _capi.SqPlan(ctx, [half, half], [(0, 0), (0, 1), (1, 1)], n_points=32, L0=[L] * 3)
    tot = _capi.SqPlan(ctx, [n], [(None, None)], n_points=32, L0=[L] * 3)
    plan.push(pos[None])
    tot.push(pos[None])
    part, _ = plan.read()
    s, _ = tot.read()
    part /= n
    s /= n
    # sum rule (structure.py:1554-1560); both sides carry kappa ~ 1e-7 per term
    tol = 1e-5 * s[0] + 3e-6 * (np.sqrt(part[0]) + np.sqrt(part[2])) + 1e-9
    assert (np.abs(part.sum(axis=0) - s[0]) <= tol).all()
    plan.close()
    tot.close()


@pytest.mark.parametrize("algo,kappa", [("factored", 1e-7), ("direct", 5e-7), ("nufft", 1e-8)])
def test_c3_sq_full_size(ctx, algo, kappa):
    """C3: S(q) on 100,000 particles, default 32^3 grid: FP32 kernels against the FP64-mode kernel."""
    from mdcraft_b200 import _capi
    from oracle import oracle

    pos, box = _frame("C3")
    n, L = len(pos), float(box[0])
    ref = _capi.SqPlan(ctx, [n], [(None, None)], n_points=32, L0=[L] * 3, precision="fp64")
    ref.push(pos[None])
    s64, _ = ref.read()
    s64 = s64[0] / n
    if algo == "factored":   # pin the FP64-mode kernel itself on a random slice of the grid
        wv = ref.wavevectors()
        idx = np.random.default_rng(0).choice(len(wv), 256, replace=False)
        rho = oracle.delta_fourier_transform_sum(wv[idx], pos.astype(np.float64))
        np.testing.assert_allclose(s64[idx], np.abs(rho) ** 2 / n, rtol=1e-10, atol=1e-12)
    ref.close()
    plan = _capi.SqPlan(ctx, [n], [(None, None)], n_points=32, L0=[L] * 3, algo=algo)
    plan.push(pos[None])
    s32, _ = plan.read()
    s32 = s32[0] / n
    plan.close()
    tol = 1e-5 * s64 + 2 * 5 * kappa * np.sqrt(s64) + 1e-12
    assert (np.abs(s32 - s64) <= tol).all(), float(np.max(np.abs(s32 - s64) / tol))
    big = s64 > (0.25 if algo == "direct" else 0.01)
    assert np.max(np.abs(s32[big] / s64[big] - 1)) < 1e-5


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_c3_full_grid_nufft_vs_oracle(ctx, precision):
    """The HEADLINE S(q) configuration (bench.py): N = 100,000, n_points = 161, q_max = 20 (Nq = 2,140,871; NUFFT with
    nf = 256, w = 12 / 16) against the CPU restatement of delta_fourier_transform_sum (structure.py:1240-1280) on 512
    random wavevectors of the full grid, two frames: 1e-5 relative in FP32 mode, 1e-10 in FP64 mode."""
    from mdcraft_b200 import _capi
    from oracle import oracle

    frames = np.stack([_frame("C3", frame=f)[0] for f in range(2)])
    n, L = frames.shape[1], float(_frame("C3")[1][0])
    plan = _capi.SqPlan(ctx, [n], [(None, None)], n_points=161, L0=[L] * 3, q_max=20.0, algo="nufft", precision=precision)
    info = plan.info()
    assert plan.n_wavevectors == 2_140_871 and info["algo"] == "nufft" and info["nf"] == 256
    assert info["w"] == (12 if precision == "fp32" else 16)
    wv = plan.wavevectors()
    plan.push(frames)
    got, nfr = plan.read()
    plan.close()
    assert nfr == 2
    # the grid itself is the reference's (structure.py:1822-1887), order included
    want_wv, _ = oracle.wavevector_grid(np.full(3, L), 161, 20.0)
    np.testing.assert_array_equal(wv, want_wv)
    idx = np.sort(np.random.default_rng(161).choice(len(wv), 512, replace=False))
    want = sum(np.abs(oracle.delta_fourier_transform_sum(wv[idx], f.astype(np.float64))) ** 2 for f in frames) / (2 * n)
    s = got[0, idx] / (2 * n)
    if precision == "fp64":
        np.testing.assert_allclose(s, want, rtol=1e-10, atol=1e-12)
    else:
        # error model of DESIGN.md 3.10 (kappa = 1e-8: the spreading kernel's aliasing), and the literal bound
        tol = 1e-5 * want + 2 * 5 * 1e-8 * np.sqrt(want) + 1e-12
        assert (np.abs(s - want) <= tol).all(), float(np.max(np.abs(s - want) / tol))
        big = want > 0.01
        assert big.sum() > 256 and np.max(np.abs(s[big] / want[big] - 1)) < 1e-5


def test_c4_partials_vs_oracle(ctx):
    """C4 (binary electrolyte, 100k cations + 100k anions) against the CPU oracle: a 256-cation slab against all anions
    (cross g(r), bit-exact), a 256-cation slab of the ++ self g(r) with exclusion, and the three partial S(q) rows on
    128 wavevectors of the FULL q_max = 20 grid (n_points = 201, Nq = 4,269,143)."""
    from mdcraft_b200 import _capi
    from oracle import oracle

    pos, box = _frame("C4")
    n, L = len(pos), float(box[0])
    half = n // 2
    cat, an = pos[:half], pos[half:]
    rng = (0.0, L / 2)
    # (structure.py:34-131) slabs of i against all j; exclusion indices are local to each group (block (1, 1))
    np.testing.assert_array_equal(_rdf_counts(ctx, cat[1000:1256], an, box, 200, rng),
                                  oracle.radial_histogram(cat[1000:1256], an, 200, rng, box))
    np.testing.assert_array_equal(_rdf_counts(ctx, cat[:256], cat, box, 200, rng, exclusion=(1, 1)),
                                  oracle.radial_histogram(cat[:256], cat, 200, rng, box, exclusion=(1, 1)))
    # the full cross g(r) restricted to those 256 cations is what the class computes for the whole group: check the
    # whole-group kernel against the slab through additivity over i (cations 1000..1255 removed = the rest)
    full = _rdf_counts(ctx, cat, an, box, 200, rng)
    rest = _rdf_counts(ctx, np.concatenate((cat[:1000], cat[1256:])), an, box, 200, rng)
    np.testing.assert_array_equal(full - rest, oracle.radial_histogram(cat[1000:1256], an, 200, rng, box))

    pairs = [(0, 0), (0, 1), (1, 1)]
    plan = _capi.SqPlan(ctx, [half, half], pairs, n_points=201, L0=[L] * 3, q_max=20.0)
    assert plan.n_wavevectors == 4_269_143 and plan.info()["algo"] == "nufft"
    wv = plan.wavevectors()
    plan.push(pos[None])
    got, _ = plan.read()
    plan.close()
    idx = np.sort(np.random.default_rng(4).choice(len(wv), 128, replace=False))
    want, nfr = oracle.sq_accumulate([pos], [half, half], wv[idx], pairs)
    assert nfr == 1
    got, want = got[:, idx] / n, want / n
    saa, sbb = want[0], want[2]
    for i, scale in enumerate((np.sqrt(saa), np.sqrt(saa) + np.sqrt(sbb), np.sqrt(sbb))):
        tol = 1e-5 * np.abs(want[i]) + 2 * 5 * 1e-8 * scale + 1e-12
        assert (np.abs(got[i] - want[i]) <= tol).all(), (i, float(np.max(np.abs(got[i] - want[i]) / tol)))
    big = saa > 0.01
    assert np.max(np.abs(got[0][big] / saa[big] - 1)) < 1e-5


def test_device_blocks_with_subrange_groups_at_full_size(ctx):
    """Frame blocks resident in HBM (trajectory.frames_block_device) consumed by groups that are SUB-RANGES of the atoms
    (cross g(r), partial S(q)) at N = 120,000: the slices are made contiguous by torch kernels on torch's stream and
    then read by the library's own (non-blocking) streams: results must equal the per-frame host protocol."""
    import torch

    from mdcraft_b200 import synthetic
    from mdcraft_b200.analysis.base import run_together
    from mdcraft_b200.analysis.structure import RadialDistributionFunction, StructureFactor

    class DeviceTrajectory(synthetic.MemoryTrajectory):
        def frames_block_device(self, lo, hi):
            return torch.from_numpy(self._root._array[lo:hi]).cuda()

    n, F = 120_000, 3
    L = float(synthetic.cubic_box_length(n, 0.8))
    arr = np.stack([synthetic.uniform_frame(n, L, 77 + f) for f in range(F)])
    dims = np.array([L, L, L, 90, 90, 90], np.float32)
    results = []
    for cls in (DeviceTrajectory, synthetic.MemoryTrajectory):
        u = synthetic.Universe(cls(arr, dims))
        a, b = u.select(0, 50_000), u.select(50_000, n)
        g = RadialDistributionFunction(a, b, n_bins=100, range_=(0.0, 0.2 * L), verbose=False, frame_block=2)
        s = StructureFactor([a, b], mode="partial", n_points=24, verbose=False, frame_block=2)
        # a sub-range with itself: the exclusion acts on the indices local to the group
        e = RadialDistributionFunction(b, n_bins=50, range_=(0.0, 0.1 * L), exclusion=(1, 1), verbose=False, frame_block=3)
        if cls is DeviceTrajectory:
            run_together([g, s])
            e.run()
        else:
            g.run(), s.run(), e.run()
        results.append((g.results.counts, s.results.ssf, e.results.counts, g.results.rdf))
    np.testing.assert_array_equal(results[0][0], results[1][0])
    np.testing.assert_allclose(results[0][1], results[1][1], rtol=1e-12, atol=1e-13)
    np.testing.assert_array_equal(results[0][2], results[1][2])
    np.testing.assert_array_equal(results[0][3], results[1][3])
    assert results[0][0].sum() > 0 and results[0][2].sum() > 0


def test_c5_million_particles(ctx):
    """C5: 1,000,000 particles: one frame of g(r) (1e12 ordered pairs) and S(q)."""
    from mdcraft_b200 import _capi
    from oracle import oracle

    pos, box = _frame("C5")
    n, L = len(pos), float(box[0])
    rng = (0.0, L / 2)
    counts = _rdf_counts(ctx, pos, None, box, 200, rng, exclusion=(1, 1))
    assert (counts % 2 == 0).all()
    assert abs(counts.sum() / (n * (n - 1.0)) - np.pi / 6) < 2e-5
    slab = _rdf_counts(ctx, pos[:64], pos, box, 200, rng)
    np.testing.assert_array_equal(slab, oracle.radial_histogram(pos[:64], pos, 200, rng, box))

    plan = _capi.SqPlan(ctx, [n], [(None, None)], n_points=32, L0=[L] * 3)
    plan.push(pos[None])
    s32, _ = plan.read()
    s32 = s32[0] / n
    wv = plan.wavevectors()
    plan.close()
    idx = np.random.default_rng(1).choice(len(wv), 96, replace=False)
    rho = oracle.delta_fourier_transform_sum(wv[idx], pos.astype(np.float64))
    want = np.abs(rho) ** 2 / n
    tol = 1e-5 * want + 2 * 5 * 1e-7 * np.sqrt(want) + 1e-12
    assert (np.abs(s32[idx] - want) <= tol).all()


@pytest.mark.parametrize("n_points,q_max,nf", [(161, 20.0, 256), (171, None, 384), (50, 9.0, 128), (24, None, 48),
                                               (343, 20.0, 768), (520, 30.0, 1024)])
def test_nufft_every_fine_grid_family(ctx, n_points, q_max, nf):
    """Every FFT radix plan of the NUFFT (2^a and 3 * 2^a fine grids, 32 ... 1024 points, shared-memory lines up to
    147 KB) against the FP64 reference-arithmetic kernel on a random subset of the grid's wavevectors, plus the
    FP64-mode NUFFT (w = 16) and the partial sum rule."""
    from mdcraft_b200 import _capi

    n, L = 3000, 40.0
    rng = np.random.default_rng(n_points)
    pos = (rng.random((2, n, 3)) * L - 5.0).astype(np.float32)     # some coordinates outside [0, L): the box is periodic
    L0 = [L, L * 1.07, L * 0.94]
    plan = _capi.SqPlan(ctx, [n], [(None, None)], n_points=n_points, L0=L0, q_max=q_max, algo="nufft")
    assert plan.info()["algo"] == "nufft" and plan.info()["nf"] == nf
    wv = plan.wavevectors()
    plan.push(pos)
    got, nfr = plan.read()
    plan.close()
    assert nfr == 2
    sel = rng.choice(wv.shape[0], size=min(2048, wv.shape[0]), replace=False)
    ref = _capi.SqPlan(ctx, [n], [(None, None)], wavevectors=wv[sel], precision="fp64")
    ref.push(pos)
    want, _ = ref.read()
    ref.close()
    # |rho|^2 summed over 2 frames; rho ~ sqrt(n): aliasing error <= 1e-8 sqrt(n) in rho
    tol = 2 * 2 * 1e-8 * np.sqrt(n) * np.sqrt(np.abs(want[0]) / 2 + 1) + 1e-9
    assert np.max(np.abs(got[0, sel] - want[0]) / tol) < 1.0
    p64 = _capi.SqPlan(ctx, [n], [(None, None)], n_points=n_points, L0=L0, q_max=q_max, algo="nufft", precision="fp64")
    p64.push(pos)
    got64, _ = p64.read()
    p64.close()
    np.testing.assert_allclose(got64[0, sel], want[0], rtol=1e-10, atol=1e-9)
    if n_points <= 171:
        pp = _capi.SqPlan(ctx, [1200, n - 1200], [(0, 0), (0, 1), (1, 1)], n_points=n_points, L0=L0, q_max=q_max,
                          algo="nufft")
        pp.push(pos)
        parts, _ = pp.read()
        pp.close()
        np.testing.assert_allclose(parts.sum(axis=0), got[0], rtol=1e-9, atol=1e-6)


def test_rdf_small_cutoff_sorted_and_culled(ctx):
    """Cutoffs well below half the box: frames are put in Morton order and distant tile pairs are skipped.  Counts
    must not change: against the same plan with sorting switched off (all frames) and against the CPU oracle
    (frame 0); self and cross g(r), block exclusion on ORIGINAL indices, a box that changes from frame to frame."""
    from mdcraft_b200 import _capi
    from oracle import oracle

    rng = np.random.default_rng(23)
    n, F = 32_768, 3
    boxes = np.array([[64.0, 66.0, 62.0, 90, 90, 90], [65.0, 65.5, 63.0, 90, 90, 90], [63.0, 67.0, 61.5, 90, 90, 90]],
                     np.float32)
    pos = (rng.random((F, n, 3)) * boxes[:, None, :3]).astype(np.float32)

    def counts(pa, pb, excl, sort, rcut, frames=F):
        old = os.environ.get("MDC_RDF_SORT")
        os.environ["MDC_RDF_SORT"] = "1" if sort else "0"
        try:
            plan = _capi.RdfPlan(ctx, 64, rcut, pa.shape[1], None if pb is None else pb.shape[1], excl)
            plan.push(pa[:frames], None if pb is None else pb[:frames], boxes[:frames], n_frames=frames)
            c, _ = plan.read()
            ms, _ = plan.last_push_stats()
            plan.close()
        finally:
            if old is None:
                del os.environ["MDC_RDF_SORT"]
            else:
                os.environ["MDC_RDF_SORT"] = old
        return c, ms

    a, b = np.ascontiguousarray(pos[:, :12_000]), np.ascontiguousarray(pos[:, 12_000:])
    for pa, pb, excl, rcut in ((pos, None, (3, 3), (0.0, 3.1)), (pos, None, None, (0.0, 3.1)),
                               (a, b, (2, 5), (0.0, 3.1)), (pos, None, (1, 1), (1.2, 4.0))):
        got, ms_sorted = counts(pa, pb, excl, True, rcut)
        plain, ms_plain = counts(pa, pb, excl, False, rcut)
        np.testing.assert_array_equal(got, plain)
        assert ms_sorted < 0.8 * ms_plain, (ms_sorted, ms_plain)      # most pairs are skipped (measured: 0.2 - 0.4)
        got0, _ = counts(pa, pb, excl, True, rcut, frames=1)
        qb = pa if pb is None else pb
        want = oracle.radial_histogram(pa[0], qb[0], 64, rcut, boxes[0], exclusion=excl)
        np.testing.assert_array_equal(got0, want)

    # a simple-cubic lattice (spacing 2) whose neighbour distances 2 and 4 sit exactly on bin and window edges: the
    # undecided-pair path of the sorted kernel, window edges r0 exclusive / r1 inclusive
    g = np.arange(32, dtype=np.float32) * 2.0
    lat = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(1, -1, 3)
    lat = np.ascontiguousarray(lat[:, rng.permutation(lat.shape[1])])
    boxes = np.array([[64.0, 64.0, 64.0, 90, 90, 90]], np.float32)
    for rcut in ((0.0, 4.0), (2.0, 4.0)):
        got, _ = counts(lat, None, None, True, rcut, frames=1)
        plain, _ = counts(lat, None, None, False, rcut, frames=1)
        np.testing.assert_array_equal(got, plain)
        want = oracle.radial_histogram(lat[0], lat[0], 64, rcut, boxes[0])
        np.testing.assert_array_equal(got, want)
        assert got[-1] >= 6 * lat.shape[1]          # the d = 4 neighbours land in the last bin (upper edge inclusive)
