"""
GPU: the CUDA path (through the C ABI and through the drop-in classes) against
the golden vectors of the real reference and against the CPU oracle.

Bars: g(r) counts bit-exact; S(q) within 1e-5 relative of the
reference in FP32 mode (plus the reference's own atol=1e-6 where S is tiny),
1e-10 in FP64 mode.
"""

import numpy as np
import pytest

from conftest import golden, universe_from

pytestmark = pytest.mark.gpu

#: per-term rms error of the FP32 S(q) kernels, measured on B200 (tools/diag_sq.py):
#: the SFU's sin/cos (3.5e-7) dominates DIRECT; FACTORED only carries FP32 roundings.
#: NUFFT: FP64 arithmetic on the exact float32 coordinates; aliasing error <= 1e-8 of a typical |rho|.
KAPPA = {"direct": 5e-7, "factored": 1e-7, "auto": 1e-7, "nufft": 1e-8}


def _sq_close(got, want, precision, algorithm="auto", pairs=((None, None),)):
    """
    FP64 mode: 1e-10 relative.  FP32 mode: 1e-5 relative PLUS the statistical floor of a sum of N
    terms that each carry an independent error kappa:  rho_a is off by ~kappa sqrt(N_a), hence
        |dS_ab| <= 2 m kappa (sqrt(S_aa) + sqrt(S_bb))        (m = 5 sigma; S_aa = S for the total S(q)),
    which is below 1e-5 S wherever S is not tiny, and below the reference's own test tolerance
    (atol = 1e-6, tests/test_analysis_structure.py:207) everywhere for the default algorithm.
    """
    if precision == "fp64":
        np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12)
        return
    pairs = [tuple(p) for p in pairs]
    self_rows = {p[0]: i for i, p in enumerate(pairs) if p[0] == p[1]}
    for i, (j, k) in enumerate(pairs):
        if j == k:
            scale = np.sqrt(np.abs(want[i]))
        elif j in self_rows and k in self_rows:
            scale = np.sqrt(np.abs(want[self_rows[j]])) + np.sqrt(np.abs(want[self_rows[k]]))
        else:   # mode="pair": the self terms are not part of the result; both are O(1)
            scale = np.full(want.shape[1], 2.0)
        tol = 1e-5 * np.abs(want[i]) + 2 * 5 * KAPPA[algorithm] * scale + 1e-12
        bad = np.abs(got[i] - want[i]) > tol
        assert not bad.any(), (i, (j, k), int(bad.sum()), float(np.max(np.abs(got[i] - want[i]) / tol)))


# --------------------------------------------------------------------------- C ABI


def test_device_is_b200_class(ctx):
    info = ctx.info()
    assert info["cc"][0] >= 10 and info["sm_count"] > 0


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_dfts_seam(ctx, precision):
    g = golden("sq_dfts.npz")
    rho = ctx.delta_fourier_transform_sum(g["qs"], g["rs"], precision)
    tol = 1e-9 if precision == "fp64" else 2e-3   # |rho| ~ sqrt(777); FP32 mode: absolute error ~1e-4
    np.testing.assert_allclose(rho, g["rho"], rtol=0, atol=tol)


def test_radial_histogram_seam(ctx):
    g = golden("rdf_radial_histogram.npz")
    h = ctx.radial_histogram(g["origin"], g["neighbors"], 10, (0, 11), g["dims"])
    np.testing.assert_array_equal(h, g["counts"])


def test_wavevector_grid_built_on_device(ctx):
    from mdcraft_b200 import _capi
    from oracle import oracle

    for dims, n_points, q_max in (((10.77, 10.77, 10.77), 32, None), ((10.77, 10.77, 10.77), 36, 20.0),
                                  ((8.0, 10.0, 6.5), 9, 5.0), ((50.0, 50.0, 50.0), 20, 1.0)):
        L0 = np.array(dims, np.float32).astype(np.float64)
        wv, _ = oracle.wavevector_grid(L0, n_points, q_max)
        plan = _capi.SqPlan(ctx, [10], [(None, None)], n_points=n_points, L0=L0, q_max=q_max)
        assert plan.n_wavevectors == wv.shape[0]
        np.testing.assert_array_equal(plan.wavevectors(), wv)
        plan.close()


def test_plan_errors(ctx):
    from mdcraft_b200 import _capi

    with pytest.raises(_capi.MdcError):
        _capi.RdfPlan(ctx, 0, (0.0, 1.0), 10)
    with pytest.raises(_capi.MdcError):
        _capi.RdfPlan(ctx, 10, (2.0, 1.0), 10)
    with pytest.raises(_capi.MdcError):
        _capi.SqPlan(ctx, [10], [(None, None)], n_points=0)
    plan = _capi.RdfPlan(ctx, 10, (0.0, 1.0), 10)
    with pytest.raises(_capi.MdcError, match="valid cell"):      # 30 + 60 < 100: no such cell
        plan.push(np.zeros((1, 10, 3), np.float32), None, np.array([[5, 5, 5, 30, 60, 100]], np.float32))
    plan.push(np.zeros((1, 10, 3), np.float32), None, np.array([[5, 5, 5, 90, 60, 90]], np.float32))   # triclinic: accepted
    with pytest.raises(ValueError):
        plan.push(np.zeros((1, 11, 3), np.float32), None, np.array([[5, 5, 5, 90, 90, 90]], np.float32))
    plan.close()


def test_histogram_bins_through_the_kernel(ctx):
    """Every pinned (value -> bin) of the reference's numba_histogram, pushed through the pair kernel
    as a two-particle system whose separation is exactly that value."""
    g = golden("histogram_pins.npz")
    box = np.array([16384, 16384, 16384, 90, 90, 90], np.float32)   # power of two: box * (inv_box * dx) == dx exactly
    checked = 0
    for c in range(int(g["n_cases"])):
        r0, r1, n = g[f"case{c}_args"]
        n = int(n)
        for x, want in zip(g[f"case{c}_values"], g[f"case{c}_bins"]):
            xf = np.float32(x)
            if float(xf) != x or x < 0:     # only separations a float32 coordinate can realise exactly
                continue
            a = np.zeros((1, 3), np.float32)
            b = np.array([[xf, 0, 0]], np.float32)
            h = ctx.radial_histogram(a, b, n, (r0, r1), box)
            nz = np.nonzero(h)[0]
            assert (nz[0] if len(nz) else -1) == want, (c, x)
            checked += 1
    assert checked > 10


# --------------------------------------------------------------------------- g(r) class


def _rdf(*args, **kw):
    from mdcraft_b200.analysis.structure import RadialDistributionFunction

    return RadialDistributionFunction(*args, verbose=False, **kw).run()


def test_rdf_self_c1_golden():
    g = golden("rdf_self_c1.npz")
    u = universe_from(g["pos"], g["dims"])
    r = _rdf(u.atoms, n_bins=int(g["n_bins"]), range_=tuple(g["range_"]), frame_block=3)
    assert r.results.counts.dtype == np.int64
    np.testing.assert_array_equal(r.results.counts, g["counts"])
    np.testing.assert_array_equal(r.results.bin_edges, g["bin_edges"])
    np.testing.assert_array_equal(r.results.bins, g["bins"])
    np.testing.assert_allclose(r.results.rdf, g["rdf"], rtol=1e-12)
    rx = _rdf(u.atoms, n_bins=int(g["n_bins"]), range_=tuple(g["range_"]), exclusion=(1, 1))
    np.testing.assert_array_equal(rx.results.counts, g["counts_excl"])
    np.testing.assert_allclose(rx.results.rdf, g["rdf_excl"], rtol=1e-12)
    # passing the same group twice is the same thing (structure.py:805)
    r2 = _rdf(u.atoms, u.atoms, n_bins=int(g["n_bins"]), range_=tuple(g["range_"]))
    np.testing.assert_array_equal(r2.results.counts, g["counts"])
    # run(start, stop, step)
    r3 = _rdf(u.atoms, n_bins=int(g["n_bins"]), range_=tuple(g["range_"]))
    r3.run(start=1, stop=4, step=2)
    from oracle import oracle

    want = oracle.rdf([g["pos"][1], g["pos"][3]], None, [g["dims"]] * 2, int(g["n_bins"]), tuple(g["range_"]))
    np.testing.assert_array_equal(r3.results.counts, want["counts"])
    np.testing.assert_allclose(r3.results.rdf, want["rdf"], rtol=1e-12)


def test_rdf_cross_exclusion_norms_golden():
    g = golden("rdf_cross_excl.npz")
    u = universe_from(g["pos"], g["dims"])
    na = int(g["sizes"][0])
    ga, gb = u.select(0, na), u.select(na, g["pos"].shape[1])
    for norm in ("rdf", "density", None):
        r = _rdf(ga, gb, n_bins=int(g["n_bins"]), range_=tuple(g["range_"]), exclusion=tuple(int(e) for e in g["exclusion"]),
                 norm=norm)
        np.testing.assert_array_equal(r.results.counts, g["counts"])
        np.testing.assert_allclose(r.results.rdf, g[f"rdf_{norm}"], rtol=1e-12)
    # asymmetric block exclusion on ONE group takes the ordered-pair path
    from oracle import oracle

    want = oracle.rdf(list(g["pos"]), None, [g["dims"]] * len(g["pos"]), 50, (0.0, 4.0), exclusion=(4, 10))
    r = _rdf(u.atoms, n_bins=50, range_=(0.0, 4.0), exclusion=(4, 10))
    np.testing.assert_array_equal(r.results.counts, want["counts"])


def test_rdf_npt_and_drop_axis_golden():
    g = golden("rdf_npt_dropaxis.npz")
    u = universe_from(g["pos"], g["dims"])
    r = _rdf(u.atoms, n_bins=int(g["n_bins"]), range_=tuple(g["range_"]), frame_block=2)
    np.testing.assert_array_equal(r.results.counts, g["counts"])
    np.testing.assert_allclose(r.results.rdf, g["rdf"], rtol=1e-12)
    rd = _rdf(u.atoms, n_bins=int(g["n_bins"]), range_=tuple(g["range_"]), drop_axis="z")
    np.testing.assert_array_equal(rd.results.counts, g["counts_drop"])
    np.testing.assert_allclose(rd.results.rdf, g["rdf_drop"], rtol=1e-12)


def test_rdf_triclinic_golden_and_vs_oracle(ctx):
    """Triclinic cells (structure.py:109-115 hands all six box values to capped_distance): the classes against the
    golden vectors of the reference classes (60/90/75 cell; NPT mixing orthorhombic and triclinic frames), and the
    functional seam against the CPU oracle at sizes where every code path of the kernel runs (ragged tiles, two groups,
    exclusion, coordinates far outside the primary cell)."""
    from mdcraft_b200.analysis.structure import radial_histogram
    from oracle import oracle

    g = golden("rdf_triclinic.npz")
    u = universe_from(g["a_pos"], g["a_dims"])
    r = _rdf(u.atoms, n_bins=80, range_=(0.0, 4.5), frame_block=2)
    np.testing.assert_array_equal(r.results.counts, g["a_counts"])
    np.testing.assert_allclose(r.results.rdf, g["a_rdf"], rtol=1e-12)
    r = _rdf(u.select(0, 180), u.select(180, 500), n_bins=57, range_=(0.4, 4.9), exclusion=(3, 5))
    np.testing.assert_array_equal(r.results.counts, g["a_cross_counts"])
    np.testing.assert_allclose(r.results.rdf, g["a_cross_rdf"], rtol=1e-12)
    u = universe_from(g["b_pos"], g["b_dims"])
    r = _rdf(u.atoms, n_bins=64, range_=(0.0, 3.9), exclusion=(1, 1), frame_block=3)    # mixed blocks: ortho + triclinic
    np.testing.assert_array_equal(r.results.counts, g["b_counts"])
    np.testing.assert_allclose(r.results.rdf, g["b_rdf"], rtol=1e-12)

    rng = np.random.default_rng(75)
    for box, n_a, n_b, excl in (([30.0, 28.0, 33.0, 60, 90, 75], 2500, None, None), ([22.0, 25.0, 21.0, 70, 110, 95], 700, 1800, (2, 3)),
                                ([18.0, 18.0, 18.0, 90, 90, 120], 1025, None, (1, 1))):
        box = np.array(box, np.float32)
        m = oracle.triclinic_vectors(box).astype(np.float64)
        a = ((rng.random((n_a, 3)) * 5 - 2) @ m).astype(np.float32)
        b = a if n_b is None else ((rng.random((n_b, 3)) * 5 - 2) @ m).astype(np.float32)
        rcut = (0.0, 0.45 * float(min(m[0, 0], m[1, 1], m[2, 2])))
        got = radial_histogram(a, b, n_bins=90, range_=rcut, dimensions=box, exclusion=excl)
        np.testing.assert_array_equal(got, oracle.radial_histogram(a, b, 90, rcut, box, exclusion=excl))
    with pytest.raises(Exception, match="valid cell"):
        radial_histogram(a, a, n_bins=10, range_=(0.0, 3.0), dimensions=np.array([10, 10, 10, 30, 100, 160], np.float32))


def _cell_widths(m):
    """Perpendicular widths of the cell with row vectors ``m`` (distance between opposite faces)."""
    vol = abs(np.linalg.det(m))
    return np.array([vol / np.linalg.norm(np.cross(m[(k + 1) % 3], m[(k + 2) % 3])) for k in range(3)])


@pytest.mark.parametrize("box,n_a,n_b,excl,n_bins,frac", [
    ([30.0, 28.0, 33.0, 60, 90, 75], 6000, None, (1, 1), 200, 0.499),      # sorted frames, relative rows, extended histograms
    ([30.0, 28.0, 33.0, 60, 90, 75], 2500, None, None, 90, 0.49),          # below the sort threshold
    ([22.0, 25.0, 21.0, 70, 110, 95], 4500, 5200, (2, 3), 128, 0.45),      # two groups, exclusion
    ([40.0, 40.0, 40.0, 50, 65, 55], 7000, None, (1, 1), 300, 0.2),        # strongly skewed cell, small cutoff (culling)
    ([18.0, 18.0, 18.0, 90, 90, 120], 4100, None, (1, 1), 64, 0.48),       # hexagonal
    ([25.0, 25.0, 25.0, 80, 95, 100], 5000, None, None, 2000, 0.3),        # narrow bins (range-tested loop)
])
def test_rdf_triclinic_filter_vs_oracle(box, n_a, n_b, excl, n_bins, frac):
    """Triclinic cells with the window below half the smallest width of the cell: these frames take the FILTER kernel
    (minimum image = wrap of the fractional differences; undecided pairs resolved with minimum_image_triclinic in
    FP64) and must give the reference's counts bit for bit (oracle.c: orc_radial_histogram_triclinic).  Coordinates far
    outside the primary cell; a lattice along the cell vectors puts separations on bin edges."""
    from mdcraft_b200.analysis.structure import radial_histogram
    from oracle import oracle

    rng = np.random.default_rng(n_a + n_bins)
    box = np.array(box, np.float32)
    m = oracle.triclinic_vectors(box).astype(np.float64)
    a = ((rng.random((n_a, 3)) * 5 - 2) @ m).astype(np.float32)
    a[: n_a // 8] = ((np.round(rng.random((n_a // 8, 3)) * 16) / 16) @ m).astype(np.float32)   # lattice points
    b = a if n_b is None else ((rng.random((n_b, 3)) * 3 - 1) @ m).astype(np.float32)
    rcut = (0.0, frac * float(_cell_widths(m).min()))
    got = radial_histogram(a, b, n_bins=n_bins, range_=rcut, dimensions=box, exclusion=excl)
    np.testing.assert_array_equal(got, oracle.radial_histogram(a, b, n_bins, rcut, box, exclusion=excl))
    # ... and it IS the filter kernel that ran (mdc_rdf_plan_info), the FP64 kernel beyond half the width
    from mdcraft_b200 import _capi

    ctx = _capi.context(0)
    for r1, filt in ((rcut[1], True), (0.62 * float(_cell_widths(m).min()), False)):
        plan = _capi.RdfPlan(ctx, 57, (0.0, r1), len(a), None if n_b is None else len(b), excl)
        plan.push(a[None], None if n_b is None else b[None], box, n_frames=1)
        info = plan.info()
        plan.close()
        assert info["triclinic"] and info["filter"] == filt and info["sorted"] == (filt and min(len(a), len(b)) >= 4096), info
    # a window that starts above zero, and one beyond half the width (exact kernel)
    for r0, r1 in ((0.3 * rcut[1], rcut[1]), (0.0, 0.62 * float(_cell_widths(m).min()))):
        got = radial_histogram(a[:3000], b[:3500], n_bins=57, range_=(r0, r1), dimensions=box, exclusion=excl)
        np.testing.assert_array_equal(got, oracle.radial_histogram(a[:3000], b[:3500], 57, (r0, r1), box, exclusion=excl))


@pytest.mark.parametrize("n,n_bins,frac", [(1, 8, 0.5), (2, 8, 0.5), (511, 33, 0.5), (512, 200, 0.5), (513, 200, 0.37),
                                           (1025, 1000, 0.5), (3000, 200, 0.5), (2049, 4096, 0.45),
                                           (4096, 200, 0.5), (4097, 64, 0.3), (5121, 300, 0.5), (6000, 3000, 0.2),
                                           (12289, 200, 0.1)])
def test_rdf_ragged_sizes_vs_oracle(n, n_bins, frac):
    """Tile-edge sizes (tiles are 512 x 512), tiny systems, many bins."""
    from oracle import oracle

    rng = np.random.default_rng(n)
    L = np.float32(max(2.0, (n / 0.8) ** (1 / 3)))
    pos = (rng.random((2, n, 3)) * L).astype(np.float32)
    dims = np.array([L, L, L, 90, 90, 90], np.float32)
    rng_ = (0.0, float(L) * frac)
    want = oracle.rdf(list(pos), None, [dims] * 2, n_bins, rng_)
    u = universe_from(pos, dims)
    r = _rdf(u.atoms, n_bins=n_bins, range_=rng_)
    np.testing.assert_array_equal(r.results.counts, want["counts"])
    if n > 2:
        na = n // 3
        want = oracle.rdf(list(pos[:, :na]), list(pos[:, na:]), [dims] * 2, n_bins, rng_)
        r = _rdf(u.select(0, na), u.select(na, n), n_bins=n_bins, range_=rng_)
        np.testing.assert_array_equal(r.results.counts, want["counts"])


@pytest.mark.parametrize("window,n_bins,excl", [((0.0, 0.5), 200, None), ((0.9, 0.5), 200, (1, 1)), ((2.5, 0.31), 77, (3, 3)),
                                                 ((0.2, 0.12), 500, None), ((0.05, 0.5), 1500, (2, 2))])
def test_rdf_sorted_frames_windows_and_changing_boxes(window, n_bins, excl):
    """Frames above the sort threshold (4096 particles): Hilbert-ordered frames, culling, extended histograms (or
    the range-tested loop where the bin range does not fit), windows that start above zero, a box that changes a
    lot from frame to frame, non-cubic boxes; self and cross g(r) with one group below the threshold."""
    from oracle import oracle

    rng = np.random.default_rng(77)
    n, F = 5000, 3
    dims = np.array([[17.0, 17.0, 17.0, 90, 90, 90], [19.5, 16.0, 21.0, 90, 90, 90], [14.0, 15.0, 14.5, 90, 90, 90]], np.float32)
    pos = (rng.random((F, n, 3)) * dims[:, None, :3]).astype(np.float32)
    pos[:, ::7] += dims[:, None, :3] * rng.integers(-2, 3, (F, len(range(0, n, 7)), 3))   # unwrapped particles
    pos = pos.astype(np.float32)
    r0, frac = window
    rng_ = (r0, float(dims[:, :3].min()) * frac) if frac * dims[:, :3].min() > r0 else (r0, r0 + 3.0)
    want = oracle.rdf(list(pos), None, list(dims), n_bins, rng_, exclusion=excl)
    u = universe_from(pos, dims)
    r = _rdf(u.atoms, n_bins=n_bins, range_=rng_, exclusion=excl)
    np.testing.assert_array_equal(r.results.counts, want["counts"])
    na = 4200
    want = oracle.rdf(list(pos[:, :na]), list(pos[:, na:]), list(dims), n_bins, rng_, exclusion=excl)
    r = _rdf(u.select(0, na), u.select(na, n), n_bins=n_bins, range_=rng_, exclusion=excl)
    np.testing.assert_array_equal(r.results.counts, want["counts"])


def test_rdf_adversarial_edge_distances():
    """Particles on a lattice whose spacings fall exactly on bin edges, coincident particles,
    separations of exactly L/2 (minimum-image ties) and unwrapped coordinates."""
    from oracle import oracle

    L = np.float32(16.0)
    g1 = np.arange(0, 16, 0.5, dtype=np.float32)
    lattice = np.stack(np.meshgrid(g1[::2], g1[::4], g1[::8], indexing="ij"), -1).reshape(-1, 3)
    extra = np.array([[0, 0, 0], [0, 0, 0], [8, 0, 0], [8, 8, 8], [-3.5, 20.25, 33.0], [15.999999, 0, 0]], np.float32)
    pos = np.concatenate([lattice, extra])[None]
    dims = np.array([L, L, L, 90, 90, 90], np.float32)
    for n_bins, rng_ in ((16, (0.0, 8.0)), (64, (0.0, 8.0)), (200, (0.0, 8.0)), (10, (1.0, 6.0)), (7, (0.5, 7.5))):
        want = oracle.rdf(list(pos), None, [dims], n_bins, rng_)
        u = universe_from(pos, dims)
        r = _rdf(u.atoms, n_bins=n_bins, range_=rng_)
        np.testing.assert_array_equal(r.results.counts, want["counts"])


def test_rdf_and_sq_residue_groupings(tmp_path):
    """groupings="residues": centres of mass are formed on the host (algorithm.center_of_mass) and take the
    same GPU path; results.save() round-trips."""
    from mdcraft_b200 import algorithm, synthetic
    from oracle import oracle

    rng = np.random.default_rng(21)
    n_res, per, L = 300, 3, 12.0
    centers = rng.random((2, n_res, 1, 3)) * L
    pos = (centers + rng.normal(scale=0.3, size=(2, n_res, per, 3))).reshape(2, -1, 3).astype(np.float32)
    masses = np.tile([16.0, 1.0, 1.0], n_res)
    dims = np.array([L, L, L, 90, 90, 90], np.float32)
    traj = synthetic.MemoryTrajectory(pos, dims)
    u = synthetic.Universe(traj, resindices=np.repeat(np.arange(n_res), per), masses=masses)
    coms, coms64 = [], []
    for i in range(2):
        u.trajectory[i]
        coms64.append(algorithm.center_of_mass(u.atoms, "residues"))
        coms.append(coms64[-1].astype(np.float32))          # capped_distance itself casts to float32
    want = oracle.rdf(coms, None, [dims] * 2, 60, (0.0, 6.0))
    r = _rdf(u.atoms, n_bins=60, range_=(0.0, 6.0), groupings="residues")
    np.testing.assert_array_equal(r.results.counts, want["counts"])
    np.testing.assert_allclose(r.results.rdf, want["rdf"], rtol=1e-12)
    r.save(str(tmp_path / "rdf"))
    saved = np.load(str(tmp_path / "rdf.npz"), allow_pickle=True)
    np.testing.assert_array_equal(saved["counts"], r.results.counts)
    # atoms of one group against residues of the same group: two different entity sets
    want2 = oracle.rdf([p for p in pos], coms, [dims] * 2, 40, (0.0, 5.0))
    r2 = _rdf(u.atoms, u.atoms, n_bins=40, range_=(0.0, 5.0), groupings=("atoms", "residues"))
    np.testing.assert_array_equal(r2.results.counts, want2["counts"])
    # S(q): the reference keeps the float64 centres of mass and evaluates q . r on them in FP64 (structure.py:1943-1944);
    # a float32 staging buffer would cost ~q |r| 6e-8 of phase and break the FP64-mode tolerance
    ws = oracle.structure_factor(coms64, [n_res], dims[:3], mode="pair", n_points=6)
    s = _sq(u.atoms, "residues", mode="pair", n_points=6)
    np.testing.assert_array_equal(s.results.wavenumbers, ws["wavenumbers"])
    np.testing.assert_allclose(s.results.ssf, ws["ssf"], rtol=1e-5, atol=1e-6)
    for algorithm_ in ("direct", "nufft"):
        s64 = _sq(u.atoms, "residues", mode="pair", n_points=6, precision="fp64", algorithm=algorithm_)
        np.testing.assert_allclose(s64.results.ssf, ws["ssf"], rtol=1e-10, atol=1e-12)


def test_backward_and_strided_runs():
    """run(step=-1) visits the frames of run() backwards (a sum: same counts), and a strided backward run that ends on
    frame 0 visits exactly range(start, -1, step) (check_slice_indices normalises its stop to -1)."""
    from mdcraft_b200.analysis.structure import RadialDistributionFunction

    g = golden("rdf_self_c1.npz")
    u = universe_from(g["pos"], g["dims"])
    kw = dict(n_bins=int(g["n_bins"]), range_=tuple(g["range_"]), verbose=False, frame_block=2)
    fwd = RadialDistributionFunction(u.atoms, **kw).run()
    bwd = RadialDistributionFunction(u.atoms, **kw).run(step=-1)
    np.testing.assert_array_equal(bwd.results.counts, fwd.results.counts)
    np.testing.assert_array_equal(bwd.frames, fwd.frames[::-1])
    n = len(u.trajectory)
    part = RadialDistributionFunction(u.atoms, **kw).run(start=n - 1, step=-2)
    np.testing.assert_array_equal(part.frames, np.arange(n - 1, -1, -2))
    same = RadialDistributionFunction(u.atoms, **kw).run(frames=list(range(n - 1, -1, -2)))
    np.testing.assert_array_equal(part.results.counts, same.results.counts)


def test_verbose_logs_the_run(caplog):
    """verbose=True (the reference's default): a start and a finish line per run with device, frames per block, the plan's
    algorithm and frames/s (the reference logs its process farm the same way, base.py:378-384, 514-515); silent otherwise."""
    import logging

    from mdcraft_b200.analysis.structure import RadialDistributionFunction, StructureFactor

    g = golden("rdf_self_c1.npz")
    u = universe_from(g["pos"], g["dims"])
    with caplog.at_level(logging.INFO, logger="mdcraft_b200.analysis"):
        RadialDistributionFunction(u.atoms, n_bins=20, range_=(0.0, 3.0), verbose=False).run()
        assert not caplog.records
        RadialDistributionFunction(u.atoms, n_bins=20, range_=(0.0, 3.0)).run(stop=3)
        StructureFactor(u.atoms, n_points=8, algorithm="nufft").run(stop=2)
    text = [r.getMessage() for r in caplog.records]
    assert len(text) == 4
    assert "RadialDistributionFunction: analysing 3 frames on cuda:0" in text[0] and "frames/s" in text[1]
    assert "algorithm nufft (nf = " in text[3]


def test_rdf_post_processing_runs():
    g = golden("rdf_self_c1.npz")
    u = universe_from(g["pos"], g["dims"])
    r = _rdf(u.atoms, n_bins=int(g["n_bins"]), range_=tuple(g["range_"]), exclusion=(1, 1))
    r.calculate_coordination_numbers(0.8, threshold=0.01)
    assert r.results.coordination_numbers.shape == (2,)
    r.calculate_pmf(300.0)
    assert r.results.pmf.shape == r.results.rdf.shape
    r.calculate_structure_factor(0.8, n_q=32)
    assert r.results.ssf.shape == (32,)


def test_radial_transforms_on_device_golden():
    """the dense (n_q x n_r) transforms evaluated by the device kernel, against the outputs of
    the reference's own functions and against the host NumPy path."""
    from mdcraft_b200.analysis import structure

    g = golden("transforms.npz")
    for tag in ("odd", "even"):
        r, gr, q = g[f"{tag}_r"], g[f"{tag}_g"], g[f"{tag}_q"]
        np.testing.assert_allclose(structure.radial_fourier_transform(r, gr - 1, q, device=0), g[f"{tag}_rft"],
                                   rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(structure.zeroth_order_hankel_transform(r, gr - 1, g[f"{tag}_ht_q"], device=0),
                                   g[f"{tag}_ht"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(structure.zeroth_order_hankel_transform(r, gr - 1, q, device=0),
                                   structure.zeroth_order_hankel_transform(r, gr - 1, q), rtol=1e-12, atol=1e-12)
        qq, s = structure.calculate_structure_factor(r, gr, True, 0.8, n_q=64, device=0)
        np.testing.assert_array_equal(qq, g[f"{tag}_q3"])
        np.testing.assert_allclose(s, g[f"{tag}_s3"], rtol=1e-12, atol=1e-12)


def test_fourier_route_matches_direct_route():
    """The consistency check rank 3 exists for: S(q) from g(r) by radial Fourier transform (device) against the
    direct S(q) of the same frames, one call each.  Configuration: a lattice with jitter larger than half its
    spacing, i.e. no Bragg peaks left and S(q) = 1 - exp(-q^2 sigma^2) (suppressed long-wavelength fluctuations).
    The two estimators differ by the truncation of g(r) at L/2, binning and the noise of single |q| shells."""
    from mdcraft_b200 import synthetic

    n, rho, F, jitter = 4000, 0.8, 6, 0.6
    L = synthetic.cubic_box_length(n, rho)
    pos = np.stack([synthetic.lj_like_frame(n, L, 900 + f, jitter=jitter) for f in range(F)]).astype(np.float32)
    u = universe_from(pos, np.array([L, L, L, 90, 90, 90], np.float32))
    r = _rdf(u.atoms, n_bins=400, range_=(0.0, L / 2), exclusion=(1, 1))
    s = _sq(u.atoms, n_points=24)
    q = s.results.wavenumbers
    sel = (q > 1.0) & (q < 8.0)
    r.calculate_structure_factor(rho, q=q[sel])
    direct, fourier = s.results.ssf[0][sel], r.results.ssf
    k = np.ones(15) / 15   # shell-average the direct estimate (single |q| shells are noisy)
    smooth = lambda y: np.convolve(y, k, mode="valid")   # noqa: E731
    assert np.max(np.abs(smooth(direct) - smooth(fourier))) < 0.15
    model = 1 - np.exp(-(smooth(q[sel]) * jitter) ** 2)
    assert np.max(np.abs(smooth(fourier) - model)) < 0.15


# --------------------------------------------------------------------------- S(q) class


def _sq(*args, **kw):
    from mdcraft_b200.analysis.structure import StructureFactor

    return StructureFactor(*args, verbose=False, **kw).run()


PRECISION_ALGO = [("fp32", "factored"), ("fp32", "direct"), ("fp32", "nufft"), ("fp64", "auto"), ("fp64", "nufft")]


@pytest.mark.parametrize("precision,algorithm", PRECISION_ALGO)
def test_sq_total_c1_golden(precision, algorithm):
    g = golden("sq_total_c1.npz")
    u = universe_from(g["pos"], g["dims"])
    r = _sq(u.atoms, precision=precision, algorithm=algorithm, frame_block=2)
    assert r.results.pairs == ((None, None),)
    np.testing.assert_array_equal(r.results.wavenumbers, g["wavenumbers"])
    _sq_close(r.results.ssf, g["ssf"], precision, algorithm)
    if precision == "fp32":
        # the FP32 bar, literally: relative error < 1e-5 on every |q| shell that is not tiny
        m = g["ssf"] > (1e-2 if algorithm in ("factored", "nufft") else 0.25)
        assert np.max(np.abs(r.results.ssf[m] / g["ssf"][m] - 1)) < 1e-5
        # and the reference's own test tolerance everywhere
        np.testing.assert_allclose(r.results.ssf, g["ssf"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("precision,algorithm", PRECISION_ALGO)
def test_sq_partial_surfaces_golden(precision, algorithm):
    g = golden("sq_partial_surfaces.npz")
    u = universe_from(g["pos"], g["dims"])
    na = int(g["sizes"][0])
    kw = dict(n_points=int(g["n_points"]), q_max=float(g["q_max"]), n_surfaces=int(g["n_surfaces"]),
              n_surface_points=int(g["n_surface_points"]), precision=precision, algorithm=algorithm)
    r = _sq([u.select(0, na), u.select(na, g["pos"].shape[1])], mode="partial", **kw)
    assert r.results.pairs == tuple(map(tuple, g["pairs"]))
    np.testing.assert_array_equal(r._wavevectors, g["wavevectors"])
    np.testing.assert_array_equal(r.results.wavenumbers, g["wavenumbers"])
    _sq_close(r.results.ssf, g["ssf"], precision, algorithm, r.results.pairs)
    # sum rule: sum of the partials equals the total S(q) of the same run
    tot = _sq(u.atoms, **kw)
    np.testing.assert_allclose(r.results.ssf.sum(axis=0), tot.results.ssf[0], rtol=1e-5, atol=3e-6)


@pytest.mark.parametrize("precision,algorithm", PRECISION_ALGO)
def test_sq_pair_ortho_golden(precision, algorithm):
    g = golden("sq_pair_ortho.npz")
    u = universe_from(g["pos"], g["dims"])
    na = int(g["sizes"][0])
    groups = [u.select(0, na), u.select(na, g["pos"].shape[1])]
    r = _sq(groups, mode="pair", n_points=int(g["n_points"]), sort=False, unique=False, precision=precision,
            algorithm=algorithm)
    assert r.results.pairs == ((0, 1),)
    np.testing.assert_array_equal(r._wavevectors, g["wavevectors"])
    _sq_close(r.results.ssf, g["ssf"], precision, algorithm, r.results.pairs)
    r2 = _sq(groups, mode="pair", n_points=int(g["n_points"]), form="trig", precision=precision, algorithm=algorithm)
    np.testing.assert_array_equal(r2.results.wavenumbers, g["wavenumbers_unique"])
    _sq_close(r2.results.ssf, g["ssf_trig_unique"], precision, algorithm, r2.results.pairs)


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_sq_explicit_golden(precision):
    g = golden("sq_explicit.npz")
    u = universe_from(g["pos"], g["dims"])
    r = _sq(u.atoms, wavevectors=g["wavevectors"], sort=False, unique=False, precision=precision)
    np.testing.assert_array_equal(r.results.wavenumbers, g["wavenumbers"])
    _sq_close(r.results.ssf, g["ssf"], precision, "direct")   # explicit wavevectors use the SFU pipeline


@pytest.mark.parametrize("algorithm", ["factored", "direct", "nufft"])
def test_sq_fp32_is_reproducible_and_frame_block_invariant(algorithm):
    g = golden("sq_total_c1.npz")
    u = universe_from(g["pos"], g["dims"])
    a = _sq(u.atoms, n_points=16, frame_block=1, algorithm=algorithm).results.ssf
    b = _sq(u.atoms, n_points=16, frame_block=1, algorithm=algorithm).results.ssf
    if algorithm == "nufft":   # FP64 reductions into the fine grid arrive in any order
        np.testing.assert_allclose(a, b, rtol=1e-11, atol=1e-13)
    else:
        np.testing.assert_array_equal(a, b)
    c = _sq(u.atoms, n_points=16, frame_block=3, algorithm=algorithm).results.ssf
    np.testing.assert_allclose(a, c, rtol=1e-12)


def test_sq_factored_patch_edges_vs_direct():
    """Grids that do not fill the 32 x 16 x 8 patches, with and without the q_max sphere; the two FP32
    algorithms and the FP64 path must agree wavevector by wavevector."""
    rng = np.random.default_rng(5)
    n, L = 3000, 15.5
    pos = (rng.random((2, n, 3)) * L).astype(np.float32)
    dims = np.array([L, L * 1.1, L * 0.9, 90, 90, 90], np.float32)
    u = universe_from(pos, dims)
    groups = [u.select(0, 1200), u.select(1200, n)]
    for n_points, q_max in ((33, None), (40, 9.0), (17, 4.0), (9, None)):
        kw = dict(mode="partial", n_points=n_points, q_max=q_max, sort=False, unique=False)
        ref = _sq(groups, precision="fp64", **kw)
        for algorithm in ("factored", "direct", "nufft"):
            got = _sq(groups, algorithm=algorithm, **kw).results.ssf
            assert got.shape == ref.results.ssf.shape
            _sq_close(got, ref.results.ssf, "fp32", algorithm, ref.results.pairs)


@pytest.mark.parametrize("algorithm", ["factored", "direct", "nufft"])
def test_sq_large_n_accuracy_vs_oracle(algorithm):
    """The coherent-error budget (SURVEY.md §7 hard part 2) at N = 50,000: one frame, a few hundred
    wavevectors of the default grid against the FP64 oracle."""
    from mdcraft_b200 import synthetic
    from oracle import oracle

    n = 50_000
    L = synthetic.cubic_box_length(n, 0.8)
    pos = synthetic.lj_like_frame(n, L, 7)[None]
    dims = np.array([L, L, L, 90, 90, 90], np.float32)
    want = oracle.structure_factor(list(pos), [n], dims[:3], n_points=7, sort=False, unique=False)
    u = universe_from(pos, dims)
    r = _sq(u.atoms, n_points=7, sort=False, unique=False, algorithm=algorithm)
    s, s0 = r.results.ssf[0], want["ssf"][0]
    _sq_close(r.results.ssf, want["ssf"], "fp32", algorithm)
    np.testing.assert_allclose(s, s0, rtol=1e-5, atol=1e-6)
    m = s0 > (0.01 if algorithm in ("factored", "nufft") else 0.25)
    assert np.max(np.abs(s[m] / s0[m] - 1)) < 1e-5
    if algorithm == "nufft":   # no per-term FP32 rounding at all: two more digits (aliasing ~1e-8 sqrt(N) in rho)
        assert np.max(np.abs(s[m] / s0[m] - 1)) < 1e-6


def test_sq_shell_means_on_device(ctx):
    """StructureFactor._conclude's |q|-shell averaging (structure.py:2000-2009) on the device against the host
    restatement of the reference loop: same windows (np.isclose membership, overlapping), means to rounding."""
    from mdcraft_b200 import _capi
    from mdcraft_b200.analysis.structure import combine_equal_wavenumbers, shell_windows

    rng = np.random.default_rng(17)
    n, L = 1500, 12.5
    pos = (rng.random((3, n, 3)) * L).astype(np.float32)
    plan = _capi.SqPlan(ctx, [600, 900], [(0, 0), (0, 1), (1, 1)], n_points=20, L0=[L, L * 1.02, L], q_max=9.0)
    plan.push(pos)
    sums, nf = plan.read()
    wn = np.linalg.norm(plan.wavevectors(), axis=1)
    uq = np.unique(wn.round(11))
    order, lo, hi = shell_windows(wn, uq)
    assert (hi - lo).max() > 1 and (lo[1:] < hi[:-1]).any()          # real shells, and overlapping windows
    plan.set_shells(order, lo, hi)
    got, nf2 = plan.read_shells(float(nf * n))
    assert nf2 == nf == 3
    want = combine_equal_wavenumbers(sums / (nf * n), wn, uq)
    np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-15)
    again, _ = plan.read_shells(float(nf * n))
    np.testing.assert_array_equal(got, again)                          # fixed reduction order
    plan.close()


def test_run_together_equals_separate_runs():
    """One pass over the trajectory feeding g(r), S(q) and a density profile (their kernels on different stream
    lanes, side by side on the device): results and bookkeeping equal those of three separate run() calls."""
    from mdcraft_b200.analysis.base import run_together
    from mdcraft_b200.analysis.profile import DensityProfile
    from mdcraft_b200.analysis.structure import RadialDistributionFunction, StructureFactor

    g = golden("sq_total_c1.npz")
    u = universe_from(g["pos"], g["dims"])
    L = float(g["dims"][0])

    def make():
        return (RadialDistributionFunction(u.atoms, n_bins=50, range_=(0.0, L / 2), exclusion=(1, 1), verbose=False,
                                           frame_block=2),
                StructureFactor(u.atoms, n_points=10, verbose=False, frame_block=3),
                DensityProfile(u.atoms, axes="z", n_bins=25, verbose=False, frame_block=2))

    sep = [a.run(start=1, step=2) for a in make()]
    tog = run_together(make(), start=1, step=2)
    np.testing.assert_array_equal(tog[0].results.counts, sep[0].results.counts)
    np.testing.assert_array_equal(tog[0].results.rdf, sep[0].results.rdf)
    np.testing.assert_array_equal(tog[1].results.ssf, sep[1].results.ssf)
    np.testing.assert_array_equal(tog[2].results.number_densities["z"], sep[2].results.number_densities["z"])
    for a, b in zip(tog, sep):
        np.testing.assert_array_equal(a.frames, b.frames)
        np.testing.assert_array_equal(a.times, b.times)
        assert a.n_frames == b.n_frames


# --------------------------------------------------------------------------- F(q, t) class


@pytest.mark.parametrize("precision,algorithm", PRECISION_ALGO)
def test_isf_golden(precision, algorithm):
    from mdcraft_b200.analysis.structure import IntermediateScatteringFunction

    g = golden("isf_small.npz")
    u = universe_from(g["pos"], g["dims"])
    n = g["pos"].shape[1]
    r = IntermediateScatteringFunction(u.atoms, n_points=6, n_lags=4, incoherent=True, verbose=False,
                                       precision=precision, algorithm=algorithm, frame_block=3).run()
    np.testing.assert_array_equal(r.results.wavenumbers, g["wavenumbers"])
    np.testing.assert_allclose(r.results.times, g["times"])
    tol = dict(rtol=1e-10, atol=1e-12) if precision == "fp64" else dict(rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(r.results.cisf, g["cisf"], **tol)
    np.testing.assert_allclose(r.results.iisf, g["iisf"], **tol)
    na = int(g["sizes"][0])
    r2 = IntermediateScatteringFunction([u.select(0, na), u.select(na, n)], mode="partial", n_points=5, n_lags=3,
                                        incoherent=True, sort=False, unique=False, verbose=False,
                                        precision=precision, algorithm=algorithm).run()
    assert r2.results.pairs == tuple(map(tuple, g["pairs_partial"]))
    np.testing.assert_allclose(r2.results.cisf, g["cisf_partial"], **tol)
    np.testing.assert_allclose(r2.results.iisf, g["iisf_partial"], **tol)


def test_isf_explicit_wavevectors_and_defaults_vs_oracle():
    from mdcraft_b200.analysis.structure import IntermediateScatteringFunction
    from oracle import oracle

    g = golden("isf_small.npz")
    u = universe_from(g["pos"], g["dims"])
    n = g["pos"].shape[1]
    wv = np.random.default_rng(8).normal(size=(40, 3)) * 2.0
    want = oracle.isf(list(g["pos"]), [n], g["dims"][:3], wavevectors=wv, incoherent=True, sort=False, unique=False)
    r = IntermediateScatteringFunction(u.atoms, wavevectors=wv, incoherent=True, sort=False, unique=False,
                                       verbose=False).run()      # n_lags defaults to every frame
    assert r.results.cisf.shape == want["cisf"].shape == (7, 1, 40)
    np.testing.assert_allclose(r.results.cisf, want["cisf"], rtol=1e-5, atol=5e-6)
    np.testing.assert_allclose(r.results.iisf, want["iisf"], rtol=1e-5, atol=5e-6)
    # coherent only, pair mode (cross term only), run(start, stop, step)
    want = oracle.isf(list(g["pos"][1:7:2]), [100, 150], g["dims"][:3], mode="pair", n_points=5, n_lags=2)
    r = IntermediateScatteringFunction([u.select(0, 100), u.select(100, n)], mode="pair", n_points=5, n_lags=2,
                                       verbose=False).run(start=1, stop=7, step=2)
    np.testing.assert_allclose(r.results.cisf, want["cisf"], rtol=1e-5, atol=2e-6)
    assert "iisf" not in r.results


# --------------------------------------------------------------------------- single-chain S(q) (§8f rank 2)


@pytest.mark.parametrize("precision,algorithm", PRECISION_ALGO)
def test_scsf_golden(precision, algorithm):
    from mdcraft_b200 import synthetic
    from mdcraft_b200.analysis.polymer import SingleChainStructureFactor
    from oracle import oracle

    g = golden("scsf_small.npz")
    M, N = int(g["n_chains"]), int(g["n_monomers"])
    traj = synthetic.MemoryTrajectory(g["pos"], g["dims"])
    u = synthetic.Universe(traj, segindices=np.repeat(np.arange(M), N))
    tol = dict(rtol=1e-10, atol=1e-12) if precision == "fp64" else dict(rtol=1e-5, atol=1e-6)
    r = SingleChainStructureFactor(u.atoms, n_chains=M, n_monomers=N, n_points=int(g["n_points"]), verbose=False,
                                   precision=precision, algorithm=algorithm, frame_block=2).run()
    np.testing.assert_array_equal(r.results.wavenumbers, g["wavenumbers"])
    np.testing.assert_allclose(r.results.scsf, g["scsf"], **tol)
    # chain structure from the topology (segments = chains), two groups, q_max
    r2 = SingleChainStructureFactor([u.select(0, 4 * N), u.select(4 * N, M * N)], n_points=int(g["n_points"]),
                                    q_max=3.0, verbose=False, precision=precision, algorithm=algorithm).run()
    assert list(r2._n_chains) == [4, M - 4] and list(r2._n_monomers) == [N, N]
    for ig, (lo, hi, m) in enumerate(((0, 4 * N, 4), (4 * N, M * N, M - 4))):
        want = oracle.scsf([p[lo:hi] for p in g["pos"]], m, N, g["dims"][:3], n_points=int(g["n_points"]), q_max=3.0)
        np.testing.assert_array_equal(r2.results.wavenumbers, want["wavenumbers"])
        np.testing.assert_allclose(r2.results.scsf[ig], want["scsf"][0], **tol)


# --------------------------------------------------------------------------- frame ingest (§8f rank 4)


@pytest.mark.parametrize("tag,name", [("ids", "dump_ids"), ("noid", "dump_noid"), ("sparse", "dump_sparse_ids"),
                                      ("triclinic", "dump_triclinic")])
def test_dump_reader_golden(tag, name):
    """LAMMPS dumps parsed on the GPU against what the reference's own reader read from the same files: positions
    bit for bit (double-rounded literals, > 2^53 mantissas, nan / inf through the host parser), boxes, steps, times."""
    import os

    from conftest import GOLDEN
    from mdcraft_b200.analysis.reader import LAMMPSDumpTrajectoryReader

    g = golden("dumps.npz")
    r = LAMMPSDumpTrajectoryReader(os.path.join(GOLDEN, name + ".lammpsdump"), dt=0.005, frame_block=2)
    assert (r.n_frames, r.n_atoms) == (5, 60)
    assert ["-" if c is None else c for c in r._conventions] == list(g[f"{tag}_conventions"])
    assert r._offsets == list(g[f"{tag}_offsets"])
    for f, ts in enumerate(r):
        np.testing.assert_array_equal(ts.positions, g[f"{tag}_positions"][f])
        np.testing.assert_array_equal(ts.dimensions, g[f"{tag}_dimensions"][f])
        assert ts.frame == f and ts.data["step"] == g[f"{tag}_steps"][f] and ts.time == g[f"{tag}_times"][f]
    np.testing.assert_array_equal(r.frames_block(0, 5), g[f"{tag}_positions"])
    if tag not in ("sparse", "triclinic"):        # the literals of frame 0 that the exact fast path refuses
        assert r.host_parsed_values() >= 5
    np.testing.assert_array_equal(r.frames_block(1, 5), g[f"{tag}_positions"][1:])
    assert r.host_parsed_values() <= 1            # %.17g, %e, %f and repr() output all take the device path
    np.testing.assert_array_equal(r[3].positions, g[f"{tag}_positions"][3])          # random access
    np.testing.assert_array_equal(r.frames_block_device(2, 4).cpu().numpy(), g[f"{tag}_positions"][2:4])
    r.close()


def test_dump_reader_feeds_the_analysis_classes(tmp_path):
    """A 3,000-atom dump with shuffled ids: the GPU reader against the CPU oracle, and g(r) / S(q) driven from the
    reader (per-frame protocol and device-resident blocks) against the same analyses on the in-memory trajectory."""
    from mdcraft_b200 import _capi, synthetic
    from mdcraft_b200.analysis.reader import LAMMPSDumpTrajectoryReader
    from oracle import oracle

    n, F, L = 3000, 7, 15.0
    rng = np.random.default_rng(12)
    pos = (rng.random((F, n, 3)) * L)
    path = tmp_path / "traj.lammpsdump"
    with open(path, "w") as fh:
        for f in range(F):
            fh.write(f"ITEM: TIMESTEP\n{100 * f}\nITEM: NUMBER OF ATOMS\n{n}\nITEM: BOX BOUNDS pp pp pp\n")
            fh.write(f"0 {L}\n0 {L}\n0 {L}\nITEM: ATOMS id type x y z\n")
            fmt = ("%d 1 %.8g %.8g %.8g\n", "%d 1 %.17g %.17g %.17g\n", "%d 1 %r %r %r\n", "%d 1 %.19e %.16e %.12f\n")[f % 4]
            for i in rng.permutation(n):
                fh.write(fmt % (i + 1, *(float(v) for v in pos[f, i])))
    want, dims, steps, _, _ = oracle.read_lammps_dump(str(path))
    r = LAMMPSDumpTrajectoryReader(str(path), parallel=True, frame_block=3)
    np.testing.assert_array_equal(r.frames_block(0, F), want)
    # 17-20 significant digits go through the checked double-double division on the device; only literals with more
    # than 19 digits (%.19e) and quotients within 2^-30 ulp of a rounding boundary are left to the host
    assert r.host_parsed_values() <= 2 * n * 3
    np.testing.assert_array_equal(r.frames_block(0, 3), want[:3])
    assert r.host_parsed_values() <= 2
    np.testing.assert_array_equal(np.stack([ts.positions for ts in r[1:6:2]]), want[1:6:2])
    u_file, u_mem = synthetic.Universe(r), universe_from(want, dims)
    calls = []
    real = r.frames_block_device
    r.frames_block_device = lambda lo, hi: (calls.append((lo, hi)), real(lo, hi))[1]
    # the analysis classes take the reader's device-resident blocks: no host staging, same results
    a = _rdf(u_file.atoms, n_bins=100, range_=(0.0, L / 2), exclusion=(1, 1))
    b = _rdf(u_mem.atoms, n_bins=100, range_=(0.0, L / 2), exclusion=(1, 1))
    assert calls == [(0, 7)] and r.ts.frame == F - 1
    np.testing.assert_array_equal(a.results.counts, b.results.counts)
    np.testing.assert_array_equal(a.results.rdf, b.results.rdf)
    np.testing.assert_array_equal(a.frames, np.arange(F))
    np.testing.assert_array_equal(a.times, 100.0 * np.arange(F))     # time = step * dt (reader.py:373)
    # two groups (cross g(r), partial S(q)), a strided slice of the trajectory, small frame blocks
    from mdcraft_b200.analysis.structure import RadialDistributionFunction, StructureFactor

    del calls[:]
    kw = dict(n_bins=64, range_=(0.5, 6.0), verbose=False, frame_block=2)
    a = RadialDistributionFunction(u_file.select(0, 1200), u_file.select(1200, n), **kw).run(start=1, stop=6, step=2)
    b = RadialDistributionFunction(u_mem.select(0, 1200), u_mem.select(1200, n), **kw).run(start=1, stop=6, step=2)
    assert calls == [(1, 2), (3, 4), (5, 6)]
    np.testing.assert_array_equal(a.results.counts, b.results.counts)
    np.testing.assert_array_equal(a.results.rdf, b.results.rdf)
    np.testing.assert_array_equal(a.frames, [1, 3, 5])
    del calls[:]
    kw = dict(mode="partial", n_points=8, verbose=False, frame_block=3)
    s1 = StructureFactor([u_file.select(0, 1200), u_file.select(1200, n)], **kw).run(start=2)
    s2 = StructureFactor([u_mem.select(0, 1200), u_mem.select(1200, n)], **kw).run(start=2)
    assert calls == [(2, 5), (5, 7)]
    np.testing.assert_array_equal(s1.results.ssf, s2.results.ssf)
    s1 = _sq(u_file.atoms, n_points=8).results.ssf
    s2 = _sq(u_mem.atoms, n_points=8).results.ssf
    np.testing.assert_array_equal(s1, s2)
    # both analyses from ONE pass over the file, blocks parsed once and consumed by each on the device
    from mdcraft_b200.analysis.base import run_together

    del calls[:]
    ga = RadialDistributionFunction(u_file.atoms, n_bins=100, range_=(0.0, L / 2), exclusion=(1, 1), verbose=False, frame_block=4)
    sa = StructureFactor(u_file.atoms, n_points=8, verbose=False, frame_block=3)
    run_together([ga, sa], start=1)
    assert calls == [(1, 4), (4, 7)]
    gb = RadialDistributionFunction(u_mem.atoms, n_bins=100, range_=(0.0, L / 2), exclusion=(1, 1), verbose=False).run(start=1)
    sb = StructureFactor(u_mem.atoms, n_points=8, verbose=False).run(start=1)
    np.testing.assert_array_equal(ga.results.counts, gb.results.counts)
    np.testing.assert_array_equal(ga.results.rdf, gb.results.rdf)
    np.testing.assert_allclose(sa.results.ssf, sb.results.ssf, rtol=1e-13)   # frames summed in other block sizes
    np.testing.assert_array_equal(ga.times, gb.times * 100.0)
    # residues groupings fall back to the per-frame protocol (centres of mass are formed on the host)
    del calls[:]
    c = _rdf(u_file.atoms, groupings="residues", n_bins=100, range_=(0.0, L / 2), exclusion=(1, 1))
    assert calls == []
    np.testing.assert_array_equal(c.results.counts, b.results.counts if False else _rdf(
        u_mem.atoms, groupings="residues", n_bins=100, range_=(0.0, L / 2), exclusion=(1, 1)).results.counts)
    # device-resident blocks straight into a plan through the C ABI
    ctx = _capi.context(0)
    plan = _capi.RdfPlan(ctx, 100, (0.0, L / 2), n, None, (1, 1))
    plan.push(real(0, 4), None, dims[:4], n_frames=4)
    plan.push(real(4, F), None, dims[4:], n_frames=F - 4)
    counts, _ = plan.read()
    np.testing.assert_array_equal(counts, _rdf(u_mem.atoms, n_bins=100, range_=(0.0, L / 2), exclusion=(1, 1)).results.counts)
    plan.close()
    r.close()


def test_triclinic_dump_to_rdf(tmp_path):
    """A LAMMPS dump with restricted triclinic boxes (xy xz yz) that change from frame to frame: GPU reader -> device
    blocks -> g(r) (filter kernel on the fractional coordinates) against the CPU oracle on what the oracle's reader
    read from the same file, counts bit for bit; and the reference's rdf normalisation through the classes."""
    from mdcraft_b200 import synthetic
    from mdcraft_b200.analysis.reader import LAMMPSDumpTrajectoryReader
    from oracle import oracle

    n, F = 4500, 3
    rng = np.random.default_rng(31)
    path = tmp_path / "tri.lammpsdump"
    with open(path, "w") as fh:
        for f in range(F):
            lx, ly, lz = 17.0 + f, 16.0 + 0.5 * f, 18.0
            xy, xz, yz = 4.0 - f, -3.0, 2.5 + 0.25 * f
            org = rng.normal(size=3)
            blo = org + [min(0, xy, xz, xy + xz), min(0, yz), 0]
            bhi = org + [lx, ly, lz] + np.array([max(0, xy, xz, xy + xz), max(0, yz), 0])
            fh.write(f"ITEM: TIMESTEP\n{10 * f}\nITEM: NUMBER OF ATOMS\n{n}\nITEM: BOX BOUNDS xy xz yz pp pp pp\n")
            fh.write("%r %r %r\n%r %r %r\n%r %r %r\n" % (float(blo[0]), float(bhi[0]), xy, float(blo[1]), float(bhi[1]), xz,
                                                      float(blo[2]), float(bhi[2]), yz))
            fh.write("ITEM: ATOMS id x y z\n")
            cell = np.array([[lx, 0, 0], [xy, ly, 0], [xz, yz, lz]])
            pos = org + rng.random((n, 3)) @ cell
            for i in rng.permutation(n):
                fh.write("%d %.9g %.9g %.9g\n" % (i + 1, *pos[i]))
    want_pos, dims, _, _, _ = oracle.read_lammps_dump(str(path))
    r = LAMMPSDumpTrajectoryReader(str(path), frame_block=2)
    np.testing.assert_array_equal(r.frames_block(0, F), want_pos)
    np.testing.assert_array_equal(np.stack([ts.dimensions for ts in r]), dims)
    assert not np.any(dims[:, 3:] == 90.0)
    u = synthetic.Universe(r)
    window = (0.0, 7.0)                       # below half the smallest width of every frame's cell
    a = _rdf(u.atoms, n_bins=140, range_=window, exclusion=(1, 1))
    want = sum(oracle.radial_histogram(want_pos[f], want_pos[f], 140, window, dims[f], exclusion=(1, 1)) for f in range(F))
    np.testing.assert_array_equal(a.results.counts, want)
    b = _rdf(universe_from(want_pos, dims).atoms, n_bins=140, range_=window, exclusion=(1, 1))
    np.testing.assert_array_equal(a.results.rdf, b.results.rdf)
    r.close()


def test_dump_reader_refuses_what_it_cannot_match(tmp_path):
    from mdcraft_b200.analysis.reader import LAMMPSDumpTrajectoryReader

    head = "ITEM: TIMESTEP\n0\nITEM: NUMBER OF ATOMS\n2\nITEM: BOX BOUNDS pp pp pp\n0 1\n0 1\n0 1\n"
    p = tmp_path / "s.lammpsdump"
    p.write_text(head + "ITEM: ATOMS id xs ys zs\n1 0.1 0.2 0.3\n2 0.4 0.5 0.6\n")
    with pytest.raises(NotImplementedError):
        LAMMPSDumpTrajectoryReader(str(p))
    p.write_text(head + "ITEM: ATOMS id x y z\n1 0.1 0.2 0.3\n2 0.4 0.5\n")
    with pytest.raises(Exception, match="fewer columns"):
        LAMMPSDumpTrajectoryReader(str(p))
    with pytest.raises(ValueError, match="Invalid convention"):
        LAMMPSDumpTrajectoryReader(str(p), conventions="q")
    with pytest.raises(ValueError, match="length 3"):
        LAMMPSDumpTrajectoryReader(str(p), conventions=["", ""])
    with pytest.raises(NotImplementedError):
        LAMMPSDumpTrajectoryReader(str(p), unwrap=True)


# --------------------------------------------------------------------------- density profiles (§8f rank 4)


def test_density_profile_golden():
    """DensityProfile on the GPU against the reference class's own results: number and charge densities are EQUAL
    (exact counts, same normalisation expressions), post-processing to rounding."""
    import warnings

    from mdcraft_b200.analysis.profile import DensityProfile

    g = golden("profiles.npz")
    pos, dims = g["pos"], g["dims"]
    n = pos.shape[1]
    u = universe_from(pos, dims)
    d = DensityProfile([u.select(0, 150), u.select(150, n)], axes="xz", n_bins=[5, 48], charges=[1.0, -0.6],
                       reduced=True, verbose=False, frame_block=4).run()
    for ax in "xz":
        np.testing.assert_array_equal(d.results.bins[ax], g[f"a_bins_{ax}"])
        np.testing.assert_array_equal(d.results.bin_edges[ax], g[f"a_edges_{ax}"])
        np.testing.assert_array_equal(d.results.number_densities[ax], g[f"a_number_{ax}"])
        np.testing.assert_array_equal(d.results.charge_densities[ax], g[f"a_charge_{ax}"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        d.calculate_surface_charge_densities(None, [1.5, 2.5], dVs=[0.1, 0.3])
        np.testing.assert_allclose(d.results.surface_charge_densities, g["a_sigma"], rtol=1e-13, equal_nan=True)
        d.calculate_potential_profiles(None, [1.5, 2.5], sigmas_q=[0.02, 0.01])
        np.testing.assert_allclose(d.results.potentials["z"], g["a_potential_integral"], rtol=1e-12, atol=1e-14)
        d.calculate_potential_profiles(None, [1.5, 2.5], sigmas_q=[0.02, 0.01], methods="matrix", V0s=[0.0, 0.2])
        np.testing.assert_allclose(d.results.potentials["z"], g["a_potential_matrix"], rtol=1e-10, atol=1e-12)
    d = DensityProfile(u.atoms, average=False, verbose=False, frame_block=2).run(start=1, stop=6, step=2)
    for ax in "xyz":
        np.testing.assert_array_equal(d.results.number_densities[ax], g[f"b_number_{ax}"])
    np.testing.assert_array_equal(d.results.times, g["b_times"])
    u2 = universe_from(pos, dims, resindices=np.arange(n) // 2, masses=1.0 + (np.arange(n) % 2) * 2.5)
    d = DensityProfile(u2.atoms, "residues", axes="y", n_bins=33, dimensions=[9.5, 11.25, 14.0],
                       dim_scales=(1.0, 0.9, 1.0), verbose=False).run()
    np.testing.assert_array_equal(d.results.bin_edges["y"], g["c_edges_y"])
    np.testing.assert_array_equal(d.results.number_densities["y"], g["c_number_y"])
    # recentring (host unwrap + centre-of-mass shift in the reference's NumPy expressions, float64 positions to the device)
    u3 = universe_from(g["pos_drift"], dims, masses=1.0 + (np.arange(n) % 3))
    groups = [u3.select(0, 150), u3.select(150, n)]
    d = DensityProfile(groups, axes="z", n_bins=40, recenter=0, verbose=False, frame_block=4).run()
    np.testing.assert_array_equal(d.results.number_densities["z"], g["d_number_z"])
    d = DensityProfile(groups, axes="xz", n_bins=[6, 40], recenter=([0, 1], [np.nan, np.nan, 3.5]), average=False,
                       verbose=False, frame_block=2).run(start=1)
    np.testing.assert_array_equal(d.results.number_densities["x"], g["e_number_x"])
    np.testing.assert_array_equal(d.results.number_densities["z"], g["e_number_z"])
    with pytest.raises(ValueError, match="Invalid group index"):
        DensityProfile(u.atoms, recenter=3)
    with pytest.raises(ValueError, match="Invalid axis"):
        DensityProfile(u.atoms, axes="xw")


def test_density_profile_large_vs_oracle(ctx):
    """100,000 particles, three axes, 1,000 bins: the device counts against the CPU restatement, exactly."""
    from mdcraft_b200 import _capi, synthetic
    from oracle import oracle

    n, L = 100_000, 37.5
    rng = np.random.default_rng(9)
    pos = (rng.normal(L / 2, L, size=(2, n, 3))).astype(np.float32)      # most particles outside the box
    dims = np.array([L, L * 0.8, L * 1.3], np.float32)
    plan = _capi.ProfilePlan(ctx, [30_000, 70_000], [0, 1, 2], [1000, 7, 333], dims.astype(np.float64))
    per = plan.push(pos, per_frame=True)
    summed, nf = plan.read()
    plan.close()
    want = oracle.density_profile_counts(pos, [30_000, 70_000], [0, 1, 2], [1000, 7, 333], dims)
    assert nf == 2
    for got, tot, w in zip(per, summed, want):
        np.testing.assert_array_equal(got, w)
        np.testing.assert_array_equal(tot, w.sum(axis=0))
        assert tot.sum() == 2 * n                                         # wrapping loses nobody


# --------------------------------------------------------------------------- NetCDF ingest (§8f rank 4)


def test_netcdf_reader_vs_scipy(tmp_path):
    """AMBER NetCDF trajectories: coordinates byte-swapped on the GPU against SciPy's independent NetCDF-3 reader
    (bit for bit), frame selections as the reference indexes them, and g(r) driven from device-resident blocks."""
    from mdcraft_b200 import synthetic
    from mdcraft_b200.openmm.file import NetCDFFile, NetCDFTrajectory
    from oracle import oracle

    rng = np.random.default_rng(31)
    F, N, L = 9, 2500, 14.0
    pos = (rng.random((F, N, 3)) * L).astype(np.float32)
    pos[0, :4, 0] = [np.float32("nan"), np.float32("inf"), -0.0, 1e-40]       # bit patterns survive the swap
    times = 2.0 * np.arange(F)
    lengths, angles = np.full((F, 3), L), np.full((F, 3), 90.0)
    for version, vel in ((2, True), (1, False)):
        path = str(tmp_path / f"traj{version}.nc")
        oracle.write_amber_netcdf(path, pos, times, lengths, angles, version=version, velocities=pos + 1 if vel else None,
                                  forces=-2 * pos if vel else None)
        want, want_t, want_l, want_a = oracle.read_amber_netcdf(path)
        f = NetCDFFile(path)
        assert (f.get_num_frames(), f.get_num_atoms()) == (F, N)
        np.testing.assert_array_equal(f.get_positions(units=False).view(np.uint32), want.view(np.uint32))
        np.testing.assert_array_equal(f.get_positions(3, units=False).view(np.uint32), want[3].view(np.uint32))
        np.testing.assert_array_equal(f.get_positions([1, 2, 5], units=False)[:, 10:], want[[1, 2, 5]][:, 10:])
        np.testing.assert_array_equal(f.get_positions(slice(2, 9, 3), units=False)[:, 10:], want[2:9:3][:, 10:])
        np.testing.assert_array_equal(f.get_positions(slice(4, 7), units=False, device=True).cpu().numpy()[:, 10:], want[4:7, 10:])
        np.testing.assert_array_equal(f.get_times(units=False), want_t)
        got_l, got_a = f.get_dimensions(units=False)
        np.testing.assert_array_equal(got_l, want_l)
        np.testing.assert_array_equal(got_a, want_a)
        np.testing.assert_array_equal(f.get_dimensions(-1, units=False)[0], want_l[-1])
        if vel:   # the other per-atom record variables (file.py:217-306) take the same path
            np.testing.assert_array_equal(f.get_velocities(units=False)[:, 10:], (pos + 1)[:, 10:])
            np.testing.assert_array_equal(f.get_forces([0, 8], units=False)[:, 10:], (-2 * pos)[[0, 8]][:, 10:])
            np.testing.assert_array_equal(f.get_forces(5, units=False, device=True).cpu().numpy(), (-2 * pos)[5])
        else:
            with pytest.warns(UserWarning, match="velocities"):
                assert f.get_velocities(units=False) is None
            with pytest.warns(UserWarning, match="forces"):
                assert f.get_forces(units=False) is None
        f.close()
    pos[0, :4, 0] = 1.0
    path = str(tmp_path / "clean.nc")
    oracle.write_amber_netcdf(path, pos, times, lengths, angles)
    traj = NetCDFTrajectory(path, frame_block=4)
    u_file = synthetic.Universe(traj)
    u_mem = universe_from(pos, np.array([L, L, L, 90, 90, 90], np.float32))
    a = _rdf(u_file.atoms, n_bins=80, range_=(0.0, L / 2), exclusion=(1, 1), frame_block=4)
    b = _rdf(u_mem.atoms, n_bins=80, range_=(0.0, L / 2), exclusion=(1, 1))
    np.testing.assert_array_equal(a.results.counts, b.results.counts)
    np.testing.assert_array_equal(a.times, times)
    np.testing.assert_array_equal(np.stack([ts.positions for ts in traj[2:7:2]]), pos[2:7:2])
    traj.close()
