"""
CPU: the oracle (oracle/oracle.c + oracle/oracle.py) against the golden vectors
produced by the REAL reference (tests/golden/make_golden.py).  This is what pins
the oracle; the GPU parity tests then compare the CUDA path with the oracle.
"""

import numpy as np
import pytest

from conftest import golden
from oracle import oracle


def test_dfts_matches_reference_kernel():
    g = golden("sq_dfts.npz")
    rho = oracle.delta_fourier_transform_sum(g["qs"], g["rs"])
    np.testing.assert_allclose(rho, g["rho"], rtol=0, atol=1e-10)


def test_sq_total_c1():
    g = golden("sq_total_c1.npz")
    out = oracle.structure_factor(list(g["pos"]), [g["pos"].shape[1]], g["dims"][:3], n_points=int(g["n_points"]))
    np.testing.assert_array_equal(out["wavenumbers"], g["wavenumbers"])
    np.testing.assert_allclose(out["ssf"], g["ssf"], rtol=1e-10, atol=1e-12)


def test_sq_partial_surfaces():
    g = golden("sq_partial_surfaces.npz")
    out = oracle.structure_factor(
        list(g["pos"]), list(g["sizes"]), g["dims"][:3], mode="partial", n_points=int(g["n_points"]),
        q_max=float(g["q_max"]), n_surfaces=int(g["n_surfaces"]), n_surface_points=int(g["n_surface_points"]),
    )
    assert tuple(map(tuple, g["pairs"])) == out["pairs"]
    np.testing.assert_array_equal(out["wavevectors"], g["wavevectors"])
    np.testing.assert_array_equal(out["wavenumbers"], g["wavenumbers"])
    np.testing.assert_allclose(out["ssf"], g["ssf"], rtol=1e-10, atol=1e-12)
    # sum rule (structure.py:1554-1560): sum of partials == total
    tot = oracle.structure_factor(
        list(g["pos"]), [int(g["sizes"].sum())], g["dims"][:3], n_points=int(g["n_points"]), q_max=float(g["q_max"]),
        n_surfaces=int(g["n_surfaces"]), n_surface_points=int(g["n_surface_points"]),
    )
    np.testing.assert_allclose(out["ssf"].sum(axis=0), tot["ssf"][0], rtol=1e-9, atol=1e-10)


def test_sq_pair_ortho_and_trig_form():
    g = golden("sq_pair_ortho.npz")
    kw = dict(mode="pair", n_points=int(g["n_points"]))
    out = oracle.structure_factor(list(g["pos"]), list(g["sizes"]), g["dims"][:3], sort=False, unique=False, **kw)
    np.testing.assert_array_equal(out["wavevectors"], g["wavevectors"])
    np.testing.assert_allclose(out["ssf"], g["ssf"], rtol=1e-10, atol=1e-11)
    out2 = oracle.structure_factor(list(g["pos"]), list(g["sizes"]), g["dims"][:3], **kw)
    np.testing.assert_array_equal(out2["wavenumbers"], g["wavenumbers_unique"])
    # exp form == trig form (the reference's two formulations)
    np.testing.assert_allclose(out2["ssf"], g["ssf_trig_unique"], rtol=1e-9, atol=1e-10)


def test_sq_explicit_wavevectors():
    g = golden("sq_explicit.npz")
    out = oracle.structure_factor(
        list(g["pos"]), [g["pos"].shape[1]], g["dims"][:3], wavevectors=g["wavevectors"], sort=False, unique=False
    )
    np.testing.assert_allclose(out["ssf"], g["ssf"], rtol=1e-10, atol=1e-12)


def test_isf_against_reference_class():
    """IntermediateScatteringFunction (structure.py:2198-2835): total with the self part, and partial."""
    g = golden("isf_small.npz")
    frames = list(g["pos"])
    n = g["pos"].shape[1]
    out = oracle.isf(frames, [n], g["dims"][:3], n_points=6, n_lags=4, incoherent=True)
    np.testing.assert_array_equal(out["wavenumbers"], g["wavenumbers"])
    np.testing.assert_allclose(out["cisf"], g["cisf"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(out["iisf"], g["iisf"], rtol=1e-9, atol=1e-11)
    # lag 0 of the self part is exactly 1 (every particle against itself), and cisf[0] is S(q)
    np.testing.assert_allclose(out["iisf"][0], 1.0, rtol=1e-12)
    sq = oracle.structure_factor(frames, [n], g["dims"][:3], n_points=6)
    np.testing.assert_allclose(out["cisf"][0, 0], sq["ssf"][0], rtol=1e-9, atol=1e-11)
    part = oracle.isf(frames, list(g["sizes"]), g["dims"][:3], mode="partial", n_points=5, n_lags=3, incoherent=True,
                      sort=False, unique=False)
    assert part["pairs"] == tuple(map(tuple, g["pairs_partial"]))
    np.testing.assert_allclose(part["cisf"], g["cisf_partial"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(part["iisf"], g["iisf_partial"], rtol=1e-9, atol=1e-11)


def test_scsf_against_reference_class():
    """polymer.SingleChainStructureFactor (polymer.py:1054-1445), one group."""
    g = golden("scsf_small.npz")
    out = oracle.scsf(list(g["pos"]), int(g["n_chains"]), int(g["n_monomers"]), g["dims"][:3], n_points=int(g["n_points"]))
    np.testing.assert_array_equal(out["wavenumbers"], g["wavenumbers"])
    np.testing.assert_allclose(out["scsf"], g["scsf"], rtol=1e-10, atol=1e-12)
    assert out["scsf"][0, 0] == pytest.approx(int(g["n_monomers"]))   # q = 0: N monomers scatter coherently


def test_histogram_formula_and_edges():
    g = golden("histogram_pins.npz")
    for c in range(int(g["n_cases"])):
        r0, r1, n = g[f"case{c}_args"]
        n = int(n)
        edges = oracle.histogram_bin_edges((r0, r1), n)
        np.testing.assert_array_equal(edges, g[f"case{c}_edges"])
        vals, want = g[f"case{c}_values"], g[f"case{c}_bins"]
        got = np.empty_like(want)
        for i, x in enumerate(vals):
            nz = np.nonzero(oracle.histogram(np.array([x]), n, edges))[0]
            got[i] = nz[0] if len(nz) else -1
        np.testing.assert_array_equal(got, want)
    # the reference itself is position dependent one ulp below an edge (oracle.c caveat)
    pd = g["position_dependence"]
    assert (pd[:, 1] + pd[:, 2] == pd[:, 0]).all()
    assert (pd[pd[:, 0] >= 4, 2] > 0).all() and (pd[pd[:, 0] < 4, 2] == 0).all()


def test_rdf_self_c1():
    g = golden("rdf_self_c1.npz")
    frames = list(g["pos"])
    dims = [g["dims"]] * len(frames)
    out = oracle.rdf(frames, None, dims, int(g["n_bins"]), tuple(g["range_"]))
    np.testing.assert_array_equal(out["counts"], g["counts"])
    np.testing.assert_array_equal(out["bin_edges"], g["bin_edges"])
    np.testing.assert_array_equal(out["bins"], g["bins"])
    np.testing.assert_allclose(out["rdf"], g["rdf"], rtol=1e-12)
    # invariants: self pairs in bin 0, everything else even
    assert out["counts"][0] >= len(frames) * frames[0].shape[0]
    assert ((out["counts"] - np.eye(1, len(out["counts"]), 0, dtype=np.int64)[0] * len(frames) * frames[0].shape[0]) % 2 == 0).all()
    ex = oracle.rdf(frames, None, dims, int(g["n_bins"]), tuple(g["range_"]), exclusion=(1, 1))
    np.testing.assert_array_equal(ex["counts"], g["counts_excl"])
    np.testing.assert_allclose(ex["rdf"], g["rdf_excl"], rtol=1e-12)


def test_rdf_cross_exclusion_and_norms():
    g = golden("rdf_cross_excl.npz")
    na = int(g["sizes"][0])
    fa = [p[:na] for p in g["pos"]]
    fb = [p[na:] for p in g["pos"]]
    dims = [g["dims"]] * len(fa)
    for norm in ("rdf", "density", None):
        out = oracle.rdf(fa, fb, dims, int(g["n_bins"]), tuple(g["range_"]), exclusion=tuple(g["exclusion"]), norm=norm)
        np.testing.assert_array_equal(out["counts"], g["counts"])
        np.testing.assert_array_equal(out["bin_edges"], g["bin_edges"])
        np.testing.assert_allclose(out["rdf"], g[f"rdf_{norm}"], rtol=1e-12)


def test_rdf_npt_and_drop_axis():
    g = golden("rdf_npt_dropaxis.npz")
    frames = list(g["pos"])
    out = oracle.rdf(frames, None, list(g["dims"]), int(g["n_bins"]), tuple(g["range_"]))
    np.testing.assert_array_equal(out["counts"], g["counts"])
    np.testing.assert_allclose(out["rdf"], g["rdf"], rtol=1e-12)
    out = oracle.rdf(frames, None, list(g["dims"]), int(g["n_bins"]), tuple(g["range_"]), drop_axis=2)
    np.testing.assert_array_equal(out["counts"], g["counts_drop"])
    np.testing.assert_allclose(out["rdf"], g["rdf_drop"], rtol=1e-12)


def test_rdf_triclinic_cells():
    """Triclinic cells (60/90/75, 50/65/55, NPT mixing orthorhombic and triclinic frames): the C restatement of
    MDAnalysis' triclinic routine under oracle.rdf against the reference classes driven over the NumPy restatement
    (tests/golden/make_golden.py triclinic): two independent statements of triclinic_vectors / _triclinic_pbc /
    minimum_image_triclinic, the reference's own binning, exclusion, ts.volume and normalisation."""
    g = golden("rdf_triclinic.npz")
    F = g["a_pos"].shape[0]
    out = oracle.rdf(list(g["a_pos"]), None, [g["a_dims"]] * F, 80, (0.0, 4.5))
    np.testing.assert_array_equal(out["counts"], g["a_counts"])
    np.testing.assert_allclose(out["rdf"], g["a_rdf"], rtol=1e-12)
    out = oracle.rdf(list(g["a_pos"][:, :180]), list(g["a_pos"][:, 180:]), [g["a_dims"]] * F, 57, (0.4, 4.9), exclusion=(3, 5))
    np.testing.assert_array_equal(out["counts"], g["a_cross_counts"])
    np.testing.assert_allclose(out["rdf"], g["a_cross_rdf"], rtol=1e-12)
    out = oracle.rdf(list(g["b_pos"]), None, list(g["b_dims"]), 64, (0.0, 3.9), exclusion=(1, 1))
    np.testing.assert_array_equal(out["counts"], g["b_counts"])
    np.testing.assert_allclose(out["rdf"], g["b_rdf"], rtol=1e-12)
    # the minimum image is a true minimum: against a search over 5^3 images in float64 on the wrapped coordinates
    box = g["a_dims"]
    m = oracle.triclinic_vectors(box).astype(np.float64)
    w = oracle._triclinic_wrap(g["a_pos"][0, :60], oracle.triclinic_vectors(box))
    dw = (w[None] - w[:, None]).astype(np.float64)             # float32 subtraction, as in the reference routine
    best = np.full((60, 60), np.inf)
    for i in range(-2, 3):
        for j in range(-2, 3):
            for k in range(-2, 3):
                d = dw + (i * m[0] + j * m[1] + k * m[2])
                best = np.minimum(best, (d * d).sum(-1))
    got = np.array([[oracle.pair_distance(a, b, box) for b in g["a_pos"][0, :60]] for a in g["a_pos"][0, :60]])
    np.testing.assert_allclose(got, np.sqrt(best), rtol=0, atol=1e-12)


def test_radial_histogram_like_the_reference_test():
    # /root/reference/tests/test_analysis_structure.py:21-46
    g = golden("rdf_radial_histogram.npz")
    h = oracle.radial_histogram(g["origin"], g["neighbors"], 10, (0, 11), g["dims"])
    np.testing.assert_array_equal(h, g["counts"])
    d = np.linalg.norm(g["neighbors"].astype(np.float64) - g["origin"].astype(np.float64), axis=1)
    np.testing.assert_array_equal(h, np.histogram(d, bins=10, range=(0, 11))[0])


def test_pair_distance_minimum_image():
    box = np.array([10.0, 12.0, 8.0], np.float32)
    a = np.array([0.5, 0.5, 0.5], np.float32)
    b = np.array([9.5, 11.5, 7.5], np.float32)
    assert oracle.pair_distance(a, b, box) == pytest.approx(np.sqrt(3.0), rel=1e-6)
    assert oracle.pair_distance(a, a, box) == 0.0
    # antisymmetry of the float32 subtraction makes the distance symmetric bit for bit
    rng = np.random.default_rng(0)
    for _ in range(100):
        p, q = (rng.random(3) * box).astype(np.float32), (rng.random(3) * box).astype(np.float32)
        assert oracle.pair_distance(p, q, box) == oracle.pair_distance(q, p, box)


@pytest.mark.parametrize("tag,name", [("ids", "dump_ids"), ("noid", "dump_noid"), ("sparse", "dump_sparse_ids"),
                                      ("triclinic", "dump_triclinic")])
def test_dump_reader_oracle_matches_the_reference_reader(tag, name):
    """oracle.read_lammps_dump against what the reference's own LAMMPSDumpTrajectoryReader read from the same
    files (tests/golden/make_golden.py dumps): positions bit for bit (NaN == NaN), boxes, steps, times, offsets."""
    import os

    from conftest import GOLDEN as GOLDEN_DIR

    g = golden("dumps.npz")
    pos, dims, steps, times, offsets = oracle.read_lammps_dump(os.path.join(GOLDEN_DIR, name + ".lammpsdump"), dt=0.005)
    np.testing.assert_array_equal(pos, g[f"{tag}_positions"])
    np.testing.assert_array_equal(dims, g[f"{tag}_dimensions"])
    np.testing.assert_array_equal(steps, g[f"{tag}_steps"])
    np.testing.assert_array_equal(times, g[f"{tag}_times"])
    np.testing.assert_array_equal(offsets, g[f"{tag}_offsets"])


def test_density_profile_oracle_matches_the_reference_class():
    """oracle.density_profile_counts, normalised as profile.py:1236-1248, against number densities the reference's
    own DensityProfile produced (tests/golden/make_golden.py profiles): positions outside the box, on its edges
    and at exact multiples of it included."""
    g = golden("profiles.npz")
    pos, dims = g["pos"], g["dims"][:3]          # float32, as the class holds them
    F, N = pos.shape[:2]
    volume = np.prod(dims)
    cx, cz = oracle.density_profile_counts(pos, [150, N - 150], [0, 2], [5, 48], dims)
    np.testing.assert_array_equal(cx.sum(axis=0) / (volume / 5 * F), g["a_number_x"])
    np.testing.assert_array_equal(cz.sum(axis=0) / (volume / 48 * F), g["a_number_z"])
    np.testing.assert_array_equal(oracle.histogram_bin_edges((0.0, float(dims[2])), 48), g["a_edges_z"])
    sel = pos[1:6:2]
    for k, ax in enumerate("xyz"):
        (c,) = oracle.density_profile_counts(sel, [N], [k], [201], dims)
        np.testing.assert_array_equal(c.transpose(1, 0, 2) / (volume / np.int64(201)), g[f"b_number_{ax}"])
