"""CPU: host-side logic of the drop-in classes, the C-ABI surface and the no-fallback rule."""

import ctypes
import os

import numpy as np
import pytest

from conftest import golden, universe_from
from mdcraft_b200 import _capi, algorithm, build, distributed
from mdcraft_b200.analysis import structure


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _capi.declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/mdcraft_b200.h but not exported"
    # and the binding covers exactly the header
    assert sorted(_capi.SIGNATURES) == declared
    assert _capi.lib().mdc_version() >= 100


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, never compute on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_capi.MdcError) as err:
        _capi.Context(0)
    assert "no CPU fallback" in str(err.value) or "CUDA" in str(err.value)
    g = golden("rdf_self_c1.npz")
    u = universe_from(g["pos"], g["dims"])
    with pytest.raises(_capi.MdcError):
        structure.RadialDistributionFunction(u.atoms, n_bins=10, range_=(0.0, 3.0), verbose=False).run()
    with pytest.raises(_capi.MdcError):
        structure.StructureFactor(u.atoms, n_points=4, verbose=False)
    # the widened rows too: transforms on a device, density profiles, the readers
    with pytest.raises(_capi.MdcError):
        structure.radial_fourier_transform(np.linspace(0.1, 1, 5), np.ones(5), np.ones(3), device=0)
    from mdcraft_b200.analysis.profile import DensityProfile
    from mdcraft_b200.analysis.reader import LAMMPSDumpTrajectoryReader

    with pytest.raises(_capi.MdcError):
        DensityProfile(u.atoms, n_bins=8, verbose=False).run()
    with pytest.raises(_capi.MdcError):
        LAMMPSDumpTrajectoryReader(os.path.join(os.path.dirname(__file__), "golden", "dump_ids.lammpsdump"))
    # nothing under mdcraft_b200/ may import the oracle
    root = os.path.dirname(os.path.abspath(build.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f


def test_constructor_validation_matches_reference_messages():
    g = golden("rdf_self_c1.npz")
    u = universe_from(g["pos"][:1], g["dims"])
    a = u.atoms
    with pytest.raises(ValueError, match="Invalid grouping 'molecules'"):
        structure.RadialDistributionFunction(a, groupings="molecules")
    with pytest.raises(ValueError, match="must be a string or an array with length 2"):
        structure.RadialDistributionFunction(a, groupings=("atoms",))
    with pytest.raises(ValueError, match="Invalid axis in 'drop_axis'"):
        structure.RadialDistributionFunction(a, drop_axis=3)
    assert structure.RadialDistributionFunction(a, drop_axis="z")._drop_axis == 2
    with pytest.raises(ValueError, match="Invalid grouping 'segments'"):
        structure.StructureFactor(a, groupings="segments")
    with pytest.raises(ValueError, match="exactly one or two atom groups"):
        structure.StructureFactor([u.select(0, 10), u.select(10, 20), u.select(20, 30)], mode="pair")
    with pytest.raises(ValueError, match="must contain all atoms"):
        structure.StructureFactor(u.select(0, 10))
    with pytest.raises(ValueError, match="'dimensions' must have length 3"):
        structure.StructureFactor(a, dimensions=[1.0, 2.0])
    u2 = universe_from(g["pos"][:1], None)
    with pytest.raises(ValueError, match="does not contain system dimension"):
        structure.RadialDistributionFunction(u2.atoms)
    with pytest.raises(ValueError, match="System dimensions were not found"):
        structure.StructureFactor(u2.atoms)


def test_bin_edges_match_reference():
    g = golden("histogram_pins.npz")
    for c in range(int(g["n_cases"])):
        r0, r1, n = g[f"case{c}_args"]
        np.testing.assert_array_equal(algorithm.histogram_bin_edges((r0, r1), int(n)), g[f"case{c}_edges"])


def test_get_closest_factors():
    assert tuple(algorithm.get_closest_factors(8, 2, reverse=True)) == (4, 2)
    assert tuple(algorithm.get_closest_factors(6, 2, reverse=True)) == (3, 2)
    assert tuple(algorithm.get_closest_factors(16, 2)) == (4, 4)
    assert tuple(algorithm.get_closest_factors(72, 2, reverse=True)) == (12, 6)
    assert tuple(algorithm.get_closest_factors(7, 2, reverse=True)) == (7, 1)
    for v in (6, 8, 12, 30, 72, 97, 128):
        assert int(np.prod(algorithm.get_closest_factors(v, 2))) == v
        assert int(np.prod(algorithm.get_closest_factors(v, 3))) == v


def test_strip_unit_plain_values():
    assert algorithm.strip_unit(2.5, "Å") == (2.5, "Å")
    assert algorithm.strip_unit([1, 2, 3], "Å")[0] == [1, 2, 3]
    assert algorithm.is_unitless(300.0)
    assert algorithm.strip_unit("12.5", "K")[0] == 12.5


def test_combine_equal_wavenumbers_is_the_reference_loop():
    rng = np.random.default_rng(3)
    for n_points, L in ((12, 10.77), (24, 50.0)):
        grid = 2 * np.pi * np.arange(n_points) / L
        wv = np.stack(np.meshgrid(grid, grid, grid), -1).reshape(-1, 3)
        wn = np.linalg.norm(wv, axis=1)
        ssf = rng.random((3, wn.shape[0]))
        uq = np.unique(wn.round(11))
        want = np.hstack([ssf[:, np.isclose(q, wn)].mean(axis=1, keepdims=True) for q in uq])
        got = structure.combine_equal_wavenumbers(ssf, wn, uq)
        np.testing.assert_allclose(got, want, rtol=1e-13, atol=0)
    # overlapping isclose windows (large |n|): membership must still be identical
    wn = np.sort(np.concatenate([1000.0 + np.arange(50) * 4e-3, [1000.1 + 1e-9, 1000.1 - 1e-9]]))
    uq = np.unique(wn.round(11))
    ssf = rng.random((2, wn.shape[0]))
    want = np.hstack([ssf[:, np.isclose(q, wn)].mean(axis=1, keepdims=True) for q in uq])
    np.testing.assert_allclose(structure.combine_equal_wavenumbers(ssf, wn, uq), want, rtol=1e-13, atol=0)


def test_frame_block_partition():
    for n in (0, 1, 7, 8, 100, 1001):
        for world in (1, 2, 3, 4, 8):
            blocks = [distributed.frame_block(n, r, world) for r in range(world)]
            covered = [i for lo, hi in blocks for i in range(lo, hi)]
            assert covered == list(range(n))
            assert max(hi - lo for lo, hi in blocks) <= -(-n // world) if n else True
    # forward and BACKWARD runs (a backward run that ends on frame 0 has stop=None, not -1), every rank's triple read
    # the way run() reads it: through slice.indices of the trajectory length
    for n in (100, 10):
        for start, stop, step in ((None, None, None), (3, 97, 4), (10, 11, 1), (5, 5, 1), (0, 100, 7), (None, None, -1),
                                  (None, None, -3), (9, None, -1), (8, 2, -2), (4, None, -1)):
            want = list(range(*slice(start, stop, step).indices(n)))
            for world in (1, 2, 3, 4, 8):
                got = []
                for r in range(world):
                    s0, s1, st = distributed.shard_slice(n, start, stop, step, rank=r, world=world)
                    got += list(range(*slice(s0, s1, st).indices(n)))
                assert got == want, (n, start, stop, step, world)


def test_post_processing_functions():
    # analytic check used by the reference's own test (tests/test_analysis_structure.py:48-59):
    # the radial Fourier transform of exp(-alpha r) / r is 4 pi / (alpha^2 + q^2)
    alpha = 3.0
    r = np.linspace(1e-8, 20, 20001)
    q = np.linspace(0.5, 5, 10)
    f = np.exp(-alpha * r) / r
    np.testing.assert_allclose(structure.radial_fourier_transform(r, f, q), 4 * np.pi / (alpha**2 + q**2), atol=4e-5)
    # coordination number of an ideal gas shell: no minima -> nan + warning
    with pytest.warns(UserWarning):
        cn = structure.calculate_coordination_numbers(np.linspace(0.1, 5, 50), np.ones(50), 0.8)
    assert np.isnan(cn).all()
    bins = np.linspace(0.05, 6, 120)
    rdf = 1 + np.cos(2 * np.pi * bins / 2.0) * np.exp(-bins / 3) * (bins > 0.8)
    rdf[bins <= 0.8] = 0.0
    cn = structure.calculate_coordination_numbers(bins, rdf, 0.8, threshold=0.05)
    assert np.isfinite(cn[0]) and cn[0] > 0
    qq, s = structure.calculate_structure_factor(bins, rdf, True, 0.8, n_q=16)
    assert qq.shape == s.shape == (16,)
    with pytest.raises(ValueError):
        structure.calculate_structure_factor(bins, rdf, False, 0.8, 0.5, 0.5, formalism="nope")


def test_synthetic_universe_protocol():
    from mdcraft_b200 import synthetic

    u = synthetic.make_config("C1", n_frames=5)
    assert u.atoms.n_atoms == 1000 and len(u.trajectory) == 5
    p0 = u.trajectory[0].positions.copy()
    assert p0.dtype == np.float32 and p0.shape == (1000, 3)
    L = u.dimensions[0]
    assert (p0 >= 0).all() and (p0 < L).all()
    sl = u.trajectory[1:5:2]
    assert [ts.frame for ts in sl] == [1, 3]
    assert u.trajectory.check_slice_indices(None, None, None) == (0, 5, 1)
    np.testing.assert_array_equal(u.trajectory[0].positions, p0)   # deterministic seeds


def test_center_of_mass_groupings():
    """`center_of_mass(group, grouping)` as the classes call it when groupings != "atoms"
    (reference: algorithm/molecule.py:16-441, einsum at :428)."""
    from mdcraft_b200 import synthetic

    rng = np.random.default_rng(11)
    n_res, per = 50, 4
    pos = rng.random((2, n_res * per, 3)).astype(np.float32) * 9
    masses = rng.random(n_res * per) + 0.5
    traj = synthetic.MemoryTrajectory(pos, np.array([9, 9, 9, 90, 90, 90], np.float32))
    u = synthetic.Universe(traj, resindices=np.repeat(np.arange(n_res), per), masses=masses)
    com = algorithm.center_of_mass(u.atoms, "residues")
    assert com.shape == (n_res, 3)
    want = (masses[:, None] * pos[0]).reshape(n_res, per, 3).sum(1) / masses.reshape(n_res, per).sum(1)[:, None]
    np.testing.assert_allclose(com, want, rtol=1e-12)
    # ragged residues take the per-group path
    res = np.concatenate([np.repeat(np.arange(10), 3), np.repeat(np.arange(10, 10 + 34), 5)])
    u2 = synthetic.Universe(traj, resindices=res, masses=masses)
    com2 = algorithm.center_of_mass(u2.atoms, "residues")
    assert com2.shape == (44, 3)
    np.testing.assert_allclose(com2[0], (masses[:3, None] * pos[0, :3]).sum(0) / masses[:3].sum(), rtol=1e-12)
    np.testing.assert_allclose(algorithm.center_of_mass(u.atoms), (masses[:, None] * pos[0]).sum(0) / masses.sum(), rtol=1e-12)
    with pytest.raises(ValueError, match="Invalid grouping"):
        algorithm.center_of_mass(u.atoms, "molecules")


def test_transforms_match_the_reference_golden():
    """radial_fourier_transform / calculate_structure_factor (host NumPy path) against outputs of the reference's
    own functions (tests/golden/make_golden.py transforms); Simpson weights reproduce scipy's rule for odd and even
    sample counts."""
    from scipy.integrate import simpson

    from conftest import golden

    g = golden("transforms.npz")
    for tag in ("odd", "even"):
        r, gr, q = g[f"{tag}_r"], g[f"{tag}_g"], g[f"{tag}_q"]
        np.testing.assert_allclose(structure.radial_fourier_transform(r, gr - 1, q), g[f"{tag}_rft"], rtol=1e-13, atol=1e-13)
        qq, s = structure.calculate_structure_factor(r, gr, True, 0.8, n_q=64)
        np.testing.assert_array_equal(qq, g[f"{tag}_q3"])
        np.testing.assert_allclose(s, g[f"{tag}_s3"], rtol=1e-13, atol=1e-13)
        qq, s = structure.calculate_structure_factor(r, gr, False, 0.5, 0.3, 0.7, n_q=48, formalism="AL")
        np.testing.assert_allclose(s, g[f"{tag}_sal"], rtol=1e-13, atol=1e-13)
        # the 2-D transform: the reference only works for length-1 q (structure.py:170-172); same numbers here
        np.testing.assert_allclose(structure.zeroth_order_hankel_transform(r, gr - 1, g[f"{tag}_ht_q"]), g[f"{tag}_ht"],
                                   rtol=1e-13, atol=1e-13)
        w = structure.simpson_weights(r)
        y = np.sin(r) * r
        np.testing.assert_allclose(w @ y, simpson(y, x=r), rtol=1e-14)


def test_dump_index_and_headers_on_the_host(tmp_path):
    """mdc_dump_open is host code (memory map, threaded newline ranking, header parsing): frame offsets, steps and
    boxes equal what the reference's reader found (tests/golden/dumps.npz); the threaded index equals the serial one."""
    import os

    from conftest import GOLDEN, golden
    from mdcraft_b200 import _capi

    g = golden("dumps.npz")
    for tag, name, has_id in (("ids", "dump_ids", True), ("noid", "dump_noid", False), ("sparse", "dump_sparse_ids", True),
                              ("triclinic", "dump_triclinic", True)):   # restricted triclinic bounds: reader.py:372-388
        d = _capi.DumpFile(os.path.join(GOLDEN, name + ".lammpsdump"))
        assert (d.n_frames, d.n_atoms, d.has_id) == (5, 60, has_id)
        np.testing.assert_array_equal(d.offsets[:-1], g[f"{tag}_offsets"])
        np.testing.assert_array_equal(d.steps, g[f"{tag}_steps"])
        np.testing.assert_array_equal(d.box, g[f"{tag}_dimensions"])
        assert ["-" if c is None else c for c in d.conventions] == list(g[f"{tag}_conventions"])
        d.close()
    # a file large enough for the threaded index (> 4 MiB), last line without a newline
    n, frames = 4000, 40
    rng = np.random.default_rng(3)
    path = tmp_path / "big.lammpsdump"
    with open(path, "w") as fh:
        for f in range(frames):
            fh.write(f"ITEM: TIMESTEP\n{f * 10}\nITEM: NUMBER OF ATOMS\n{n}\nITEM: BOX BOUNDS pp pp pp\n0 10\n0 11\n0 12.5\n")
            fh.write("ITEM: ATOMS id x y z\n")
            body = "\n".join("%d %g %g %g" % (i + 1, *rng.random(3) * 10) for i in range(n))
            fh.write(body + ("\n" if f < frames - 1 else ""))
    serial = _capi.DumpFile(str(path), n_threads=1)
    threaded = _capi.DumpFile(str(path), n_threads=7)
    assert serial.n_frames == threaded.n_frames == frames
    np.testing.assert_array_equal(serial.offsets, threaded.offsets)
    assert serial.offsets[-1] == os.path.getsize(path)
    np.testing.assert_array_equal(threaded.steps, np.arange(frames) * 10)
    np.testing.assert_array_equal(threaded.box[0], np.array([10, 11, 12.5, 90, 90, 90], np.float32))
    serial.close()
    threaded.close()
    # layouts the reader refuses, with the reference's message where it has one
    bad = tmp_path / "bad.lammpsdump"
    bad.write_text("ITEM: TIMESTEP\n0\nITEM: NUMBER OF ATOMS\n1\nITEM: BOX BOUNDS abc origin pp pp pp\n1 0 0 0\n0 1 0 0\n0 0 1 0\n"
                   "ITEM: ATOMS id x y z\n1 0 0 0\n")
    with pytest.raises(_capi.MdcError, match="general triclinic"):
        _capi.DumpFile(str(bad))
    bad.write_text("ITEM: TIMESTEP\n0\nITEM: NUMBER OF ATOMS\n1\nITEM: BOX BOUNDS xy xz yz pp pp pp\n0 1 0\n0 1\n0 1 0\n"
                   "ITEM: ATOMS id x y z\n1 0 0 0\n")
    with pytest.raises(_capi.MdcError, match="tilt factor missing"):
        _capi.DumpFile(str(bad))
    bad.write_text("ITEM: TIMESTEP\n0\nITEM: NUMBER OF ATOMS\n2\nITEM: BOX BOUNDS pp pp pp\n0 1\n0 1\n0 1\n"
                   "ITEM: ATOMS id x y z\n1 0 0 0\n2 0 0 0\n"
                   "ITEM: TIMESTEP\n5\nITEM: NUMBER OF ATOMS\n3\nITEM: BOX BOUNDS pp pp pp\n0 1\n0 1\n0 1\n"
                   "ITEM: ATOMS id x y z\n1 0 0 0\n2 0 0 0\n")
    with pytest.raises(_capi.MdcError, match="Timestep 5 has 3 atoms, but the topology has 2 atoms"):
        _capi.DumpFile(str(bad))
    with pytest.raises(_capi.MdcError, match="convention 'u'"):
        _capi.DumpFile(os.path.join(GOLDEN, "dump_ids.lammpsdump"), conventions=["u", "u", "u"])


def test_profile_post_processing_matches_the_reference_golden():
    """calculate_surface_charge_density / calculate_potential_profile (host NumPy / SciPy) on the reference's own
    charge density profile against the reference's own results (integral, matrix and periodic matrix methods)."""
    import warnings

    from conftest import golden
    from mdcraft_b200.analysis import profile

    g = golden("profiles.npz")
    bins, rho_q, L = g["a_bins_z"].copy(), g["a_charge_z"], float(g["dims"][2])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sigma = profile.calculate_surface_charge_density(bins, rho_q, 2.5, L=L, dV=0.3, reduced=True)
        np.testing.assert_allclose(sigma, g["a_sigma"][1], rtol=1e-13)
        kw = dict(L=L, sigma_q=0.01, reduced=True)
        np.testing.assert_allclose(profile.calculate_potential_profile(bins, rho_q, 2.5, **kw), g["a_potential_integral"],
                                   rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(profile.calculate_potential_profile(bins, rho_q, 2.5, method="matrix", V0=0.2, **kw),
                                   g["a_potential_matrix"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(profile.calculate_potential_profile(bins, rho_q, 2.5, method="matrix", pbc=True, **kw),
                                   g["a_potential_matrix_pbc"], rtol=1e-10, atol=1e-12)
    with pytest.raises(ValueError, match="one-dimensional"):
        profile.calculate_surface_charge_density(bins[None], rho_q)
    with pytest.raises(ValueError, match="'dielectric' must be specified"):
        profile.calculate_surface_charge_density(bins, rho_q, dV=1.0)


def test_netcdf_header_on_the_host(tmp_path):
    """mdc_ncdf_open is host code: NetCDF-3 classic and 64-bit-offset headers, record layout with and without other
    record variables, against SciPy's reader."""
    from mdcraft_b200 import _capi
    from oracle import oracle

    rng = np.random.default_rng(2)
    F, N = 6, 41
    pos = rng.normal(scale=20, size=(F, N, 3)).astype(np.float32)
    times = 0.25 * np.arange(F)
    lengths = np.column_stack((30 + 0.1 * np.arange(F), np.full(F, 31.5), np.full(F, 29.0)))
    angles = np.full((F, 3), 90.0)
    for version, box, vel in ((1, True, False), (2, True, True), (2, False, False)):
        path = str(tmp_path / f"t{version}{int(box)}{int(vel)}.nc")
        oracle.write_amber_netcdf(path, pos, times, lengths if box else None, angles if box else None, version=version,
                                  velocities=pos * 0.5 if vel else None)
        _, want_t, want_l, want_a = oracle.read_amber_netcdf(path)
        nc = _capi.NcdfFile(path)
        assert (nc.n_frames, nc.n_atoms, nc.has_box, nc.has_time) == (F, N, box, True)
        np.testing.assert_array_equal(nc.times.astype(np.float32), want_t)
        if box:
            np.testing.assert_array_equal(nc.cell_lengths, want_l)
            np.testing.assert_array_equal(nc.cell_angles, want_a)
        nc.close()
    bad = tmp_path / "bad.nc"
    bad.write_bytes(b"CDF\x05" + bytes(64))
    with pytest.raises(_capi.MdcError, match="classic"):
        _capi.NcdfFile(str(bad))


def test_generate_spherical_wavevectors_matches_the_reference():
    """structure.py:1436-1507, golden from the reference function (tests/golden/make_golden.py triclinic)."""
    from conftest import golden

    g = golden("rdf_triclinic.npz")
    cell = g["w_cell"]
    np.testing.assert_array_equal(structure.generate_spherical_wavevectors(cell, n_points=(3, 4, 5)), g["w_points"])
    np.testing.assert_array_equal(
        structure.generate_spherical_wavevectors(np.array([11.0, 12.5, 10.0, 60, 90, 75]), n_points=4), g["w_params"])
    q, k = structure.generate_spherical_wavevectors(cell, q_max=2.5, wavenumbers=True)
    np.testing.assert_array_equal(q, g["w_qmax"])
    np.testing.assert_array_equal(k, g["w_qmax_norm"])
    with pytest.raises(ValueError):
        structure.generate_spherical_wavevectors(cell, n_points=(3, 4))


class _CountingAnalysis:
    """Host-only analysis on the in-repo AnalysisBase protocol (no kernels): records what run_together feeds it."""

    def __new__(cls, universe, verbose=False):
        from mdcraft_b200.analysis.base import SerialAnalysisBase

        class Impl(SerialAnalysisBase):
            def __init__(self, universe, verbose):
                super().__init__(universe.trajectory, verbose)
                self.universe = universe
                self._device, self._frame_block = 0, 1

            def _prepare(self):
                self.seen, self.sums = [], []

            def _single_frame(self):
                self.seen.append(self._ts.frame)
                self.sums.append(float(self.universe.atoms.positions.sum()))

            def _conclude(self):
                self.results.total = float(np.sum(self.sums))

        return Impl(universe, verbose)


def test_run_together_steps_different_trajectories_in_lockstep(caplog):
    """run_together over analyses of DIFFERENT trajectories: frame i of each in the same turn, the bookkeeping of a
    plain run(); unequal frame counts are refused; verbose=True logs a start and a finish line per analysis."""
    import logging

    from mdcraft_b200 import synthetic
    from mdcraft_b200.analysis.base import run_together

    rng = np.random.default_rng(3)
    dims = np.array([5, 5, 5, 90, 90, 90], np.float32)
    ua = synthetic.Universe(synthetic.MemoryTrajectory(rng.random((7, 11, 3)).astype(np.float32), dims))
    ub = synthetic.Universe(synthetic.MemoryTrajectory(rng.random((7, 5, 3)).astype(np.float32), dims, dt=2.0))
    a, a2, b = _CountingAnalysis(ua), _CountingAnalysis(ua), _CountingAnalysis(ub, verbose=True)
    with caplog.at_level(logging.INFO, logger="mdcraft_b200.analysis"):
        run_together([a, b, a2], start=1, step=2)
    assert a.seen == a2.seen == b.seen == [1, 3, 5]
    np.testing.assert_array_equal(a.frames, [1, 3, 5])
    np.testing.assert_array_equal(b.times, [2.0, 6.0, 10.0])
    solo = _CountingAnalysis(ub).run(start=1, step=2)
    assert solo.results.total == b.results.total and solo.seen == b.seen
    assert [r.getMessage().split(":")[0] for r in caplog.records] == ["Impl", "Impl"] and "3 frames" in caplog.records[0].getMessage()
    uc = synthetic.Universe(synthetic.MemoryTrajectory(rng.random((6, 5, 3)).astype(np.float32), dims))
    with pytest.raises(ValueError, match="same number of selected frames"):
        run_together([a, _CountingAnalysis(uc)])


def test_run_log_names_the_kernel_path_of_each_plan(caplog):
    """The finish line of the verbose run log (base.py:514-515 logs the farm's wall time; here: device, frames/s and what
    the plan ran): S(q) plans report their algorithm, g(r) plans the kernel path of mdc_rdf_plan_info."""
    import logging

    from mdcraft_b200 import synthetic

    rng = np.random.default_rng(4)
    dims = np.array([5, 5, 5, 90, 90, 90], np.float32)
    u = synthetic.Universe(synthetic.MemoryTrajectory(rng.random((3, 9, 3)).astype(np.float32), dims))

    class FakePlan:
        def __init__(self, info):
            self._info = info

        def info(self):
            return self._info

    cases = (({"algo": "nufft", "nf": 256, "w": 12}, "algorithm nufft (nf = 256, w = 12)"),
             ({"algo": "factored", "nf": 0, "w": 0}, "algorithm factored"),
             ({"filter": True, "sorted": True, "ext_words": 367, "copies_interleaved": True, "relative_rows": True,
               "triclinic": False, "queue": 1024, "frames": 3},
              "filter kernel, Hilbert-ordered frames, relative rows, one histogram copy per bank"),
             ({"filter": True, "sorted": False, "ext_words": 640, "copies_interleaved": False, "relative_rows": False,
               "triclinic": True, "queue": 4096, "frames": 3}, "filter kernel, triclinic cell, extended histograms"),
             ({"filter": False, "sorted": False, "ext_words": 0, "copies_interleaved": False, "relative_rows": False,
               "triclinic": True, "queue": 0, "frames": 3}, "FP64 kernel for every pair, triclinic cell"))
    for info, want in cases:
        a = _CountingAnalysis(u, verbose=True)
        a._plan = FakePlan(info)
        caplog.clear()
        with caplog.at_level(logging.INFO, logger="mdcraft_b200.analysis"):
            a.run()
        assert caplog.records[-1].getMessage().endswith(want), caplog.records[-1].getMessage()
