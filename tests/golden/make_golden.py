"""
Generates the golden fixtures in this directory by running the REAL reference
(`/root/reference/src/mdcraft/analysis/structure.py` and
`algorithm/accelerated.py`, unmodified) in the build container.

    python tests/golden/make_golden.py

The reference cannot be imported normally here (MDAnalysis, pint, matplotlib ...
are not installed), so `oracle/ref_loader.py` registers stub modules and drives
the real `StructureFactor` / `RadialDistributionFunction` classes on the
in-memory Universe of `mdcraft_b200.synthetic`.  S(q) outputs are 100 % reference
arithmetic.  For g(r) the only non-reference code on the path is the
`capped_distance` stub (`oracle.capped_distance`, a NumPy restatement of
MDAnalysis' brute-force orthorhombic routine); binning, exclusion, accumulation
and normalisation are the reference's.

Every fixture stores its inputs (positions, boxes, arguments) next to the
reference outputs, so the tests never need `/root/reference`.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from mdcraft_b200 import synthetic  # noqa: E402
from oracle import ref_loader  # noqa: E402


def universe_from(pos, dims, resindices=None, masses=None):
    traj = synthetic.MemoryTrajectory(np.asarray(pos, np.float32), np.asarray(dims, np.float32))
    return synthetic.Universe(traj, resindices=resindices, masses=masses)


def lj_frames(n, n_frames, rho, seed0, scale=None):
    L = synthetic.cubic_box_length(n, rho)
    pos = np.stack([synthetic.lj_like_frame(n, L, seed0 + i) for i in range(n_frames)])
    dims = np.array([L, L, L, 90, 90, 90], np.float32)
    if scale is not None:
        scale = np.asarray(scale, np.float32)
        pos = (pos * scale).astype(np.float32)
        dims[:3] = dims[:3] * scale
        pos = np.where(pos >= dims[:3], np.float32(0), pos).astype(np.float32)
    return pos, dims


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def main():
    st, acc = ref_loader.load()
    rng = np.random.default_rng(20261019)

    # ------------------------------------------------------------------ S(q)
    # (a) C1: total S(q), default 32^3 grid, N = 1000, 3 frames
    pos, dims = lj_frames(1000, 3, 0.8, 1000)
    u = universe_from(pos, dims)
    r = st.StructureFactor(u.atoms, parallel=True, verbose=False).run()
    save("sq_total_c1.npz", pos=pos, dims=dims, ssf=r.results.ssf, wavenumbers=r.results.wavenumbers,
         n_points=32)

    # (b) partial S(q), two groups, q_max filter, off-lattice surfaces
    pos, dims = lj_frames(600, 3, 0.8, 1100)
    u = universe_from(pos, dims)
    ga, gb = u.select(0, 250), u.select(250, 600)
    r = st.StructureFactor([ga, gb], mode="partial", n_points=14, q_max=6.0, n_surfaces=2, n_surface_points=6,
                           verbose=False).run()
    save("sq_partial_surfaces.npz", pos=pos, dims=dims, sizes=np.array([250, 350]), ssf=r.results.ssf,
         wavenumbers=r.results.wavenumbers, pairs=np.array(r.results.pairs), n_points=14, q_max=6.0,
         n_surfaces=2, n_surface_points=6, wavevectors=r._wavevectors)

    # (c) pair mode, orthorhombic box, unsorted / non-unique output, trig form
    pos, dims = lj_frames(512, 2, 0.8, 1200, scale=(1.0, 1.25, 0.8))
    u = universe_from(pos, dims)
    ga, gb = u.select(0, 200), u.select(200, 512)
    r = st.StructureFactor([ga, gb], mode="pair", n_points=7, sort=False, unique=False, verbose=False).run()
    r2 = st.StructureFactor([ga, gb], mode="pair", n_points=7, form="trig", verbose=False).run()
    save("sq_pair_ortho.npz", pos=pos, dims=dims, sizes=np.array([200, 312]), ssf=r.results.ssf,
         wavenumbers=r.results.wavenumbers, ssf_trig_unique=r2.results.ssf, wavenumbers_unique=r2.results.wavenumbers,
         n_points=7, wavevectors=r._wavevectors)

    # (d) explicit wavevectors
    pos, dims = lj_frames(300, 2, 0.8, 1300)
    u = universe_from(pos, dims)
    wv = rng.normal(size=(97, 3)) * 2.5
    r = st.StructureFactor(u.atoms, wavevectors=wv, sort=False, unique=False, verbose=False).run()
    save("sq_explicit.npz", pos=pos, dims=dims, wavevectors=wv, ssf=r.results.ssf, wavenumbers=r.results.wavenumbers)

    # (e) the functional seam: delta_fourier_transform_sum
    import numba

    dfts = numba.njit(st.delta_fourier_transform_sum, fastmath=True)
    qs = rng.normal(size=(64, 3)) * 3.0
    rs = rng.random((777, 3)) * 9.0
    save("sq_dfts.npz", qs=qs, rs=rs, rho=dfts(qs, rs))

    # (e2) intermediate scattering functions: total with the self part, and partial
    n, F = 250, 7
    L = synthetic.cubic_box_length(n, 0.8)
    base = synthetic.lj_like_frame(n, L, 1600).astype(np.float64)
    walk = np.cumsum(np.random.default_rng(1601).normal(scale=0.08, size=(F, n, 3)), axis=0)
    pos = np.mod(base[None] + walk, float(L)).astype(np.float32)
    dims = np.array([L, L, L, 90, 90, 90], np.float32)
    u = universe_from(pos, dims)
    r = st.IntermediateScatteringFunction(u.atoms, n_points=6, n_lags=4, incoherent=True, verbose=False).run()
    r2 = st.IntermediateScatteringFunction([u.select(0, 100), u.select(100, n)], mode="partial", n_points=5, n_lags=3,
                                           incoherent=True, sort=False, unique=False, verbose=False).run()
    save("isf_small.npz", pos=pos, dims=dims, cisf=r.results.cisf, iisf=r.results.iisf, times=r.results.times,
         wavenumbers=r.results.wavenumbers, sizes=np.array([100, 150]), cisf_partial=r2.results.cisf,
         iisf_partial=r2.results.iisf, pairs_partial=np.array(r2.results.pairs))

    # (e3) single-chain structure factor: 12 chains of 20 monomers (random walks)
    pol = ref_loader.load_polymer()
    M, N, F = 12, 20, 3
    L = 14.0
    steps = np.random.default_rng(1700).normal(scale=0.6, size=(F, M, N, 3))
    starts = np.random.default_rng(1701).random((F, M, 1, 3)) * L
    pos = np.mod(starts + np.cumsum(steps, axis=2), L).reshape(F, M * N, 3).astype(np.float32)
    dims = np.array([L, L, L, 90, 90, 90], np.float32)
    u = universe_from(pos, dims)
    r = pol.SingleChainStructureFactor(u.atoms, n_chains=np.array([M]), n_monomers=np.array([N]), n_points=9,
                                       verbose=False).run()
    save("scsf_small.npz", pos=pos, dims=dims, n_chains=M, n_monomers=N, n_points=9, scsf=r.results.scsf,
         wavenumbers=r.results.wavenumbers)

    # ------------------------------------------------------------------ g(r)
    # (f) C1: self-RDF, 200 bins to L/2, with and without exclusion=(1, 1)
    pos, dims = lj_frames(1000, 4, 0.8, 1000)
    u = universe_from(pos, dims)
    half = float(dims[0]) / 2
    r = st.RadialDistributionFunction(u.atoms, n_bins=200, range_=(0.0, half), verbose=False).run()
    rx = st.RadialDistributionFunction(u.atoms, n_bins=200, range_=(0.0, half), exclusion=(1, 1), verbose=False).run()
    save("rdf_self_c1.npz", pos=pos, dims=dims, range_=np.array([0.0, half]), n_bins=200,
         counts=r.results.counts, rdf=r.results.rdf, bins=r.results.bins, bin_edges=r.results.bin_edges,
         counts_excl=rx.results.counts, rdf_excl=rx.results.rdf)

    # (g) cross-RDF, r0 > 0, block exclusion (4, 10), default 201 bins; norm variants
    pos, dims = lj_frames(700, 3, 0.8, 1400)
    u = universe_from(pos, dims)
    ga, gb = u.select(0, 200), u.select(200, 700)
    out = {}
    for norm in ("rdf", "density", None):
        r = st.RadialDistributionFunction(ga, gb, n_bins=201, range_=(0.35, 4.1), exclusion=(4, 10), norm=norm,
                                          verbose=False).run()
        out[f"rdf_{norm}"] = r.results.rdf
        out["counts"] = r.results.counts
        out["bin_edges"] = r.results.bin_edges
    save("rdf_cross_excl.npz", pos=pos, dims=dims, sizes=np.array([200, 500]), range_=np.array([0.35, 4.1]),
         n_bins=201, exclusion=np.array([4, 10]), **out)

    # (h) NPT (box changes every frame), orthorhombic, drop_axis
    n, F = 400, 4
    boxes = np.stack([np.array([7.5 + 0.3 * i, 8.0 - 0.2 * i, 9.0 + 0.1 * i, 90, 90, 90], np.float32) for i in range(F)])
    pos = np.stack([(np.random.default_rng(1500 + i).random((n, 3)) * boxes[i, :3]).astype(np.float32) for i in range(F)])
    u = universe_from(pos, boxes)
    r = st.RadialDistributionFunction(u.atoms, n_bins=64, range_=(0.0, 3.5), verbose=False).run()
    rd = st.RadialDistributionFunction(u.atoms, n_bins=64, range_=(0.0, 3.5), drop_axis=2, verbose=False).run()
    save("rdf_npt_dropaxis.npz", pos=pos, dims=boxes, range_=np.array([0.0, 3.5]), n_bins=64,
         counts=r.results.counts, rdf=r.results.rdf, counts_drop=rd.results.counts, rdf_drop=rd.results.rdf)

    # (i) the functional seam, as the reference's own test uses it
    #     (tests/test_analysis_structure.py:21-46): neighbours at known radii around the box centre
    L = 20.0
    origin = np.full(3, L / 2, np.float32)
    norm_ = rng.random(1000) * (L / 2)
    v = rng.normal(size=(1000, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    neighbors = (origin + norm_[:, None] * v).astype(np.float32)
    box = np.array([L, L, L, 90, 90, 90], np.float32)
    h = st.radial_histogram(origin, neighbors, n_bins=10, range_=(0, 11), dimensions=box)
    save("rdf_radial_histogram.npz", origin=origin, neighbors=neighbors, dims=box, counts=h)

    # (j) the bin formula itself: numba_histogram on length-1 arrays (scalar loop) for random values and
    #     for values within a few ulps of every edge, plus numba_histogram_bin_edges
    cases = [(0.0, 5.386094808578491, 50), (0.0, 15.0, 201), (0.3, 7.7, 200), (1.0, 11.0, 10), (0.25, 9.3, 1000),
             (0.0, 17.09976, 200)]
    hist = {}
    for c, (r0, r1, n) in enumerate(cases):
        edges = acc.numba_histogram_bin_edges(np.asarray((r0, r1), dtype=float), n)
        vals = list(rng.uniform(r0 - 0.05, r1 + 0.05, 300))
        for e in edges[:: max(1, n // 40)]:
            x = e
            for _ in range(3):
                vals.append(x)
                x = np.nextafter(x, np.inf)
            x = e
            for _ in range(3):
                x = np.nextafter(x, -np.inf)
                vals.append(x)
        vals += [r1, np.nextafter(r1, -np.inf), np.nextafter(r1, np.inf), r0, r0 - 1e-17, r0 - 2.3e-16]
        vals = np.array(vals)
        bins = np.empty(vals.shape[0], dtype=np.int64)
        for i, x in enumerate(vals):
            nz = np.nonzero(acc.numba_histogram(np.array([x]), n, edges))[0]
            bins[i] = nz[0] if len(nz) else -1
        hist[f"case{c}_args"] = np.array([r0, r1, n])
        hist[f"case{c}_edges"] = edges
        hist[f"case{c}_values"] = vals
        hist[f"case{c}_bins"] = bins
    # the reference's position dependence at a bin edge (see oracle.c): np.full(L, just-below-edge)
    edges = acc.numba_histogram_bin_edges(np.asarray((0.0, 15.0), dtype=float), 201)
    below = np.nextafter(edges[50], -np.inf)
    hist["position_dependence"] = np.array(
        [[L_, *acc.numba_histogram(np.full(L_, below), 201, edges)[49:51]] for L_ in (1, 2, 3, 4, 8)]
    )
    save("histogram_pins.npz", n_cases=len(cases), **hist)


def transforms():
    """g(r)-derived transforms: the reference's radial_fourier_transform,
    zeroth_order_hankel_transform and calculate_structure_factor on a synthetic liquid-like g(r), for an odd and
    an even number of samples (SciPy's Simpson rule treats them differently) and with q = 0 present."""
    st, _ = ref_loader.load()
    out = {}
    for tag, n in (("odd", 201), ("even", 200)):
        r = (np.arange(n) + 0.5) * (8.0 / n)
        g = 1 + 1.8 * np.cos(2 * np.pi * (r - 1.0) / 0.95) * np.exp(-(r - 1.0) / 1.3) * (r > 0.85)
        g[r <= 0.85] = 0.0
        q = np.concatenate(([0.0], np.linspace(0.3, 25.0, 300)))
        out[f"{tag}_r"], out[f"{tag}_g"], out[f"{tag}_q"] = r, g, q
        out[f"{tag}_rft"] = st.radial_fourier_transform(r, g - 1, q)
        # The reference's 2-D transform (structure.py:170-172) multiplies q * r elementwise instead of forming the
        # outer product and then tests `0 in q`: it raises for every q except a length-1 array, the form pinned here.
        out[f"{tag}_ht_q"] = np.array([2.5, 7.0])
        out[f"{tag}_ht"] = np.array([float(st.zeroth_order_hankel_transform(r, g - 1, np.array([qq])))
                                     for qq in (2.5, 7.0)])
        out[f"{tag}_q3"], out[f"{tag}_s3"] = st.calculate_structure_factor(r, g, True, 0.8, n_q=64)
        out[f"{tag}_qal"], out[f"{tag}_sal"] = st.calculate_structure_factor(r, g, False, 0.5, 0.3, 0.7, n_q=48,
                                                                            formalism="AL")
    save("transforms.npz", **out)


def write_dump(path, pos, ids, box_lo, box_hi, steps, columns, fmt, extra=None, literal=None, tilt=None):
    """A LAMMPS text dump: pos (F, N, 3) written with the per-frame format functions fmt[f](value) -> str.
    tilt (F, 3) = (xy, xz, yz): restricted triclinic boxes, box_lo / box_hi are then the BOUNDS LAMMPS writes."""
    F, N, _ = pos.shape
    with open(path, "w") as fh:
        for f in range(F):
            kind = "xy xz yz pp pp pp" if tilt is not None else "pp pp pp"
            fh.write(f"ITEM: TIMESTEP\n{steps[f]}\nITEM: NUMBER OF ATOMS\n{N}\nITEM: BOX BOUNDS {kind}\n")
            for ax in range(3):
                t = f" {float(tilt[f][ax])!r}" if tilt is not None else ""
                fh.write(f"{float(box_lo[f][ax])!r} {float(box_hi[f][ax])!r}{t}\n")
            fh.write("ITEM: ATOMS " + " ".join(columns) + "\n")
            for i in range(N):
                fields = []
                for c in columns:
                    if c == "id":
                        fields.append(str(ids[f][i]))
                    elif c in ("x", "y", "z", "xu", "yu", "zu"):
                        ax = "xyz".index(c[0])
                        fields.append((literal or {}).get((f, i, ax)) or fmt[f](pos[f, i, ax]))
                    else:
                        fields.append(extra(c, f, i))
                fh.write(" ".join(fields) + (" \n" if i % 7 == 3 else "\n"))   # some trailing blanks


def dumps():
    """Frame ingest: small LAMMPS dumps and what the reference's LAMMPSDumpTrajectoryReader
    (reader.py, unmodified, on the ReaderBase stub of oracle/ref_loader.py) reads from them."""
    rd = ref_loader.load_reader()
    rng = np.random.default_rng(77)
    out = {}
    F, N = 5, 60
    pos = rng.normal(scale=8.0, size=(F, N, 3))
    # number syntax the device's exact fast path must get right or hand to the host parser: a float32 midpoint
    # nudged by less than half a double ulp (double rounding), > 2^53 mantissas, long digit strings, odd but valid
    # float() literals, non-finite values
    lits = ["0", "-0", "1e-30", "123456789.125", "1.00000005960464477625798673798840354720596224069595336914062",
            "9007199254740993", "0.1e1", "1E+2", "-.5", "5.", "+3.25", "1e23", "8.5e-23", "123456789012345678901234567890e-25",
            "0.000000000000000000000000000001234", "nan", "-inf", "1e400", "4.35", "-7.0000000000000000000001"]
    literal = {(0, i, i % 3): lit for i, lit in enumerate(lits)}
    ids = np.stack([rng.permutation(N) + 1 for _ in range(F)])
    lo = rng.normal(size=(F, 3)) - 10.0
    hi = lo + 20.0 + rng.random((F, 3))
    steps = [0, 500, 1000, 1500, 2500]
    fmts = [lambda v: "%g" % v, lambda v: "%.17g" % v, lambda v: "%20.15e" % v, lambda v: "%+.8f" % v, lambda v: repr(float(v))]
    extra = lambda c, f, i: {"type": str(1 + i % 3), "vx": "%g" % (0.1 * i), "q": "-1.5e-3"}[c]   # noqa: E731
    # (a) ids shuffled per frame, extra columns around the coordinates
    a = os.path.join(HERE, "dump_ids.lammpsdump")
    write_dump(a, pos, ids, lo, hi, steps, ["id", "type", "x", "y", "vx", "z", "q"], fmts, extra, literal)
    # (b) no id column, unwrapped names
    b = os.path.join(HERE, "dump_noid.lammpsdump")
    write_dump(b, pos, ids, lo, hi, steps, ["type", "xu", "yu", "zu"], fmts, extra, literal)
    # (c) ids that are not a contiguous range (a dumped sub-group)
    c = os.path.join(HERE, "dump_sparse_ids.lammpsdump")
    sparse = np.stack([rng.permutation(np.arange(N) * 7 + 3) for _ in range(F)])
    write_dump(c, pos, sparse, lo, hi, steps, ["x", "y", "z", "id"], fmts, extra)
    # (d) restricted triclinic boxes (ITEM: BOX BOUNDS xy xz yz): bounds = extent of the tilted box, as LAMMPS writes them
    d = os.path.join(HERE, "dump_triclinic.lammpsdump")
    rng_t = np.random.default_rng(78)
    edge = 18.0 + 4.0 * rng_t.random((F, 3))
    tilt = (rng_t.random((F, 3)) - 0.5) * 0.9 * edge[:, [0, 0, 1]]
    tilt[1] = [0.0, -3.5, 0.0]                      # one tilt only
    tilt[2, 0] = 0.0
    org = rng_t.normal(size=(F, 3)) - 5.0
    xy, xz, yz = tilt[:, 0], tilt[:, 1], tilt[:, 2]
    zero = np.zeros(F)
    tlo = org + np.stack([np.min([zero, xy, xz, xy + xz], axis=0), np.minimum(0, yz), zero], axis=1)
    thi = org + edge + np.stack([np.max([zero, xy, xz, xy + xz], axis=0), np.maximum(0, yz), zero], axis=1)
    write_dump(d, pos, ids, tlo, thi, steps, ["id", "x", "y", "z"], fmts, extra, tilt=tilt)
    for tag, path in (("ids", a), ("noid", b), ("sparse", c), ("triclinic", d)):
        r = rd.LAMMPSDumpTrajectoryReader(path, dt=0.005)
        got_pos, got_dim, got_step, got_time = [], [], [], []
        for f in range(r.n_frames):
            ts = r[f]
            got_pos.append(ts.positions.copy())
            got_dim.append(ts.dimensions)
            got_step.append(ts.data["step"])
            got_time.append(ts.data["time"])
        out[f"{tag}_positions"] = np.stack(got_pos)
        out[f"{tag}_dimensions"] = np.stack(got_dim)
        out[f"{tag}_steps"] = np.array(got_step)
        out[f"{tag}_times"] = np.array(got_time)
        out[f"{tag}_offsets"] = np.array(r._offsets)
        out[f"{tag}_conventions"] = np.array(["-" if c_ is None else c_ for c_ in r._conventions])
        r.close()
    save("dumps.npz", **out)


def profiles():
    """DensityProfile: the reference class (profile.py, unmodified) on in-memory trajectories
    whose positions partly lie outside the box (so that wrap() matters), plus its post-processing."""
    import warnings

    pr = ref_loader.load_profile()
    rng = np.random.default_rng(4242)
    out = {}
    F, N = 6, 400
    dims = np.array([9.5, 11.25, 14.0, 90, 90, 90], np.float32)
    pos = (rng.random((F, N, 3)) * (dims[:3] * 1.6) - dims[:3] * 0.3).astype(np.float32)
    pos[0, :3, 2] = [0.0, dims[2], -0.0]                       # on the edges
    pos[1, :2, 0] = [np.float32(dims[0]) * 2, -dims[0]]        # exact multiples of the box
    # slab-like charge separation along z so that the potential is not flat
    pos[:, :150, 2] = (rng.normal(3.0, 0.8, size=(F, 150))).astype(np.float32)
    pos[:, 150:, 2] = (rng.normal(10.5, 1.1, size=(F, N - 150))).astype(np.float32)
    out["pos"], out["dims"] = pos, dims
    u = universe_from(pos, dims)
    # (a) two groups, two axes, per-axis bin counts, charges, reduced units
    d = pr.DensityProfile([u.select(0, 150), u.select(150, N)], axes="xz", n_bins=[5, 48], charges=[1.0, -0.6],
                          reduced=True, verbose=False).run()
    for ax in "xz":
        out[f"a_bins_{ax}"], out[f"a_edges_{ax}"] = d.results.bins[ax], d.results.bin_edges[ax]
        out[f"a_number_{ax}"], out[f"a_charge_{ax}"] = d.results.number_densities[ax], d.results.charge_densities[ax]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # all axes at once: with a subset the reference zips the per-axis arrays against the wrong entries
        # (profile.py:1377-1380 pairs axes[i] with self._dielectrics[i], not with its relative index) and yields NaN
        d.calculate_surface_charge_densities(None, [1.5, 2.5], dVs=[0.1, 0.3])
        out["a_sigma"] = d.results.surface_charge_densities.copy()
        d.calculate_potential_profiles(None, [1.5, 2.5], sigmas_q=[0.02, 0.01])
        out["a_potential_integral"] = d.results.potentials["z"].copy()
        d.calculate_potential_profiles(None, [1.5, 2.5], sigmas_q=[0.02, 0.01], methods="matrix", V0s=[0.0, 0.2])
        out["a_potential_matrix"] = d.results.potentials["z"].copy()
        d.calculate_potential_profiles(None, [1.5, 2.5], sigmas_q=[0.02, 0.01], methods="matrix", pbcs=True)
        out["a_potential_matrix_pbc"] = d.results.potentials["z"].copy()
    # (b) one group, default bins on all axes, every frame kept, run(start, stop, step)
    d = pr.DensityProfile(u.atoms, average=False, verbose=False).run(start=1, stop=6, step=2)
    for ax in "xyz":
        out[f"b_number_{ax}"] = d.results.number_densities[ax]
    out["b_times"] = d.results.times
    # (c) residues (pairs of atoms, unequal masses), explicit dimensions and scale factors
    u2 = universe_from(pos, dims, resindices=np.arange(N) // 2, masses=1.0 + (np.arange(N) % 2) * 2.5)
    d = pr.DensityProfile(u2.atoms, "residues", axes="y", n_bins=33, dimensions=[9.5, 11.25, 14.0],
                          dim_scales=(1.0, 0.9, 1.0), verbose=False).run()
    out["c_number_y"], out["c_edges_y"] = d.results.number_densities["y"], d.results.bin_edges["y"]
    # (d) recentring: a drifting slab (the whole system translates and crosses the periodic boundary), recentred on
    # group 0 with the default target (box centre), and on both groups with an explicit target (NaN = leave that axis)
    drift = (np.arange(F)[:, None, None] * np.array([0.0, 0.0, 2.7])).astype(np.float32)
    pos_d = np.mod(pos + drift, dims[:3]).astype(np.float32)
    out["pos_drift"] = pos_d
    u3 = universe_from(pos_d, dims, masses=1.0 + (np.arange(N) % 3))
    d = pr.DensityProfile([u3.select(0, 150), u3.select(150, N)], axes="z", n_bins=40, recenter=0, verbose=False).run()
    out["d_number_z"] = d.results.number_densities["z"]
    d = pr.DensityProfile([u3.select(0, 150), u3.select(150, N)], axes="xz", n_bins=[6, 40],
                          recenter=([0, 1], [np.nan, np.nan, 3.5]), average=False, verbose=False).run(start=1)
    out["e_number_x"], out["e_number_z"] = d.results.number_densities["x"], d.results.number_densities["z"]
    save("profiles.npz", **out)


def triclinic():
    """g(r) in triclinic cells: the reference classes (binning, exclusion, accumulation, normalisation, ts.volume) over
    the restated triclinic capped_distance (oracle.capped_distance: triclinic_vectors, _triclinic_pbc,
    minimum_image_triclinic).  Cells: 60/90/75 (as VERDICT r01 asks), a strongly skewed 50/65/55, an NPT sequence that
    mixes orthorhombic and triclinic frames; positions in and outside the primary cell."""
    from oracle import oracle

    st, _ = ref_loader.load()
    rng = np.random.default_rng(60_90_75)
    out = {}
    n, F = 500, 3
    box = np.array([11.0, 12.5, 10.0, 60, 90, 75], np.float32)
    m = oracle.triclinic_vectors(box).astype(np.float64)
    pos = ((rng.random((F, n, 3)) * 2.2 - 0.6) @ m).astype(np.float32)          # fractional coordinates in [-0.6, 1.6)
    u = universe_from(pos, box)
    r = st.RadialDistributionFunction(u.atoms, n_bins=80, range_=(0.0, 4.5), verbose=False).run()
    out.update(a_pos=pos, a_dims=box, a_counts=r.results.counts, a_rdf=r.results.rdf)
    r = st.RadialDistributionFunction(u.select(0, 180), u.select(180, n), n_bins=57, range_=(0.4, 4.9), exclusion=(3, 5),
                                      verbose=False).run()
    out.update(a_cross_counts=r.results.counts, a_cross_rdf=r.results.rdf)
    boxes = np.array([[9.0, 9.5, 10.0, 50, 65, 55], [9.2, 9.4, 10.1, 90, 90, 90], [9.1, 9.6, 9.9, 90, 101.5, 90],
                      [9.3, 9.3, 10.2, 120, 90, 90]], np.float32)
    posb = np.stack([((rng.random((n, 3)) * 1.4 - 0.2) @ oracle.triclinic_vectors(b).astype(np.float64)).astype(np.float32)
                     for b in boxes])
    u = universe_from(posb, boxes)
    r = st.RadialDistributionFunction(u.atoms, n_bins=64, range_=(0.0, 3.9), exclusion=(1, 1), verbose=False).run()
    out.update(b_pos=posb, b_dims=boxes, b_counts=r.results.counts, b_rdf=r.results.rdf)
    # generate_spherical_wavevectors (structure.py:1436-1507) of the same cell: box vectors / lattice parameters, q_max
    cell = oracle.triclinic_vectors(box).astype(np.float64)
    out["w_cell"] = cell
    out["w_points"] = st.generate_spherical_wavevectors(cell, n_points=(3, 4, 5))
    out["w_params"] = st.generate_spherical_wavevectors(np.array([11.0, 12.5, 10.0, 60, 90, 75]), n_points=4)
    q, k = st.generate_spherical_wavevectors(cell, q_max=2.5, wavenumbers=True)
    out["w_qmax"], out["w_qmax_norm"] = q, k
    save("rdf_triclinic.npz", **out)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "triclinic"):
        triclinic()
    if which in ("all", "profiles"):
        profiles()
    if which in ("all", "main"):
        main()
    if which in ("all", "transforms"):
        transforms()
    if which in ("all", "dumps"):
        dumps()
