#!/usr/bin/env python
"""
bench.py — frames/s of the S(q) + g(r) hot path at 100,000 particles.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1] + configs[2], SURVEY.md §8d C2 + C3):
  * g(r): N = 100,000 uniform ("GCMe-like", rho* = 2.5, L = 34.2), self-RDF, 200 bins to L/2,
          exclusion = (1, 1)  -> 1e10 ordered pairs per frame;
  * S(q): N = 100,000 lattice + jitter ("LJ-like", rho* = 0.8, L = 50), the FULL reciprocal grid to
          q_max = 20 / sigma (n_points = 161 -> Nq = 2,140,871 wavevectors) -> 2.14e11 terms per frame.
A "step" is one frame block (--frames, default 2) through BOTH analyses.  `value` = frames/s with the
frames already resident in HBM; `e2e` = the same through the C ABI with pinned HOST buffers (H2D of the
frames and D2H of the step's results inside the timed region).  Frames shard across ranks (weak scaling:
every rank processes its own --frames per step); accumulators are all-reduced once at the end.

Rooflines.  `roofline` describes the kernel with the largest share of the timed step: the g(r)
pair kernel (issue-rate bound, 20 algorithmic issue slots per pair-distance).  S(q) on this grid runs as a gridding NUFFT
(`roofline_sq`, HBM bound: fine grid + pruned FFT intermediates); the two O(Nq N) kernels it replaced are timed on
one frame each and reported against their issue-rate rooflines measured on the box in this run
(`roofline_sfu_direct`: 2 SFU ops per term, the contract kernel of BASELINE.json's north_star;
`roofline_fp32_factored`: 4 FP32 FMAs per term).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_PART = 100_000
GR = dict(rho=2.5, n_bins=200, seed0=2000)
SQ = dict(rho=0.8, n_points=161, q_max=20.0, seed0=3000)
METRIC = "frames/sec for S(q) and g(r) at 1e5 particles"
#: dram__bytes_read.sum + dram__bytes_write.sum of the four NUFFT kernels (spread, z / y / x FFT passes) per frame from
#: one `ncu --set full` capture (profiles/sq_nufft_r01.txt: 2,706 MB per 2-frame launch); clearing the fine grid adds
#: 0.27e9 that ncu does not attribute to a kernel
NUFFT_DRAM_BYTES_PER_FRAME = 1.353e9
UNIT = "frames/s"


def workload_config(frames, n_gpus):
    return {
        "workload": "C2 g(r) (N=1e5 uniform rho*=2.5, 200 bins to L/2, exclusion (1,1)) + "
                    "C3 S(q) (N=1e5 LJ-like rho*=0.8, full lattice grid to q_max=20/sigma, Nq=2,140,871)",
        "n_particles": N_PART,
        "frames_per_step_per_gpu": frames,
        "sharding": f"frames x{n_gpus}",
        "l2": "256 MiB device buffer written between timed steps (L2 flush)",
    }


# ----------------------------------------------------------------------------- inputs


def make_frames(frames, rank):
    from mdcraft_b200 import synthetic

    Lg = synthetic.cubic_box_length(N_PART, GR["rho"])
    Ls = synthetic.cubic_box_length(N_PART, SQ["rho"])
    off = rank * 1000
    gr = np.stack([synthetic.uniform_frame(N_PART, Lg, GR["seed0"] + off + i) for i in range(frames)])
    sq = np.stack([synthetic.lj_like_frame(N_PART, Ls, SQ["seed0"] + off + i) for i in range(frames)])
    box_g = np.array([Lg, Lg, Lg, 90, 90, 90], np.float32)
    return gr, box_g, sq, float(Ls)


# ----------------------------------------------------------------------------- clocks


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self._stop, self._t = device, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits"],
                    capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


# ----------------------------------------------------------------------------- CPU baseline (oracle)


def cpu_baseline(seconds_target=10.0):
    """
    The CPU restatement of the reference path (oracle/oracle.c, OpenMP over all host cores) on a bounded
    sample of the same workload: S(q) on a slab of the wavevector grid, g(r) on a slab of i-particles,
    both against all 100,000 particles of one frame; scaled to frames/s by the exact work ratio.
    """
    from mdcraft_b200 import synthetic
    from oracle import oracle

    cores = oracle.num_threads()
    Ls = synthetic.cubic_box_length(N_PART, SQ["rho"])
    pos_s = synthetic.lj_like_frame(N_PART, Ls, SQ["seed0"]).astype(np.float64)
    wv, _ = oracle.wavevector_grid(np.full(3, float(Ls)), 64, SQ["q_max"])
    nq_full = 2_140_871
    wv = wv[-min(2048 * cores, len(wv)):]
    nq = len(wv)
    oracle.delta_fourier_transform_sum(wv[:cores], pos_s[:1000])   # warm the thread pool
    t0 = time.perf_counter()
    oracle.delta_fourier_transform_sum(wv, pos_s)
    t_sq = time.perf_counter() - t0
    terms_per_s = nq * N_PART / t_sq

    Lg = synthetic.cubic_box_length(N_PART, GR["rho"])
    pos_g = synthetic.uniform_frame(N_PART, Lg, GR["seed0"])
    box = np.array([Lg, Lg, Lg, 90, 90, 90], np.float32)
    na = 1024 * cores
    oracle.radial_histogram(pos_g[:cores], pos_g[:2000], GR["n_bins"], (0.0, float(Lg) / 2), box)
    t0 = time.perf_counter()
    oracle.radial_histogram(pos_g[:na], pos_g, GR["n_bins"], (0.0, float(Lg) / 2), box, exclusion=(1, 1))
    t_gr = time.perf_counter() - t0
    pairs_per_s = na * N_PART / t_gr

    t_frame = nq_full * N_PART / terms_per_s + float(N_PART) ** 2 / pairs_per_s
    return {
        "value": 1.0 / t_frame,
        "unit": UNIT,
        "cores": cores,
        "kind": "port",
        "sample": f"S(q): {nq} of {nq_full} wavevectors x 1e5 particles in {t_sq:.2f} s "
                  f"({terms_per_s / 1e6:.0f} M terms/s); g(r): {na} x 1e5 ordered pairs in {t_gr:.2f} s "
                  f"({pairs_per_s / 1e6:.0f} M pairs/s); one frame = 2.14e11 terms + 1e10 ordered pairs",
        "sq_frames_per_s": terms_per_s / (nq_full * N_PART),
        "gr_frames_per_s": pairs_per_s / float(N_PART) ** 2,
    }


def run_reference(args, rank, emit=print):
    """--impl reference: the reference's CPU path (oracle port; /root/reference does not exist on the
    GPU box and needs MDAnalysis), all host threads, each step a bounded sample of the workload."""
    if rank != 0:
        return
    from oracle import oracle

    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core it is allowed to
    oracle.set_num_threads(len(os.sched_getaffinity(0)))
    vals = []
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline()
        if i >= args.warmup:
            vals.append(cb)
    v = float(np.mean([c["value"] for c in vals]))
    cb = dict(vals[-1], value=v)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * args.frames / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.frames, args.gpus),
        "cpu_baseline": cb, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))


# ----------------------------------------------------------------------------- native arm


def run_native(args, rank, world, emit=print):
    import torch

    from mdcraft_b200 import _capi

    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # one context (= one pair of streams) per analysis: g(r) is issue bound, the NUFFT L2 / shared-memory bound, and
    # on separate streams the device runs them side by side (the analysis classes do the same through their lanes)
    ctx = _capi.Context(local)
    ctx_s = ctx if args.one_stream else _capi.Context(local, high_priority=True)
    F = args.frames
    gr_host, box_g, sq_host, Ls = make_frames(F, rank)
    Lg = float(box_g[0])

    rdf = _capi.RdfPlan(ctx, GR["n_bins"], (0.0, Lg / 2), N_PART, None, (1, 1))
    if ctx_s is not ctx:
        rdf.set_blocks_per_sm(3)   # of the 4 that fit: the free quarter of every SM hosts the S(q) kernels
    sqp = _capi.SqPlan(ctx_s, [N_PART], [(None, None)], n_points=SQ["n_points"], L0=[Ls] * 3, q_max=SQ["q_max"],
                       algo=args.sq_algo)
    nq = sqp.n_wavevectors
    gr_dev = torch.from_numpy(gr_host).cuda()
    sq_dev = torch.from_numpy(sq_host).cuda()
    gr_pin = torch.from_numpy(gr_host).pin_memory()
    sq_pin = torch.from_numpy(sq_host).pin_memory()
    # the step's S(q) result is what StructureFactor.results.ssf holds: the |q|-shell means (unique=True), formed on the
    # device; the windows are per-plan bookkeeping, set up once outside the timed region
    from mdcraft_b200.analysis.structure import shell_windows

    wavenumbers = np.linalg.norm(sqp.wavevectors(), axis=1)
    unique_q = np.unique(wavenumbers.round(11))
    sqp.set_shells(*shell_windows(wavenumbers, unique_q))
    n_shells = int(unique_q.shape[0])
    ssf_pin = torch.empty((1, n_shells), dtype=torch.float64).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        ctx.synchronize()
        ctx_s.synchronize()
        if world > 1:
            torch.distributed.barrier()

    def step(device_resident):
        if device_resident:
            rdf.push(gr_dev, None, box_g, n_frames=F)
            sqp.push(sq_dev, n_frames=F)
        else:
            rdf.push(gr_pin.numpy(), None, box_g, n_frames=F)
            sqp.push(sq_pin.numpy(), n_frames=F)
            rdf.read()
            sqp.read_shells(float(N_PART), out=ssf_pin.numpy())

    def timed(device_resident, steps):
        """Device time of `steps` steps: CUDA events on the library's compute stream around each step
        (L2 flushed in between, outside the events), wall clock around the whole region as a cross-check."""
        launches, dev_ms, sq_ms, gr_ms = 0, 0.0, 0.0, 0.0
        barrier()
        w0 = time.perf_counter()
        for _ in range(steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            ctx.mark(0)
            ctx_s.mark(0)
            step(device_resident)
            ctx.mark(1)
            ctx_s.mark(1)
            # both start events are recorded on idle streams (barrier above): the step ends when the later one does
            dev_ms += max(ctx.elapsed_ms(), ctx_s.elapsed_ms()) if ctx_s is not ctx else ctx.elapsed_ms()
            a, la = rdf.last_push_stats()
            b, lb = sqp.last_push_stats()
            gr_ms, sq_ms, launches = gr_ms + a, sq_ms + b, launches + la + lb
        barrier()
        wall = time.perf_counter() - w0
        return dev_ms, wall, launches, sq_ms, gr_ms

    for _ in range(max(args.warmup, 3)):
        step(True)
    barrier()
    with ClockSampler(local) as clk:
        dev_ms, wall, launches, sq_ms, gr_ms = timed(True, args.steps)
    clocks = clk.summary()
    for _ in range(2):
        step(False)
    e2e_ms, e2e_wall, _, _, _ = timed(False, args.steps)

    # max over ranks
    t = torch.tensor([dev_ms, e2e_ms, sq_ms, gr_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    dev_ms, e2e_ms, sq_ms, gr_ms = t.tolist()

    total_frames = F * args.steps * world
    value = total_frames / (dev_ms * 1e-3)
    e2e = total_frames / (e2e_ms * 1e-3)

    def alone(push, plan, reps=3):
        """Device time (ms per push of F frames) of one analysis' dominant kernels with the other analysis idle."""
        ms = 0.0
        for _ in range(reps):
            flush.fill_(1)
            torch.cuda.synchronize()
            push()
            ctx.synchronize()
            ctx_s.synchronize()
            ms += plan.last_push_stats()[0]
        return ms / reps

    rdf.set_blocks_per_sm(0)       # by itself g(r) takes the whole SM
    gr_alone_ms = alone(lambda: rdf.push(gr_dev, None, box_g, n_frames=F), rdf)
    sq_alone_ms = alone(lambda: sqp.push(sq_dev, n_frames=F), sqp)

    # secondary figure: the reference's DEFAULT grid (n_points = 32, Nq = 32,768) on the same frames, device resident
    sq_default_fps = None
    if not args.no_extras:
        sqd = _capi.SqPlan(ctx, [N_PART], [(None, None)], n_points=32, L0=[Ls] * 3, algo=args.sq_algo)
        rep = 16
        for _ in range(3):
            sqd.push(sq_dev, n_frames=F)
        ctx.synchronize()
        ctx.mark(0)
        for _ in range(rep):
            sqd.push(sq_dev, n_frames=F)
        ctx.mark(1)
        sq_default_fps = F * rep / (ctx.elapsed_ms() * 1e-3)
        sqd.close()

    sm = ctx.info()["sm_count"]
    used = sqp.info()

    def term_roofline(algo, ms, frames):
        """Issue-rate roofline of one of the O(Nq N) lattice kernels from its device time over `frames` frames."""
        ach = float(nq) * N_PART * frames / (ms * 1e-3)
        if algo == "direct":
            pk = ctx.probe_rate("mufu_sincos_pairs")      # measured on this box, this run
            return {"bound": "sfu", "kernel": "sq_lattice_direct", "achieved": ach / 1e9, "peak": pk / 1e9,
                    "unit": "G(q,particle) terms/s", "frac": ach / pk, "frames_per_s": frames * 1e3 / ms,
                    "traffic": 5.8e7,
                    "peak_source": f"measured in this run: MUFU.SIN+MUFU.COS pairs/s (mdc_probe_rate); nominal {sm} SMs x 8 pairs/clk",
                    "algorithmic": "2 SFU ops + ~10 FP32/INT issue slots per term"}
        pk = ctx.probe_rate("ffma") / 4.0                 # 4 FP32 FMAs (2 packed FFMA2) per term
        return {"bound": "fp32", "kernel": "sq_lattice_factored", "achieved": ach / 1e9, "peak": pk / 1e9,
                "unit": "G(q,particle) terms/s", "frac": ach / pk, "frames_per_s": frames * 1e3 / ms,
                "traffic": 6.8e8,
                "peak_source": (f"measured in this run: FFMA/s / 4 (mdc_probe_rate); nominal {sm} SMs x 128 FMA/clk / 4; "
                                f"FFMA2 in this kernel's operand pattern sustains {ctx.probe_rate('ffma2_loop') / 2e9:.0f} G terms/s"),
                "algorithmic": "one complex FMA = 4 FP32 FMAs (2 x fma.rn.f32x2) per term; 56 table phasors per 4096 terms"}

    def one_frame(algo):
        """One frame of the full grid through an O(Nq N) kernel (L2 flushed first): its device time in ms."""
        sqx = _capi.SqPlan(ctx, [N_PART], [(None, None)], n_points=SQ["n_points"], L0=[Ls] * 3, q_max=SQ["q_max"],
                           algo=algo)
        sqx.push(sq_dev[:1], n_frames=1)
        ctx.synchronize()
        flush.fill_(1)
        torch.cuda.synchronize()
        sqx.push(sq_dev[:1], n_frames=1)
        ctx.synchronize()
        ms, _ = sqx.last_push_stats()
        sqx.close()
        return ms

    # secondary figures: the kernels the NUFFT replaced, on the same full grid against their issue-rate rooflines
    # (direct = the contract kernel of BASELINE.json's north_star: one fused sin/cos per term on the SFU)
    direct = factored = None
    if rank == 0 and not args.no_extras:
        if used["algo"] != "direct":
            direct = term_roofline("direct", one_frame("direct"), 1)
        if used["algo"] != "factored":
            factored = term_roofline("factored", one_frame("factored"), 1)

    # the once-per-run exchange step: all-reduce of the accumulators over NVLink (outside the timed steps)
    if world > 1:
        from mdcraft_b200 import distributed as mdist

        mdist.all_reduce_device(rdf.accumulator(), device=local)
        mdist.all_reduce_device(sqp.accumulator(), device=local)

    if rank == 0:
        frames_rank = F * args.steps
        if used["algo"] == "nufft":
            # HBM roofline of the NUFFT chain (memset, spread, three pruned FFT passes, pair accumulation):
            # algorithmic bytes = every buffer written once and read once, 16 B per complex FP64
            nf, n = used["nf"], SQ["n_points"]
            ncp = (n + 1) & ~1
            alg = 16.0 * (nf ** 3                                   # clear the fine grid
                          + 2 * nf ** 3                            # spread: read-modify-write of the fine grid
                          + nf ** 3 + nf * nf * ncp                # z pass
                          + nf * nf * ncp + nf * n * ncp           # y pass
                          + nf * n * ncp + nq                      # x pass + deconvolution
                          + 2 * nq)                                # rho -> |rho|^2 accumulated into ssf
            peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
            if os.path.exists(peaks_file):
                peak_gbs, peak_src = float(json.load(open(peaks_file))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy bandwidth, burst)"
            else:
                peak_gbs, peak_src = 6650.0, "of fallback (B200_PROFILING.md): 6.65 TB/s"
            ach = alg * F / (sq_alone_ms * 1e-3) / 1e9
            roof_sq = {"bound": "hbm", "measured": "S(q) alone (in the timed steps these kernels share the device with g(r))",
                       "kernel": "nufft_spread + nufft_fft_z + nufft_fft_s<0> + nufft_fft_s<1> (+ memset, sq_pair_accum)",
                       "achieved": ach, "peak": peak_gbs, "unit": "GB/s", "frac": ach / peak_gbs,
                       "traffic": NUFFT_DRAM_BYTES_PER_FRAME, "peak_source": peak_src,
                       "algorithmic": f"{alg / 1e6:.0f} MB per frame: fine grid {nf}^3 complex FP64 cleared, spread into (w = {used['w']}: "
                                      f"{used['w'] ** 3} FP64 complex reductions per particle, L2 resident) and read once; pruned "
                                      f"intermediates {nf}^2 x {n} and {nf} x {n}^2 written and read once; {nq} densities"}
        else:
            roof_sq = term_roofline(used["algo"], sq_alone_ms, F)
        pairs = 0.5 * N_PART * (N_PART - 1) * F * args.steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 (S(q): NUFFT grid, FFTs, sums) + f32 filter / f64 exact (g(r) distances)",
            "data": "synthetic", "config": workload_config(F, world), "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT,
                    "h2d_bytes_per_step": int(gr_host.nbytes + sq_host.nbytes + F * 24),
                    "d2h_bytes_per_step": int(GR["n_bins"] * 8 + n_shells * 8)},
            "gpu_launches": launches,
            "sq_algo": args.sq_algo, "sq_algo_used": used,
            "roofline_sq": roof_sq,
            "roofline_sfu_direct": direct,
            "roofline_fp32_factored": factored,
            # each analysis by itself (the other idle), and the kernel time each spent inside the timed steps
            "sq_frames_per_s": F / (sq_alone_ms * 1e-3),
            "sq_default_grid_frames_per_s": sq_default_fps,
            "gr_frames_per_s": F / (gr_alone_ms * 1e-3),
            "gr_pairs_per_s": 0.5 * N_PART * (N_PART - 1) * F / (gr_alone_ms * 1e-3),
            "streams": "one" if ctx_s is ctx else "g(r) and S(q) on separate streams (side by side)",
            "timed_kernel_ms_per_step": {"gr": gr_ms / args.steps, "sq": sq_ms / args.steps},
            "wall_s": wall,
        }
        # g(r): SURVEY.md 8(d) fixes the algorithmic cost of one pair-distance at ~20 FP32/INT-pipe instructions (3 sub,
        # 3 convert, r^2, sqrt, bin, range test, increment); the ceiling is the issue rate, 4 schedulers x 32 lanes per SM
        # per clock, divided by that.  The kernel's own count is reported next to it and is NOT the denominator: it was
        # 19.9 at the start of the round and is 15.1 in the hot loop now (SASS: 121 executed per 8 pairs), 15.6 per
        # algorithmic pair over the whole kernel (tile staging, culling tests, undecided-pair path; 19 % of the pairs are
        # culled and never evaluated; ncu, profiles/rdf_fast_r01.txt).
        SLOTS = 20.0
        gr_peak = sm * 4 * 32 * (clocks["sm_mhz"] or 1965.0) * 1e6 / SLOTS
        line["roofline_gr"] = {
            "bound": "issue", "kernel": "rdf_pairs_fast", "achieved": pairs / (gr_ms * 1e-3) / 1e9,
            "peak": gr_peak / 1e9, "unit": "G pair-distances/s", "frac": pairs / (gr_ms * 1e-3) / gr_peak,
            "frac_alone": 0.5 * N_PART * (N_PART - 1) * F / (gr_alone_ms * 1e-3) / gr_peak,
            "traffic": 3.1e6, "peak_source": f"{sm} SMs x 4 schedulers x 32 lanes x SM clock / {SLOTS:g} issue slots per pair "
                                             "(SURVEY.md 8(d) algorithmic cost)",
            "algorithmic": "20 FP32/INT issue slots + 1 SFU op (sqrt) + 1 shared-memory atomic per pair-distance "
                           "(SURVEY.md 8(d)); 24 B of positions per particle per frame (f32 + u32 fixed point)",
            "kernel_slots_per_pair": {"hot_loop_sass": 15.1, "whole_kernel_ncu_per_algorithmic_pair": 15.6},
        }
        # the headline roofline: the kernel with the largest share of the step's WORK (device time of each analysis by
        # itself; side by side on two streams their durations overlap and stretch, so shares of the wall time do not
        # add up).  Its `achieved` / `frac` are those of the timed steps, `frac_alone` that of the kernel alone.
        share_gr = gr_alone_ms / (gr_alone_ms + sq_alone_ms)
        line["roofline"] = dict(line["roofline_gr"] if share_gr >= 0.5 else roof_sq)
        line["roofline"]["share_of_step"] = max(share_gr, 1.0 - share_gr)
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline()
        emit(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=2, help="frames per step per GPU")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--one-stream", action="store_true",
                    help="queue g(r) and S(q) on ONE context (the two analyses then run back to back, not side by side)")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the secondary figures (default grid, one frame each through the direct / factored S(q) "
                         "kernels): only the timed steps launch kernels, which is what the ncu launch list is taken from")
    ap.add_argument("--sq-algo", default="auto", choices=["auto", "direct", "factored", "nufft"],
                    help="S(q) lattice kernel: auto (the plan's cost model: NUFFT on this grid), or force one")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    # stdout carries the ONE JSON line and nothing else: while the run is in progress, file descriptor 1 points at
    # stderr, so that banners written by native libraries (e.g. "NCCL version ...") cannot precede the line
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    lines = []
    emit = lines.append
    try:
        if args.impl == "reference":
            run_reference(args, rank, emit)
        else:
            run_native(args, rank, world, emit)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    for line in lines:
        print(line, flush=True)


if __name__ == "__main__":
    main()
