#!/usr/bin/env python
"""Diagnostic: where does the FP32 S(q) error come from at large N?  (GPU box)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdcraft_b200 import _capi, synthetic
from oracle import oracle

ctx = _capi.Context(0)
n = 50_000
L = synthetic.cubic_box_length(n, 0.8)
for kind in ("lj", "uniform"):
    pos = (synthetic.lj_like_frame(n, L, 7) if kind == "lj" else synthetic.uniform_frame(n, L, 7))[None]
    wv, wn = oracle.wavevector_grid(np.full(3, float(L)), 7, None)
    rho0 = oracle.delta_fourier_transform_sum(wv, pos[0].astype(np.float64))
    for prec, algo in (("fp32", "direct"), ("fp32", "factored"), ("fp64", "auto")):
        plan = _capi.SqPlan(ctx, [n], [(None, None)], n_points=7, L0=[float(L)] * 3, precision=prec, algo=algo)
        plan.push(pos)
        ssf, _ = plan.read()
        s0 = np.abs(rho0) ** 2
        err = ssf[0] - s0
        # implied error in rho: dS = 2 |rho| d|rho|
        drho = err / (2 * np.maximum(np.abs(rho0), 1e-30))
        idx = np.argsort(-np.abs(drho))[:6]
        print(kind, prec, algo, "rms d|rho| =", np.sqrt(np.mean(drho[1:] ** 2)), " expected MUFU-limited ~", 3.46e-7 * np.sqrt(n / 2))
        for i in idx:
            a = np.rint(wv[i] * L / (2 * np.pi)).astype(int)
            print("   q idx", tuple(a), "S/N", s0[i] / n, "d|rho|", drho[i], "rel dS", err[i] / s0[i])
        plan.close()
