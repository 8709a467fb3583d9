#!/usr/bin/env python
"""On-box microbenchmarks behind the issue-rate rooflines: writes gpurun_out/probe.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdcraft_b200 import _capi  # noqa: E402

ctx = _capi.Context(0)
info = ctx.info()
out = {"device": info, "rates_per_s": {}, "per_clk_per_sm_at_max_clock": {}}
clk = info["sm_clock_khz"] * 1e3
for name in _capi.PROBES:
    r = ctx.probe_rate(name)
    out["rates_per_s"][name] = r
    out["per_clk_per_sm_at_max_clock"][name] = r / clk / info["sm_count"]
out["sincos_pipeline"] = ctx.probe_sincos(0)
out["sincos_single_constant"] = ctx.probe_sincos(1)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/probe.json", "w") as fh:
    json.dump(out, fh, indent=1)
print(json.dumps(out, indent=1))
