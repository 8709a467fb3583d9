"""GPU: device time of every stage of the NUFFT chain while the g(r) pair kernel runs next to it (3 blocks per SM), and alone.
usage: MDC_NUFFT_PROFILE=1 python tools/confined_profile.py [alone]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdcraft_b200 import _capi, synthetic  # noqa: E402

N, F = 100_000, 48
alone = len(sys.argv) > 1 and sys.argv[1] == "alone"
ctx, ctx_s = _capi.Context(0), _capi.Context(0, high_priority=True)
Lg, Ls = float(synthetic.cubic_box_length(N, 2.5)), float(synthetic.cubic_box_length(N, 0.8))
g = torch.from_numpy(np.stack([synthetic.uniform_frame(N, Lg, 2000 + i) for i in range(F)])).cuda()
s = torch.from_numpy(np.stack([synthetic.lj_like_frame(N, Ls, 3000 + i) for i in range(F)])).cuda()
box = np.array([Lg, Lg, Lg, 90, 90, 90], np.float32)
rdf = _capi.RdfPlan(ctx, 200, (0.0, Lg / 2), N, None, (1, 1))
rdf.set_blocks_per_sm(3)
sq = _capi.SqPlan(ctx_s, [N], [(None, None)], n_points=161, L0=[Ls] * 3, q_max=20.0)
for rep in range(3):
    ctx.mark(0)
    ctx_s.mark(0)
    if not alone:
        rdf.push(g, None, box, n_frames=F)
    sq.push(s, n_frames=F)
    ctx.mark(1)
    ctx_s.mark(1)
    ctx.synchronize()
    ctx_s.synchronize()
    print("rep", rep, "g(r) stream ms", round(ctx.elapsed_ms(), 2), "S(q) stream ms", round(ctx_s.elapsed_ms(), 2),
          "kernels", rdf.last_push_stats()[0] if not alone else None, sq.last_push_stats()[0])
