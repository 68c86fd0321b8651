"""Shared-t reverse step, device time per launch at several step indices (A/B of guide / table-staging changes):
    [SO3D_LIB_PATH=...] python tests/tools/probe_pstep_t.py [log2_rows]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import diffusion_extensions_b200 as dx
from diffusion_extensions_b200 import ops

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 24)
dev = torch.device("cuda:0")
torch.manual_seed(0)
R = ops.quat_to_rmat(torch.randn(n, 4, device=dev))
pred = torch.zeros(n, 3, device=dev)
proc = dx.SO3Diffusion(None).to(dev)
_, post, t_range = proc.tables()
sched = (proc.sqrt_recip_alphas_cumprod, proc.sqrt_recipm1_alphas_cumprod, proc.posterior_mean_coef1, proc.posterior_mean_coef2)
out = {"lib": os.path.basename(os.environ.get("SO3D_LIB_PATH", "") or "shipped")}
for t in (1, 5, 20, 100, 300, 500, 700, 900, 990, 999):
    fn = lambda: ops.p_sample_fused(R, pred, t_range[t:t + 1], *sched, post_cdf=post, seed=1, rng_offset=t)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    out[f"t{t}_ms"] = round(e0.elapsed_time(e1) / 10, 4)
print(json.dumps(out))
