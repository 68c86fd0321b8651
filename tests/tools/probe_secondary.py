"""Device time of the secondary per-row kernels (per-row sampler, SE(3) per-row-t reverse step, noising with the noise
output, reverse step with the x0_hat output) -- A/B of SO3D_ROW_LANES=1 / 2:  python tests/tools/probe_secondary.py [log2_rows] [tag]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import diffusion_extensions_b200 as dx
from diffusion_extensions_b200 import ops

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
tag = sys.argv[2] if len(sys.argv) > 2 else "default"
n = 1 << lg
dev = torch.device("cuda:0")
torch.manual_seed(0)
R = ops.quat_to_rmat(torch.randn(n, 4, device=dev))
v = torch.randn(n, 3, device=dev) * 0.3
proc = dx.SO3Diffusion(None).to(dev)
fwd, post, t_range = proc.tables()
fg, pg = proc.guides()
tt = torch.randint(0, 1000, (n,), device=dev)
sched = (proc.sqrt_recip_alphas_cumprod, proc.sqrt_recipm1_alphas_cumprod, proc.posterior_mean_coef1, proc.posterior_mean_coef2)
sig = (0.5 * proc.posterior_log_variance_clipped).exp().contiguous()
cases = [
    ("sample per-row", 56, lambda: ops.igso3_sample(fwd, (n,), row_idx=tt, seed=1, rng_offset=3, guide=fg)),
    ("se3 p_sample per-row t", 128, lambda: ops.se3_p_sample_fused(R, v, v, v, tt, *sched, sig, 75.0, post_cdf=post, seed=1, rng_offset=5, post_guide=pg)),
    ("q_sample + noise + score", 140, lambda: ops.q_sample_fused(R, tt, proc.sqrt_alphas_cumprod, proc.sqrt_one_minus_alphas_cumprod, fwd, seed=1, rng_offset=1, guide=fg, want_noise=True, want_score=True)),
    ("p_sample per-row t + x0_hat", 128, lambda: ops.p_sample_fused(R, v, tt, *sched, post_cdf=post, seed=1, rng_offset=2, post_guide=pg, want_x0_hat=True)),
]
for name, bytes_per_row, fn in cases:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({"tag": tag, "op": name, "rows": n, "ms": round(ms, 4), "GBps": round(n * bytes_per_row / ms / 1e6, 1), "frac_hbm": round(n * bytes_per_row / ms / 1e6 / 6551.4, 3)}))
