"""One op of probe_engine's list, a few launches (for ncu captures):  python tests/tools/probe_one.py q_sample|q_sample_score|p_sample|p_sample_rows|auto|closed|loop [log2_rows]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import diffusion_extensions_b200 as dx
from diffusion_extensions_b200 import ops

what = sys.argv[1]
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 22)
dev = torch.device("cuda:0")
torch.manual_seed(0)
R = ops.quat_to_rmat(torch.randn(n, 4, device=dev))
proc = dx.SO3Diffusion(None).to(dev)
fwd, post, t_range = proc.tables()
fwd_guide, post_guide = proc.guides()
tt = torch.randint(0, 1000, (n,), device=dev)
pred = torch.zeros(n, 3, device=dev)
eps = torch.exp(torch.empty(n, device=dev).uniform_(-5.05, 0.0))
sched = (proc.sqrt_recip_alphas_cumprod, proc.sqrt_recipm1_alphas_cumprod, proc.posterior_mean_coef1, proc.posterior_mean_coef2)
fns = {
    "q_sample": lambda: ops.q_sample_fused(R, tt, proc.sqrt_alphas_cumprod, proc.sqrt_one_minus_alphas_cumprod, fwd, seed=1, rng_offset=1, guide=fwd_guide),
    "q_sample_score": lambda: ops.q_sample_fused(R, tt, proc.sqrt_alphas_cumprod, proc.sqrt_one_minus_alphas_cumprod, fwd, seed=1, rng_offset=1, guide=fwd_guide,
                                                 want_score=True),
    "p_sample": lambda: ops.p_sample_fused(R, pred, t_range[500:501], *sched, post_cdf=post, seed=1, rng_offset=1),
    "p_sample_rows": lambda: ops.p_sample_fused(R, pred, tt, *sched, post_cdf=post, seed=1, rng_offset=1, post_guide=post_guide),
    "auto": lambda: ops.igso3_logp_score(R, eps, mode="auto"),
    "closed": lambda: ops.igso3_logp_score(R, eps, mode="closed"),
    "series_small": lambda: ops.igso3_logp_score(R, eps, mode="series", L=2000),      # n = 2^12: one warp per rotation
    # the one-launch reverse process, 40 steps (an ncu replay of all 1000 would take minutes)
    "loop": lambda: ops.p_sample_loop_fused(R, None, 539, 500, *sched, post, post_guide, seed=1, rng_offset=0),
}
for _ in range(2 if what == "loop" else 6):
    fns[what]()
torch.cuda.synchronize()
