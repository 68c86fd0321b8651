"""Device-timed rates of the row-engine kernels (prints one JSON line per op; no asserts).

    python tests/tools/probe_engine.py [log2_rows] [tag]

Used for A/B runs of engine changes (e.g. SO3D_ENGINE=cta vs the default warp-autonomous schedule).
Every op runs on 2^log2_rows rows (default 24: inputs larger than L2), 3 warm-up + 10 timed launches.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import diffusion_extensions_b200 as dx
from diffusion_extensions_b200 import ops

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
tag = sys.argv[2] if len(sys.argv) > 2 else os.environ.get("SO3D_ENGINE", "warp")
n = 1 << lg
dev = torch.device("cuda:0")
torch.manual_seed(0)
HBM = 6551.4
try:
    with open(os.path.join(os.path.dirname(__file__), "..", "..", "MEASURED_PEAKS.json")) as f:
        HBM = float(json.load(f)["hbm_gbs"])
except Exception:
    pass

R = ops.quat_to_rmat(torch.randn(n, 4, device=dev))
R2 = ops.quat_to_rmat(torch.randn(n, 4, device=dev))
v = torch.randn(n, 3, device=dev) * 0.3
eps = torch.exp(torch.empty(n, device=dev).uniform_(-5.05, 0.0))
proc = dx.SO3Diffusion(None).to(dev)
fwd, post, t_range = proc.tables()
fwd_guide, post_guide = proc.guides()
tt = torch.randint(0, 1000, (n,), device=dev)
t1 = t_range[500:501]
pred = torch.zeros(n, 3, device=dev)
sc = torch.rand(n, device=dev) + 0.5


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


sigma = (0.5 * proc.posterior_log_variance_clipped).exp().contiguous()
sched = (proc.sqrt_recip_alphas_cumprod, proc.sqrt_recipm1_alphas_cumprod, proc.posterior_mean_coef1, proc.posterior_mean_coef2)
cases = [
    ("p_sample shared t", 84, lambda: ops.p_sample_fused(R, pred, t1, *sched, post_cdf=post, seed=1, rng_offset=1)),
    ("p_sample per-row t", 92, lambda: ops.p_sample_fused(R, pred, tt, *sched, post_cdf=post, seed=1, rng_offset=1, post_guide=post_guide)),
    ("q_sample per-row t", 92, lambda: ops.q_sample_fused(R, tt, proc.sqrt_alphas_cumprod, proc.sqrt_one_minus_alphas_cumprod, fwd, seed=1, rng_offset=1, guide=fwd_guide)),
    ("q_sample + score", 104, lambda: ops.q_sample_fused(R, tt, proc.sqrt_alphas_cumprod, proc.sqrt_one_minus_alphas_cumprod, fwd, seed=1, rng_offset=1, guide=fwd_guide, want_score=True)),
    ("se3 q_sample per-row t", 128, lambda: ops.se3_q_sample_fused(R, v, tt, proc.sqrt_alphas_cumprod, proc.sqrt_one_minus_alphas_cumprod, fwd, 75.0, seed=1, rng_offset=1, guide=fwd_guide)),
    ("se3 p_sample shared t", 120, lambda: ops.se3_p_sample_fused(R, v, pred, pred, t1, *sched, sigma, 75.0, post_cdf=post, seed=1, rng_offset=1)),
    ("score auto", 56, lambda: ops.igso3_logp_score(R, eps, mode="auto")),
    ("score closed", 56, lambda: ops.igso3_logp_score(R, eps, mode="closed")),
    ("sample shared row", 48, lambda: ops.igso3_sample(fwd, (n,), row=500, seed=1, rng_offset=1)),
    ("log_rmat", 72, lambda: ops.log_rmat(R)),
    ("log_vec", 48, lambda: ops.log_vec(R)),
    ("exp_vec", 48, lambda: ops.exp_vec(v)),
    ("so3_scale", 76, lambda: ops.so3_scale(R, sc)),
    ("compose", 108, lambda: ops.compose(R, R2)),
    ("rmat_dist", 76, lambda: ops.rmat_dist(R, R2)),
]
for name, bytes_per_row, fn in cases:
    try:
        ms = timeit(fn)
    except Exception as e:  # keep going: this is a diagnostic
        print(json.dumps({"tag": tag, "op": name, "error": str(e)[:200]}))
        continue
    gbs = n * bytes_per_row / (ms * 1e-3) / 1e9
    print(json.dumps({"tag": tag, "op": name, "rows": n, "ms": round(ms, 4), "rows_per_s": n / (ms * 1e-3), "bytes_per_row": bytes_per_row,
                      "GBps": round(gbs, 1), "frac_hbm": round(gbs / HBM, 3)}))
