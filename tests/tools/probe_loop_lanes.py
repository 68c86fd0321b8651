"""A/B of the one-launch reverse process: one row per thread vs two rows per thread with packed FP32 (SO3D_LOOP_LANES=1|2):
python tests/tools/probe_loop_lanes.py [log2_rows] [steps] -> one JSON line per variant (ms, particle-steps/s, checksum)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import json, os, sys, torch
sys.path.insert(0, %r)
import diffusion_extensions_b200 as dx
from diffusion_extensions_b200 import ops
n, steps = 1 << int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
torch.manual_seed(0)
R = ops.quat_to_rmat(torch.randn(n, 4, device=dev))
p = dx.SO3Diffusion(None).to(dev)
_, post, _ = p.tables(); pg = p.guides()[1]
sched = (p.sqrt_recip_alphas_cumprod, p.sqrt_recipm1_alphas_cumprod, p.posterior_mean_coef1, p.posterior_mean_coef2)
out = torch.empty_like(R)
fn = lambda: ops.p_sample_loop_fused(R, None, 999, 1000 - steps, *sched, post, pg, seed=7, rng_offset=0, out=out)
fn(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fn(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(json.dumps({"lanes": os.environ.get("SO3D_LOOP_LANES", "2"), "rows": n, "steps": steps, "ms": round(ms, 3),
                  "particle_steps_per_s": n * steps / (ms * 1e-3), "checksum": float(out.double().sum())}))
''' % ROOT
lg = sys.argv[1] if len(sys.argv) > 1 else "24"
steps = sys.argv[2] if len(sys.argv) > 2 else "100"
for lanes in ("1", "2"):
    env = dict(os.environ, SO3D_LOOP_LANES=lanes)
    p = subprocess.run([sys.executable, "-c", CHILD, lg, steps], env=env, capture_output=True, text=True, timeout=600)
    print(p.stdout.strip() or json.dumps({"lanes": lanes, "error": p.stderr[-400:]}), flush=True)
