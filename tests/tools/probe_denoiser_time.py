"""Device-timed RotPredict-fused reverse step on 2^24 particles:  python tests/tools/probe_denoiser_time.py [log2_rows]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import diffusion_extensions_b200 as dx
from diffusion_extensions_b200 import ops

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 24)
dev = torch.device("cuda:0")
torch.manual_seed(1234)
net = dx.RotPredict(out_type="skewvec").to(dev)
proc = dx.SO3Diffusion(net).to(dev)
proc.tables()
x = ops.quat_to_rmat(torch.randn(n, 4, device=dev))
t = proc.tables()[2][500:501]
with torch.no_grad():
    for _ in range(3):
        proc.p_sample(x, t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        proc.p_sample(x, t)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(json.dumps({"lib": os.path.basename(os.environ.get("SO3D_LIB_PATH", "shipped")), "rows": n, "ms_per_step": round(ms, 4), "particle_steps_per_s": n / (ms * 1e-3),
                  "frac_of_xu_floor_1p95ms": round(1.95 * (n / (1 << 24)) / ms, 3)}))
