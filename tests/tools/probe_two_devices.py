"""ONE process driving two GPUs (the usual layout is one process per GPU): the kernels' per-device attributes
(dynamic shared-memory limit, occupancy, SM count) must be configured on each device the process touches.
Runs the large-shared-memory ops on cuda:0, then on cuda:1, and compares the results bit for bit.  Exits non-zero on mismatch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import diffusion_extensions_b200 as dx
from diffusion_extensions_b200 import ops

assert torch.cuda.device_count() >= 2, "needs two GPUs"
torch.manual_seed(0)
n = 5000
q = torch.randn(n, 4)
pred = torch.randn(n, 3) * 0.3
net_cpu = dx.RotPredict(out_type="skewvec")
outs = []
for d in (0, 1, 0):
    dev = torch.device("cuda", d)
    net = dx.RotPredict(out_type="skewvec").to(dev)
    net.load_state_dict(net_cpu.state_dict())
    proc = dx.SO3Diffusion(net).to(dev)
    x = ops.quat_to_rmat(q.to(dev))
    _, post, t_range = proc.tables()
    sched = (proc.sqrt_recip_alphas_cumprod, proc.sqrt_recipm1_alphas_cumprod, proc.posterior_mean_coef1, proc.posterior_mean_coef2)
    a = ops.p_sample_fused(x, pred.to(dev), t_range[500:501], *sched, post_cdf=post, seed=3, rng_offset=1)        # 53 KB of shared memory
    blob, c1 = net.packed(proc.num_timesteps)
    b = ops.rotpredict_p_sample_fused(x, blob, c1, t_range[500:501], *sched, post_cdf=post, seed=3, rng_offset=1)  # 170 KB
    c = ops.q_sample_fused(x, torch.full((n,), 321, device=dev), proc.sqrt_alphas_cumprod, proc.sqrt_one_minus_alphas_cumprod, proc.tables()[0],
                           seed=3, rng_offset=2, guide=proc.guides()[0])["x_t"]
    torch.cuda.synchronize(dev)
    outs.append((a.cpu(), b.cpu(), c.cpu()))
ok = all(torch.equal(outs[0][k], outs[1][k]) and torch.equal(outs[0][k], outs[2][k]) for k in range(3))
print({"two_devices_one_process_ok": ok})
sys.exit(0 if ok else 1)
