"""Diagnostic run of the fused RotPredict + reverse-step kernel on a GPU box (prints, no asserts)."""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

import diffusion_extensions_b200 as dx
from diffusion_extensions_b200 import ops

torch.manual_seed(0)
dev = torch.device("cuda:0")
net = dx.RotPredict(out_type="skewvec").to(dev)
proc = dx.SO3Diffusion(net).to(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
x = ops.quat_to_rmat(torch.randn(n, 4, device=dev))
for tval in (500, 0, 999):
    t = torch.tensor([tval], device=dev)
    blob, c1 = net.packed(proc.num_timesteps)
    pred = ops.rotpredict_p_sample_fused(x, blob, c1, t, proc.sqrt_recip_alphas_cumprod, proc.sqrt_recipm1_alphas_cumprod,
                                         proc.posterior_mean_coef1, proc.posterior_mean_coef2, want_out=False, want_pred=True)
    torch.cuda.synchronize()
    net64 = dx.RotPredict(out_type="skewvec").double()
    net64.load_state_dict({k: v.double().cpu() for k, v in net.state_dict().items()})
    ref = net64(x.double().cpu(), t.cpu().expand(n)).detach()
    ref32 = net(x, t.expand(n)).detach().cpu().double()
    err = (pred.cpu().double() - ref).abs()
    print(f"t={tval} n={n}: |ref| max {ref.abs().max():.4f}  fused err max {err.max():.3e} mean {err.mean():.3e};  torch fp32 err max {(ref32 - ref).abs().max():.3e}")
    if err.max() > 1e-3:
        bad = err.max(dim=1).values.argmax().item()
        print("  worst row", bad, pred[bad].cpu().numpy(), ref[bad].numpy())
        print("  first rows", pred[:3].cpu().numpy(), ref[:3].numpy())
# full step vs the two-kernel path with the same draws
ops.manual_seed(7)
t = torch.tensor([300], device=dev)
a = proc.p_sample(x, t)
ops.manual_seed(7)
proc.fuse_denoiser = False
b = proc.p_sample(x, t)
torch.cuda.synchronize()
d = (a - b).abs().max().item()
print("p_sample fused vs unfused max abs diff", d)

# timing
if len(sys.argv) > 2:
    nb = int(sys.argv[2])
    xb = ops.quat_to_rmat(torch.randn(nb, 4, device=dev))
    t = torch.tensor([500], device=dev)
    for fuse in (True, False):
        proc.fuse_denoiser = fuse
        for _ in range(3):
            y = proc.p_sample(xb, t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10 if fuse else 3
        e0.record()
        for _ in range(reps):
            y = proc.p_sample(xb, t)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"fuse={fuse} n={nb}: {ms:.3f} ms/step  {nb / ms * 1e3:.3e} particle-steps/s")
    with torch.no_grad():
        torch.backends.cuda.matmul.allow_tf32 = True
        proc.fuse_denoiser = False
        for _ in range(2):
            y = proc.p_sample(xb, t)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            y = proc.p_sample(xb, t)
        e1.record()
        torch.cuda.synchronize()
        print(f"unfused, torch allow_tf32: {e0.elapsed_time(e1) / 3:.3f} ms/step")
