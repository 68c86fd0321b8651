"""Small, deterministic exercise of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck /
initcheck):  compute-sanitizer --tool racecheck python tests/tools/sanitize_target.py

Ragged and unaligned sizes (both engine schedules, TMA path and the cooperative fallback), the warp-autonomous
schedule's mbarrier / acq_rel stage recycling (per-row sampler, the one-row fallbacks) and the two-row kernels' per-warp
input slices (q_sample, per-row-t reverse step, SE(3) noising), the
shared-t kernels with the CDF row staged in shared memory, the tcgen05 / TMEM denoiser step and its multi-step loop,
the all-pairs MMD kernel, the table builders.  Sizes are tiny: the tools slow kernels down 10-100x."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import diffusion_extensions_b200 as dx
from diffusion_extensions_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
proc = dx.SO3Diffusion(None, timesteps=40).to(dev)          # a 40-step schedule keeps the table builders short
fwd, post, t_range = proc.tables()
fg, pg = proc.guides()
sched = (proc.sqrt_recip_alphas_cumprod, proc.sqrt_recipm1_alphas_cumprod, proc.posterior_mean_coef1, proc.posterior_mean_coef2)
T = proc.num_timesteps
sizes = [1, 255, 257, 1300, 5000]
only = [a for a in sys.argv[1:] if not a.isdigit()]          # sections: rows, denoiser_step, denoiser_loop1, denoiser_loop5
if any(a.isdigit() for a in sys.argv[1:]):
    sizes = [int(a) for a in sys.argv[1:] if a.isdigit()]
want = lambda sec: not only or sec in only
for n in (sizes if want("rows") else []):
    q = torch.randn(n + 1, 4, device=dev)
    Rfull = ops.quat_to_rmat(q)
    for R in (Rfull[:n], Rfull[1:]):                        # 16-byte aligned base / a view starting 36 B in (no TMA: cooperative path)
        R = R if R.data_ptr() % 16 == 0 else R              # (views are passed as they are; ops make them contiguous only if strided)
        v = torch.randn(n, 3, device=dev) * 0.3
        tt = torch.randint(0, T, (n,), device=dev)
        t1 = t_range[T // 2:T // 2 + 1]
        eps = torch.rand(n, device=dev) * 0.9 + 0.05
        ops.log_rmat(R); ops.log_vec(R); ops.rmat_to_aa(R); ops.exp_vec(v); ops.so3_scale(R, eps); ops.compose(R, R, trans_a=True)
        ops.rmat_dist(R, R); ops.rmat_to_quat(R); ops.so3_lerp(R, R, eps)
        ops.igso3_logp_score(R, eps, mode="series", L=96); ops.igso3_logp_score(R, eps, mode="auto"); ops.igso3_logp_score(R, eps, mode="series_adaptive", L=200)
        ops.q_sample_fused(R, tt, proc.sqrt_alphas_cumprod, proc.sqrt_one_minus_alphas_cumprod, fwd, seed=1, rng_offset=1, guide=fg)
        ops.q_sample_fused(R, tt, proc.sqrt_alphas_cumprod, proc.sqrt_one_minus_alphas_cumprod, fwd, seed=1, rng_offset=1, guide=fg, want_noise=True, want_score=True)
        ops.p_sample_fused(R, v, t1, *sched, post_cdf=post, seed=1, rng_offset=2)
        ops.p_sample_fused(R, v, tt, *sched, post_cdf=post, seed=1, rng_offset=2, post_guide=pg, want_x0_hat=True)
        ops.p_sample_fused(R, v, tt, *sched, post_cdf=post, seed=1, rng_offset=2, post_guide=pg)   # two rows per thread, per-warp input slices
        ops.q_sample_fused(R, tt, proc.sqrt_alphas_cumprod, proc.sqrt_one_minus_alphas_cumprod, fwd, seed=1, rng_offset=1, guide=fg, want_score=True)
        ops.igso3_sample(fwd, (n,), row=3, seed=1, rng_offset=3)
        ops.igso3_sample(fwd, (n,), row_idx=tt, seed=1, rng_offset=3, guide=fg)
        sig = (0.5 * proc.posterior_log_variance_clipped).exp().contiguous()
        ops.se3_q_sample_fused(R, v, tt, proc.sqrt_alphas_cumprod, proc.sqrt_one_minus_alphas_cumprod, fwd, 75.0, seed=1, rng_offset=4, guide=fg)
        ops.se3_p_sample_fused(R, v, v, v, t1, *sched, sig, 75.0, post_cdf=post, seed=1, rng_offset=5)
        ops.se3_p_sample_fused(R, v, v, v, tt, *sched, sig, 75.0, post_cdf=post, seed=1, rng_offset=5, post_guide=pg)
        if hasattr(ops, "p_sample_loop_fused"):
            ops.p_sample_loop_fused(R.contiguous(), None, T - 1, 0, *sched, post, pg, seed=1, rng_offset=100)
            ops.p_sample_loop_fused(R.contiguous(), v, T - 1, T - 6, *sched, post, pg, seed=1, rng_offset=100)
    ops.bingham_sample(torch.eye(4, device=dev), (n,), seed=2, rng_offset=0, want_rmat=True)
    ops.pair_kernel_sums(Rfull[:n], Rfull[1:], "gaussian")
    ops.pair_kernel_sums(Rfull[:n], None, "cosine")
torch.cuda.synchronize()
# tcgen05 / TMEM denoiser: one step, and the multi-step loop (5 steps inside one launch)
net = dx.RotPredict(out_type="skewvec").to(dev)
pn = dx.SO3Diffusion(net, timesteps=40).to(dev)
for n in (1, 300, 4099):
    x = ops.quat_to_rmat(torch.randn(n, 4, device=dev))
    blob, c1 = net.packed(pn.num_timesteps)
    if want("denoiser_step"):
        ops.rotpredict_p_sample_fused(x, blob, c1, t_range[20:21], *sched, post_cdf=post, seed=3, rng_offset=7)
        torch.cuda.synchronize()
    if want("denoiser_loop1"):
        ops.rotpredict_p_sample_loop(x, blob, c1, 6, 6, *sched, post, 1234)
        torch.cuda.synchronize()
    if want("denoiser_loop5"):
        ops.rotpredict_p_sample_loop(x, blob, c1, 6, 2, *sched, post, 1234)
        torch.cuda.synchronize()
    print("denoiser n =", n, "ok", flush=True)
torch.cuda.synchronize()
print("sanitize target done")
