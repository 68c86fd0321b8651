#!/usr/bin/env python
"""Multi-GPU check of the sharded hot path on real devices (NCCL), launched by torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py

Every rank runs the fused kernels on its contiguous shard; rank 0 additionally runs the whole batch on one GPU and
the shards (gathered over NCCL) must match it BIT FOR BIT (the Philox counters are global row indices).  Also checks
the loss / statistics reductions, the sharded MMD (one all-reduce of three doubles), a DDP training step and the
sharded reverse loop.  Not a pytest file (needs >= 2 GPUs); the CPU tier covers the same host logic on gloo
(tests/test_parallel_gloo.py).  Prints one JSON line on rank 0 and exits non-zero on any mismatch."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import diffusion_extensions_b200 as dx  # noqa: E402
from diffusion_extensions_b200 import parallel as P  # noqa: E402


def gather_rows(x, n_global, world):
    """all_gather of ragged contiguous shards -> the global tensor (on every rank)."""
    sizes = [P.shard_bounds(n_global, r, world) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros(pad, *x.shape[1:], device=x.device, dtype=x.dtype)
    buf[: x.shape[0]] = x
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)])


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    rank, world = P.init_from_env(device=dev)
    assert world >= 2, "launch with torchrun on >= 2 GPUs"
    dx._lib.load()
    n = 100_003  # ragged shards on purpose
    g = torch.Generator(device=dev).manual_seed(1234)  # identical global data on every rank
    x0 = dx.ops.quat_to_rmat(torch.randn(n, 4, device=dev, generator=g))
    t = torch.randint(0, 1000, (n,), device=dev, generator=g)
    pred = torch.randn(n, 3, device=dev, generator=g) * 0.3
    report = {"world": world, "n": n}

    proc = dx.SO3Diffusion(None).to(dev)
    lo, hi = P.attach(proc, n)
    fwd, post, t_range = proc.tables()
    fg, pg = proc.guides()
    a = (proc.sqrt_alphas_cumprod, proc.sqrt_one_minus_alphas_cumprod, fwd)
    sched = (proc.sqrt_recip_alphas_cumprod, proc.sqrt_recipm1_alphas_cumprod, proc.posterior_mean_coef1, proc.posterior_mean_coef2)

    # ---- forward noising + target, reverse step (shared and per-row t), sampler: shards == single GPU, bit for bit
    mine = dx.ops.q_sample_fused(x0[lo:hi], t[lo:hi], *a, seed=7, rng_offset=3, row_offset=lo, guide=fg)
    full = dx.ops.q_sample_fused(x0, t, *a, seed=7, rng_offset=3, row_offset=0, guide=fg)
    ok = torch.equal(gather_rows(mine["x_t"], n, world), full["x_t"]) and torch.equal(gather_rows(mine["target"], n, world), full["target"])
    report["q_sample_bit_identical"] = bool(ok)
    for name, tt, tl in (("shared_t", t_range[500:501], t_range[500:501]), ("per_row_t", t, t[lo:hi])):
        mine_p = dx.ops.p_sample_fused(x0[lo:hi], pred[lo:hi], tl, *sched, post_cdf=post, post_guide=pg, seed=8, rng_offset=5, row_offset=lo)
        full_p = dx.ops.p_sample_fused(x0, pred, tt, *sched, post_cdf=post, post_guide=pg, seed=8, rng_offset=5, row_offset=0)
        report[f"p_sample_{name}_bit_identical"] = bool(torch.equal(gather_rows(mine_p, n, world), full_p))
    se3 = dx.SE3Diffusion(None).to(dev)
    s0 = torch.randn(n, 3, device=dev, generator=g) * 10
    m3 = dx.ops.se3_q_sample_fused(x0[lo:hi], s0[lo:hi], t[lo:hi], *a, 75.0, seed=9, rng_offset=1, row_offset=lo, guide=fg)
    f3 = dx.ops.se3_q_sample_fused(x0, s0, t, *a, 75.0, seed=9, rng_offset=1, guide=fg)
    report["se3_q_sample_bit_identical"] = bool(all(torch.equal(gather_rows(m3[k], n, world), f3[k]) for k in ("rot", "shift", "target_rot", "target_shift")))

    # ---- score evaluation on shards (no randomness): identical to the single-GPU launch
    eps = torch.exp(torch.empty(n, device=dev).uniform_(-5.0, 0.0, generator=g))
    lp_m, sc_m, _ = dx.ops.igso3_logp_score(x0[lo:hi], eps[lo:hi], mode="auto")
    lp_f, sc_f, _ = dx.ops.igso3_logp_score(x0, eps, mode="auto")
    report["score_bit_identical"] = bool(torch.equal(gather_rows(lp_m, n, world), lp_f) and torch.equal(gather_rows(sc_m, n, world), sc_f))

    # ---- reductions: global loss over ragged shards == single-process mse; rotation statistics
    y = torch.randn(n, 3, device=dev, generator=g)
    loss = P.global_loss((pred[lo:hi] - y[lo:hi]) ** 2)
    report["global_loss_abs_err"] = abs(float(loss) - float(torch.nn.functional.mse_loss(pred, y)))
    st = P.rotation_statistics(x0[lo:hi])
    report["stats_count_ok"] = int(st["count"]) == n

    # ---- MMD over the ranks (tile pairs dealt round-robin, three doubles all-reduced) == single-GPU MMD
    X, Y = x0[:20000], full["x_t"][:17000]
    mm = float(P.mmd_sharded(X, Y))
    single = float(P.mmd_from_sums(dx.ops.pair_kernel_sums(X, Y), len(X), len(Y)))
    report["mmd_sharded"] = mm
    report["mmd_abs_err_vs_single"] = abs(mm - single)

    # ---- one DDP training step of a small denoiser on the shard + the sharded reverse loop (20 steps)
    net = torch.nn.Sequential(torch.nn.Linear(9, 64), torch.nn.SiLU(), torch.nn.Linear(64, 3)).to(dev)

    class Denoiser(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.net = net

        def forward(self, x, tt):
            return self.net(x.flatten(-2))

    proc.denoise_fn = P.wrap_denoiser(Denoiser().to(dev), dev)
    opt = torch.optim.Adam(proc.denoise_fn.parameters(), lr=1e-3)
    m = n - n % world
    lo2, hi2 = P.shard_bounds(m, rank, world)
    proc.row_offset = lo2
    l0 = float(P.train_step_sharded(proc, x0[lo2:hi2], opt))
    l1 = float(P.train_step_sharded(proc, x0[lo2:hi2], opt))
    w = torch.cat([p.detach().flatten() for p in net.parameters()])
    w0 = w.clone()
    dist.broadcast(w0, 0)
    report["ddp_weights_in_sync"] = bool(torch.equal(w, w0))
    report["train_losses"] = [l0, l1]
    xs = P.sample_sharded(proc, 8192, steps=20)
    report["sample_sharded_finite"] = bool(torch.isfinite(xs).all()) and xs.shape[0] == P.shard_bounds(8192, rank, world)[1] - P.shard_bounds(8192, rank, world)[0]

    flags = [v for k, v in report.items() if isinstance(v, bool)]
    good = all(flags) and report["global_loss_abs_err"] < 1e-6 and report["mmd_abs_err_vs_single"] < 1e-12 and all(map(lambda v: v == v, report["train_losses"]))
    tot = torch.tensor([1.0 if good else 0.0], device=dev)
    dist.all_reduce(tot, op=dist.ReduceOp.MIN)
    report["ok"] = bool(tot.item() == 1.0)
    if rank == 0:
        print(json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if report["ok"] else 1)


if __name__ == "__main__":
    main()
