#!/usr/bin/env python
"""Benchmark of the SO(3) diffusion hot path (BASELINE.json metric: IGSO(3) score evals/sec &
reverse-diffusion particle-steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n ROWS] [--L 2000]

Headline workload (BASELINE configs[1]): one *step* = one fused launch evaluating log-density and
score of the IGSO(3) truncated series (exactly L = 2000 terms per evaluation, no early exit) on
n = 2^24 random rotations per GPU with per-row eps (E-set of SURVEY 8d).  `value` is device-timed with
the inputs resident in HBM; `e2e` goes through the public Python API from pinned HOST buffers and
back.  Secondary numbers (fused reverse step, fused forward noising, closed-form/auto evaluator)
are reported under "extra" with their own HBM rooflines.  Rank 0 prints ONE JSON line.

--impl reference times the torch-CPU port of the reference's algorithm (oracle/ref_port.py; the
pure-Python reference cannot travel to the GPU box) on the host cores for the same metric.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "igso3_score_evals_per_sec"
UNIT = "evals/s"
SEED = 1234

# ---- work of the series kernel per term ------------------------------------
# Executed: 7 FP32 instructions (1 FMUL, 4 FFMA, 2 FADD = 11 flop) + 1 MUFU.EX2 + 1 uniform constant load
# (LDCU.64) = 9 issue slots.  SURVEY 8(d) fixed the algorithmic unit for recurrence-based variants at
# >= 10 FP32 lane-instructions per term (bound 1.86e9 evals/s/GPU at 128 lanes/clk/SM); `roofline.frac` uses
# that unit, and the executed mix is reported next to it.
ALGO_LANE_INSTR_PER_TERM = 10
FP32_INSTR_PER_TERM = 7
MUFU_PER_TERM = 1
UNIFORM_PER_TERM = 1
FLOP_PER_TERM = 11
BYTES_PER_EVAL = 56          # 36 R + 4 eps in, 4 logp + 12 score out
BYTES_PER_PARTICLE_STEP = 84 # 36 x_t + 12 pred in, 36 out
BYTES_PER_QSAMPLE = 92       # 36 x0 + 8 t in, 36 x_t + 12 target out
# DRAM traffic of the series kernel per evaluation from the ncu --set full capture (profiles/r02p_series_full.md:
# 169.07 MB read + 43.80 MB written for 4 194 304 evaluations; the rest of the 16 B/eval of results is still in L2
# when the kernel ends) -- equal to the algorithmic 40 B/eval of inputs: nothing is re-read.
NCU_DRAM_BYTES_PER_EVAL = (169.070080e6 + 43.798272e6) / 4194304
MMD_LANE_INSTR_PER_PAIR = 41  # executed FP32-pipe instructions per pair of the all-pairs kernel (SASS count)


def peaks():
    p = {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)", "sm_max_mhz": 1965.0}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update(hbm_gbs=float(m["hbm_gbs"]), sm_max_mhz=float(m.get("sm_max_mhz", 1965.0)), source="measured (MEASURED_PEAKS.json)")
    except Exception:
        pass
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def make_eset(n, device, seed):
    """E-set of SURVEY 8(d): eps log-uniform in [6.4e-3, 1], omega = eps*sqrt2*k, k~U[0,4] (<= 3.0)."""
    import diffusion_extensions_b200 as dx

    g = torch.Generator(device=device).manual_seed(seed)
    eps = torch.exp(torch.empty(n, device=device).uniform_(math.log(6.4e-3), 0.0, generator=g))
    k = torch.empty(n, device=device).uniform_(0.0, 4.0, generator=g)
    omega = torch.clamp(eps * math.sqrt(2.0) * k, max=3.0)
    axis = torch.randn(n, 3, device=device, generator=g)
    R = dx.ops.aa_to_rmat(axis, omega)
    return R, eps


def time_loop(fn, steps, warmup, dist_on):
    """W untimed + K timed calls, bracketed by barrier + synchronize; device time via CUDA events."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if dist_on:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    if dist_on:
        torch.distributed.barrier()
    ms = ev0.elapsed_time(ev1)
    if dist_on:
        tt = torch.tensor([ms], device="cuda")
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        ms = tt.item()
    return ms


def cpu_baseline(seconds_target=15.0):
    """The reference's CPU path (torch port, all host threads) on a bounded sample of the same
    workload: IGSO3(eps).log_prob(R) + autograd score, per-row eps, closed form in fp64."""
    from oracle import ref_port as P

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = 1 << 18
    g = torch.Generator().manual_seed(SEED)
    eps = torch.exp(torch.empty(n).uniform_(math.log(6.4e-3), 0.0, generator=g))
    k = torch.empty(n).uniform_(0.0, 4.0, generator=g)
    ang = torch.clamp(eps * math.sqrt(2.0) * k, min=1e-4, max=3.0)
    axis = torch.randn(n, 3, generator=g)
    axis = axis / axis.norm(dim=-1, keepdim=True)
    K = P.hat(axis)
    R = torch.eye(3) + torch.sin(ang)[:, None, None] * K + (1 - torch.cos(ang))[:, None, None] * (K @ K)
    P.score_via_autograd(R, eps)  # warm-up (thread pool, allocator)
    t0 = time.perf_counter()
    reps = 0
    while True:
        P.score_via_autograd(R, eps)
        reps += 1
        dt = time.perf_counter() - t0
        if dt > seconds_target:
            break
    return {"value": n * reps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{reps} x 2^18 rotations, per-row eps, reference closed-form fp64 log_prob + autograd score (oracle/ref_port.py), {dt:.1f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_port as P

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = 1 << 18
    g = torch.Generator().manual_seed(SEED)
    eps = torch.exp(torch.empty(n).uniform_(math.log(6.4e-3), 0.0, generator=g))
    k = torch.empty(n).uniform_(0.0, 4.0, generator=g)
    ang = torch.clamp(eps * math.sqrt(2.0) * k, min=1e-4, max=3.0)
    axis = torch.randn(n, 3, generator=g)
    axis = axis / axis.norm(dim=-1, keepdim=True)
    K = P.hat(axis)
    R = torch.eye(3) + torch.sin(ang)[:, None, None] * K + (1 - torch.cos(ang))[:, None, None] * (K @ K)
    for _ in range(args.warmup):
        P.score_via_autograd(R, eps)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        P.score_via_autograd(R, eps)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    sample = f"each step = 2^18 rotations (bounded sample of the 2^24-row step), per-row eps, reference closed-form fp64 log_prob + autograd score"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": "igso3_logp_score E-set, CPU reference port", "rows_per_step": n, "series_terms": None,
                                         "evaluator": "closed form fp64 (the reference has no series)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_RESULT_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on
    stdout when NCCL_DEBUG is set on the box), so file descriptor 1 is pointed at stderr for the whole run and the
    result line goes to a saved copy of the original stdout."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1 << 24, help="rotations per GPU per step")
    ap.add_argument("--L", type=int, default=2000, help="series truncation")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary kernels")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (profiling runs)")
    ap.add_argument("--no-numa-bind", action="store_true", help="multi-GPU runs: leave the ranks' CPU affinity alone (A/B of the e2e leg)")
    ap.add_argument("--e2e-chunk", type=int, default=1 << 21, help="rows per chunk of the host-buffer pipeline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return run_reference(args)

    import diffusion_extensions_b200 as dx

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist_on = world > 1
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa = None
    if dist_on and not args.no_numa_bind:
        # one process per GPU: keep it (and the pinned staging buffers it is about to allocate) on the GPU's NUMA node
        numa = dx.parallel.bind_to_gpu_numa_node(local)
    if dist_on:
        torch.distributed.init_process_group("nccl", device_id=device)
    dx._lib.load()
    n, L = args.n, args.L
    pk = peaks()

    # ---- inputs: this rank's shard of the global batch (rows are independent: no exchange) -----
    R, eps = make_eset(n, device, SEED + rank)
    logp = torch.empty(n, device=device)
    score = torch.empty(n, 3, device=device)
    lib_call, ptr = dx._lib.call, dx._lib.ptr

    def step_series():
        lib_call("so3d_igso3_logp_score_f32", ptr(R), ptr(eps), 1, ptr(logp), ptr(score), None, n, dx._lib.MODE_SERIES, L, device=device)

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        step_series()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    ms = time_loop(step_series, args.steps, 0, dist_on)
    clocks = sampler.stop() if sampler else None
    evals_per_s = world * n * args.steps / (ms * 1e-3)
    per_gpu = evals_per_s / world

    # ---- e2e: public host-buffer API, pinned host inputs -> host results -------------------------
    # dx.ops.HostScorePipeline is the call a user with host-resident batches makes: it overlaps the H2D copy,
    # the kernel and the D2H copy chunk by chunk.  Staging buffers are allocated once outside the timed
    # region; every timed step moves all n x 40 B in and n x 16 B out across PCIe.
    n_e2e = n
    e2e = None
    e2e_launches = 0
    if not args.no_e2e:
        hR = torch.empty(n_e2e, 3, 3, pin_memory=True).copy_(R[:n_e2e].cpu())
        heps = torch.empty(n_e2e, pin_memory=True).copy_(eps[:n_e2e].cpu())
        hlogp = torch.empty(n_e2e, pin_memory=True)
        hscore = torch.empty(n_e2e, 3, pin_memory=True)
        pipe = dx.ops.HostScorePipeline(device, chunk_rows=args.e2e_chunk, depth=3)

        # a stream of batches: every step moves its own 40 B/row in and 16 B/row out; the pipeline is joined to the
        # timing stream once, after the last step (the next batch's uploads overlap the previous batch's drain)
        e2e_steps = max(3, min(args.steps, 10))
        count = {"k": 0}

        def step_e2e():
            count["k"] += 1
            pipe.run(hR, heps, hlogp, hscore, mode="series", L=L, wait=False, join=False)
            if count["k"] in (2, 2 + e2e_steps):      # last warm-up step / last timed step: join before the event is recorded
                pipe.finish()

        ms_e2e = time_loop(step_e2e, e2e_steps, 2, dist_on)
        e2e_launches = pipe.launches
        # the results really are on the host: spot-check them against the device-resident run
        torch.cuda.synchronize()
        assert torch.equal(hlogp[:4096], logp[:4096].cpu()) and torch.equal(hscore[-4096:], score[-4096:].cpu())
        # what bounds this leg: the pinned H2D copy of the step's inputs alone (and with the D2H of the results running
        # the other way at the same time), measured live on this box with the same buffers
        dR = torch.empty_like(R[:n_e2e])
        s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()

        def copy_up():
            dR.copy_(hR, non_blocking=True)
            eps.copy_(heps, non_blocking=True)

        def copy_duplex():
            with torch.cuda.stream(s_up):
                copy_up()
            with torch.cuda.stream(s_down):
                hlogp.copy_(logp, non_blocking=True)
                hscore.copy_(score, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s_up)
            torch.cuda.current_stream().wait_stream(s_down)

        ms_up = time_loop(copy_up, 3, 1, dist_on) / 3
        ms_duplex = time_loop(copy_duplex, 3, 1, dist_on) / 3
        del dR
        pcie = {"h2d_alone_gbs": n_e2e * 40 / (ms_up * 1e-3) / 1e9, "h2d_alone_ms": ms_up, "duplex_ms": ms_duplex,
                "frac_of_duplex_copy_floor": ms_duplex / (ms_e2e / e2e_steps),
                "note": "floor of a step = max(kernel, duplex copy); the copies are PCIe-bound, the kernel is not"}
        e2e = {"value": world * n_e2e * e2e_steps / (ms_e2e * 1e-3), "unit": UNIT, "pcie_floor": pcie, "h2d_bytes_per_step": n_e2e * 40, "d2h_bytes_per_step": n_e2e * 16,
               "ms_per_step": ms_e2e / e2e_steps, "api": "ops.HostScorePipeline.run(join=False) x steps + finish() (3-stream chunked overlap, batches streamed back to back)", "chunk_rows": args.e2e_chunk, "numa_bind_rank0": numa,
               "pcie_gbs": {"h2d": n_e2e * 40 / (ms_e2e / e2e_steps * 1e-3) / 1e9, "d2h": n_e2e * 16 / (ms_e2e / e2e_steps * 1e-3) / 1e9}}

    # ---- secondary kernels ----------------------------------------------------------------------
    extra = {}
    if not args.no_extra:
        def rate(fn, rows, reps=5):
            m = time_loop(fn, reps, 3, dist_on)
            return world * rows * reps / (m * 1e-3), m / reps

        for mode_name, mode in (("auto", dx._lib.MODE_AUTO), ("series_adaptive", dx._lib.MODE_SERIES_ADAPTIVE)):
            v, m = rate(lambda: lib_call("so3d_igso3_logp_score_f32", ptr(R), ptr(eps), 1, ptr(logp), ptr(score), None, n, mode, L, device=device), n)
            extra[f"score_evals_per_sec_{mode_name}"] = {"value": v, "ms_per_step": m}
            if mode_name == "auto":
                gbs = v / world * BYTES_PER_EVAL / 1e9
                extra[f"score_evals_per_sec_{mode_name}"]["roofline"] = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"]}
        proc = dx.SO3Diffusion(None).to(device)
        proc.row_offset = rank * n
        fwd, post, t_range = proc.tables()
        fwd_guide, post_guide = proc.guides()
        pred = torch.zeros(n, 3, device=device)
        xa, xb = R.clone(), torch.empty_like(R)
        t_step = t_range[500:501]
        state = {"off": 0}

        def step_rev():
            state["off"] += 1
            lib_call("so3d_p_sample_f32", ptr(xa), ptr(pred), ptr(t_step), 0, ptr(proc.sqrt_recip_alphas_cumprod), ptr(proc.sqrt_recipm1_alphas_cumprod),
                     ptr(proc.posterior_mean_coef1), ptr(proc.posterior_mean_coef2), 1000, ptr(post), None, ptr(dx.ops.cdf_grid(device)[2]), SEED, state["off"],
                     rank * n, ptr(xb), None, n, device=device)

        v, m = rate(step_rev, n, 10)
        gbs = v / world * BYTES_PER_PARTICLE_STEP / 1e9
        extra["reverse_particle_steps_per_sec"] = {"value": v, "ms_per_step": m, "note": "fused p_sample kernel, pred = 0 (denoiser excluded), shared t = 500",
                                                   "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"]}}
        tt = torch.randint(0, 1000, (n,), device=device)
        tgt = torch.empty(n, 3, device=device)

        def step_fwd():
            state["off"] += 1
            lib_call("so3d_q_sample_f32", ptr(xa), ptr(tt), ptr(proc.sqrt_alphas_cumprod), ptr(proc.sqrt_one_minus_alphas_cumprod), 1000, ptr(fwd),
                     ptr(fwd_guide), ptr(dx.ops.cdf_grid(device)[2]), SEED, state["off"], rank * n, ptr(xb), ptr(tgt), None, None, n, device=device)

        v, m = rate(step_fwd, n, 10)
        gbs = v / world * BYTES_PER_QSAMPLE / 1e9
        extra["noised_rotations_per_sec"] = {"value": v, "ms_per_step": m, "note": "fused q_sample + skewvec target, per-row t",
                                             "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"]}}

        # north_star's "fused IGSO(3) score + noising": the same launch also evaluates the score of the drawn noise
        # (d log f / d omega along the axis, closed form / series by eps): 36 + 8 B in, 36 + 12 + 12 B out per rotation
        scr = torch.empty(n, 3, device=device)

        def step_fwd_score():
            state["off"] += 1
            lib_call("so3d_q_sample_f32", ptr(xa), ptr(tt), ptr(proc.sqrt_alphas_cumprod), ptr(proc.sqrt_one_minus_alphas_cumprod), 1000, ptr(fwd),
                     ptr(fwd_guide), ptr(dx.ops.cdf_grid(device)[2]), SEED, state["off"], rank * n, ptr(xb), ptr(tgt), None, ptr(scr), n, device=device)

        v, m = rate(step_fwd_score, n, 10)
        gbs = v / world * (BYTES_PER_QSAMPLE + 12) / 1e9
        extra["noised_rotations_with_score_per_sec"] = {"value": v, "ms_per_step": m, "note": "fused q_sample + skewvec target + IGSO(3) score of the noise, per-row t",
                                                        "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"]}}
        del scr

        # BASELINE configs[2]: the whole reverse process, 1000 fused steps over this GPU's share of 2^24 particles
        # (pred = 0: the denoiser is excluded); x ping-pongs between two buffers, t and the Philox offset
        # change every step, nothing synchronises with the host inside the loop.
        n_loop = max(1, (1 << 24) // world)
        la, lb = R[:n_loop].clone(), torch.empty_like(R[:n_loop])
        pz = pred[:n_loop]

        def reverse_loop():
            a, b = la, lb
            for i in reversed(range(1000)):
                lib_call("so3d_p_sample_f32", ptr(a), ptr(pz), ptr(t_range[i:i + 1]), 0, ptr(proc.sqrt_recip_alphas_cumprod),
                         ptr(proc.sqrt_recipm1_alphas_cumprod), ptr(proc.posterior_mean_coef1), ptr(proc.posterior_mean_coef2), 1000, ptr(post),
                         None, ptr(dx.ops.cdf_grid(device)[2]), SEED, 1000 + i, rank * n_loop, ptr(b), None, n_loop, device=device)
                a, b = b, a

        m = time_loop(reverse_loop, 1, 1, dist_on)
        v = world * n_loop * 1000 / (m * 1e-3)
        gbs = v / world * BYTES_PER_PARTICLE_STEP / 1e9
        extra["reverse_loop_1000_steps"] = {"value": v, "unit": "particle-steps/s", "seconds": m * 1e-3, "particles": world * n_loop, "steps": 1000,
                                            "note": "BASELINE configs[2]: 1000 fused reverse steps x 2^24 particles (split over the GPUs), pred = 0",
                                            "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"]}}
        del la, lb

        # one reverse step WITH the reference's RotPredict denoiser (so3_train.py:11-49) inside the kernel
        # (tcgen05 tf32 3-term split), against the two-kernel route (stock PyTorch MLP + fused step).
        torch.manual_seed(SEED)
        net = dx.RotPredict(out_type="skewvec").to(device)
        procn = dx.SO3Diffusion(net).to(device)
        procn.row_offset = rank * n
        procn.tables()
        with torch.no_grad():
            vd, md = rate(lambda: procn.p_sample(xa, t_step), n, 5)
            procn.fuse_denoiser = False
            nu = min(n, 1 << 22)  # the stock route materialises (n, 65) activations: bounded
            vu, mu = rate(lambda: procn.p_sample(xa[:nu], t_step), nu, 3)
        flop_ps = 2 * 65 * (65 * 4 + 3)                       # algorithmic MLP flops per particle-step (fp32 semantics)
        extra["reverse_with_denoiser_particle_steps_per_sec"] = {
            "value": vd, "ms_per_step": md, "note": "RotPredict (65-wide, 5 layers) + reverse step in ONE tcgen05 kernel, shared t = 500",
            "stock_mlp_plus_fused_step": {"value": vu, "ms_per_step": mu, "rows": nu},
            "mlp_tflops_algorithmic": vd / world * flop_ps * 1e-12}
        del net, procn

        # the launch-bound ends of the same paths, as ONE CUDA graph each (device-resident Philox seed):
        # the reference's own sampling size (bingham_test.py:25: 20 000 particles x 1000 steps) and its toy training step
        # (so3_train.py:65-76: batch 256, RotPredict + Adam).  Wall clock, per call, on this rank.
        if rank == 0:
            try:  # secondary, wall-clock legs built on graph capture: a failure here must not cost the headline line
                torch.manual_seed(SEED)
                gnet = dx.RotPredict(out_type="skewvec").to(device)
                gproc = dx.SO3Diffusion(gnet).to(device)
                loops = {}
                for mode, use_graph, one_launch in (("eager", False, False), ("cuda_graph", True, False), ("one_launch", True, True)):
                    gproc.fused_loop = one_launch
                    gproc.p_sample_loop((20000,), cuda_graph=use_graph)          # warm-up / capture
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for _ in range(3):
                        gproc.p_sample_loop((20000,), cuda_graph=use_graph)
                    torch.cuda.synchronize()
                    loops[mode] = (time.perf_counter() - t0) / 3 * 1e3
                extra["reverse_loop_20000_particles_ms"] = {**loops, "unit": "ms per 1000-step loop (wall clock)",
                                                            "note": "RotPredict + reverse step fused: one launch per step from the host (eager), the same launches as one "
                                                                    "captured CUDA graph, and all 1000 steps inside ONE launch (so3d_rotpredict_p_sample_loop_f32)"}
                gopt = torch.optim.Adam(gnet.parameters(), lr=1e-3, capturable=True)
                xb256 = R[:256].contiguous()
                steps_ms = {}

                def eager_step():
                    gopt.zero_grad(set_to_none=True)
                    gproc(xb256).backward()
                    gopt.step()

                graphed = gproc.make_graphed_train_step(gopt, xb256)
                for name, fn in (("eager", eager_step), ("cuda_graph", lambda: graphed(xb256))):
                    for _ in range(10):
                        fn()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for _ in range(200):
                        fn()
                    torch.cuda.synchronize()
                    steps_ms[name] = (time.perf_counter() - t0) / 200 * 1e3
                extra["train_step_batch256_ms"] = {**steps_ms, "unit": "ms per step (wall clock)",
                                                   "note": "SO3Diffusion('skewvec') + RotPredict + Adam; eager vs SO3Diffusion.make_graphed_train_step"}
                del gnet, gproc, gopt, graphed
            except Exception as e:  # recorded, not hidden
                extra["graph_legs_error"] = repr(e)[:300]
        if dist_on:
            torch.distributed.barrier()

        # MMD two-sample statistic at bingham_test.py:29's size (20 000 vs 20 000 rotations), one fused
        # all-pairs launch; multi-GPU: tile pairs dealt round-robin, three doubles all-reduced.
        nm = 20000
        mx, my = R[:nm].contiguous(), xb[:nm].contiguous()
        sums = torch.zeros(3, dtype=torch.float64, device=device)

        def step_mmd():
            sums.copy_(dx.ops.pair_kernel_sums(mx, my, "gaussian", shard=rank, nshards=world))
            if dist_on:
                torch.distributed.all_reduce(sums)

        mm = time_loop(step_mmd, 5, 3, dist_on) / 5
        pairs_logical = 3 * nm * nm                      # k(X,X), k(Y,Y), k(X,Y) as the reference evaluates them
        tiles = (nm + 255) // 256
        pairs_computed = (tiles * (tiles + 1) + tiles * tiles) * 65536   # lower-triangular tile pairs for the two self sums
        sm_ = torch.cuda.get_device_properties(device).multi_processor_count
        lane_peak = world * sm_ * 128 * pk["sm_max_mhz"] * 1e6
        extra["mmd_pairs_per_sec"] = {"value": pairs_logical / (mm * 1e-3), "unit": "kernel evaluations/s (as the reference counts them)", "ms_per_mmd": mm,
                                      "n": [nm, nm], "note": "util.MMD with rmat_gaussian_kernel, fused all-pairs kernel (bingham_test.py:29 size)",
                                      "roofline": {"bound": "fp32", "achieved": pairs_computed * MMD_LANE_INSTR_PER_PAIR / (mm * 1e-3) * 1e-12,
                                                   "peak": lane_peak * 1e-12, "unit": "T lane-instr/s",
                                                   "frac": pairs_computed * MMD_LANE_INSTR_PER_PAIR / (mm * 1e-3) / lane_peak}}

    if rank == 0:
        sm = torch.cuda.get_device_properties(device).multi_processor_count
        ghz = pk["sm_max_mhz"] * 1e-3
        issue_peak = sm * 128 * ghz * 1e9                      # FP32 lane-instructions/s (= 4 warp-instr/clk/SM)
        mufu_peak = sm * 16 * ghz * 1e9
        fp32_peak_tflops = sm * 128 * 2 * ghz * 1e-3          # FFMA = 2 flop/lane/clk
        ach = per_gpu * L * ALGO_LANE_INSTR_PER_TERM
        # time per term implied by each resource for the executed mix, in SMSP clocks per warp-term:
        #   issue 9 slots; XU pipe 8 (MUFU at 4 lanes/clk/SMSP); FP32 operand reads 8.8 (measured per-form costs:
        #   1-register 1.0, 2-register 1.09, 3-register 1.5 clk, profiles/microbench/pipes.cu)
        clk_per_term = sm * 4 * ghz * 1e9 * 32 / (per_gpu * L)
        roofline = {
            "bound": "fp32", "achieved": ach * 1e-12, "peak": issue_peak * 1e-12, "unit": "T lane-instr/s", "frac": ach / issue_peak,
            "traffic": NCU_DRAM_BYTES_PER_EVAL * n, "traffic_source": "ncu --set full, profiles/r02p_series_full.md (bytes/eval x rows per launch)",
            "definition": "SURVEY 8(d): FP32-issue bound with 10 lane-instr/term (1.86e9 evals/s/GPU); executed mix below",
            "executed_per_term": {"fp32_instr": FP32_INSTR_PER_TERM, "mufu": MUFU_PER_TERM, "uniform_ldc": UNIFORM_PER_TERM, "flop": FLOP_PER_TERM},
            "clk_per_warp_term": clk_per_term, "clk_floor_issue": 9.0, "clk_floor_xu": 8.0, "clk_floor_fp32_operands": 8.8,
            "frac_of_executed_mix_floor": 9.0 / clk_per_term,
            "executed_tflops": per_gpu * L * FLOP_PER_TERM * 1e-12, "fp32_peak_tflops": fp32_peak_tflops,
            "mufu_frac": per_gpu * L * MUFU_PER_TERM / mufu_peak,
            "naive_3mufu_roofline_evals_per_s": mufu_peak / (3 * L), "frac_of_naive_3mufu_roofline": per_gpu / (mufu_peak / (3 * L)),
            "hbm_gbs": per_gpu * BYTES_PER_EVAL / 1e9,
            "peak_source": f"derived: {sm} SMs x 128 FP32 lanes x {pk['sm_max_mhz']:.0f} MHz ({pk['source']}); HBM {pk['hbm_gbs']} GB/s",
        }
        line = {
            "metric": METRIC, "value": evals_per_s, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "igso3_logp_score series L=2000, E-set (BASELINE configs[1])", "rows_per_gpu_per_step": n, "series_terms": L,
                       "eps": "per-row, log-uniform [6.4e-3,1]", "l2": "inputs (640 MiB/GPU) larger than L2", "parallelism": f"batch-sharded x{world}, no collective in the data path"},
            "roofline": roofline, "e2e": e2e, "gpu_launches": args.steps, "clocks": clocks, "extra": extra,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline()
        emit(line)
    if dist_on:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
