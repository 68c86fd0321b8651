"""Training step of BASELINE configs[0]/[3] (so3_train.py / bingham_train.py: SO3Diffusion('skewvec') + RotPredict + Adam):
device-timed steps/s of the drop-in on cuda:0 for several batch sizes, eager and captured in a CUDA graph, and the
reference's CPU path (oracle/ref_port.py, the reference's own op sequence) at batch 256.  Diagnostic, prints JSON lines.

    python tests/tools/probe_train.py [--cpu] [--graph]

(--graph: also time SO3Diffusion.make_graphed_train_step, the step captured as one CUDA graph.)
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F

import diffusion_extensions_b200 as dx
from diffusion_extensions_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)


def gpu_leg(B, graph):
    net = dx.RotPredict(out_type="skewvec").to(dev)
    proc = dx.SO3Diffusion(net).to(dev)
    proc.tables(); proc.guides()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=graph)
    x0 = ops.quat_to_rmat(torch.randn(B, 4, device=dev))

    def step():
        opt.zero_grad(set_to_none=True)
        loss = proc(x0)
        loss.backward()
        opt.step()
        return loss

    run = step
    if graph:  # whole-step capture (SO3Diffusion.make_graphed_train_step: device-resident noise seed)
        stepper = proc.make_graphed_train_step(opt, x0)
        run = lambda: stepper(x0)
        loss = run()
    else:
        for _ in range(5):
            loss = step()
    torch.cuda.synchronize()
    reps = 200 if B <= 65536 else 20
    for _ in range(10):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1) / reps
    return {"leg": "gpu", "batch": B, "cuda_graph": graph, "ms_per_step": round(ms, 4), "steps_per_s": round(1e3 / ms, 1), "samples_per_s": B * 1e3 / ms,
            "wall_ms_per_step": round(1e3 * wall / reps, 4), "loss": float(loss)}


for B in (256, 4096, 65536, 1 << 20):
    for graph in ((False, True) if "--graph" in sys.argv else (False,)):
        try:
            print(json.dumps(gpu_leg(B, graph)), flush=True)
        except Exception as e:
            import traceback
            print(json.dumps({"leg": "gpu", "batch": B, "cuda_graph": graph, "error": repr(e)[:300], "tb": traceback.format_exc()[-1500:]}), flush=True)

if "--cpu" in sys.argv:
    from oracle import ref_port as P

    torch.set_num_threads(os.cpu_count() or 1)
    net = dx.RotPredict(out_type="skewvec")
    port = P.SO3DiffusionPort(net)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    B = 256
    q = torch.randn(B, 4)
    q = q / q.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    x0 = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                      2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1).reshape(B, 3, 3)

    def cpu_step():
        t = torch.randint(0, 1000, (B,))
        opt.zero_grad()
        x_noisy, target = port.p_losses_inputs(x0, t)
        loss = F.mse_loss(net(x_noisy, t), target)
        loss.backward()
        opt.step()

    cpu_step()
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < 10:
        cpu_step()
        n += 1
    dt = time.perf_counter() - t0
    print(json.dumps({"leg": "cpu reference port", "batch": B, "cores": os.cpu_count(), "ms_per_step": round(1e3 * dt / n, 2), "steps_per_s": round(n / dt, 2)}))
