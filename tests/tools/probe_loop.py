"""Wall time of the whole reverse loop (1000 steps) at the reference's sampling size (bingham_test.py:25: 20 000 particles)
and a few others, eager vs one captured CUDA graph, fused tensor-core denoiser vs stock-PyTorch denoiser.  Prints JSON lines."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import diffusion_extensions_b200 as dx

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = dx.RotPredict(out_type="skewvec").to(dev)
proc = dx.SO3Diffusion(net).to(dev)
for n in (1024, 20000, 1 << 18, 1 << 22):
    for fuse in (True, False):
        proc.fuse_denoiser = fuse
        for graph in ((False, True, "one launch") if fuse else (False, True)):
            proc.fused_loop = graph == "one launch"
            proc.p_sample_loop((n,), cuda_graph=bool(graph))       # warm-up / capture
            torch.cuda.synchronize()
            reps = 3
            t0 = time.perf_counter()
            for _ in range(reps):
                x = proc.p_sample_loop((n,), cuda_graph=bool(graph))
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / reps
            print(json.dumps({"particles": n, "denoiser": "fused tcgen05" if fuse else "stock torch", "mode": {False: "eager", True: "cuda graph"}.get(graph, graph), "loop_ms": round(1e3 * dt, 2),
                              "particle_steps_per_s": n * 1000 / dt}), flush=True)
