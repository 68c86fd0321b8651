#!/bin/bash
# r05k: two-row auto score: closed form for both rows + one warp vote for the series override (lib) vs per-row branches (novote)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "score or logp or bench_size" 2>&1 | tail -2
for v in novote "" novote ""; do
  if [ -z "$v" ]; then lib=""; tag=vote; else lib=build/variants/libso3d_$v.so; tag=$v; fi
  SO3D_LIB_PATH=$lib timeout 300 python tests/tools/probe_engine.py 24 $tag 2>&1 | grep -E "\"score" >> gpurun_out/r05k_probe.txt
done
cut -c1-170 gpurun_out/r05k_probe.txt
python - <<'PY'
import os, subprocess, sys
code = r"""
import torch, sys
sys.path.insert(0, '.')
import diffusion_extensions_b200 as dx
from diffusion_extensions_b200 import ops
torch.manual_seed(0)
n = (1 << 20) + 77
R = ops.quat_to_rmat(torch.randn(n, 4, device='cuda'))
eps = torch.exp(torch.empty(n, device='cuda').uniform_(-3.0, 1.5))
l, s, _ = ops.igso3_logp_score(R, eps, mode='auto')
print(float(l.double().sum()), float(s.double().abs().sum()), int(torch.isfinite(l).all()))
"""
outs = []
for lib in ("build/variants/libso3d_novote.so", ""):
    env = dict(os.environ); env["SO3D_LIB_PATH"] = lib
    outs.append(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True).stdout.strip())
print("auto checksums (eps up to e^1.5)", outs, "IDENTICAL" if outs[0] == outs[1] and outs[0] else "DIFFERENT")
PY
