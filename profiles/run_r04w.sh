#!/bin/bash
# r04w: auto = closed form first, series override for eps > 1 (new) vs branch first (base = HEAD + rolled loop); head = committed HEAD
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "score or logp or q_sample or noising or smoke or series" 2>&1 | tail -3
for v in base "" base ""; do
  if [ -z "$v" ]; then lib=""; tag=new; else lib=build/variants/libso3d_$v.so; tag=$v; fi
  SO3D_LIB_PATH=$lib timeout 300 python tests/tools/probe_engine.py 24 $tag 2>&1 | grep -E "score" >> gpurun_out/r04w_probe.txt
done
cut -c1-175 gpurun_out/r04w_probe.txt
python - <<'PY'
# bit-identity of auto above eps = 1 between the two builds (the rolled loop runs the same operations per term)
import os, subprocess, sys, json
code = r"""
import torch, sys
sys.path.insert(0, '.')
import diffusion_extensions_b200 as dx
from diffusion_extensions_b200 import ops
torch.manual_seed(0)
n = 1 << 20
R = ops.quat_to_rmat(torch.randn(n, 4, device='cuda'))
eps = torch.exp(torch.empty(n, device='cuda').uniform_(-1.0, 1.5))
l, s, _ = ops.igso3_logp_score(R, eps, mode='auto')
print(float(l.double().sum()), float(s.double().abs().sum()), int(torch.isfinite(l).all()))
"""
outs = []
for lib in ("build/variants/libso3d_base.so", ""):
    env = dict(os.environ); env["SO3D_LIB_PATH"] = lib
    outs.append(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True).stdout.strip())
print("auto eps>1 checksums", outs, "IDENTICAL" if outs[0] == outs[1] and outs[0] else "DIFFERENT")
PY
