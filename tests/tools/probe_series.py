"""A/B of series-kernel builds:  python tests/tools/probe_series.py [log2_rows] -- one JSON line per library under
build/variants/ (plus the shipped one): ms per launch of the L = 2000 series (modes series / series_pure / series_adaptive)
on the bench's E-set, and a checksum of the results (all builds must agree bit for bit)."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CHILD = r'''
import json, os, sys, math, torch
sys.path.insert(0, %r)
import diffusion_extensions_b200 as dx
import bench
n = 1 << int(sys.argv[1])
dev = torch.device("cuda:0")
dx._lib.load()
R, eps = bench.make_eset(n, dev, 1234)
logp = torch.empty(n, device=dev); score = torch.empty(n, 3, device=dev)
call, ptr = dx._lib.call, dx._lib.ptr
out = {"lib": os.path.basename(os.environ.get("SO3D_LIB_PATH", "shipped"))}
for name, mode in (("series", 0), ("series_pure", 4), ("series_adaptive", 3)):
    fn = lambda: call("so3d_igso3_logp_score_f32", ptr(R), ptr(eps), 1, ptr(logp), ptr(score), None, n, mode, 2000, device=dev)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    out[name + "_ms"] = round(e0.elapsed_time(e1) / 10, 4)
    out[name + "_sum"] = [float(logp.double().sum()), float(score.double().abs().sum())]
print(json.dumps(out))
''' % ROOT

lg = sys.argv[1] if len(sys.argv) > 1 else "24"
libs = [None] + sorted(glob.glob(os.path.join(ROOT, "build", "variants", "libso3d_*.so")))
for lib in libs:
    env = dict(os.environ)
    if lib:
        env["SO3D_LIB_PATH"] = lib
    p = subprocess.run([sys.executable, "-c", CHILD, lg], env=env, capture_output=True, text=True, timeout=300)
    print(p.stdout.strip() or json.dumps({"lib": lib, "error": p.stderr[-300:]}), flush=True)
