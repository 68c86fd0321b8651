"""Per-shard time of the all-pairs MMD kernel at bingham_test.py:29's size, for 1 and 8 shards (one GPU plays shard 0):
python tests/tools/probe_mmd.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from diffusion_extensions_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
X = ops.quat_to_rmat(torch.randn(20000, 4, device=dev))
Y = ops.quat_to_rmat(torch.randn(20000, 4, device=dev))
for nshards in (1, 2, 4, 8):
    fn = lambda: ops.pair_kernel_sums(X, Y, "gaussian", shard=0, nshards=nshards)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(json.dumps({"nshards": nshards, "ms_per_shard_call": round(ms, 4), "ideal_ms": None if nshards == 1 else "t1/nshards"}))
full = ops.pair_kernel_sums(X, Y)
parts = sum(ops.pair_kernel_sums(X, Y, shard=s, nshards=8) for s in range(8))
print(json.dumps({"sum_of_8_shards_equals_single": bool(torch.equal(parts, full)), "rel_diff": float(((parts - full).abs() / full).max())}))
