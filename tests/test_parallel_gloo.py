"""World-size-2 tests of the multi-GPU host logic (diffusion_extensions_b200/parallel.py) on the gloo
backend, CPU only: partitioning, the loss / statistics reductions, DDP of the denoiser, and the
shard-invariance of the counter-based draws (through the host build of the kernels' Philox code)."""
import ctypes
import os
import socket
import subprocess

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffusion_extensions_b200 import parallel as P

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _host_math():
    src = os.path.join(HERE, "host_math", "host_math.cpp")
    lib = os.path.join(HERE, "host_math", "_host_math.so")
    hdr = os.path.join(HERE, "..", "diffusion_extensions_b200", "csrc", "so3d_math.cuh")
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-march=native", "-shared", "-fPIC", "-o", lib, src])
    return ctypes.CDLL(lib)


def _draw(hm, seed, row0, offset, n):
    axis = np.empty((n, 3), np.float32)
    u = np.empty(n, np.float32)
    F = ctypes.POINTER(ctypes.c_float)
    hm.hm_draw(ctypes.c_ulonglong(seed), ctypes.c_ulonglong(row0), ctypes.c_ulonglong(offset), axis.ctypes.data_as(F), u.ctypes.data_as(F), ctypes.c_long(n))
    return axis, u


def _worker(rank, world, port, n, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    r, w = P.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and P.world_info() == (rank, world)
    torch.manual_seed(0)  # same global data on every rank
    x = torch.randn(n, 3)
    y = torch.randn(n, 3)
    # ---- partitioning: contiguous, disjoint, covers [0, n)
    xs, lo = P.shard_rows(x)
    lo2, hi2 = P.shard_bounds(n, rank, world)
    assert lo == lo2 and xs.shape[0] == hi2 - lo2
    sizes = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([xs.shape[0]]))
    assert sum(int(s) for s in sizes) == n

    # ---- global loss == single-process mse_loss even with ragged shards (diffusion.py:362)
    ys, _ = P.shard_rows(y)
    loss = P.global_loss((xs - ys) ** 2)
    want = torch.nn.functional.mse_loss(x, y)
    assert abs(float(loss) - float(want)) < 1e-6

    # ---- rotation statistics over the global batch
    q = torch.randn(n, 4)
    q = q / q.norm(dim=-1, keepdim=True)
    a, b, c, d = q.unbind(-1)
    R = torch.stack([1 - 2 * (c * c + d * d), 2 * (b * c - d * a), 2 * (b * d + c * a),
                     2 * (b * c + d * a), 1 - 2 * (b * b + d * d), 2 * (c * d - b * a),
                     2 * (b * d - c * a), 2 * (c * d + b * a), 1 - 2 * (b * b + c * c)], -1).reshape(n, 3, 3)
    st = P.rotation_statistics(P.shard_rows(R)[0])
    ang = torch.acos(torch.clamp((R.diagonal(dim1=-2, dim2=-1).sum(-1) - 1) / 2, -1, 1))
    assert int(st["count"]) == n and abs(float(st["mean_angle"]) - float(ang.mean())) < 1e-5
    assert torch.allclose(st["mean_matrix"], R.mean(0), atol=1e-5)

    # ---- DDP on the denoiser: averaged shard gradients == full-batch gradient for equal shards
    m = n - n % world
    net = torch.nn.Sequential(torch.nn.Linear(3, 16), torch.nn.SiLU(), torch.nn.Linear(16, 3))
    ref = torch.nn.Sequential(torch.nn.Linear(3, 16), torch.nn.SiLU(), torch.nn.Linear(16, 3))
    ref.load_state_dict(net.state_dict())
    ddp = P.wrap_denoiser(net)
    lo3, hi3 = P.shard_bounds(m, rank, world)
    ((ddp(x[lo3:hi3]) - y[lo3:hi3]) ** 2).mean().backward()
    ((ref(x[:m]) - y[:m]) ** 2).mean().backward()
    for pa, pb in zip(net.parameters(), ref.parameters()):
        assert torch.allclose(pa.grad, pb.grad, atol=1e-6)

    # ---- counter-based draws do not depend on the sharding: shard draws == rows [lo, hi) of the global draw
    hm = _host_math()
    full_axis, full_u = _draw(hm, 1234, 0, 7, n)
    axis, u = _draw(hm, 1234, lo, 7, hi2 - lo2)
    assert np.array_equal(axis, full_axis[lo2:hi2]) and np.array_equal(u, full_u[lo2:hi2])

    # ---- MMD over the ranks: round-robin tile pairs + one all-reduce of three doubles == single-process MMD.
    # (The per-shard sums come from the fp64 oracle here; on the GPU box the fused kernel takes its place.)
    from oracle import so3_oracle as O

    rng = np.random.default_rng(5)
    A = O.random_rotations(300 + n % 7, rng)[0]
    B = O.random_rotations(515, rng, 1.0)[0]

    def oracle_pair_sums(Xs, Ys, kernel, shard, nshards):
        out = np.zeros(3)
        T = P.PAIR_TILE
        for sec, bi, bj, w in P.pair_tiles_of_shard(len(Xs), len(Ys), shard, nshards):
            a = (Ys if sec == 1 else Xs)[bi * T:(bi + 1) * T]
            b = (Xs if sec == 0 else Ys)[bj * T:(bj + 1) * T]
            out[sec] += w * O.rmat_gaussian_kernel(a[:, None], b[None]).sum()
        return torch.from_numpy(out)

    got = float(P.mmd_sharded(A, B, "gaussian", pair_sums=oracle_pair_sums))
    assert abs(got - O.mmd(A, B)) < 1e-12
    cover = sorted(tp for sh in range(world) for tp in P.pair_tiles_of_shard(len(A), len(B), sh, world))
    assert cover == sorted(P.pair_tiles_of_shard(len(A), len(B), 0, 1))

    # ---- train_step_sharded places the shard itself (row_offset, per-GLOBAL-row t) and, with ragged shards, weights the
    # local loss so that DDP's plain average is the global-mean gradient.  Stand-in process: the fused noising launch is
    # replaced by a CPU function of (global row, t), everything else is the real host logic.
    class FakeProcess:
        num_timesteps = 1000
        row_offset = -1

        def __init__(self, fn):
            self.denoise_fn, self.seen = fn, None

        def noise_and_target(self, x0, t):
            rows = torch.arange(self.row_offset, self.row_offset + x0.shape[0], dtype=torch.float32)
            self.seen = (self.row_offset, t.clone())
            return {"x_t": x0 + 0.001 * t[:, None].float(), "target": torch.sin(rows)[:, None] * torch.ones(1, 3)}

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(3, 3)

        def forward(self, xx, tt):
            return self.lin(xx)

    torch.manual_seed(3)
    netd, netr = Net(), Net()
    netr.load_state_dict(netd.state_dict())
    fp_ = FakeProcess(P.wrap_denoiser(netd))
    opt = torch.optim.SGD(fp_.denoise_fn.parameters(), lr=0.0)
    loss = P.train_step_sharded(fp_, x[lo2:hi2], opt, n_global=n)
    assert fp_.row_offset == lo2 and fp_.seen[0] == lo2
    single = FakeProcess(netr)
    single.row_offset = 0
    gen = torch.Generator().manual_seed(torch.initial_seed() & 0x7FFFFFFFFFFFFFFF)
    t_all = torch.randint(0, 1000, (n,), generator=gen)
    assert torch.equal(fp_.seen[1], t_all[lo2:hi2])            # t of a row does not depend on the partition
    out = single.noise_and_target(x, t_all)
    sq = (netr(out["x_t"], t_all) - out["target"]) ** 2
    sq.mean().backward()
    assert abs(float(loss) - float(sq.mean())) < 1e-6
    for pa, pb in zip(netd.parameters(), netr.parameters()):    # ragged shards (n = 1001): still the global-mean gradient
        assert torch.allclose(pa.grad, pb.grad, atol=1e-6)
    with pytest.raises(ValueError):
        P.train_step_sharded(fp_, x[: (hi2 - lo2) + 1], opt, n_global=n)

    # ---- attach(): row_offset bookkeeping on a stand-in process object
    class Proc:
        row_offset = 0

    p = Proc()
    assert P.attach(p, n) == (lo2, hi2) and p.row_offset == lo2
    dist.barrier()
    with open(os.path.join(tmp, f"ok{rank}"), "w") as f:
        f.write("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1001, 64])
def test_world2_gloo(tmp_path, n):
    _host_math()  # build once, not concurrently in the workers
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_shard_bounds_properties():
    for n in (0, 1, 7, 8, 1000, 2 ** 24):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = P.shard_bounds(n, r, world)
                assert lo == prev and hi >= lo and hi - lo in (n // world, n // world + 1)
                prev = hi
            assert prev == n
    with pytest.raises(ValueError):
        P.shard_bounds(10, 2, 2)


def test_bind_to_gpu_numa_node_with_fake_sysfs(tmp_path):
    """Host logic of the NUMA binding on a fabricated sysfs tree: parses cpulist, intersects with the allowed CPUs,
    sets the affinity, and degrades to {"bound": False} (never raises) when the tree is missing or the node is -1."""
    import os

    from diffusion_extensions_b200 import parallel as P

    assert P._parse_cpulist("0-2,5,7-8\n") == {0, 1, 2, 5, 7, 8}
    before = os.sched_getaffinity(0)
    try:
        one = min(before)
        bdf = "0000:1b:00.0"
        (tmp_path / "bus" / "pci" / "devices" / bdf).mkdir(parents=True)
        (tmp_path / "bus" / "pci" / "devices" / bdf / "numa_node").write_text("1\n")
        (tmp_path / "devices" / "system" / "node" / "node1").mkdir(parents=True)
        (tmp_path / "devices" / "system" / "node" / "node1" / "cpulist").write_text(f"{one}\n")
        info = P.bind_to_gpu_numa_node(0, sysfs=str(tmp_path), bdf=bdf)
        assert info == {"bound": True, "node": 1, "cpus": 1, "pci": bdf}
        assert os.sched_getaffinity(0) == {one}
        os.sched_setaffinity(0, before)
        (tmp_path / "bus" / "pci" / "devices" / bdf / "numa_node").write_text("-1\n")
        assert P.bind_to_gpu_numa_node(0, sysfs=str(tmp_path), bdf=bdf)["bound"] is False
        assert P.bind_to_gpu_numa_node(0, sysfs=str(tmp_path / "nowhere"), bdf=bdf)["bound"] is False
        assert os.sched_getaffinity(0) == before
    finally:
        os.sched_setaffinity(0, before)
