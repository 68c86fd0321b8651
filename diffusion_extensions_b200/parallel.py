"""Multi-GPU harness: one process per GPU, batch rows partitioned across ranks.

Every op of the SO(3) hot path is row-independent, so the data path needs NO collective: each
rank owns a contiguous shard of the global batch and runs the fused kernels on it.  What does cross ranks:

  * the loss / statistics reductions (a handful of floats per step) -- `all_reduce_stats`, `global_mean`;
  * the gradient all-reduce of the small denoiser network -- stock `DistributedDataParallel` (`wrap_denoiser`).

Both go through `torch.distributed` (NCCL over NVLink on the GPU box; gloo in the CPU test tier).  The
Philox counter of every random draw is the GLOBAL row index (`row_offset` of the shard + local row), so a
sharded run draws exactly the numbers a single-GPU run would: results are bit-identical for any world size.

The reference has no distributed code at all; this module is the B200-side addition behind the
same `SO3Diffusion` API (`process.row_offset` is the only coupling).
"""
import os

import torch
import torch.distributed as dist


# ---------------------------------------------------------------------------------------------
# process group
# ---------------------------------------------------------------------------------------------
def init_from_env(backend=None, device=None, bind_numa=True):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world).
    With WORLD_SIZE unset or 1 nothing is initialised.  bind_numa: with several ranks and a CUDA `device`, pin the
    process to that GPU's NUMA node first (bind_to_gpu_numa_node)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and bind_numa and device is not None and torch.device(device).type == "cuda" and torch.cuda.is_available():
        idx = torch.device(device).index
        bind_to_gpu_numa_node(torch.cuda.current_device() if idx is None else idx)
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl" and device is not None:
            kwargs["device_id"] = torch.device(device)
        dist.init_process_group(backend, **kwargs)
    return rank, world


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index, sysfs="/sys", bdf=None):
    """Pin this process's CPU affinity to the NUMA node the GPU hangs off (sysfs: the PCI device's `numa_node`, the
    node's `cpulist`), so that pinned host staging buffers allocated afterwards are first-touched on that node and the
    H2D / D2H copies of the host-buffer pipeline do not cross the socket interconnect.  One process per GPU on an
    8-GPU box otherwise lands wherever the scheduler puts it; the end-to-end (host-buffer) rate then scales far worse
    than the device-resident one (profiles/r01r_bench_n8.json: 2.4x at 8 GPUs).  Returns a dict describing what was
    done; never raises (no sysfs / single node / no permission -> {"bound": False, "why": ...})."""
    try:
        if bdf is None:  # (tests pass the PCI address directly)
            props = torch.cuda.get_device_properties(device_index)
            bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(os.path.join(sysfs, "bus", "pci", "devices", bdf, "numa_node")) as f:
            node = int(f.read().strip())
        if node < 0:
            return {"bound": False, "why": "no NUMA affinity reported for " + bdf}
        with open(os.path.join(sysfs, "devices", "system", "node", f"node{node}", "cpulist")) as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if not target:
            return {"bound": False, "why": f"node {node} has no allowed CPUs"}
        os.sched_setaffinity(0, target)
        try:
            torch.set_num_threads(max(1, min(torch.get_num_threads(), len(target))))
        except Exception:
            pass
        return {"bound": True, "node": node, "cpus": len(target), "pci": bdf}
    except Exception as e:  # diagnostic helper: the data path does not depend on it
        return {"bound": False, "why": repr(e)[:120]}


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


# ---------------------------------------------------------------------------------------------
# partitioning
# ---------------------------------------------------------------------------------------------
def shard_bounds(n, rank, world):
    """Contiguous shard [lo, hi) of n rows for `rank`: the first n % world ranks get one extra row."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"invalid rank/world: {rank}/{world}")
    base, extra = divmod(int(n), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rows(x, rank=None, world=None):
    """This rank's contiguous slice of a globally indexed tensor (dim 0) and its global row offset."""
    if rank is None or world is None:
        rank, world = world_info()
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi], lo


def attach(process, n_global, rank=None, world=None):
    """Point a SO3Diffusion at this rank's shard of an n_global-row batch: sets `process.row_offset` (the
    global index of the shard's first row, which keys the Philox draws) and returns (lo, hi)."""
    if rank is None or world is None:
        rank, world = world_info()
    lo, hi = shard_bounds(n_global, rank, world)
    process.row_offset = lo
    return lo, hi


# ---------------------------------------------------------------------------------------------
# reductions (the only collectives of the path)
# ---------------------------------------------------------------------------------------------
def all_reduce_stats(stats):
    """Sum a small 1-D tensor of statistics over all ranks (in place; no-op for a single process)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def global_mean(local_sum, local_count):
    """Mean over the global batch from per-shard sums and counts: one all-reduce of two numbers.
    `local_sum` is a 0-d tensor (kept on its device, no host sync), `local_count` an int."""
    buf = torch.stack([local_sum.detach().to(torch.float32).reshape(()), torch.tensor(float(local_count), device=local_sum.device)])
    all_reduce_stats(buf)
    return buf[0] / buf[1]


def global_loss(per_row_sq_err):
    """The reference's `F.mse_loss(pred, target)` (diffusion.py:362) over the GLOBAL batch: shards may be
    ragged, so the mean is sum / count reduced over ranks, not a mean of per-rank means."""
    return global_mean(per_row_sq_err.sum(), per_row_sq_err.numel())


def wrap_denoiser(module, device=None):
    """DistributedDataParallel around the denoiser (gradient all-reduce, bucketed by DDP); identity for a
    single process.  The SO(3) kernels themselves hold no parameters."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return module
    from torch.nn.parallel import DistributedDataParallel as DDP

    if device is not None and torch.device(device).type == "cuda":
        return DDP(module.to(device), device_ids=[torch.device(device).index])
    return DDP(module)


def rotation_statistics(x):
    """Per-shard sufficient statistics of a batch of rotations for global reporting:
    [count, sum of rotation angle, sum of angle^2, sum of the 9 matrix entries] -> all-reduced."""
    tr = x[..., 0, 0] + x[..., 1, 1] + x[..., 2, 2]
    ang = torch.acos(torch.clamp((tr - 1) * 0.5, -1.0, 1.0))
    stats = torch.cat([torch.tensor([float(ang.numel())], device=x.device), ang.sum().reshape(1), (ang * ang).sum().reshape(1),
                       x.reshape(-1, 9).sum(0)]).to(torch.float32)
    all_reduce_stats(stats)
    n = stats[0]
    return {"count": n, "mean_angle": stats[1] / n, "std_angle": torch.sqrt(torch.clamp(stats[2] / n - (stats[1] / n) ** 2, min=0)),
            "mean_matrix": (stats[3:] / n).reshape(3, 3)}


# ---------------------------------------------------------------------------------------------
# MMD two-sample statistic over the ranks (util.py:254-285)
# ---------------------------------------------------------------------------------------------
PAIR_TILE = 256  # rows per tile of the all-pairs kernel (csrc/so3d_pairwise.cu)


def pair_tiles_of_shard(nx, ny, shard, nshards):
    """The 256 x 256 tile pairs shard `shard` of `nshards` sums, in the kernel's order: the flat list
    [X-X lower triangle | Y-Y lower triangle | X-Y all] dealt round-robin.  -> list of (section, bi, bj, weight),
    section 0/1/2 = XX/YY/XY, weight 2 for the off-diagonal tiles of the symmetric sections."""
    tx, ty = -(-nx // PAIR_TILE), -(-ny // PAIR_TILE)
    flat = [(0, i, j, 1 if i == j else 2) for i in range(tx) for j in range(i + 1)]
    flat += [(1, i, j, 1 if i == j else 2) for i in range(ty) for j in range(i + 1)]
    flat += [(2, i, j, 1) for i in range(tx) for j in range(ty)]
    return flat[shard::nshards]


def mmd_from_sums(sums, nx, ny):
    """util.py:281-285: mean k(X,X) + mean k(Y,Y) - 2 mean k(X,Y) from the three all-pairs sums."""
    return sums[0] / (nx ** 2) + sums[1] / (ny ** 2) - sums[2] * (2 / (nx * ny))


def mmd_sharded(X, Y, kernel="gaussian", pair_sums=None):
    """MMD(X, Y) with every rank holding the full sample sets and summing its round-robin share of the tile pairs;
    one all-reduce of three doubles.  `pair_sums(X, Y, kernel, shard=, nshards=)` defaults to the fused kernel."""
    rank, world = world_info()
    if pair_sums is None:
        from . import ops

        pair_sums = ops.pair_kernel_sums
    sums = pair_sums(X, Y, kernel, shard=rank, nshards=world)
    all_reduce_stats(sums)
    return mmd_from_sums(sums, len(X), len(Y))


# ---------------------------------------------------------------------------------------------
# sharded drivers
# ---------------------------------------------------------------------------------------------
@torch.no_grad()
def sample_sharded(process, n_global, init="igso3_1", steps=None):
    """BASELINE config 3: n_global particles partitioned over the ranks, each rank running the fused reverse
    step on its shard for all timesteps.  No communication inside the loop.  Returns the local shard."""
    lo, hi = attach(process, n_global)
    if steps is None:
        return process.p_sample_loop((hi - lo,), init=init)
    from . import ops
    from .distributions import IsotropicGaussianSO3

    device = process.betas.device
    x = IsotropicGaussianSO3(torch.ones([], device=device)).sample((hi - lo,), row_offset=lo)
    _, _, t_range = process.tables()
    for i in reversed(range(process.num_timesteps - steps, process.num_timesteps)):
        x = process.p_sample(x, t_range[i:i + 1])
    return x


def _step_indices(process, n_global, lo, hi, device):
    """The step index of every GLOBAL row, drawn identically on all ranks from a generator private to `process`
    (seeded from torch's seed at first use), then sliced to this rank's shard: t of a row does not depend on how the
    batch is partitioned, and ranks seeded alike stay in lock-step without consuming torch's default stream."""
    gen = process.__dict__.get("_t_generator")
    if gen is None or gen.device != torch.device(device):
        gen = torch.Generator(device=device)
        gen.manual_seed(torch.initial_seed() & 0x7FFFFFFFFFFFFFFF)
        process.__dict__["_t_generator"] = gen
    t = torch.randint(0, process.num_timesteps, (n_global,), device=device, generator=gen)
    return t[lo:hi]


def train_step_sharded(process, x0_local, optimizer, n_global=None):
    """One data-parallel training step on this rank's shard: fused noising + target, denoiser forward/backward
    (DDP all-reduces the gradients), optimizer step; returns the GLOBAL mean loss (one 2-float all-reduce).

    The shard is placed in the global batch here (no separate `attach()` call to forget): `n_global` rows split by
    `shard_bounds`; without it the shards are taken to be equal (n_global = world x local rows), which is checked.
    Noise is keyed by the global row index (`process.row_offset`) and t is drawn per GLOBAL row, so the step's draws are
    the same for any world size.  With ragged shards DDP's plain average of per-rank gradients is not the global-mean
    gradient; the local loss is therefore weighted by b_local x world / n_global, which makes it exact."""
    rank, world = world_info()
    b = x0_local.shape[0]
    if n_global is None:
        n_global = world * b
    lo, hi = shard_bounds(n_global, rank, world)
    if hi - lo != b:
        raise ValueError(f"rank {rank}: shard has {b} rows but rows [{lo}, {hi}) of a {n_global}-row batch are expected "
                         "(pass n_global for ragged shards; equal shards otherwise)")
    process.row_offset = lo
    optimizer.zero_grad(set_to_none=True)
    t = _step_indices(process, n_global, lo, hi, x0_local.device)
    fused = process.noise_and_target(x0_local, t)
    pred = process.denoise_fn(fused["x_t"], t)
    sq = (pred - fused["target"]) ** 2
    loss = sq.mean() * (b * world / n_global)  # == sq.mean() for equal shards
    loss.backward()
    optimizer.step()
    return global_loss(sq.detach())
