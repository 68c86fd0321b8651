"""GPU parity tests for SURVEY 8(f-4): the RotPredict denoiser (so3_train.py:11-49) fused with the reverse step
(diffusion.py:308-326) in one tensor-core kernel (so3d_rotpredict_p_sample_f32), called through the Python drop-in
(-> C ABI of libso3d.so).

Checked against (a) golden vectors produced by the unmodified reference (tests/golden/rotpredict.npz), (b) the
fp64 oracle restatement of the network, (c) the two-kernel route (stock PyTorch MLP + so3d_p_sample_f32) with the
same Philox draws.  Tolerances: network outputs <= 1e-5 absolute vs fp64 when both sides use the same fp32 time
features (the 3-term tf32 split measured 6e-8); <= 1e-4 absolute vs the reference's own fp32 forward over the whole
step range (the sinusoidal features' conditioning: see tests/test_oracle_golden.py::test_rotpredict_forward);
reverse-step rotations <= 1e-5 rad geodesic vs the oracle step algebra applied to the kernel's own prediction.
"""
import numpy as np
import pytest
import torch

from oracle import so3_oracle as O

pytestmark = pytest.mark.gpu

LAYERS = (0, 2, 4, 6, 8)


@pytest.fixture(scope="module")
def dx(cuda_device):
    import diffusion_extensions_b200 as pkg

    pkg._lib.load()
    return pkg


def golden_net(dx, g, device):
    net = dx.RotPredict(out_type="skewvec")
    net.load_state_dict({f"net.{i}.{k}": torch.as_tensor(g[f"net_{i}_{k}"]) for i in LAYERS for k in ("weight", "bias")})
    return net.to(device)


def params64(net):
    return ([net.net[i].weight.detach().cpu().double().numpy() for i in LAYERS],
            [net.net[i].bias.detach().cpu().double().numpy() for i in LAYERS])


def fused_pred(dx, proc, net, x, tval, want_out=False, post=None, **kw):
    blob, c1 = net.packed(proc.num_timesteps)
    t = torch.tensor([tval], device=x.device)
    return dx.ops.rotpredict_p_sample_fused(x, blob, c1, t, proc.sqrt_recip_alphas_cumprod, proc.sqrt_recipm1_alphas_cumprod,
                                            proc.posterior_mean_coef1, proc.posterior_mean_coef2, post_cdf=post,
                                            want_out=want_out, want_pred=True, **kw)


def forward64_same_features(net, x, tval):
    """fp64 forward that takes the time features the package itself feeds the kernel (fp32, from the device)."""
    w, b = params64(net)
    emb = net.time_embedding(torch.tensor([tval], device=x.device)).double().cpu().numpy()
    h = np.concatenate([x.double().cpu().numpy().reshape(-1, 9), np.broadcast_to(emb, (x.shape[0], emb.shape[1]))], axis=-1)
    for i in range(5):
        h = h @ w[i].T + b[i]
        if i < 4:
            h = h / (1.0 + np.exp(-h))
    return h


def test_state_dict_names_match_reference(dx, golden):
    g = golden("rotpredict")
    want = {f"net.{i}.{k}" for i in LAYERS for k in ("weight", "bias")}
    assert set(dx.RotPredict(out_type="skewvec").state_dict().keys()) == want
    assert {k.replace("_", ".", 2) for k in g if k.startswith("net_")} == want


def test_fused_prediction_against_reference_golden(dx, cuda_device, golden):
    g = golden("rotpredict")
    net = golden_net(dx, g, cuda_device)
    proc = dx.SO3Diffusion(net).to(cuda_device)
    x = torch.as_tensor(g["x"]).to(cuda_device)
    for k, tv in enumerate(g["t_shared"]):
        pred = fused_pred(dx, proc, net, x, int(tv)).cpu().numpy()
        assert np.max(np.abs(pred - g["pred_shared"][k])) < (5e-6 if tv <= 1 else 1e-4), tv
        # the stock forward of the drop-in module (training path) agrees with the reference too
        stock = net(x, torch.tensor([int(tv)], device=cuda_device)).detach().cpu().numpy()
        assert np.max(np.abs(stock - g["pred_shared"][k])) < (5e-6 if tv <= 1 else 1e-4), tv
    stock = net(x, torch.as_tensor(g["t_row"]).to(cuda_device)).detach().cpu().numpy()
    assert np.max(np.abs(stock - g["pred_row"])) < 1e-4


def test_fused_reverse_mean_against_reference_golden(dx, cuda_device, golden):
    """Posterior mean with the denoiser inside the kernel vs the reference's p_mean_variance (t = 1, 300; the
    reference's own so3_scale error makes t = 999 meaningless, quirk Q5)."""
    g = golden("rotpredict")
    net = golden_net(dx, g, cuda_device)
    proc = dx.SO3Diffusion(net).to(cuda_device)
    x = torch.as_tensor(g["x"]).to(cuda_device)
    for k, tv in enumerate(g["t_mean"]):
        if tv > 300:
            continue
        blob, c1 = net.packed(proc.num_timesteps)
        mean = dx.ops.rotpredict_p_sample_fused(x, blob, c1, torch.tensor([int(tv)], device=cuda_device), proc.sqrt_recip_alphas_cumprod,
                                                proc.sqrt_recipm1_alphas_cumprod, proc.posterior_mean_coef1, proc.posterior_mean_coef2,
                                                post_cdf=None)
        err = O.geodesic_angle(mean.cpu().numpy().astype(np.float64), g["mean_shared"][k].astype(np.float64))
        assert np.max(err) < 1e-4, (tv, np.max(err))


@pytest.mark.parametrize("n", [1, 2, 127, 128, 129, 1000, 4097, 70001])
def test_fused_prediction_against_fp64(dx, cuda_device, n):
    """Ragged and multi-tile sizes, large-ish weights (activations of size ~3): the tf32 hi/lo split must hold
    fp32-level accuracy (<= 1e-5 abs) against an fp64 forward given the same time features."""
    torch.manual_seed(n)
    net = dx.RotPredict(out_type="skewvec").to(cuda_device)
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(3.0)
    proc = dx.SO3Diffusion(net).to(cuda_device)
    x = dx.ops.quat_to_rmat(torch.randn(n, 4, device=cuda_device))
    for tv in (0, 517, 999):
        pred = fused_pred(dx, proc, net, x, tv).cpu().numpy().astype(np.float64)
        want = forward64_same_features(net, x, tv)
        scale = max(1.0, np.max(np.abs(want)))
        assert pred.shape == (n, 3)
        assert np.max(np.abs(pred - want)) < 1e-5 * scale, (n, tv, np.max(np.abs(pred - want)))


def test_fused_step_equals_step_algebra_on_its_own_prediction(dx, cuda_device):
    """out and pred from ONE launch: out must be the oracle's reverse-step mean of (x_t, pred) (t with and without
    noise switched off), <= 1e-5 rad."""
    torch.manual_seed(3)
    n = 3000
    net = dx.RotPredict(out_type="skewvec").to(cuda_device)
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(2.0)
    proc = dx.SO3Diffusion(net).to(cuda_device)
    x = dx.ops.quat_to_rmat(torch.randn(n, 4, device=cuda_device))
    s = O.schedule_buffers(1000)
    for tv in (0, 1, 40, 300):
        out, pred = fused_pred(dx, proc, net, x, tv, want_out=True)
        want = O.p_sample_mean(x.double().cpu().numpy(), pred.double().cpu().numpy(), s["sqrt_recip_alphas_cumprod"][tv],
                               s["sqrt_recipm1_alphas_cumprod"][tv], s["posterior_mean_coef1"][tv], s["posterior_mean_coef2"][tv])
        err = O.geodesic_angle(out.double().cpu().numpy(), want)
        assert np.max(err) < 1e-5, (tv, np.max(err))


def test_fused_p_sample_matches_two_kernel_route(dx, cuda_device):
    """SO3Diffusion.p_sample with the denoiser fused vs stock MLP + so3d_p_sample_f32 at the same seed: identical
    Philox draws, predictions equal to ~1e-6, so samples agree to <= 1e-5 rad geodesic; t = 0 adds no noise."""
    torch.manual_seed(5)
    n = 5000
    net = dx.RotPredict(out_type="skewvec").to(cuda_device)
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(2.0)
    proc = dx.SO3Diffusion(net).to(cuda_device)
    x = dx.ops.quat_to_rmat(torch.randn(n, 4, device=cuda_device))
    for tv in (0, 1, 250, 600):
        t = torch.tensor([tv], device=cuda_device)
        dx.ops.manual_seed(11)
        proc.fuse_denoiser = True
        a = proc.p_sample(x, t)
        dx.ops.manual_seed(11)
        proc.fuse_denoiser = False
        b = proc.p_sample(x, t)
        err = O.geodesic_angle(a.double().cpu().numpy(), b.double().cpu().numpy())
        # the step's gain on the prediction is sqrt_recipm1[t] * coef1-ish <= 1 for these t
        assert np.max(err) < 1e-5, (tv, np.max(err))
        ortho = (a.transpose(-1, -2) @ a - torch.eye(3, device=cuda_device)).abs().max().item()
        assert ortho < 5e-6
    proc.fuse_denoiser = True
    dx.ops.manual_seed(11)
    a0 = proc.p_sample(x, torch.tensor([0], device=cuda_device))
    m0 = proc.p_mean_variance(x, torch.tensor([0], device=cuda_device))[0]
    assert np.max(O.geodesic_angle(a0.double().cpu().numpy(), m0.double().cpu().numpy())) < 1e-5


def test_fused_draws_are_shard_invariant(dx, cuda_device):
    """row_offset makes a shard's draws those of the same global rows."""
    torch.manual_seed(9)
    n = 2048
    net = dx.RotPredict(out_type="skewvec").to(cuda_device)
    proc = dx.SO3Diffusion(net).to(cuda_device)
    _, post, _ = proc.tables()
    x = dx.ops.quat_to_rmat(torch.randn(n, 4, device=cuda_device))
    blob, c1 = net.packed(proc.num_timesteps)
    t = torch.tensor([400], device=cuda_device)
    args = (proc.sqrt_recip_alphas_cumprod, proc.sqrt_recipm1_alphas_cumprod, proc.posterior_mean_coef1, proc.posterior_mean_coef2)
    full = dx.ops.rotpredict_p_sample_fused(x, blob, c1, t, *args, post_cdf=post, seed=5, rng_offset=17)
    h = 1000  # not a tile multiple
    lo = dx.ops.rotpredict_p_sample_fused(x[:h].contiguous(), blob, c1, t, *args, post_cdf=post, seed=5, rng_offset=17)
    hi = dx.ops.rotpredict_p_sample_fused(x[h:].contiguous(), blob, c1, t, *args, post_cdf=post, seed=5, rng_offset=17, row_offset=h)
    assert torch.equal(torch.cat([lo, hi]), full)


def test_packed_weights_follow_parameter_updates(dx, cuda_device):
    torch.manual_seed(1)
    net = dx.RotPredict(out_type="skewvec").to(cuda_device)
    proc = dx.SO3Diffusion(net).to(cuda_device)
    x = dx.ops.quat_to_rmat(torch.randn(256, 4, device=cuda_device))
    p0 = fused_pred(dx, proc, net, x, 100)
    with torch.no_grad():
        net.net[8].bias.add_(1.0)  # in-place update, as an optimizer step does
    p1 = fused_pred(dx, proc, net, x, 100)
    assert torch.allclose(p1, p0 + 1.0, atol=1e-5)


def test_p_sample_loop_with_fused_denoiser(dx, cuda_device):
    """so3_test.py:24-33: the full reverse loop (shortened schedule) stays on SO(3) and is reproducible."""
    torch.manual_seed(2)
    net = dx.RotPredict(out_type="skewvec").to(cuda_device)
    proc = dx.SO3Diffusion(net, timesteps=50).to(cuda_device)
    dx.ops.manual_seed(21)
    a = proc.p_sample_loop((512,))
    dx.ops.manual_seed(21)
    b = proc.p_sample_loop((512,))
    assert torch.equal(a, b)
    assert a.shape == (512, 3, 3) and torch.isfinite(a).all()
    assert (a.transpose(-1, -2) @ a - torch.eye(3, device=cuda_device)).abs().max().item() < 5e-6
    assert (torch.linalg.det(a) - 1).abs().max().item() < 1e-5


@pytest.mark.parametrize("route", ["one_launch", "graph_fused", "graph_stock"])
def test_p_sample_loop_as_cuda_graph(dx, cuda_device, route):
    """p_sample_loop(cuda_graph=True) runs the whole reverse process without per-step launches from the host:
    "one_launch" -- the fused tensor-core kernel iterates all T steps itself (so3d_rotpredict_p_sample_loop_f32);
    "graph_fused" / "graph_stock" -- one captured CUDA graph of T launches whose kernels read the Philox seed from device
    memory, with the fused denoiser or the stock-PyTorch denoiser.  Each must (1) equal, bit for bit, the eager loop run
    with that seed by value, (2) draw fresh noise on the next call, (3) follow a weight update."""
    torch.manual_seed(3)
    T, n = 40, 700
    fuse = route != "graph_stock"
    net = dx.RotPredict(out_type="skewvec").to(cuda_device)
    proc = dx.SO3Diffusion(net, timesteps=T).to(cuda_device)
    proc.fuse_denoiser = fuse
    proc.fused_loop = route == "one_launch"
    sched = (proc.sqrt_recip_alphas_cumprod, proc.sqrt_recipm1_alphas_cumprod, proc.posterior_mean_coef1, proc.posterior_mean_coef2)

    def eager_with_seed(seed_start):
        dx.ops.manual_seed(seed_start)
        x = dx.IsotropicGaussianSO3(torch.ones([], device=cuda_device)).sample((n,), row_offset=0)   # the loop's init draw
        seed, off = dx.ops.rng.next()                                                                # the graph's seed draw
        mixed = (seed + 0x9E3779B97F4A7C15 * (off + 1)) & 0xFFFFFFFFFFFFFFFF
        _, post, t_range = proc.tables()
        with torch.no_grad():
            for i in reversed(range(T)):
                t = t_range[i:i + 1]
                if fuse:
                    blob, c1 = net.packed(T)
                    x = dx.ops.rotpredict_p_sample_fused(x, blob, c1, t, *sched, post_cdf=post, seed=mixed, rng_offset=i)
                else:
                    x = dx.ops.p_sample_fused(x, net(x, t.expand(n)), t, *sched, post_cdf=post, seed=mixed, rng_offset=i)
        return x

    dx.ops.manual_seed(31)
    a = proc.p_sample_loop((n,), cuda_graph=True)
    assert torch.equal(a, eager_with_seed(31))
    dx.ops.manual_seed(31)
    a2 = proc.p_sample_loop((n,), cuda_graph=True)          # replay of the cached graph, same seed
    b = proc.p_sample_loop((n,), cuda_graph=True)           # replay, next seed: fresh noise
    assert torch.equal(a, a2) and not torch.equal(a, b)
    assert len(getattr(proc, "_loop_graphs", {})) == (0 if route == "one_launch" else 1)
    assert (b.transpose(-1, -2) @ b - torch.eye(3, device=cuda_device)).abs().max().item() < 5e-6
    with torch.no_grad():
        net.net[0].weight.mul_(1.05)                         # a weight update: the fused route must re-pack and re-capture
    dx.ops.manual_seed(32)
    c = proc.p_sample_loop((n,), cuda_graph=True)
    assert torch.equal(c, eager_with_seed(32))


def test_errors(dx, cuda_device):
    net = dx.RotPredict(out_type="skewvec").to(cuda_device)
    proc = dx.SO3Diffusion(net).to(cuda_device)
    x = dx.ops.quat_to_rmat(torch.randn(8, 4, device=cuda_device))
    blob, c1 = net.packed(proc.num_timesteps)
    args = (proc.sqrt_recip_alphas_cumprod, proc.sqrt_recipm1_alphas_cumprod, proc.posterior_mean_coef1, proc.posterior_mean_coef2)
    with pytest.raises(ValueError):
        dx.ops.rotpredict_p_sample_fused(x, blob, c1, torch.zeros(8, dtype=torch.int64, device=cuda_device), *args)
    with pytest.raises(ValueError):
        dx.ops.rotpredict_p_sample_fused(x, blob[:-1], c1, torch.zeros(1, dtype=torch.int64, device=cuda_device), *args)
    # the reference's default head (6-D -> six2rmat, so3_train.py:12,46-47) runs as stock PyTorch and is never fused
    rm = dx.RotPredict().to(cuda_device)
    assert rm.out_type == "rotmat" and rm.d_out == 6 and not rm.fusable()
    out = rm(x, torch.zeros(8, dtype=torch.int64, device=cuda_device))
    assert out.shape == (8, 3, 3)
    eye = torch.eye(3, device=cuda_device).expand(8, 3, 3)
    assert torch.allclose(out @ out.transpose(-1, -2), eye, atol=1e-5)
    with pytest.raises(RuntimeError):
        dx.RotPredict(out_type="euler")
