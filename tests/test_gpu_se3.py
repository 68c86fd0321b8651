"""GPU parity tests of the SE(3) arm: SE3Diffusion / ProjectedSE3Diffusion / IGSO3xR3 / se3_scale through
the Python drop-in (-> C ABI), against golden vectors produced by the reference's SE3Diffusion and the fp64 oracle.
Tolerances: rotations <= 1e-5 (entries; geodesic <= 1e-5 rad), translations <= 1e-5 relative to max(1, |x|)."""
import math

import numpy as np
import pytest
import torch

from oracle import so3_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dx(cuda_device):
    import diffusion_extensions_b200 as pkg

    pkg._lib.load()
    return pkg


def dev(a, device, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a)).to(device=device, dtype=dtype)


def host(t):
    return t.detach().cpu().numpy().astype(np.float64)


def relshift(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0))


def test_se3_against_reference_golden(dx, cuda_device, golden):
    g = golden("se3")
    d = lambda k: dev(g[k], cuda_device)
    A, G = dx.AffineT, dx.AffineGrad
    proc = dx.SE3Diffusion(None).to(cuda_device)
    assert proc.shift_scale == float(g["shift_scale"])
    t, tr = dev(g["t"], cuda_device, torch.int64), dev(g["t_rev"], cuda_device, torch.int64)
    x0 = A(d("rot0"), d("shift0"))
    # q_sample with the reference's noise (diffusion.py:498-506)
    x_t = proc.q_sample(x0, t, noise=A(d("noise_rot"), d("noise_shift")))
    assert np.max(O.geodesic_angle(host(x_t.rot), g["xt_rot"].astype(np.float64))) < 1e-5
    assert relshift(host(x_t.shift), g["xt_shift"]) < 1e-5
    # reverse algebra on the reference's x_t; rows where the reference's own fp32 log is accurate
    ok = O.rmat_to_aa(g["xt_rot"].astype(np.float64))[1][:, 0] < 3.0
    ok2 = ok & (O.rmat_to_aa(g["recon_rot"].astype(np.float64))[1][:, 0] < 3.0)
    xt = A(d("xt_rot"), d("xt_shift"))
    pred = G(d("pred_rot"), d("pred_shift"))
    rec = proc.predict_start_from_noise(xt, tr, pred)
    assert np.max(np.abs(host(rec.rot) - g["recon_rot"])[ok]) < 2e-5 and relshift(host(rec.shift), g["recon_shift"]) < 1e-5
    pm, pv, plv = proc.q_posterior(A(d("recon_rot"), d("recon_shift")), xt, tr)
    assert np.max(np.abs(host(pm.rot) - g["post_rot"])[ok2]) < 2e-5 and relshift(host(pm.shift), g["post_shift"]) < 1e-5
    assert np.array_equal(pv.cpu().numpy(), g["post_var"]) and np.array_equal(plv.cpu().numpy(), g["post_logvar"])
    proc.denoise_fn = lambda x, tt: pred
    mean, _, _ = proc.p_mean_variance(xt, tr)                      # ONE fused launch for both halves
    assert np.max(np.abs(host(mean.rot) - g["pm_rot"])[ok2]) < 3e-5 and relshift(host(mean.shift), g["pm_shift"]) < 1e-5
    # ... and against the fp64 oracle on every row
    s = O.schedule_buffers(1000)
    trn = g["t_rev"]
    wr, ws = O.se3_p_mean(g["xt_rot"], g["xt_shift"], g["pred_rot"], g["pred_shift"], s["sqrt_recip_alphas_cumprod"][trn],
                          s["sqrt_recipm1_alphas_cumprod"][trn], s["posterior_mean_coef1"][trn], s["posterior_mean_coef2"][trn])
    assert np.max(O.geodesic_angle(host(mean.rot), wr)) < 1e-5 and relshift(host(mean.shift), ws) < 1e-5
    # se3_scale and the Euler helpers (util.py:382-422)
    sc = dx.util.se3_scale(x0, d("scal"))
    ok3 = O.rmat_to_aa(g["rot0"].astype(np.float64))[1][:, 0] < 3.0
    assert np.max(np.abs(host(sc.rot) - g["scaled_rot"])[ok3]) < 1e-5 and relshift(host(sc.shift), g["scaled_shift"]) < 1e-6
    e = d("eul")
    Re = dx.util.euler_to_rmat(*e.unbind(-1))
    assert np.max(np.abs(host(Re) - g["eul_rmat"])) < 1e-6
    assert np.max(np.abs(host(torch.stack(dx.util.rmat_to_euler(Re), -1)) - g["eul_back"])) < 1e-5
    assert len(x0) == 96 and x0[3:5].rot.shape == (2, 3, 3) and x0.shape == (96, 3)


@pytest.mark.parametrize("shape", [(1,), (257,), (8, 33), (4099,)])
def test_se3_fused_forward_noising(dx, cuda_device, shape):
    """Fused draw + q_sample + targets: the rotation half is bit-identical to the SO(3) kernel at the same Philox
    state, the translation half satisfies shift_t = a_t shift0 + eps_t shift_scale z with z the returned target,
    and everything equals the oracle algebra applied to the implied noise."""
    proc = dx.SE3Diffusion(None, shift_scale=75.0).to(cuda_device)
    so3 = dx.SO3Diffusion(None).to(cuda_device)
    n = int(np.prod(shape))
    R0 = dev(O.random_rotations(n, np.random.default_rng(n), math.pi)[0], cuda_device).reshape(*shape, 3, 3)
    s0 = torch.randn(*shape, 3, device=cuda_device) * 20
    t = torch.randint(0, 1000, shape[:1], device=cuda_device)
    fwd, _, _ = proc.tables()
    args = (proc.sqrt_alphas_cumprod, proc.sqrt_one_minus_alphas_cumprod, fwd)
    out = dx.ops.se3_q_sample_fused(R0, s0, t, *args, 75.0, seed=5, rng_offset=7, guide=proc.guides()[0])
    t_rows = dx.ops._rows_t(t, shape, cuda_device)
    ref = dx.ops.q_sample_fused(R0, t_rows, *args, seed=5, rng_offset=7, guide=so3.guides()[0])
    assert torch.equal(out["rot"], ref["x_t"]) and torch.equal(out["target_rot"], ref["target"])
    assert out["shift"].shape == (*shape, 3) and out["target_shift"].shape == (*shape, 3)
    a = host(proc.sqrt_alphas_cumprod[t_rows])[..., None]
    e = host(proc.sqrt_one_minus_alphas_cumprod[t_rows])[..., None]
    want = a * host(s0) + e * 75.0 * host(out["target_shift"])
    mag = np.maximum(1.0, np.abs(a * host(s0)) + np.abs(e * 75.0 * host(out["target_shift"])))  # the two terms may cancel
    assert np.max(np.abs(host(out["shift"]) - want) / mag) < 1e-6
    # sharding invariance of both halves
    if len(shape) == 1 and n > 300:
        part = dx.ops.se3_q_sample_fused(R0[300:], s0[300:], t[300:], *args, 75.0, seed=5, rng_offset=7, row_offset=300, guide=proc.guides()[0])
        for k in ("rot", "shift", "target_rot", "target_shift"):
            assert torch.equal(part[k], out[k][300:]), k


def test_se3_translation_noise_statistics(dx, cuda_device):
    n = 1 << 20
    proc = dx.SE3Diffusion(None, shift_scale=3.0).to(cuda_device)
    R0 = dx.ops.quat_to_rmat(torch.randn(n, 4, device=cuda_device))
    t = torch.full((n,), 700, device=cuda_device)
    x_t, tgt = proc.noise_and_target(dx.AffineT(R0, torch.zeros(n, 3, device=cuda_device)), t)
    z = tgt.shift_g.double()
    assert float(z.mean(0).abs().max()) < 4e-3 and float((z.var(0) - 1).abs().max()) < 6e-3
    assert float((torch.corrcoef(z.T) - torch.eye(3, device=cuda_device)).abs().max()) < 4e-3
    eps = float(proc.sqrt_one_minus_alphas_cumprod[700])
    assert abs(float(x_t.shift.double().std()) / (eps * 3.0) - 1) < 3e-3
    # translation noise is independent of the rotation noise of the same row (separate Philox blocks)
    assert float(torch.corrcoef(torch.cat([z, tgt.rot_g.double()], 1).T)[:3, 3:].abs().max()) < 4e-3
    # a second call advances the Philox offset: fresh draws
    _, tgt2 = proc.noise_and_target(dx.AffineT(R0, torch.zeros(n, 3, device=cuda_device)), t)
    assert not torch.equal(tgt2.shift_g, tgt.shift_g)


@pytest.mark.parametrize("shared_t", [True, False])
def test_se3_fused_reverse_step(dx, cuda_device, shared_t):
    """Fused SE(3) reverse step: rotation bit-identical to the SO(3) kernel at the same Philox state; translation
    = posterior mean + sigma_t shift_scale z, z standard normal; t == 0 rows receive no noise."""
    n = 1 << 16
    proc = dx.SE3Diffusion(None, shift_scale=75.0).to(cuda_device)
    Rt = dx.ops.quat_to_rmat(torch.randn(n, 4, device=cuda_device))
    st = torch.randn(n, 3, device=cuda_device) * 30
    pr, ps = torch.randn(n, 3, device=cuda_device) * 0.3, torch.randn(n, 3, device=cuda_device)
    # t < 600 keeps sqrt_recip_alphas_cumprod <= 1.7: beyond, the scaled angle s * theta is ill-conditioned in fp32
    #  and a 1e-5 rad comparison with the fp64 oracle is meaningless
    t = torch.tensor([400], device=cuda_device) if shared_t else torch.randint(0, 600, (n,), device=cuda_device)
    if not shared_t:
        t[:5] = 0
    sched = (proc.sqrt_recip_alphas_cumprod, proc.sqrt_recipm1_alphas_cumprod, proc.posterior_mean_coef1, proc.posterior_mean_coef2)
    _, post, _ = proc.tables()
    pg = None if shared_t else proc.guides()[1]
    rot, shift = dx.ops.se3_p_sample_fused(Rt, st, pr, ps, t, *sched, proc._sigma(), 75.0, post_cdf=post, seed=9, rng_offset=2, post_guide=pg)
    ref = dx.ops.p_sample_fused(Rt, pr, t, *sched, post_cdf=post, seed=9, rng_offset=2, post_guide=pg)
    assert torch.equal(rot, ref)
    mrot, mshift = dx.ops.se3_p_sample_fused(Rt, st, pr, ps, t, *sched, proc._sigma(), 75.0, post_cdf=None)
    tn = np.broadcast_to(t.cpu().numpy(), (n,))
    s = O.schedule_buffers(1000)
    wr, ws = O.se3_p_mean(host(Rt), host(st), host(pr), host(ps), s["sqrt_recip_alphas_cumprod"][tn], s["sqrt_recipm1_alphas_cumprod"][tn],
                          s["posterior_mean_coef1"][tn], s["posterior_mean_coef2"][tn])
    assert np.max(O.geodesic_angle(host(mrot), wr)) < 1e-5
    # shift0_hat = recip shift_t - recipm1 pred amplifies fp32 rounding of shift_t by recip (up to 2e4 at t -> T)
    c1 = s["posterior_mean_coef1"][tn][:, None]
    scale = np.maximum(1.0, c1 * (np.abs(host(st)) * s["sqrt_recip_alphas_cumprod"][tn][:, None] + np.abs(host(ps)) * s["sqrt_recipm1_alphas_cumprod"][tn][:, None])
                       + np.abs(host(st)) * s["posterior_mean_coef2"][tn][:, None])
    assert np.max(np.abs(host(mshift) - ws) / scale) < 2e-6
    sig = host(proc._sigma())[tn][:, None] * 75.0
    z = (host(shift) - host(mshift)) / np.where(tn[:, None] == 0, 1.0, sig)
    live = tn != 0
    assert np.all(z[~live] == 0.0) and (shared_t or (~live).sum() >= 5)
    assert torch.equal(rot[~torch.as_tensor(live)], mrot[~torch.as_tensor(live)])
    zz = z[live]
    assert abs(zz.mean()) < 0.01 and abs(zz.var() - 1) < 0.02


def test_se3_module_end_to_end(dx, cuda_device):
    """SE3Diffusion / ProjectedSE3Diffusion with a small denoiser on (B, N_res) frames: loss is finite and
    differentiable w.r.t. the denoiser, p_sample and p_sample_loop run, state dict matches the SO(3) buffers."""
    B, Nres = 6, 17

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(13, 6)

        def forward(self, x, t):
            tt = t.reshape(tuple(t.shape) + (1,) * (x.shift.dim() - t.dim())).expand(x.shift.shape[:-1] + (1,)).float() / 1000
            h = self.lin(torch.cat([x.rot.flatten(-2), x.shift / 75.0, tt], -1))
            return dx.AffineGrad(h[..., :3], h[..., 3:])

    net = Net().to(cuda_device)
    proc = dx.SE3Diffusion(net).to(cuda_device)
    assert set(k for k, _ in proc.named_buffers()) >= {"betas", "sqrt_alphas_cumprod", "posterior_mean_coef2", "identity"}
    x = dx.AffineT(dx.ops.quat_to_rmat(torch.randn(B, Nres, 4, device=cuda_device)), torch.randn(B, Nres, 3, device=cuda_device) * 10)
    loss = proc(x)
    loss.backward()
    assert torch.isfinite(loss) and net.lin.weight.grad.abs().sum() > 0
    t = torch.full((B,), 321, device=cuda_device)
    y = proc.p_sample(x, t)
    assert y.rot.shape == (B, Nres, 3, 3) and y.shift.shape == (B, Nres, 3) and torch.isfinite(y.rot).all() and torch.isfinite(y.shift).all()
    ortho = y.rot.transpose(-1, -2) @ y.rot - torch.eye(3, device=cuda_device)
    assert float(ortho.abs().max()) < 1e-5
    small = dx.SE3Diffusion(net, timesteps=12).to(cuda_device)
    out = small.p_sample_loop((3, Nres))
    assert out.rot.shape == (3, Nres, 3, 3) and torch.isfinite(out.shift).all()
    pp = dx.ProjectedSE3Diffusion(net, timesteps=12).to(cuda_device)
    assert "identity" in pp.state_dict()                                   # diffusion.py:529
    assert torch.isfinite(pp(x, lambda a: a))
    out = pp.p_sample_loop((3, Nres), lambda a: a)
    assert out.rot.shape == (3, Nres, 3, 3)
    with pytest.raises(RuntimeError):
        dx.SE3Diffusion(net, loss_type="nope").to(cuda_device)(x)
    # IGSO3xR3 (distributions.py:84-110)
    d = dx.IGSO3xR3(torch.full((B,), 0.3, device=cuda_device), shift_scale=2.0)
    smp = d.sample()
    assert smp.rot.shape == (B, 3, 3) and smp.shift.shape == (B, 3)
    assert d.log_prob(smp).shape == (B, 3)
