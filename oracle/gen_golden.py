"""Generate golden input/output vectors by importing the UNMODIFIED reference from /root/reference.

Run in the build container only (the reference does not travel to the GPU box):

    cd /tmp && python /root/repo/oracle/gen_golden.py

Outputs small .npz fixtures into tests/golden/.  The reference modules are imported from where they
lie; nothing is copied.  Importing the submodule creates ./results in the cwd, so run from /tmp.
All reference calls run on CPU in the reference's own dtypes (float32 unless it upcasts itself).
"""
import math
import os
import sys
import warnings

import numpy as np
import torch

REF = os.environ.get("SO3D_REFERENCE", "/root/reference")
sys.path[:0] = [REF, os.path.join(REF, "denoising-diffusion-pytorch")]
warnings.filterwarnings("ignore")

import util as rutil  # noqa: E402  (reference util.py)
import distributions as rdist  # noqa: E402
import diffusion as rdiff  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(1)


def rand_rot(n, gen, max_angle=math.pi):
    """Random rotations built in fp64 Rodrigues and cast to fp32 (inputs, not reference outputs)."""
    axis = torch.randn(n, 3, generator=gen, dtype=torch.float64)
    axis = axis / axis.norm(dim=-1, keepdim=True)
    ang = torch.rand(n, generator=gen, dtype=torch.float64) * max_angle
    k = torch.zeros(n, 3, 3, dtype=torch.float64)
    k[:, 0, 1], k[:, 0, 2], k[:, 1, 0] = -axis[:, 2], axis[:, 1], axis[:, 2]
    k[:, 1, 2], k[:, 2, 0], k[:, 2, 1] = -axis[:, 0], -axis[:, 1], axis[:, 0]
    r = torch.eye(3, dtype=torch.float64) + torch.sin(ang)[:, None, None] * k + (1 - torch.cos(ang))[:, None, None] * (k @ k)
    return r.float(), axis.float(), ang.float()


def save(name, **arrs):
    arrs = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrs.items()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
    print(name, {k: v.shape for k, v in arrs.items()})


def main():
    g = torch.Generator().manual_seed(1234)

    # ---- L0: util.py ------------------------------------------------------------------------
    n = 192
    R, axis, ang = rand_rot(n, g, 3.0)
    R2, _, _ = rand_rot(n, g, 3.0)
    small, _, small_ang = rand_rot(32, g, 1e-3)
    Rall = torch.cat([R, small, torch.eye(3)[None]])
    log_all = rutil.log_rmat(Rall)
    ax_o, ang_o = rutil.rmat_to_aa(R)
    scal = torch.exp(torch.rand(n, generator=g) * math.log(40.0) - math.log(20.0))  # [0.05, 2]
    q = torch.randn(n, 4, generator=g)
    v3 = torch.randn(n, 3, generator=g)
    w = torch.rand(n, 1, generator=g)
    axes_in = torch.randn(n, 3, generator=g) * 3.0  # un-normalised on purpose (util.py:201)
    save(
        "util_l0",
        R=R, R2=R2, Rall=Rall, log_all=log_all, axis=ax_o, angle=ang_o, true_axis=axis, true_angle=ang,
        scalars=scal, scaled=rutil.so3_scale(R, scal),
        axes_in=axes_in, ang_in=ang[:, None], aa_rmat=rutil.aa_to_rmat(axes_in, ang[:, None]),
        quat=q, quat_rmat=rutil.quat_to_rmat(q),
        vec=v3, skew=rutil.vec2skew(v3), vee=rutil.skew2vec(rutil.vec2skew(v3)),
        lerp_w=w, lerp=rutil.so3_lerp(R, R2, w), dist=rutil.rmat_dist(R, R2),
        expvec=torch.matrix_exp(rutil.vec2skew(v3)),
        pi_z=rutil.log_rmat(torch.diag(torch.tensor([-1.0, -1.0, 1.0]))[None]),  # util.py:507-512
    )

    # ---- L1: IGSO3 density, table, sampler, log_prob, autograd score ---------------------------
    eps_list = [0.05, 0.1, 0.2, 0.5, 1.0]
    omega = torch.cat([torch.zeros(1), torch.logspace(-4, 0, 40), torch.linspace(1.0, math.pi, 24)])
    dens, traps, samples, axes_draw, u_draw, logps, grads, lp_R = [], [], [], [], [], [], [], []
    for e in eps_list:
        d = rdist.IsotropicGaussianSO3(torch.tensor(e))
        dens.append(d._eps_ft(omega))
        traps.append(d.trap[:, 0])
        torch.manual_seed(777)
        samples.append(d.sample((256,)))
        torch.manual_seed(777)
        axes_draw.append(torch.randn((256, 3)))
        u_draw.append(torch.rand((256,)))
        # log_prob + autograd wrt the matrix entries (distributions.py:186-190)
        a = torch.randn(96, 3, generator=g)
        an = (torch.rand(96, 1, generator=g) * min(3.0, 5.0 * e)).clamp(min=1e-3)
        Rm = rutil.aa_to_rmat(a, an).detach().requires_grad_(True)
        lp = d.log_prob(Rm)
        (gr,) = torch.autograd.grad(lp.sum(), Rm)
        lp_R.append(Rm.detach()); logps.append(lp.detach()); grads.append(gr)
    d0 = rdist.IsotropicGaussianSO3(torch.tensor(0.5))
    # small-eps quirk (Q3/D5): density zeroed beyond omega > 709 eps^2/pi, table built from it
    dq = rdist.IsotropicGaussianSO3(torch.tensor(0.0064))
    om_q = torch.linspace(0.0, 0.05, 64)
    # batched eps (training path): table (999, B) and the column-0 gather bug Q1
    eps_b = torch.tensor([0.9, 0.3, 0.08, 0.5, 0.02, 0.0064])
    db = rdist.IsotropicGaussianSO3(eps_b)
    torch.manual_seed(4242)
    samp_b = db.sample()
    torch.manual_seed(4242)
    axes_b = torch.randn((6, 3)); u_b = torch.rand((6,))
    save(
        "igso3",
        eps_list=np.array(eps_list, dtype=np.float32), omega=omega, density=torch.stack(dens),
        grid_loc=math.pi * torch.linspace(0, 1.0, 1000) ** 3.0,
        grid_haar=(1 - (math.pi * torch.linspace(0, 1.0, 1000) ** 3.0).cos()) / math.pi,
        trap=torch.stack(traps), trap_loc=d0.trap_loc[:, 0],
        samples=torch.stack(samples), axes_draw=torch.stack(axes_draw), u_draw=torch.stack(u_draw),
        lp_R=torch.stack(lp_R), logp=torch.stack(logps), logp_grad=torch.stack(grads),
        q_eps=np.float32(0.0064), q_omega=om_q, q_density=dq._eps_ft(om_q), q_trap=dq.trap[:, 0],
        b_eps=eps_b, b_trap=db.trap, b_samples=samp_b, b_axes=axes_b, b_u=u_b,
    )

    # ---- L2: schedule + SO3Diffusion ----------------------------------------------------------
    proc = rdiff.SO3Diffusion(None)
    bufs = {k: v for k, v in proc.named_buffers()}
    save("schedule", **bufs)

    B = 128
    x0, _, _ = rand_rot(B, g)
    t = torch.randint(0, 1000, (B,), generator=g)
    t[:4] = torch.tensor([0, 1, 998, 999])
    eps_t = bufs["sqrt_one_minus_alphas_cumprod"][t]
    # noise drawn per-row from scalar-eps distributions (avoids the Q1 batched gather bug)
    noise = torch.stack([rdist.IsotropicGaussianSO3(e).sample()[0] for e in eps_t])  # scalar eps -> (1,3,3)
    x_t = proc.q_sample(x0, t, noise=noise)
    target = rutil.skew2vec(rutil.log_rmat(noise)) * (1 / eps_t)[..., None]  # diffusion.py:355
    pred = torch.randn(B, 3, generator=g) * 0.5
    # reverse-step algebra; restrict to t where so3_scale's matrix_exp is still accurate (Q5)
    t_rev = torch.randint(0, 600, (B,), generator=g)
    t_rev[:2] = torch.tensor([0, 1])
    x_recon = proc.predict_start_from_noise(x_t, t_rev, pred)
    post_mean, post_var, post_logvar = proc.q_posterior(x_recon, x_t, t_rev)
    proc.denoise_fn = lambda x, tt: pred
    mean_pm, _, _ = proc.p_mean_variance(x_t, t_rev, clip_denoised=False)
    # full p_sample at shared t (scalar-eps sampler path), with the draws recorded
    steps = [999, 500, 100, 1, 0]
    ps_out, ps_axes, ps_u = [], [], []
    for s in steps:
        tt = torch.full((B,), s, dtype=torch.long)
        torch.manual_seed(99 + s)
        ps_out.append(proc.p_sample(x_t, tt))
        torch.manual_seed(99 + s)
        ps_axes.append(torch.randn((B, 3))); ps_u.append(torch.rand((B,)))
    save(
        "diffusion",
        x0=x0, t=t, eps_t=eps_t, noise=noise, x_t=x_t, target=target, pred=pred, t_rev=t_rev,
        x_recon=x_recon, post_mean=post_mean, post_var=post_var, post_logvar=post_logvar, mean_pm=mean_pm,
        ps_steps=np.array(steps), ps_out=torch.stack(ps_out), ps_axes=torch.stack(ps_axes), ps_u=torch.stack(ps_u),
    )


def main_eval():
    """Evaluation / data-side fixtures: Bingham.rsample with its normals recorded, and the
    reference's MMD (util.py:254-285) with both SO(3) kernels, un-chunked and chunked."""
    cov3 = torch.tensor([[1.0, 0, 0, 0], [0, 1.0, 0.9, 0.9], [0, 0.9, 1.0, 0.9], [0, 0.9, 0.9, 1.0]])  # bingham_train.py:66-71
    cov1 = torch.diag(torch.tensor([1000.0, 0.1, 0.1, 0.1]))                                            # bingham_train.py:55
    loc = torch.zeros(4)
    out = {}
    for name, cov, n in (("c3", cov3, 300), ("c1", cov1, 333)):
        d = rdist.Bingham(loc, covariance_matrix=cov)
        torch.manual_seed(2024)
        q = d.sample((n,))
        torch.manual_seed(2024)
        z = torch.empty(n, 4).normal_()  # what MultivariateNormal.rsample draws (_standard_normal)
        chk = z @ d.scale_tril.T
        chk = chk / chk.norm(dim=-1, keepdim=True)
        assert torch.allclose(chk, q, atol=1e-6), "normals do not reproduce the reference's Bingham draw"
        out.update({f"{name}_cov": cov, f"{name}_tril": d.scale_tril, f"{name}_z": z, f"{name}_q": q, f"{name}_R": rutil.quat_to_rmat(q)})
    X, Y = out["c3_R"], out["c1_R"]
    out["mmd_gauss"] = rutil.MMD(X, Y, rutil.rmat_gaussian_kernel)
    out["mmd_gauss_chunk128"] = rutil.MMD(X, Y, rutil.rmat_gaussian_kernel, chunksize=128)
    out["mmd_cos"] = rutil.MMD(X, Y, rutil.rmat_cosine_kernel)
    out["mmd_gauss_same"] = rutil.MMD(X[:150], X[150:], rutil.rmat_gaussian_kernel)
    out["ker_gauss_xy"] = rutil.rmat_gaussian_kernel(X[:64].unsqueeze(0), Y[:48].unsqueeze(1))
    out["ker_cos_xy"] = rutil.rmat_cosine_kernel(X[:64].unsqueeze(0), Y[:48].unsqueeze(1))
    out["cos_dist"] = rutil.rmat_cosine_dist(X[:64], Y[:64])
    out["test_same"] = np.array(rutil.Ker_2samp_test(X[:150], X[150:], rutil.rmat_gaussian_kernel))
    out["test_diff"] = np.array(rutil.Ker_2samp_test(X, Y[:300], rutil.rmat_gaussian_kernel))
    out["logp_same"] = np.array(rutil.Ker_2samp_log_prob(X[:150], X[150:], rutil.rmat_gaussian_kernel))
    save("evaluation", **out)


def main_se3():
    """SE(3) arm fixtures: the reference's SE3Diffusion forward / reverse algebra on AffineT batches
    with explicit noise, IGSO3xR3 draws with the underlying normals recorded, se3_scale and the Euler helpers."""
    g = torch.Generator().manual_seed(4321)
    proc = rdiff.SE3Diffusion(None)
    bufs = {k: v for k, v in proc.named_buffers()}
    B, ss = 96, proc.shift_scale
    rot0, _, _ = rand_rot(B, g)
    x0 = rutil.AffineT(rot0, torch.randn(B, 3, generator=g) * 10.0)
    t = torch.randint(0, 1000, (B,), generator=g)
    t[:4] = torch.tensor([0, 1, 998, 999])
    eps_t = bufs["sqrt_one_minus_alphas_cumprod"][t]
    torch.manual_seed(4322)  # the reference's sampler draws from the GLOBAL generators (distributions.py:35,38): seed them, so the fixture regenerates bit for bit
    noise_rot = torch.stack([rdist.IsotropicGaussianSO3(e).sample()[0] for e in eps_t])
    z = torch.randn(B, 3, generator=g)
    noise = rutil.AffineT(noise_rot, z * (eps_t * ss)[:, None])
    x_t = proc.q_sample(x0, t, noise=noise)
    tgt_shift = noise.shift * (1 / (eps_t * ss))[..., None]                               # diffusion.py:514
    tgt_rot = rutil.skew2vec(rutil.log_rmat(noise.rot)) * (1 / eps_t)[..., None]          # diffusion.py:515
    pred = rutil.AffineGrad(torch.randn(B, 3, generator=g) * 0.5, torch.randn(B, 3, generator=g))
    t_rev = torch.randint(0, 600, (B,), generator=g)
    t_rev[:2] = torch.tensor([0, 1])
    x_recon = proc.predict_start_from_noise(x_t, t_rev, pred)
    post_mean, post_var, post_logvar = proc.q_posterior(x_recon, x_t, t_rev)
    proc.denoise_fn = lambda x, tt: pred
    mean_pm, _, _ = proc.p_mean_variance(x_t, t_rev, clip_denoised=False)
    # IGSO3xR3 with a scalar eps and a batched mean (what SE3Diffusion.p_sample builds, diffusion.py:482): the
    # rotation noise is ONE draw shared by the whole batch (axes = randn((3,)), quirk Q12), the shift noise is per row
    sig = torch.tensor(0.25)
    d = rdist.IGSO3xR3(eps=sig, mean=post_mean, shift_scale=ss)
    torch.manual_seed(55)
    smp = d.sample()
    torch.manual_seed(55)
    ax = torch.randn((3,)); u = torch.rand(()); zz = torch.empty(B, 3).normal_()
    assert torch.allclose(smp.shift, post_mean.shift + zz * sig * ss, atol=1e-4), "normals do not reproduce the reference's shift draw"
    scal = torch.rand(B, generator=g) * 1.5 + 0.1
    sc = rutil.se3_scale(x0, scal)
    eul = torch.rand(B, 3, generator=g) * 2 - 1
    R_e = rutil.euler_to_rmat(*eul.unbind(-1))
    save(
        "se3",
        shift_scale=np.float32(ss), rot0=x0.rot, shift0=x0.shift, t=t, eps_t=eps_t, noise_rot=noise.rot, noise_shift=noise.shift, z=z,
        xt_rot=x_t.rot, xt_shift=x_t.shift, tgt_rot=tgt_rot, tgt_shift=tgt_shift, pred_rot=pred.rot_g, pred_shift=pred.shift_g, t_rev=t_rev,
        recon_rot=x_recon.rot, recon_shift=x_recon.shift, post_rot=post_mean.rot, post_shift=post_mean.shift, post_var=post_var,
        post_logvar=post_logvar, pm_rot=mean_pm.rot, pm_shift=mean_pm.shift,
        smp_sigma=sig, smp_rot=smp.rot, smp_shift=smp.shift, smp_axis=ax, smp_u=u, smp_z=zz,
        scal=scal, scaled_rot=sc.rot, scaled_shift=sc.shift, eul=eul, eul_rmat=R_e, eul_back=torch.stack(rutil.rmat_to_euler(R_e), -1),
    )


def main_denoiser():
    """RotPredict denoiser fixtures: the reference's so3_train.RotPredict (out_type 'skewvec') with
    seeded weights, evaluated for a shared step index (so3_test.py:31) and per-row steps, and the reference's
    SO3Diffusion.p_mean_variance driven by it (diffusion.py:308-313).  models.py / prot_util.py import packages
    that are not installed, so those imports are stubbed; the stubs are never called."""
    import types

    for name in ("se3_transformer_pytorch", "se3_transformer_pytorch.se3_transformer_pytorch", "Bio", "Bio.PDB", "wandb"):
        sys.modules.setdefault(name, types.ModuleType(name))
    m = sys.modules["se3_transformer_pytorch.se3_transformer_pytorch"]
    m.LinearSE3 = m.Fiber = m.NormSE3 = object
    sys.modules["se3_transformer_pytorch"].SE3Transformer = object
    import so3_train as rtrain  # reference so3_train.py

    torch.manual_seed(77)
    net = rtrain.RotPredict(out_type="skewvec")
    with torch.no_grad():  # trained-network-sized activations instead of the near-zero default init
        for p_ in net.parameters():
            p_.mul_(2.5)
    g = torch.Generator().manual_seed(78)
    n = 300
    R, _, _ = rand_rot(n, g)
    out = {k.replace(".", "_"): v for k, v in net.state_dict().items()}
    with torch.no_grad():
        shared = [0, 1, 500, 999]
        out["t_shared"] = np.array(shared)
        out["pred_shared"] = torch.stack([net(R, torch.tensor([tv])) for tv in shared])
        t_row = torch.randint(0, 1000, (n,), generator=g)
        out["t_row"] = t_row
        out["pred_row"] = net(R, t_row)
        proc = rdiff.SO3Diffusion(net)
        out["mean_shared"] = torch.stack([proc.p_mean_variance(R, torch.full((n,), tv, dtype=torch.long), clip_denoised=False)[0] for tv in (1, 300, 999)])
        out["t_mean"] = np.array([1, 300, 999])
    save("rotpredict", x=R, **out)


if __name__ == "__main__":
    todo = [a for a in sys.argv[1:] if a in ("core", "eval", "se3", "denoiser")] or ["core", "eval", "se3", "denoiser"]
    for name in todo:
        {"core": main, "eval": main_eval, "se3": main_se3, "denoiser": main_denoiser}[name]()
