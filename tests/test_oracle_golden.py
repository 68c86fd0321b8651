"""Pin the CPU oracle (oracle/so3_oracle.py) against golden vectors produced by the UNMODIFIED
reference (oracle/gen_golden.py, run in the build container).  CPU only.

Tolerances: the golden outputs are the reference's float32 results (its own rounding included),
the oracle is float64, so agreement is limited by the reference's fp32 error, stated per test.
"""
import math

import numpy as np

from oracle import so3_oracle as O


def rel_err(a, b, floor=0.0):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), floor if floor > 0 else 1e-300))


def test_hat_vee(golden):
    g = golden("util_l0")
    assert np.array_equal(O.vec2skew(g["vec"]), g["skew"])
    assert np.array_equal(O.skew2vec(g["skew"]), g["vee"])


def test_log_rmat(golden):
    g = golden("util_l0")
    got = O.log_rmat(g["Rall"].astype(np.float64), reference_quirks=True)
    assert np.max(np.abs(got - g["log_all"])) < 2e-6  # reference fp32 log, angles <= 3.0
    # default (robust) mode agrees on this set too (no exactly-pi inputs)
    got2 = O.log_rmat(g["Rall"].astype(np.float64))
    assert np.max(np.abs(got2 - g["log_all"])) < 2e-6
    # identity row -> exactly zero (util.py:500)
    assert np.all(got[-1] == 0)


def test_log_rmat_pi_about_z(golden):
    g = golden("util_l0")
    r = np.diag([-1.0, -1.0, 1.0])[None]
    got = O.log_rmat(r)
    # reference known answer util.py:507-512: pi about z (sign of the axis is arbitrary at pi)
    assert np.allclose(np.abs(got), np.abs(g["pi_z"]), atol=1e-6)
    assert np.isclose(abs(O.log_vec(r)[0, 2]), math.pi)


def test_rmat_to_aa(golden):
    g = golden("util_l0")
    axis, ang = O.rmat_to_aa(g["R"])
    assert np.max(np.abs(ang - g["angle"])) < 2e-6
    assert np.max(np.abs(axis - g["axis"])) < 5e-6
    # and against the fp64 construction parameters
    assert np.max(np.abs(ang[:, 0] - g["true_angle"])) < 1e-6


def test_aa_to_rmat(golden):
    g = golden("util_l0")
    got = O.aa_to_rmat(g["axes_in"], g["ang_in"])
    assert np.max(np.abs(got - g["aa_rmat"])) < 2e-6  # matrix_exp + SVD vs Rodrigues: 4e-7
    assert np.max(np.abs(O.exp_vec(g["vec"]) - g["expvec"])) < 2e-6


def test_so3_scale(golden):
    g = golden("util_l0")
    got = O.so3_scale(g["R"], g["scalars"])
    assert np.max(np.abs(got - g["scaled"])) < 5e-6


def test_quat_lerp_dist(golden):
    g = golden("util_l0")
    assert np.max(np.abs(O.quat_to_rmat(g["quat"]) - g["quat_rmat"])) < 1e-6
    # A^T B can come close to pi where the reference's fp32 log loses the axis: 2e-5
    assert np.max(np.abs(O.so3_lerp(g["R"], g["R2"], g["lerp_w"]) - g["lerp"])) < 2e-5
    assert np.max(np.abs(O.rmat_dist(g["R"], g["R2"]) - g["dist"])) < 5e-6
    q = O.rmat_to_quat_canonical(g["R"])
    assert np.max(np.abs(O.quat_to_rmat(q) - g["R"])) < 1e-6


def test_grid(golden):
    g = golden("igso3")
    locs, haar = O.grid_f32()
    # ATen's vectorised linspace rounds a few entries differently from the scalar restatement
    assert np.max(np.abs(locs - g["grid_loc"])) < 1e-6
    # the fp32 cos differs by an ulp between libms; 1-cos amplifies it near 0.  Production code
    # takes this table from torch on the host, exactly like the reference; here only closeness.
    assert np.max(np.abs(haar - g["grid_haar"])) < 5e-7
    assert np.array_equal(g["grid_loc"][1:], g["trap_loc"])


def test_density_closed_vs_reference(golden):
    g = golden("igso3")
    for e, ref in zip(g["eps_list"], g["density"]):
        got = O.igso3_closed(g["omega"].astype(np.float64), float(e))
        lo = 0 if e >= 0.2 else 1  # the reference's omega==0 limit is NaN below eps = 0.167 (Q3)
        # D5: the reference zeroes the density for omega > 709 eps^2/pi (0*inf = NaN -> 0); compare the
        # stable form only below that cut-off, the quirk form (below) everywhere
        ok = g["omega"] <= 0.99 * 709.0 * float(e) ** 2 / math.pi
        ok[:lo] = False
        assert rel_err(got[ok], ref[ok], 1e-30) < 3e-7, e  # fp32 rounding of the reference's output
        if lo:
            assert np.isnan(ref[0]) and np.isclose(got[0], O.igso3_series(0.0, float(e))[0][0], rtol=1e-12)
        got_q = O.igso3_closed(g["omega"].astype(np.float64), float(e), reference_quirks=True)
        ok = np.isfinite(got_q) & (ref > 0)
        if e >= 0.2:  # the reference's omega==0 limit overflows below eps = 0.167 (Q3)
            assert rel_err(got_q, ref, 1e-30) < 3e-7
        else:
            assert rel_err(got_q[1:], ref[1:], 1e-30) < 3e-7


def test_density_series_vs_reference(golden):
    """The reference closed form equals the fp64 series to <= 2e-7 for eps in [0.05, 1];
    at eps = 1 the 3-image closed form itself is only good to ~2e-5."""
    g = golden("igso3")
    for e, ref in zip(g["eps_list"], g["density"]):
        f, _ = O.igso3_series(g["omega"].astype(np.float64), float(e))
        tol = 3e-5 if e >= 1.0 else 5e-7
        # compare where the density is not vanishing (fp32 output of the reference keeps rel. precision)
        keep = ref > 1e-9 * np.nanmax(ref)  # beyond that the fp64 series itself is cancellation noise
        keep[0] = e >= 0.2
        assert rel_err(f[keep], ref[keep]) < tol, e


def test_series_forms_agree():
    om = np.linspace(0.05, 3.1, 50)
    for e in (0.05, 0.3, 1.0, 2.0):
        f1, g1 = O.igso3_series(om, e)
        f2, g2 = O.igso3_series_sincot(om, e)
        keep = f1 > 1e-6 * f1.max()
        assert rel_err(f1[keep], f2[keep]) < 1e-9
        assert np.max(np.abs(g1 - g2)[keep]) < 1e-7 * np.max(np.abs(g1[keep]))
    # omega = 0 limit: sum (2l+1)^2 exp(-l(l+1) eps^2)
    l = np.arange(2000.0)
    assert np.isclose(O.igso3_series(0.0, 0.5)[0][0], ((2 * l + 1) ** 2 * np.exp(-l * (l + 1) * 0.25)).sum(), rtol=1e-13)


def test_closed_dlog_matches_series():
    om = np.linspace(0.01, 3.0, 60)
    for e in (0.05, 0.2, 0.5):
        fs, gs = O.igso3_series(om, e)
        gc = O.igso3_closed_dlog(om, e)
        keep = fs > 1e-6 * fs.max()
        assert np.max((np.abs(gs - gc) / np.maximum(np.abs(gs), 1e-3))[keep]) < 1e-6


def test_small_eps_quirk(golden):
    g = golden("igso3")
    e = float(g["q_eps"])
    om = g["q_omega"].astype(np.float64)
    got_q = O.igso3_closed(om, e, reference_quirks=True)
    ref = g["q_density"]
    assert np.array_equal(got_q[1:] == 0, ref[1:] == 0)  # same zeroed region (omega > 709 eps^2/pi)
    nz = ref > 0
    nz[0] = False
    assert rel_err(got_q[nz], ref[nz]) < 3e-7
    # the stable form is not zeroed there and matches the series
    stable = O.igso3_closed(om[1:], e)
    ser, _ = O.igso3_series(om[1:], e)
    keep = ser > 1e-200
    assert rel_err(stable[keep], ser[keep]) < 1e-9


def test_cdf_table(golden):
    g = golden("igso3")
    trap, loc = O.igso3_cdf_table(g["eps_list"], g["grid_loc"], g["grid_haar"])
    assert np.array_equal(loc, g["trap_loc"])
    assert np.max(np.abs(trap - g["trap"])) <= 1.2e-7  # ~1 ulp at 1.0
    # quirk table at the schedule's minimum eps
    trq, _ = O.igso3_cdf_table([float(g["q_eps"])], g["grid_loc"], g["grid_haar"], reference_quirks=True)
    assert np.max(np.abs(trq[0] - g["q_trap"])) <= 1.2e-7
    # batched-eps constructor (999, B) layout is the transpose of ours
    trb, _ = O.igso3_cdf_table(g["b_eps"], g["grid_loc"], g["grid_haar"], reference_quirks=True)
    assert np.max(np.abs(trb.T - g["b_trap"])) <= 1.2e-7


def test_sampler_given_draws(golden):
    g = golden("igso3")
    for k, e in enumerate(g["eps_list"]):
        r, ang = O.igso3_sample_given(g["u_draw"][k], g["axes_draw"][k], g["trap"][k], g["trap_loc"])
        # angle of the reference sample
        ref_ang = O.rmat_to_aa(g["samples"][k])[1][:, 0]
        assert np.max(np.abs(ang - ref_ang)) < 1e-5, e   # north-star tolerance: 1e-5 rad
        assert np.max(np.abs(r - g["samples"][k])) < 5e-6


def test_sampler_batched_eps_quirk(golden):
    """Q1: with batched eps the reference gathers the lerp endpoints from column 0."""
    g = golden("igso3")
    tr = g["b_trap"].T
    loc = g["trap_loc"]
    ang_q = O.igso3_angle_from_uniform(g["b_u"], tr, loc, trap_row_for_weight=tr[0])
    r_q = O.aa_to_rmat(g["b_axes"].astype(np.float64), ang_q.astype(np.float64)[:, None])
    assert np.max(np.abs(r_q - g["b_samples"])) < 5e-6
    # row 0 is unaffected by the bug
    ang = O.igso3_angle_from_uniform(g["b_u"], tr, loc)
    assert ang[0] == ang_q[0]


def test_log_prob_and_ambient_grad(golden):
    g = golden("igso3")
    for k, e in enumerate(g["eps_list"]):
        R = g["lp_R"][k].astype(np.float64)
        lp = O.igso3_log_prob(R, float(e))
        assert np.max(np.abs(lp - g["logp"][k])) < 2e-5 + 1e-6 * np.max(np.abs(g["logp"][k]))
        _, ang = O.rmat_to_aa(R)
        gd = O.igso3_closed_dlog(ang[:, 0], float(e))
        amb = O.log_prob_ambient_grad(R, gd)
        ref = g["logp_grad"][k]
        scale = np.max(np.abs(ref), axis=(-1, -2), keepdims=True)
        # autograd through the reference's fp32 log_rmat: ~1e-4 relative noise at small angles
        assert np.max(np.abs(amb - ref) / scale) < 2e-3, e


def test_schedule(golden):
    g = golden("schedule")
    bufs = O.schedule_buffers(1000)
    for k, v in bufs.items():
        assert np.array_equal(v, g[k]), k


def test_diffusion_algebra(golden):
    g = golden("diffusion")
    s = golden("schedule")
    t = g["t"]
    x_t = O.q_sample(g["x0"], s["sqrt_alphas_cumprod"][t], g["noise"])
    assert np.max(np.abs(x_t - g["x_t"])) < 5e-6
    tgt = O.skewvec_target(g["noise"], g["eps_t"])
    assert np.max(np.abs(tgt - g["target"]) / np.maximum(np.abs(g["target"]), 1.0)) < 2e-5
    tr = g["t_rev"]
    xr = O.predict_start_from_noise(g["x_t"], g["pred"], s["sqrt_recip_alphas_cumprod"][tr], s["sqrt_recipm1_alphas_cumprod"][tr])
    # reference so3_scale error grows with the scalar (Q5); t_rev < 600 keeps scalars <= ~1.7
    ok = O.rmat_to_aa(g["x_t"])[1][:, 0] < 3.0   # the reference's fp32 log degrades towards pi
    assert np.max(np.abs(xr - g["x_recon"])[ok]) < 2e-5
    pm = O.q_posterior_mean(g["x_recon"], g["x_t"], s["posterior_mean_coef1"][tr], s["posterior_mean_coef2"][tr])
    ok2 = ok & (O.rmat_to_aa(g["x_recon"])[1][:, 0] < 3.0)
    assert np.max(np.abs(pm - g["post_mean"])[ok2]) < 2e-5
    mm = O.p_sample_mean(g["x_t"], g["pred"], s["sqrt_recip_alphas_cumprod"][tr], s["sqrt_recipm1_alphas_cumprod"][tr],
                         s["posterior_mean_coef1"][tr], s["posterior_mean_coef2"][tr])
    assert np.max(np.abs(mm - g["mean_pm"])[ok2]) < 3e-5
    assert np.array_equal(s["posterior_variance"][tr], g["post_var"])


# ---------------------------------------------------------------------------------------------
# evaluation / data side: MMD with the SO(3) kernels, Bingham draws
# ---------------------------------------------------------------------------------------------
def test_pair_kernels_and_mmd(golden):
    g = golden("evaluation")
    X, Y = g["c3_R"], g["c1_R"]
    kg = O.rmat_gaussian_kernel(X[:64][None], Y[:48][:, None])
    assert np.max(np.abs(kg - g["ker_gauss_xy"])) < 2e-6          # reference fp32 atan2 / exp
    kc = O.rmat_cosine_kernel(X[:64][None], Y[:48][:, None])
    assert np.max(np.abs(kc - g["ker_cos_xy"])) < 1e-6
    assert np.max(np.abs((1 - O.rmat_cosine_kernel(X[:64], Y[:64])) - g["cos_dist"])) < 1e-6
    # MMD: the reference sums ~1e5 fp32 kernel values per term, so agreement is ~1e-6 absolute
    assert abs(O.mmd(X, Y) - float(g["mmd_gauss"])) < 3e-6
    assert abs(O.mmd(X, Y, chunk=128) - float(g["mmd_gauss_chunk128"])) < 3e-6
    assert abs(O.mmd(X, Y, O.rmat_cosine_kernel) - float(g["mmd_cos"])) < 3e-6
    assert abs(O.mmd(X[:150], X[150:]) - float(g["mmd_gauss_same"])) < 3e-6
    assert float(g["mmd_gauss"]) > 0.1 > float(g["mmd_gauss_same"]) > 0.0
    assert bool(g["test_same"]) and not bool(g["test_diff"])


def test_bingham_given_normals(golden):
    g = golden("evaluation")
    for name in ("c3", "c1"):
        q = O.bingham_sample_given(g[name + "_z"], g[name + "_tril"])
        assert np.max(np.abs(q - g[name + "_q"])) < 1e-6
        assert np.max(np.abs(O.quat_to_rmat(q) - g[name + "_R"])) < 2e-6


# ---------------------------------------------------------------------------------------------
# SE(3) arm
# ---------------------------------------------------------------------------------------------
def test_se3_diffusion_algebra(golden):
    g = golden("se3")
    s = golden("schedule")
    t, tr, ss = g["t"], g["t_rev"], float(g["shift_scale"])
    assert ss == 75.0
    rot, shift = O.se3_q_sample(g["rot0"], g["shift0"], s["sqrt_alphas_cumprod"][t], g["noise_rot"], g["noise_shift"])
    assert np.max(np.abs(rot - g["xt_rot"])) < 5e-6
    assert np.max(np.abs(shift - g["xt_shift"])) < 1e-5 * max(1.0, np.abs(g["xt_shift"]).max())
    tgt_rot, tgt_shift = O.se3_targets(g["noise_rot"], g["noise_shift"], g["eps_t"], ss)
    assert np.max(np.abs(tgt_rot - g["tgt_rot"]) / np.maximum(np.abs(g["tgt_rot"]), 1.0)) < 2e-5
    assert np.max(np.abs(tgt_shift - g["tgt_shift"])) < 2e-6 and np.max(np.abs(g["tgt_shift"] - g["z"])) < 2e-6
    ok = O.rmat_to_aa(g["xt_rot"])[1][:, 0] < 3.0
    r0, s0 = O.se3_predict_start(g["xt_rot"], g["xt_shift"], g["pred_rot"], g["pred_shift"], s["sqrt_recip_alphas_cumprod"][tr],
                                 s["sqrt_recipm1_alphas_cumprod"][tr])
    assert np.max(np.abs(r0 - g["recon_rot"])[ok]) < 2e-5
    assert np.max(np.abs(s0 - g["recon_shift"]) / np.maximum(np.abs(g["recon_shift"]), 1.0)) < 1e-5
    ok2 = ok & (O.rmat_to_aa(g["recon_rot"])[1][:, 0] < 3.0)
    pr, ps = O.se3_posterior_mean(g["recon_rot"], g["recon_shift"], g["xt_rot"], g["xt_shift"], s["posterior_mean_coef1"][tr],
                                  s["posterior_mean_coef2"][tr])
    assert np.max(np.abs(pr - g["post_rot"])[ok2]) < 2e-5
    assert np.max(np.abs(ps - g["post_shift"]) / np.maximum(np.abs(g["post_shift"]), 1.0)) < 1e-5
    mr, ms = O.se3_p_mean(g["xt_rot"], g["xt_shift"], g["pred_rot"], g["pred_shift"], s["sqrt_recip_alphas_cumprod"][tr],
                          s["sqrt_recipm1_alphas_cumprod"][tr], s["posterior_mean_coef1"][tr], s["posterior_mean_coef2"][tr])
    assert np.max(np.abs(mr - g["pm_rot"])[ok2]) < 3e-5
    assert np.max(np.abs(ms - g["pm_shift"]) / np.maximum(np.abs(g["pm_shift"]), 1.0)) < 1e-5
    # IGSO3xR3 draw (distributions.py:84-110): shift = mean.shift + sigma shift_scale z with the recorded normals
    assert np.max(np.abs(g["post_shift"] + g["smp_z"] * float(g["smp_sigma"]) * ss - g["smp_shift"])) < 1e-4
    # se3_scale (util.py:382-385)
    sr, sh = O.se3_scale(g["rot0"], g["shift0"], g["scal"])
    assert np.max(np.abs(sh - g["scaled_shift"])) < 1e-5
    ok3 = O.rmat_to_aa(g["rot0"])[1][:, 0] < 3.0
    assert np.max(np.abs(sr - g["scaled_rot"])[ok3]) < 1e-5


def rotpredict_params(g):
    return [g[f"net_{i}_weight"] for i in (0, 2, 4, 6, 8)], [g[f"net_{i}_bias"] for i in (0, 2, 4, 6, 8)]


def test_rotpredict_forward(golden):
    """RotPredict (so3_train.py:11-49) for shared and per-row step indices; the golden outputs are the reference's
    fp32 forward (error ~1e-6 on outputs of size ~1).  The time features are sin/cos(t * f_j) with f_j an fp32
    exp(): a 1-ulp difference between exp implementations (numpy / ATen CPU / CUDA) moves the phase by t * 6e-8,
    so the pin is 5e-6 at small t and 1e-4 over the whole range -- a conditioning property of the reference's
    embedding, not of this restatement."""
    g = golden("rotpredict")
    w, b = rotpredict_params(g)
    for k, tv in enumerate(g["t_shared"]):
        got = O.rotpredict_forward(w, b, g["x"], np.array([tv]))
        assert np.max(np.abs(got - g["pred_shared"][k])) < (5e-6 if tv <= 1 else 1e-4)
    got = O.rotpredict_forward(w, b, g["x"], g["t_row"])
    assert np.max(np.abs(got - g["pred_row"])) < 1e-4
    assert np.max(np.abs(g["pred_row"])) > 0.05  # the fixture is not a near-zero network


def test_rotpredict_reverse_mean(golden):
    """SO3Diffusion.p_mean_variance driven by RotPredict (diffusion.py:308-313): oracle denoiser + oracle step
    algebra against the reference's model mean.  Tolerance: the reference's so3_scale error grows with the scale
    (quirk Q5), so t = 999 (scale 20291) is compared loosely;
    at t = 300 the bound is the time-embedding conditioning above times the step's gain."""
    g = golden("rotpredict")
    w, b = rotpredict_params(g)
    s = O.schedule_buffers(1000)
    for k, tv in enumerate(g["t_mean"]):
        pred = O.rotpredict_forward(w, b, g["x"], np.array([tv]))
        want = O.p_sample_mean(g["x"].astype(np.float64), pred, s["sqrt_recip_alphas_cumprod"][tv], s["sqrt_recipm1_alphas_cumprod"][tv],
                               s["posterior_mean_coef1"][tv], s["posterior_mean_coef2"][tv])
        err = O.geodesic_angle(want, g["mean_shared"][k].astype(np.float64))
        assert np.max(err) < (1e-4 if tv <= 300 else 0.2), (tv, np.max(err))
