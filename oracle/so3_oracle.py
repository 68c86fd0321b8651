"""CPU oracle for the SO(3) manifold-diffusion hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a numpy (float64 unless stated) restatement of the arithmetic the reference
(qazwsxal/diffusion-extensions @ f100885d) performs on its hot path.  It is the checker the
parity tests compare the sm_100a kernels against.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline leg may import it; the product package never does.

Pinning: the reference ships no tests or golden vectors, so this oracle is
pinned against outputs of the *reference itself*, imported unmodified in the build container by
``oracle/gen_golden.py`` and committed under ``tests/golden/`` (see ``tests/test_oracle_golden.py``).

Every function cites the reference ``file:line`` it follows.  Where the reference is numerically
broken the default here is the mathematically intended
result and ``reference_quirks=True`` reproduces the reference's behaviour.
"""
from __future__ import annotations

import math

import numpy as np

PI = math.pi
N_GRID = 1000  # distributions.py:15


# ----------------------------------------------------------------------------------------------
# L0: hat / vee, log / exp, axis-angle, scale, lerp, quaternion   (util.py)
# ----------------------------------------------------------------------------------------------
def vec2skew(v):
    """hat map, util.py:87-92:  (x,y,z) -> [[0,-z,y],[z,0,-x],[-y,x,0]]."""
    v = np.asarray(v)
    out = np.zeros(v.shape[:-1] + (3, 3), dtype=v.dtype)
    x, y, z = v[..., 0], v[..., 1], v[..., 2]
    out[..., 0, 1] = -z
    out[..., 0, 2] = y
    out[..., 1, 0] = z
    out[..., 1, 2] = -x
    out[..., 2, 0] = -y
    out[..., 2, 1] = x
    return out


def skew2vec(m):
    """vee map, util.py:79-84:  (m[2,1], -m[2,0], m[1,0])."""
    m = np.asarray(m)
    return np.stack((m[..., 2, 1], -m[..., 2, 0], m[..., 1, 0]), axis=-1)


def _sin_cos_angle(r):
    """s = |vee(R-R^T)|/2, c = (tr R - 1)/2, theta = atan2(s, c)   (util.py:165-169)."""
    r = np.asarray(r, dtype=np.float64)
    a = r - np.swapaxes(r, -1, -2)
    v = skew2vec(a)
    s = np.linalg.norm(v, axis=-1) / 2
    c = (np.trace(r, axis1=-2, axis2=-1) - 1) / 2
    return a, v, s, c, np.arctan2(s, c)


def _axis_from_symmetric(r, c, v):
    """Axis from the symmetric part (R+R^T)/2 = c I + (1-c) n n^T  (SURVEY A.4 robust branch).

    Used where the reference's skew-part formula loses the axis (theta -> pi).  Sign is chosen so
    that n . vee(R - R^T) >= 0, i.e. consistent with the generic branch.
    """
    with np.errstate(all="ignore"):
        return _axis_from_symmetric_impl(r, c, v)


def _axis_from_symmetric_impl(r, c, v):
    sym = (r + np.swapaxes(r, -1, -2)) / 2
    d = np.stack([sym[..., i, i] for i in range(3)], axis=-1)
    k = np.argmax(d, axis=-1)
    one_m_c = np.maximum(1 - c, 1e-300)
    nk = np.sqrt(np.maximum((np.take_along_axis(d, k[..., None], -1)[..., 0] - c) / one_m_c, 0.0))
    row = np.take_along_axis(sym, k[..., None, None].repeat(3, -1), -2)[..., 0, :]
    n = row / (one_m_c * np.maximum(nk, 1e-300))[..., None]
    np.put_along_axis(n, k[..., None], nk[..., None], -1)
    n = n / np.maximum(np.linalg.norm(n, axis=-1, keepdims=True), 1e-300)
    sgn = np.where((n * v).sum(-1) < 0, -1.0, 1.0)
    return n * sgn[..., None]


NEAR_PI_COS = -0.9  # c below this: take the axis from the symmetric part


def log_vec(r, reference_quirks=False):
    """vee(log R) as a 3-vector.  util.py:164-192 (generic branch: theta/(2 s) * vee(R-R^T)).

    reference_quirks=False: exact-pi / near-pi rows use the symmetric-part axis (fixes Q4);
    True: generic formula everywhere (the reference's eigh fallback only fires on NaN rows, which
    for inputs away from exactly pi never happens).
    """
    a, v, s, c, th = _sin_cos_angle(r)
    with np.errstate(divide="ignore", invalid="ignore"):
        scale = np.where(s > 0, th / (2 * s), 0.5)
    scale = np.where(th == 0.0, 0.0, scale)  # util.py:174
    out = scale[..., None] * v
    if not reference_quirks:
        near = c < NEAR_PI_COS
        if np.any(near):
            n = _axis_from_symmetric(np.asarray(r, dtype=np.float64), c, v)
            out = np.where(near[..., None], th[..., None] * n, out)
    return out


def log_rmat(r, reference_quirks=False):
    """util.py:164-192, returned as the 3x3 skew matrix."""
    return vec2skew(log_vec(r, reference_quirks))


DEFAULT_AXIS = np.array([0.0, 0.0, 1.0])


def rmat_to_aa(r, reference_quirks=False):
    """util.py:208-219.  angle keeps a trailing dim of 1.  Identity rows: the reference returns a
    NaN axis (Q10); the default here is the fixed axis (0,0,1) with angle 0."""
    v = log_vec(r, reference_quirks)
    ang = np.linalg.norm(v, axis=-1, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        axis = v / ang
    if not reference_quirks:
        axis = np.where(ang > 0, axis, DEFAULT_AXIS)
    return axis, ang


def rodrigues(axis_unit, ang):
    """R = I + sin(a) K + (1-cos(a)) K^2, K = hat(axis)  (replaces util.py:204-205)."""
    axis_unit = np.asarray(axis_unit, dtype=np.float64)
    ang = np.asarray(ang, dtype=np.float64)
    k = vec2skew(axis_unit)
    k2 = k @ k
    eye = np.eye(3)
    return eye + np.sin(ang)[..., None, None] * k + (1 - np.cos(ang))[..., None, None] * k2


def aa_to_rmat(axis, ang):
    """util.py:195-205: normalise the axis, exp(hat(axis) * ang) (matrix_exp + SVD projection in
    the reference == Rodrigues to 4e-7).  `ang` has a trailing dim of 1 like the reference."""
    axis = np.asarray(axis, dtype=np.float64)
    ang = np.asarray(ang, dtype=np.float64)
    n = axis / np.linalg.norm(axis, axis=-1, keepdims=True)
    return rodrigues(n, ang[..., 0])


def exp_vec(v):
    """exp(hat(v)) for a rotation vector v (diffusion.py:294 matrix_exp(vec2skew(.)))."""
    v = np.asarray(v, dtype=np.float64)
    th = np.linalg.norm(v, axis=-1)
    n = np.where(th[..., None] > 0, v / np.maximum(th, 1e-300)[..., None], DEFAULT_AXIS)
    return rodrigues(n, th)


def so3_scale(r, scalars, reference_quirks=False):
    """util.py:349-361: exp(s * log R) == rotation by s*theta about the axis of R."""
    v = log_vec(r, reference_quirks)
    s = np.asarray(scalars, dtype=np.float64)
    return exp_vec(v * s[..., None])


def so3_lerp(a, b, w):
    """util.py:325-338: A @ aa_to_rmat(axis(A^T B), w * angle(A^T B)).  w has a trailing dim 1."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    c = np.swapaxes(a, -1, -2) @ b
    axis, ang = rmat_to_aa(c)
    return a @ rodrigues(axis, (np.asarray(w, dtype=np.float64) * ang)[..., 0])


def rmat_dist(a, b):
    """util.py:315-322: Frobenius norm of log(A^T B) = sqrt(2) * theta."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    v = log_vec(np.swapaxes(a, -1, -2) @ b)
    return math.sqrt(2.0) * np.linalg.norm(v, axis=-1)


def rmat_gaussian_kernel(a, b):
    """util.py:128-134: exp(-rmat_dist(a, b)), broadcasting over batch dims."""
    return np.exp(-rmat_dist(a, b))


def rmat_cosine_kernel(a, b):
    """util.py:136-150: (tr(b^T a) - 1) / 2."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return ((a * b).sum(axis=(-1, -2)) - 1.0) / 2.0


def pair_kernel_sums(x, y, kernel=rmat_gaussian_kernel, chunk=512):
    """The three outer-product sums of util.py:254-285: sum k(X,X), sum k(Y,Y), sum k(X,Y) (float64), evaluated in
    chunk x chunk blocks like the reference's chunked branch (util.py:265-278)."""
    def total(p, q):
        acc = 0.0
        for i in range(0, len(p), chunk):
            for j in range(0, len(q), chunk):
                acc += float(kernel(p[None, i:i + chunk], q[j:j + chunk, None]).sum())
        return acc

    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    return np.array([total(x, x), total(y, y), total(x, y)])


def mmd(x, y, kernel=rmat_gaussian_kernel, chunk=512):
    """util.py:254-285 MMD (biased V-statistic): mean k(X,X) + mean k(Y,Y) - 2 mean k(X,Y)."""
    sxx, syy, sxy = pair_kernel_sums(x, y, kernel, chunk)
    lx, ly = len(x), len(y)
    return sxx / lx ** 2 + syy / ly ** 2 - 2.0 * sxy / (lx * ly)


def bingham_sample_given(z, scale_tril):
    """distributions.py:113-127 Bingham.rsample with the standard-normal draws z (...,4) given:
    vals = L z (MultivariateNormal.rsample with loc = 0), out = vals / |vals|  (unit quaternions, real first)."""
    z = np.asarray(z, dtype=np.float64)
    vals = z @ np.asarray(scale_tril, dtype=np.float64).T
    return vals / np.linalg.norm(vals, axis=-1, keepdims=True)


def quat_to_rmat(q):
    """util.py:222-252: real-first (r,i,j,k), un-normalised input allowed (two_s = 2/|q|^2)."""
    q = np.asarray(q, dtype=np.float64)
    r, i, j, k = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    two_s = 2.0 / (q * q).sum(-1)
    o = np.stack(
        (
            1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
            two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
            two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j),
        ),
        axis=-1,
    )
    return o.reshape(q.shape[:-1] + (3, 3))


def rmat_to_quat_canonical(r):
    """No reference counterpart.  Unit quaternion (real-first, r >= 0) from
    axis-angle: q = (cos(theta/2), sin(theta/2) n).  Round-trip oracle only."""
    axis, ang = rmat_to_aa(r)
    h = ang / 2
    return np.concatenate((np.cos(h), np.sin(h) * axis), axis=-1)


# ----------------------------------------------------------------------------------------------
# L1: IGSO(3) density, score, CDF table, sampler   (distributions.py)
# ----------------------------------------------------------------------------------------------
def igso3_series(omega, eps, L=2000):
    """fp64 truncated series (the 'fp64 reference series' of the north star):
        f  = sum_{l<L} (2l+1) exp(-l(l+1) eps^2) chi_l(omega),
        chi_l(w) = sin((l+1/2) w)/sin(w/2) = 1 + 2 sum_{m<=l} cos(m w)   (chi_l(0) = 2l+1)
    Returns (f, dlogf/domega).  Evaluated in the character (chi) form, which has no 0/0 at
    omega = 0 and no cancellation against cot(omega/2)."""
    om = np.atleast_1d(np.asarray(omega, dtype=np.float64))
    ep = np.atleast_1d(np.asarray(eps, dtype=np.float64))
    om, ep = np.broadcast_arrays(om, ep)
    f = np.zeros(om.shape)
    df = np.zeros(om.shape)
    chi = np.ones(om.shape)
    dchi = np.zeros(om.shape)
    v = ep * ep
    f += chi  # l = 0 term: weight 1
    for l in range(1, L):
        chi = chi + 2 * np.cos(l * om)
        dchi = dchi - 2 * l * np.sin(l * om)
        w = (2 * l + 1) * np.exp(-l * (l + 1) * v)
        f += w * chi
        df += w * dchi
        if np.all(l * (l + 1) * v > 745.0):  # every further weight is exactly 0.0 in fp64
            break
    return f, df / f


def igso3_series_sincot(omega, eps, L=2000):
    """The series exactly as SURVEY A.1 writes it (sin/sin and C/S - cot/2 forms); used only to
    cross-check igso3_series away from omega = 0."""
    om = np.asarray(omega, dtype=np.float64)
    ep = np.asarray(eps, dtype=np.float64)
    l = np.arange(L, dtype=np.float64).reshape((L,) + (1,) * np.ndim(om))
    a = (2 * l + 1) * np.exp(-l * (l + 1) * ep * ep)
    s = (a * np.sin((l + 0.5) * om)).sum(0)
    c = (a * (l + 0.5) * np.cos((l + 0.5) * om)).sum(0)
    return s / np.sin(om / 2), c / s - 0.5 / np.tan(om / 2)


def igso3_closed(omega, eps, reference_quirks=False):
    """Closed (Poisson-dual, 3-image) density, distributions.py:53-72, fp64.

    reference_quirks=True follows distributions.py:56-71 literally: the overflowing
    exp(-pi^2/v) * exp(+-pi t/v) products, inf/NaN -> 0, and the t == 0 limit expression (which is
    itself NaN for eps < 0.167, Q3).  Default: algebraically identical stable form
    exp(-pi (pi -+ t)/v) and the exact t -> 0 limit of the same 3-image expression.
    Returns float64 (the reference casts to float32 on return, :72)."""
    t = np.asarray(omega, dtype=np.float64)
    v = np.asarray(eps, dtype=np.float64) ** 2
    t, v = np.broadcast_arrays(t, v)
    with np.errstate(all="ignore"):
        pref = math.sqrt(PI) * v ** (-1.5) * np.exp(v / 4) * np.exp(-((t / 2) ** 2) / v)
        if reference_quirks:
            inner = t - np.exp(-(PI ** 2) / v) * ((t - 2 * PI) * np.exp(PI * t / v) + (t + 2 * PI) * np.exp(-PI * t / v))
        else:
            inner = t - (t - 2 * PI) * np.exp(-PI * (PI - t) / v) - (t + 2 * PI) * np.exp(-PI * (PI + t) / v)
        vals = pref * inner / (2 * np.sin(t / 2))
        vals = np.where(np.isinf(vals), 0.0, vals)
        vals = np.where(np.isnan(vals), 0.0, vals)
        if reference_quirks:
            lim = math.sqrt(PI) * (v * np.exp(2 * PI ** 2 / v) - 2 * v * np.exp(PI ** 2 / v)
                                   + 4 * PI ** 2 * v * np.exp(PI ** 2 / v)) * np.exp(v / 4 - (2 * PI ** 2) / v) / v ** 2.5
        else:
            e1 = np.exp(-(PI ** 2) / v)
            lim = math.sqrt(PI) * v ** (-1.5) * np.exp(v / 4) * (1 - 2 * e1 + 4 * PI ** 2 * e1 / v)
        vals = np.where(t == 0, lim, vals)
    return vals


def igso3_closed_dlog(omega, eps):
    """d/d omega of log igso3_closed (stable form), fp64.  No reference counterpart (D2): the
    reference only obtains it through autograd of log_prob (distributions.py:186-190)."""
    t = np.asarray(omega, dtype=np.float64)
    v = np.asarray(eps, dtype=np.float64) ** 2
    t, v = np.broadcast_arrays(t, v)
    e1 = np.exp(-PI * (PI - t) / v)
    e2 = np.exp(-PI * (PI + t) / v)
    b = t - (t - 2 * PI) * e1 - (t + 2 * PI) * e2
    db = 1 - e1 * (1 + (PI / v) * (t - 2 * PI)) - e2 * (1 - (PI / v) * (t + 2 * PI))
    with np.errstate(all="ignore"):
        g = -t / (2 * v) + db / b - 0.5 / np.tan(t / 2)
    return np.where(t == 0, 0.0, g)


def grid_f32():
    """The 1000-point cubic grid of distributions.py:15 in the reference's own float32 arithmetic
    (torch.linspace(0,1,1000)**3 * pi).  Returns (locs[1000] f32, haar_w[1000] f32) with
    haar_w = (1 - cos(loc))/pi as evaluated in float32 at :21."""
    # torch.linspace(0, 1, 1000) in fp32: start + i*step for the lower half, end - (n-1-i)*step
    # for the upper half (ATen RangeFactories).  Restated here; checked against the golden file.
    n = N_GRID
    step = np.float32(1.0) / np.float32(n - 1)
    i = np.arange(n)
    lo = (np.float32(0.0) + step * i.astype(np.float32)).astype(np.float32)
    hi = (np.float32(1.0) - step * (n - 1 - i).astype(np.float32)).astype(np.float32)
    lin = np.where(i < n // 2, lo, hi).astype(np.float32)
    locs = (np.float32(PI) * (lin * lin * lin).astype(np.float32)).astype(np.float32)
    haar = ((np.float32(1.0) - np.cos(locs, dtype=np.float32)) / np.float32(PI)).astype(np.float32)
    return locs, haar


def igso3_cdf_table(eps, locs=None, haar=None, reference_quirks=False):
    """distributions.py:15-30, following the reference's dtype at every step.

    eps: (E,) float.  Returns (trap[E, 999] f32, trap_loc[999] f32).  NB the reference's `trap` is
    laid out (999, *E); this is its transpose (one contiguous CDF row per eps).
      :19-21  pdf = float32(f_eps(loc) in fp64) * haar_w      (fp32 multiply)
      :23     pdf[loc == 0] = 0
      :26-28  cumsum of (dloc * (pdf[k] + pdf[k+1]) / 2) in fp32 values; ATen's CPU cumsum
              accumulates float in double and rounds each prefix to float
      :29     divide by the last entry."""
    if locs is None or haar is None:
        locs, haar = grid_f32()
    eps = np.atleast_1d(np.asarray(eps, dtype=np.float32))
    dens = igso3_closed(locs[None, :].astype(np.float64), eps[:, None].astype(np.float64), reference_quirks)
    pdf = dens.astype(np.float32) * haar[None, :]
    pdf[:, locs == 0] = 0.0
    sums = pdf[:, :-1] + pdf[:, 1:]
    dloc = np.diff(locs).astype(np.float32)
    inc = (dloc[None, :] * sums / np.float32(2.0)).astype(np.float32)
    trap = np.cumsum(inc.astype(np.float64), axis=1).astype(np.float32)
    trap = (trap / trap[:, -1:]).astype(np.float32)
    return trap, locs[1:].copy()


def igso3_angle_from_uniform(u, trap_row, trap_loc, trap_row_for_weight=None):
    """distributions.py:38-49 in float32: inverse-CDF lookup + lerp.

    u: (n,) f32 in [0,1);  trap_row: (n, 999) or (999,) CDF values for each sample.
    trap_row_for_weight: the reference's Q1 bug gathers the lerp endpoints from column 0 of a
    batched table; pass that row here to reproduce it (reference_quirks)."""
    u = np.asarray(u, dtype=np.float32)
    tr = np.asarray(trap_row, dtype=np.float32)
    if tr.ndim == 1:
        tr = np.broadcast_to(tr, u.shape + tr.shape)
    idx1 = (tr <= u[..., None]).sum(-1)
    idx0 = np.maximum(idx1 - 1, 0)
    wt = tr if trap_row_for_weight is None else np.broadcast_to(np.asarray(trap_row_for_weight, np.float32), tr.shape)
    # idx1 == 999 cannot happen for u < 1 since trap[-1] == 1 exactly; clip for safety like gather would fail
    i1 = np.minimum(idx1, tr.shape[-1] - 1)
    t0 = np.take_along_axis(wt, idx0[..., None], -1)[..., 0]
    t1 = np.take_along_axis(wt, i1[..., None], -1)[..., 0]
    diff = np.maximum(t1 - t0, np.float32(1e-6))
    w = np.clip((u - t0) / diff, np.float32(0), np.float32(1)).astype(np.float32)
    a0 = np.asarray(trap_loc, np.float32)[idx0]
    a1 = np.asarray(trap_loc, np.float32)[i1]
    # torch.lerp(start, end, w): w < 0.5 ? start + w*(end-start) : end - (end-start)*(1-w)
    d = (a1 - a0).astype(np.float32)
    ang = np.where(w < np.float32(0.5), a0 + w * d, a1 - d * (np.float32(1) - w)).astype(np.float32)
    return ang


def igso3_sample_given(u, axes, trap_row, trap_loc, mean=None):
    """distributions.py:33-51 with the random draws supplied: R = mean @ Rodrigues(axes/|axes|, angle)."""
    ang = igso3_angle_from_uniform(u, trap_row, trap_loc).astype(np.float64)
    r = aa_to_rmat(np.asarray(axes, np.float64), ang[..., None])
    if mean is not None:
        r = np.asarray(mean, np.float64) @ r
    return r, ang


def igso3_log_prob(r, eps, kind="closed", L=2000):
    """distributions.py:74-77: log f_eps(angle(R)), shape (..., 1).  (No Haar factor.)"""
    _, ang = rmat_to_aa(r)
    f = igso3_closed(ang, eps) if kind == "closed" else igso3_series(ang, eps, L)[0].reshape(ang.shape)
    return np.log(f)


def igso3_score(r, eps, kind="series", L=2000):
    """score(R; eps) = dlogf/domega * axis  (SURVEY D2).  Returns (logp (...,), score (...,3))."""
    axis, ang = rmat_to_aa(r)
    a = ang[..., 0]
    if kind == "series":
        f, g = igso3_series(a, eps, L)
        f = f.reshape(a.shape)
        g = g.reshape(a.shape)
    else:
        f, g = igso3_closed(a, eps), igso3_closed_dlog(a, eps)
    return np.log(f), g[..., None] * axis


def log_prob_ambient_grad(r, g):
    """what autograd of log_prob w.r.t. the 9 matrix entries yields through
    util.py:165-176:  dlogf/dR = g * [ c/(4 s) (R - R^T) - (s/2) I ] / (s^2 + c^2)."""
    a, v, s, c, th = _sin_cos_angle(r)
    eye = np.eye(3)
    den = s * s + c * c
    return np.asarray(g)[..., None, None] * ((c / (4 * s))[..., None, None] * a - (s / 2)[..., None, None] * eye) / den[..., None, None]


# ----------------------------------------------------------------------------------------------
# L2: schedule + SO3Diffusion step algebra   (diffusion.py, submodule helpers)
# ----------------------------------------------------------------------------------------------
def cosine_beta_schedule(timesteps, s=0.008):
    """denoising_diffusion_pytorch.py:278-288."""
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    ac = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return np.clip(betas, 0, 0.999)


def schedule_buffers(timesteps=1000, betas=None):
    """diffusion.py:57-92: the 12 float32 schedule buffers, keyed by the reference's names."""
    betas = cosine_beta_schedule(timesteps) if betas is None else np.asarray(betas, dtype=np.float64)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    pv = betas * (1.0 - ac_prev) / (1.0 - ac)
    buf = {
        "betas": betas,
        "alphas_cumprod": ac,
        "alphas_cumprod_prev": ac_prev,
        "sqrt_alphas_cumprod": np.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": np.sqrt(1.0 - ac),
        "log_one_minus_alphas_cumprod": np.log(1.0 - ac),
        "sqrt_recip_alphas_cumprod": np.sqrt(1.0 / ac),
        "sqrt_recipm1_alphas_cumprod": np.sqrt(1.0 / ac - 1),
        "posterior_variance": pv,
        "posterior_log_variance_clipped": np.log(np.maximum(pv, 1e-20)),
        "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (1.0 - ac),
        "posterior_mean_coef2": (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
    }
    return {k: v.astype(np.float32) for k, v in buf.items()}


def q_sample(x0, scale, noise):
    """diffusion.py:339-346: so3_scale(x0, sqrt_alphas_cumprod[t]) @ noise."""
    return so3_scale(x0, scale) @ np.asarray(noise, np.float64)


def skewvec_target(noise, eps):
    """diffusion.py:355: vee(log noise) / eps."""
    return log_vec(noise) / np.asarray(eps, np.float64)[..., None]


def predict_start_from_noise(x_t, pred, sqrt_recip_ac, sqrt_recipm1_ac):
    """diffusion.py:291-297: so3_scale(x_t, 1/sqrt(abar)) @ exp(hat(pred * sqrt(1/abar - 1)))^T."""
    xt_term = so3_scale(x_t, sqrt_recip_ac)
    noise_term = exp_vec(np.asarray(pred, np.float64) * np.asarray(sqrt_recipm1_ac, np.float64)[..., None])
    return xt_term @ np.swapaxes(noise_term, -1, -2)


def q_posterior_mean(x_start, x_t, coef1, coef2):
    """diffusion.py:299-302: so3_scale(x0_hat, c1) @ so3_scale(x_t, c2)."""
    return so3_scale(x_start, coef1) @ so3_scale(x_t, coef2)


def p_sample_mean(x_t, pred, sqrt_recip_ac, sqrt_recipm1_ac, coef1, coef2):
    """diffusion.py:308-313: model mean of the reverse step."""
    x_recon = predict_start_from_noise(x_t, pred, sqrt_recip_ac, sqrt_recipm1_ac)
    return q_posterior_mean(x_recon, x_t, coef1, coef2)


def p_sample(x_t, pred, sqrt_recip_ac, sqrt_recipm1_ac, coef1, coef2, noise=None):
    """diffusion.py:315-326: mean @ noise (noise = None at t == 0)."""
    m = p_sample_mean(x_t, pred, sqrt_recip_ac, sqrt_recipm1_ac, coef1, coef2)
    return m if noise is None else m @ np.asarray(noise, np.float64)


# ----------------------------------------------------------------------------------------------
# RotPredict denoiser (so3_train.py:11-49, models.py:13-25), SURVEY 8f-4
# ----------------------------------------------------------------------------------------------
def sinusoidal_pos_emb(t, dim):
    """models.py:13-25: cat(sin(t f_j), cos(t f_j)), f_j = exp(-j log(1e4)/(dim/2 - 1)).  The reference builds
    f_j and the products in float32 (torch.arange -> exp), which is reproduced here before the float64 sin/cos:
    at t = 999 a float64 frequency would move the phase by 3e-5 rad."""
    half = dim // 2
    scale = np.float32(math.log(10000) / (half - 1))
    freq = np.exp(np.arange(half, dtype=np.float32) * -scale).astype(np.float32)
    arg = (np.asarray(t, np.float32)[:, None] * freq[None, :]).astype(np.float64)
    return np.concatenate([np.sin(arg), np.cos(arg)], axis=-1)


def rotpredict_forward(weights, biases, x, t):
    """so3_train.py:39-49 with out_type 'skewvec': net(cat(flatten(x), time_embedding(t))), five Linear layers
    (weights[i]: out x in) with SiLU between them (so3_train.py:26-36).  t: (n,) or (1,) (expanded, :42-43)."""
    x = np.asarray(x, np.float64)
    n = x.shape[0]
    d_model = np.asarray(weights[0]).shape[1]
    emb = sinusoidal_pos_emb(np.asarray(t).reshape(-1), d_model - 9)
    if emb.shape[0] == 1:
        emb = np.broadcast_to(emb, (n, emb.shape[1]))
    h = np.concatenate([x.reshape(n, 9), emb], axis=-1)
    for i, (w, b) in enumerate(zip(weights, biases)):
        h = h @ np.asarray(w, np.float64).T + np.asarray(b, np.float64)
        if i + 1 < len(weights):
            h = h / (1.0 + np.exp(-h))  # SiLU
    return h


# ----------------------------------------------------------------------------------------------
# helpers for tests / benchmarks
# ----------------------------------------------------------------------------------------------
# ----------------------------------------------------------------------------------------------
# SE(3) arm (diffusion.py:432-523, util.py:382-385, distributions.py:84-110): (rotation, translation) pairs
# ----------------------------------------------------------------------------------------------
def se3_scale(rot, shift, scalars):
    """util.py:382-385: so3_scale on the rotation, shift * scalars[..., None]."""
    scalars = np.asarray(scalars, dtype=np.float64)
    return so3_scale(rot, scalars), np.asarray(shift, dtype=np.float64) * scalars[..., None]


def se3_q_sample(rot0, shift0, scale, noise_rot, noise_shift):
    """diffusion.py:498-506: (so3_scale(rot0, a) @ noise_rot, a shift0 + noise_shift)."""
    r, s = se3_scale(rot0, shift0, scale)
    return r @ np.asarray(noise_rot, dtype=np.float64), s + np.asarray(noise_shift, dtype=np.float64)


def se3_targets(noise_rot, noise_shift, eps, shift_scale):
    """diffusion.py:514-515: (vee(log noise_rot)/eps, noise_shift/(eps shift_scale))."""
    eps = np.asarray(eps, dtype=np.float64)
    return skewvec_target(noise_rot, eps), np.asarray(noise_shift, dtype=np.float64) / (eps * shift_scale)[..., None]


def se3_predict_start(rot_t, shift_t, pred_rot, pred_shift, sqrt_recip_ac, sqrt_recipm1_ac):
    """diffusion.py:444-455."""
    rot = predict_start_from_noise(rot_t, pred_rot, sqrt_recip_ac, sqrt_recipm1_ac)
    a = np.asarray(sqrt_recip_ac, dtype=np.float64)[..., None]
    b = np.asarray(sqrt_recipm1_ac, dtype=np.float64)[..., None]
    return rot, np.asarray(shift_t, dtype=np.float64) * a - np.asarray(pred_shift, dtype=np.float64) * b


def se3_posterior_mean(rot0, shift0, rot_t, shift_t, coef1, coef2):
    """diffusion.py:457-460."""
    c1 = np.asarray(coef1, dtype=np.float64)[..., None]
    c2 = np.asarray(coef2, dtype=np.float64)[..., None]
    return q_posterior_mean(rot0, rot_t, coef1, coef2), c1 * np.asarray(shift0, dtype=np.float64) + c2 * np.asarray(shift_t, dtype=np.float64)


def se3_p_mean(rot_t, shift_t, pred_rot, pred_shift, sqrt_recip_ac, sqrt_recipm1_ac, coef1, coef2):
    """diffusion.py:466-471: posterior mean given the denoiser output."""
    r0, s0 = se3_predict_start(rot_t, shift_t, pred_rot, pred_shift, sqrt_recip_ac, sqrt_recipm1_ac)
    return se3_posterior_mean(r0, s0, rot_t, shift_t, coef1, coef2)


def random_rotations(n, rng, max_angle=PI):
    axis = rng.standard_normal((n, 3))
    axis /= np.linalg.norm(axis, axis=-1, keepdims=True)
    ang = rng.uniform(0, max_angle, n)
    return rodrigues(axis, ang), axis, ang


def geodesic_angle(a, b):
    """Angle of A^T B in fp64 (robust), for error reporting."""
    m = np.swapaxes(np.asarray(a, np.float64), -1, -2) @ np.asarray(b, np.float64)
    _, _, s, c, th = _sin_cos_angle(m)
    return th
