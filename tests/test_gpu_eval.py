"""GPU parity tests for the evaluation / data-side rows of SURVEY 8(f): the fused all-pairs MMD (util.py:254-285
with rmat_gaussian_kernel / rmat_cosine_kernel) and the fused Bingham sampler (distributions.py:113-127), called
through the Python drop-in (-> C ABI of libso3d.so), against the fp64 oracle and the golden vectors produced by
the reference.  Tolerances: kernel values <= 1e-6 absolute (theta accurate to ~2e-7 rad), MMD <= 1e-6 absolute
(means of kernel values in [0, 1], sums accumulated in fp64), Bingham quaternions <= 1e-6 given the same normals.
"""
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


def dev(a, device):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).to(device)


def rots(n, seed, max_angle=math.pi):
    return O.random_rotations(n, np.random.default_rng(seed), max_angle)[0].astype(np.float32)


def test_mmd_against_reference_golden(dx, cuda_device, golden):
    g = golden("evaluation")
    U = dx.util
    X, Y = dev(g["c3_R"], cuda_device), dev(g["c1_R"], cuda_device)
    assert abs(U.MMD(X, Y, U.rmat_gaussian_kernel).item() - float(g["mmd_gauss"])) < 3e-6
    assert abs(U.MMD(X, Y, U.rmat_gaussian_kernel, chunksize=128).item() - float(g["mmd_gauss_chunk128"])) < 3e-6
    assert abs(U.MMD(X, Y, U.rmat_cosine_kernel).item() - float(g["mmd_cos"])) < 3e-6
    assert abs(U.MMD(X[:150], X[150:], U.rmat_gaussian_kernel).item() - float(g["mmd_gauss_same"])) < 3e-6
    # elementwise (broadcasting) kernels, as the reference's chunked route calls them
    kg = U.rmat_gaussian_kernel(X[:64].unsqueeze(0), Y[:48].unsqueeze(1))
    assert kg.shape == (48, 64) and np.max(np.abs(kg.cpu().numpy() - g["ker_gauss_xy"])) < 2e-6
    kc = U.rmat_cosine_kernel(X[:64].unsqueeze(0), Y[:48].unsqueeze(1))
    assert np.max(np.abs(kc.cpu().numpy() - g["ker_cos_xy"])) < 1e-6
    assert np.max(np.abs(U.rmat_cosine_dist(X[:64], Y[:64]).cpu().numpy() - g["cos_dist"])) < 1e-6
    # two-sample test wrappers (util.py:287-312)
    assert U.Ker_2samp_test(X[:150], X[150:], U.rmat_gaussian_kernel) == bool(g["test_same"])
    assert U.Ker_2samp_test(X, Y[:300], U.rmat_gaussian_kernel) == bool(g["test_diff"])
    assert abs(U.Ker_2samp_log_prob(X[:150], X[150:], U.rmat_gaussian_kernel) - float(g["logp_same"])) < 1e-4


@pytest.mark.parametrize("nx,ny", [(1, 1), (5, 300), (256, 256), (257, 511), (700, 1)])
@pytest.mark.parametrize("kernel", ["gaussian", "cosine"])
def test_pair_sums_against_oracle(dx, cuda_device, nx, ny, kernel):
    """All three sums against the fp64 oracle on ragged / single-tile / multi-tile sizes, including pairs at
    angle 0 (duplicates), near 0 and near pi where an acos-of-trace formulation would lose 3 digits."""
    X, Y = rots(nx, 10 + nx), rots(ny, 20 + ny)
    if nx >= 5:
        X[1] = X[0]                                                        # theta = 0 exactly
        X[2] = (O.rodrigues(np.array([[0.0, 0.0, 1.0]]), np.array([1e-4]))[0] @ X[0]).astype(np.float32)
        X[3] = (O.rodrigues(np.array([[1.0, 0.0, 0.0]]), np.array([math.pi - 1e-4]))[0] @ X[0]).astype(np.float32)
        X[4] = (O.rodrigues(np.array([[0.0, 1.0, 0.0]]), np.array([math.pi]))[0] @ X[0]).astype(np.float32)
    ker = O.rmat_gaussian_kernel if kernel == "gaussian" else O.rmat_cosine_kernel
    want = O.pair_kernel_sums(X, Y, ker)
    got = dx.ops.pair_kernel_sums(dev(X, cuda_device), dev(Y, cuda_device), kernel).cpu().numpy()
    assert got.dtype == np.float64
    scale = np.array([nx * nx, ny * ny, nx * ny], dtype=np.float64)
    assert np.max(np.abs(got - want) / scale) < 1e-6, (got, want)


def test_pair_values_near_zero_and_pi(dx, cuda_device):
    """Single pairs (nx = ny = 1) expose the per-pair value: exp(-sqrt(2) theta) to 1e-6 for theta from 1e-6 to pi."""
    base = rots(1, 5)[0].astype(np.float64)
    for theta in (0.0, 1e-6, 1e-4, 1e-2, 0.5, 1.5707963, 3.0, 3.1414, math.pi):
        B = (base @ O.rodrigues(np.array([[0.6, 0.0, 0.8]]), np.array([theta]))[0]).astype(np.float32)
        got = dx.ops.pair_kernel_sums(dev(base[None], cuda_device), dev(B[None], cuda_device), "gaussian")[2].item()
        assert abs(got - math.exp(-math.sqrt(2.0) * theta)) < 1e-6, theta
        gotc = dx.ops.pair_kernel_sums(dev(base[None], cuda_device), dev(B[None], cuda_device), "cosine")[2].item()
        assert abs(gotc - math.cos(theta)) < 1e-6, theta


def test_pair_sums_shards_add_up_and_are_deterministic(dx, cuda_device):
    X, Y = dev(rots(1500, 1), cuda_device), dev(rots(1100, 2, 1.0), cuda_device)
    full = dx.ops.pair_kernel_sums(X, Y)
    assert torch.equal(full, dx.ops.pair_kernel_sums(X, Y))  # fixed summation order
    for nshards in (2, 3, 8):
        parts = sum(dx.ops.pair_kernel_sums(X, Y, shard=s, nshards=nshards) for s in range(nshards))
        assert torch.allclose(parts, full, rtol=1e-12, atol=0)
    # Y = None: only the X-X sum
    only = dx.ops.pair_kernel_sums(X, None)
    assert only[0].item() == full[0].item() and only[1].item() == 0.0 and only[2].item() == 0.0
    # MMD of a set with itself is 0; symmetric in its arguments
    U = dx.util
    assert abs(U.MMD(X, X, U.rmat_gaussian_kernel).item()) < 1e-7
    assert abs(U.MMD(X, Y, U.rmat_gaussian_kernel).item() - U.MMD(Y, X, U.rmat_gaussian_kernel).item()) < 1e-7


def test_mmd_fused_equals_generic_callable_route(dx, cuda_device):
    """A user-supplied kernel callable takes the reference's (chunked) outer-product route through the elementwise
    ops; it must agree with the fused launch."""
    U = dx.util
    X, Y = dev(rots(700, 3), cuda_device), dev(rots(650, 4, 0.7), cuda_device)
    fused = U.MMD(X, Y, U.rmat_gaussian_kernel).item()
    generic = U.MMD(X, Y, lambda a, b: U.rmat_gaussian_kernel(a, b), chunksize=256).item()
    assert abs(fused - generic) < 2e-6
    assert abs(fused - O.mmd(X.cpu().numpy(), Y.cpu().numpy())) < 1e-6


def test_mmd_full_size_properties(dx, cuda_device):
    """bingham_test.py:29 size (20 000 vs 20 000 = 1.2e9 pair evaluations incl. the two self sums): size-independent
    properties -- MMD(X, X) = 0, symmetry, shard additivity, and separation of two different Bingham laws."""
    D = dx.distributions
    loc = torch.zeros(4, device=cuda_device)
    cov3 = torch.tensor([[1.0, 0, 0, 0], [0, 1.0, 0.9, 0.9], [0, 0.9, 1.0, 0.9], [0, 0.9, 0.9, 1.0]], device=cuda_device)
    dx.manual_seed(7)
    A = D.Bingham(loc, covariance_matrix=cov3).sample_rmat((20_000,))
    B = D.Bingham(loc, covariance_matrix=cov3).sample_rmat((20_000,))
    C = D.Bingham(loc, covariance_matrix=torch.eye(4, device=cuda_device)).sample_rmat((20_000,))
    U = dx.util
    same, diff = U.MMD(A, B, U.rmat_gaussian_kernel).item(), U.MMD(A, C, U.rmat_gaussian_kernel).item()
    assert abs(U.MMD(A, A, U.rmat_gaussian_kernel).item()) < 1e-7
    assert 0.0 <= same < 1e-3 < 0.01 < diff
    assert U.Ker_2samp_test(A, B, U.rmat_gaussian_kernel) and not U.Ker_2samp_test(A, C, U.rmat_gaussian_kernel)
    full = dx.ops.pair_kernel_sums(A, C)
    parts = sum(dx.ops.pair_kernel_sums(A, C, shard=s, nshards=4) for s in range(4))
    assert torch.allclose(parts, full, rtol=1e-12, atol=0)


# ---------------------------------------------------------------------------------------------
# Bingham (distributions.py:113-127)
# ---------------------------------------------------------------------------------------------
def test_bingham_given_normals_against_reference_golden(dx, cuda_device, golden):
    g = golden("evaluation")
    D = dx.distributions
    for name in ("c3", "c1"):
        d = D.Bingham(torch.zeros(4, device=cuda_device), covariance_matrix=dev(g[name + "_cov"], cuda_device))
        assert np.max(np.abs(d.scale_tril.cpu().numpy() - g[name + "_tril"])) < 1e-5 * np.abs(g[name + "_tril"]).max()
        z = dev(g[name + "_z"], cuda_device)
        q = d.rsample((z.shape[0],), z=z)
        assert q.shape == (z.shape[0], 4)
        assert np.max(np.abs(q.cpu().numpy() - g[name + "_q"])) < 2e-6
        R = d.sample_rmat((z.shape[0],), z=z)
        assert np.max(np.abs(R.cpu().numpy() - g[name + "_R"])) < 3e-6
        # unaligned view of the normals takes the scalar-load path
        zz = torch.empty(z.numel() + 1, device=cuda_device)[1:].view_as(z).copy_(z)
        assert torch.equal(d.rsample((z.shape[0],), z=zz), q)


def test_bingham_device_draws(dx, cuda_device):
    """Device Philox normals: unit quaternions whose second-moment matrix matches a numpy Monte-Carlo estimate of
    the same law, reproducible per seed, independent of sharding (row_offset), consistent between the quaternion
    and the fused rotation-matrix outputs."""
    D = dx.distributions
    cov = np.array([[1.0, 0, 0, 0], [0, 1.0, 0.9, 0.9], [0, 0.9, 1.0, 0.9], [0, 0.9, 0.9, 1.0]])
    d = D.Bingham(torch.zeros(4, device=cuda_device), covariance_matrix=dev(cov, cuda_device))
    n = 1 << 20
    dx.manual_seed(11)
    q = d.sample((n,))
    dx.manual_seed(11)
    q2 = d.sample((n,))
    assert torch.equal(q, q2) and torch.isfinite(q).all()
    assert float((q.norm(dim=-1) - 1).abs().max()) < 1e-6
    ref = O.bingham_sample_given(np.random.default_rng(0).standard_normal((n, 4)), np.linalg.cholesky(cov))
    m_got = (q.double().T @ q.double() / n).cpu().numpy()
    assert np.max(np.abs(m_got - ref.T @ ref / n)) < 5e-3
    # the underlying normals are standard: with an identity covariance q is uniform on S^3 -> E[q q^T] = I/4
    e = D.Bingham(torch.zeros(4, device=cuda_device), covariance_matrix=torch.eye(4, device=cuda_device)).sample((n,))
    assert np.max(np.abs((e.double().T @ e.double() / n).cpu().numpy() - np.eye(4) / 4)) < 2e-3
    # sharding invariance and quaternion/matrix consistency through the explicit-state op
    st = d.scale_tril.contiguous()
    full_q, full_R = dx.ops.bingham_sample(st, (5000,), seed=3, rng_offset=1, want_quat=True, want_rmat=True)
    part = dx.ops.bingham_sample(st, (5000 - 1234,), seed=3, rng_offset=1, row_offset=1234)
    assert torch.equal(part, full_q[1234:])
    assert float((dx.util.quat_to_rmat(full_q) - full_R).abs().max()) < 1e-6
    # empty and batched sample shapes
    assert d.sample((0,)).shape == (0, 4) and d.sample((3, 5)).shape == (3, 5, 4) and d.sample_rmat((2, 2)).shape == (2, 2, 3, 3)
