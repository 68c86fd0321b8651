"""Host-side helpers of the util drop-in that the reference's scripts import (util.py:62-76, 340-346, 426-481).
Plain-torch helpers are checked on CPU against the unmodified formulas; so3_bezier (GPU) against the fp64 oracle."""
import math

import numpy as np
import pytest
import torch

from oracle import so3_oracle as O


@pytest.fixture(scope="module")
def util():
    from diffusion_extensions_b200 import util as U

    return U


def test_six_d_representation(util):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(64, 6, generator=g)
    R = util.six2rmat(x)
    assert R.shape == (64, 3, 3)
    eye = torch.eye(3).expand(64, 3, 3)
    assert torch.allclose(R @ R.transpose(-1, -2), eye, atol=1e-5) and torch.allclose(torch.linalg.det(R), torch.ones(64), atol=1e-5)
    assert torch.allclose(R[:, 0], x[:, :3] / x[:, :3].norm(dim=-1, keepdim=True), atol=1e-6)   # util.py:70: rows, not columns
    six = util.rmat2six(R)
    assert six.shape == (64, 6) and torch.equal(six, torch.cat([R[:, 0], R[:, 1]], dim=-1))
    assert torch.allclose(util.six2rmat(six), R, atol=1e-5)                                      # round trip on rotations


def test_masked_mean_cycle_identity_init_from_dict(util):
    t = torch.arange(12, dtype=torch.float32).reshape(2, 3, 2)
    mask = torch.tensor([[True, False, True], [False, False, False]])
    m = util.masked_mean(t.clone(), mask, dim=1)
    assert torch.equal(m, torch.tensor([[2.0, 3.0], [0.0, 0.0]]))
    it = util.cycle([1, 2, 3])
    assert [next(it) for _ in range(7)] == [1, 2, 3, 1, 2, 3, 1]
    assert util.identity(t) is t

    class A:
        def __init__(self, lr, width=3):
            self.lr, self.width = lr, width

    class B:
        def __init__(self, width, depth=2):
            self.width, self.depth = width, depth

    a, b = util.init_from_dict({"lr": 0.1, "width": 7, "unused": None}, A, B)
    assert (a.lr, a.width, b.width, b.depth) == (0.1, 7, 7, 2)


def test_to_device_keeps_structure(util):
    x, y = torch.zeros(2), torch.ones(3)
    tr = util.AffineT(torch.eye(3)[None], torch.zeros(1, 3))
    out = util.to_device(torch.device("cpu"), x, (y, [x]), tr)
    assert torch.equal(out[0], x) and torch.equal(out[1][0], y) and torch.equal(out[1][1][0], x) and isinstance(out[2], util.AffineT)
    with pytest.raises(NotImplementedError):
        util.to_device(torch.device("cpu"), 3.0)


@pytest.mark.gpu
def test_so3_bezier_against_oracle(util, cuda_device):
    rng = np.random.default_rng(5)
    n = 1000
    Rs = [O.random_rotations(n, rng, 2.5)[0].astype(np.float32) for _ in range(3)]
    w = rng.uniform(0, 1, (n, 1)).astype(np.float32)
    dev = [torch.from_numpy(r).to(cuda_device) for r in Rs]
    wd = torch.from_numpy(w).to(cuda_device)
    two = util.so3_bezier(dev[0], dev[1], weight=wd).cpu().numpy()
    assert np.max(np.abs(two - O.so3_lerp(Rs[0], Rs[1], w))) < 3e-6
    three = util.so3_bezier(*dev, weight=wd).cpu().numpy()
    ref = O.so3_lerp(O.so3_lerp(Rs[0], Rs[1], w), O.so3_lerp(Rs[1], Rs[2], w), w)
    assert np.max(np.abs(three - ref)) < 1e-5
    with pytest.raises(ValueError):
        util.so3_bezier(dev[0], weight=wd)
