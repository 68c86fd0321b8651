"""Pin oracle/ref_port.py (the torch-CPU port timed as the CPU baseline) against the golden vectors
generated from the unmodified reference.  CPU only."""
import numpy as np
import torch

from oracle import ref_port as P


def t(a):
    return torch.tensor(a)


def test_port_util(golden):
    g = golden("util_l0")
    assert torch.allclose(P.log_rmat(t(g["Rall"])), t(g["log_all"]), atol=1e-6)
    ax, an = P.rmat_to_aa(t(g["R"]))
    assert torch.allclose(an, t(g["angle"]), atol=1e-6) and torch.allclose(ax, t(g["axis"]), atol=1e-6)
    assert torch.allclose(P.aa_to_rmat(t(g["axes_in"]), t(g["ang_in"])), t(g["aa_rmat"]), atol=1e-6)
    assert torch.allclose(P.so3_scale(t(g["R"]), t(g["scalars"])), t(g["scaled"]), atol=1e-6)


def test_port_igso3(golden):
    g = golden("igso3")
    for k, e in enumerate(g["eps_list"]):
        d = P.IGSO3(torch.tensor(float(e)))
        ref = t(g["density"][k])
        got = d.density(t(g["omega"]))
        ok = torch.isfinite(ref)
        assert torch.equal(got[ok], ref[ok])
        assert torch.equal(d.trap[:, 0], t(g["trap"][k]))
        torch.manual_seed(777)
        assert torch.allclose(d.sample((256,)), t(g["samples"][k]), atol=1e-6)
        lp, gr = P.score_via_autograd(t(g["lp_R"][k]), torch.tensor(float(e)))
        assert torch.allclose(lp, t(g["logp"][k]), atol=1e-5)
        assert torch.allclose(gr, t(g["logp_grad"][k]), rtol=1e-3, atol=1e-3 * float(np.abs(g["logp_grad"][k]).max()))
    # per-row eps extension agrees with the scalar path
    R = t(g["lp_R"][3])
    lp_rows, _ = P.score_via_autograd(R, torch.full((R.shape[0],), float(g["eps_list"][3])))
    assert torch.allclose(lp_rows, t(g["logp"][3]), atol=1e-5)


def test_port_diffusion(golden):
    g = golden("diffusion")
    p = P.SO3DiffusionPort(lambda x, tt: t(g["pred"]))
    assert torch.allclose(p.q_sample(t(g["x0"]), t(g["t"]), t(g["noise"])), t(g["x_t"]), atol=1e-6)
    for k, step in enumerate(g["ps_steps"]):
        tt = torch.full((128,), int(step), dtype=torch.long)
        torch.manual_seed(99 + int(step))
        out = p.p_sample(t(g["x_t"]), tt)
        assert torch.allclose(out, t(g["ps_out"][k]), atol=2e-5), step
