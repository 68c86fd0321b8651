"""The reference's skew-vector denoiser (``RotPredict``, so3_train.py:11-49 / bingham_train.py:9-47) and its
tensor-core fusion with the reverse step.

``RotPredict`` here has the reference's constructor, parameter names (``net.0.weight`` ... ``net.8.bias``: state
dicts load both ways) and forward (stock PyTorch: the training path needs autograd).  For sampling,
``SO3Diffusion.p_sample`` recognises a ``RotPredict(out_type="skewvec")`` denoiser called with a step index shared
by the batch (so3_test.py:31, diffusion.py:328-337) and runs denoiser + reverse step as ONE kernel
(``so3d_rotpredict_p_sample_f32``: tcgen05 tf32 MMAs with a 3-term split, activations resident in tensor memory).
"""
import math

import torch
import torch.nn as nn

from . import ops

D_MODEL = 65


class SinusoidalPosEmb(nn.Module):
    """models.py:13-25."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, x):
        half_dim = self.dim // 2
        emb = math.log(10000) / (half_dim - 1)
        emb = torch.exp(torch.arange(half_dim, device=x.device) * -emb)
        emb = x[:, None] * emb[None, :]
        return torch.cat((emb.sin(), emb.cos()), dim=-1)


class RotPredict(nn.Module):
    """so3_train.py:11-49, same constructor defaults.  in_type 'rotmat' (9 matrix entries + sinusoidal time features);
    out_type 'rotmat' (the reference's default: 6 outputs mapped to a rotation by `six2rmat`, util.py:67-76) or
    'skewvec' (3 outputs -- what every reference script passes, so3_train.py:60, and the head the fused sampling kernel
    implements; the 6-D head runs as stock PyTorch)."""

    def __init__(self, d_model=D_MODEL, out_type="rotmat", in_type="rotmat"):
        super().__init__()
        if in_type != "rotmat":
            raise ValueError("only in_type='rotmat' exists in the reference")
        if out_type not in ("skewvec", "rotmat"):
            raise RuntimeError(f"Unexpected out_type: {out_type}")  # the reference builds this error but forgets to raise it
        self.in_type, self.out_type = in_type, out_type
        self.d_out = 3 if out_type == "skewvec" else 6
        self.time_embedding = SinusoidalPosEmb(d_model - 9)
        self.net = nn.Sequential(
            nn.Linear(d_model, d_model), nn.SiLU(),
            nn.Linear(d_model, d_model), nn.SiLU(),
            nn.Linear(d_model, d_model), nn.SiLU(),
            nn.Linear(d_model, d_model), nn.SiLU(),
            nn.Linear(d_model, self.d_out),
        )
        self._packed = None  # (key, blob, c1_table)

    def forward(self, x: torch.Tensor, t: torch.Tensor):
        x_flat = torch.flatten(x, start_dim=-2)
        t_emb = self.time_embedding(t)
        if t_emb.shape[0] == 1:
            t_emb = t_emb.expand(x_flat.shape[0], -1)
        out = self.net(torch.cat((x_flat, t_emb), dim=-1))
        if self.out_type == "rotmat":
            from .util import six2rmat

            out = six2rmat(out)
        return out

    # ---- fused sampling path ------------------------------------------------------------------
    def fusable(self):
        lin = self.net[0]
        return self.out_type == "skewvec" and lin.in_features == D_MODEL and lin.weight.is_cuda and lin.weight.dtype == torch.float32

    def _linears(self):
        return [self.net[i] for i in (0, 2, 4, 6, 8)]

    def packed(self, num_timesteps):
        """(blob, c1_table) for the fused kernel, rebuilt when a parameter changed (optimizer step, load_state_dict)."""
        lins = self._linears()
        key = (num_timesteps,) + tuple((p.data_ptr(), p._version) for l in lins for p in (l.weight, l.bias))
        if self._packed is None or self._packed[0] != key:
            with torch.no_grad():
                dev = lins[0].weight.device
                blob = ops.rotpredict_pack([l.weight for l in lins], [l.bias for l in lins])
                # time embedding of every step folded into layer 1's bias: b1 + W1[:, 9:] @ emb(t)
                steps = torch.arange(num_timesteps, device=dev)
                emb = self.time_embedding(steps).to(torch.float64)
                w1 = lins[0].weight.detach().to(torch.float64)
                c1 = (emb @ w1[:, 9:].T + lins[0].bias.detach().to(torch.float64)).to(torch.float32).contiguous()
            self._packed = (key, blob, c1)
        return self._packed[1], self._packed[2]
