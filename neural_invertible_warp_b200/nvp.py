"""Host-side mirror of the reference's NVP warp network (model/nvp/nvp_ndr.py:229-468).

``DeformNetwork`` keeps the reference's parameter names (``lin{b}_{a,b}_0.{weight_g,weight_v,bias}``,
``lin{b}_{a,b}_1.{weight,bias}``, ``lin{b}_c.{weight,bias}``), constructor signature and
initialisation, so reference checkpoints load unchanged.  ``forward`` resolves the weight-norm
re-parametrisation and folds the per-image latent code into per-image first-layer biases with a
few B-sized PyTorch ops (autograd handles their backward), then evaluates all points in one CUDA
kernel (csrc/nvp.cu).
"""
import math

import torch
from torch import nn

from . import functional as F

N_FREQ = 6
HID = 128


def pack_effective(p, code, n_blocks=3, n_freq=N_FREQ, prefix=""):
    """(state-dict-style params, code [B,D]) -> (wpack [3*5508], code_bias [3,2,B,128]).

    w = g * v / ||v||_row is torch's legacy ``nn.utils.weight_norm`` (dim=0) used at
    model/nvp/nvp_ndr.py:291-292,335-336; code_b = lin_c(code) + code is :382.
    """
    chunks, biases = [], []
    for b in range(n_blocks):
        cb = torch.addmm(p[f"{prefix}lin{b}_c.bias"], code, p[f"{prefix}lin{b}_c.weight"].t()) + code
        for part, emb in (("a", 2 * (1 + 2 * n_freq)), ("b", 1 + 2 * n_freq)):
            name = f"{prefix}lin{b}_{part}_0"
            if name + ".weight_g" in p:
                v, g = p[name + ".weight_v"], p[name + ".weight_g"]
                w0 = v * (g / v.norm(dim=1, keepdim=True))
            else:
                w0 = p[name + ".weight"]
            chunks += [w0[:, :emb].reshape(-1), p[f"{prefix}lin{b}_{part}_1.weight"].reshape(-1),
                       p[f"{prefix}lin{b}_{part}_1.bias"].reshape(-1)]
            biases.append(torch.addmm(p[name + ".bias"], cb, w0[:, emb:].t()))
    B = code.shape[0]
    return torch.cat(chunks), torch.stack(biases).view(n_blocks, 2, B, -1)


class DeformNetwork(nn.Module):
    """Drop-in for ``model.nvp.nvp_ndr.DeformNetwork`` restricted to what the target models
    instantiate (barf_inn_llff.py:54-55, pose_models/inn.py:23-27): d_in=3, n_blocks=3,
    n_layers=1, skip_in=[], multires=6, weight_norm=True, softplus.  Anything else raises."""

    def __init__(self, d_feature, d_in, d_out_1, d_out_2, n_blocks, d_hidden, n_layers, skip_in=(4,),
                 multires=0, weight_norm=True, actfn="softplus"):
        super().__init__()
        ok = (d_in == 3 and d_out_1 == 1 and d_out_2 == 3 and n_blocks == 3 and d_hidden == HID and n_layers == 1
              and len(tuple(skip_in)) == 0 and multires == N_FREQ and weight_norm and actfn == "softplus")
        if not ok:
            raise RuntimeError("niw_b200 DeformNetwork: only the configuration used by barf_inn_llff / barf_inn_dtu "
                               "is implemented in CUDA (3 blocks, hidden 128, 6 bands, weight-norm, softplus)")
        self.n_blocks, self.d_feature = n_blocks, d_feature
        for b in range(n_blocks):
            for part, ori in (("a", 2), ("b", 1)):
                emb = ori * (1 + 2 * multires)
                n_out = d_out_1 if part == "a" else d_out_2
                lin0 = nn.Linear(emb + d_feature, d_hidden)
                nn.init.constant_(lin0.bias, 0.0)
                nn.init.normal_(lin0.weight[:, :ori], 0.0, math.sqrt(2) / math.sqrt(d_hidden))
                nn.init.constant_(lin0.weight[:, ori:], 0.0)
                lin0 = nn.utils.weight_norm(lin0)
                lin1 = nn.Linear(d_hidden, n_out)
                nn.init.constant_(lin1.bias, 0.0)
                nn.init.constant_(lin1.weight, 0.0)
                setattr(self, f"lin{b}_{part}_0", lin0)
                setattr(self, f"lin{b}_{part}_1", lin1)
        for b in range(n_blocks):
            lin = nn.Linear(d_feature, d_feature)
            nn.init.constant_(lin.bias, 0.0)
            nn.init.constant_(lin.weight, 0.0)
            setattr(self, f"lin{b}_c", lin)

    def _params(self):
        p = {}
        for b in range(self.n_blocks):
            for part in ("a", "b"):
                l0, l1 = getattr(self, f"lin{b}_{part}_0"), getattr(self, f"lin{b}_{part}_1")
                p[f"lin{b}_{part}_0.weight_g"], p[f"lin{b}_{part}_0.weight_v"] = l0.weight_g, l0.weight_v
                p[f"lin{b}_{part}_0.bias"] = l0.bias
                p[f"lin{b}_{part}_1.weight"], p[f"lin{b}_{part}_1.bias"] = l1.weight, l1.bias
            lc = getattr(self, f"lin{b}_c")
            p[f"lin{b}_c.weight"], p[f"lin{b}_c.bias"] = lc.weight, lc.bias
        return p

    def forward(self, deformation_code, input_pts, alpha_ratio=0):
        """deformation_code [B,D], input_pts [B,P,1,3] -> [B,P,1,3]  (nvp_ndr.py:365)."""
        squeeze = input_pts.dim() == 4
        pts = input_pts[:, :, 0] if squeeze else input_pts
        wpack, code_bias = pack_effective(self._params(), deformation_code, self.n_blocks)
        out = F.nvp_warp(wpack, code_bias, pts.detach(), alpha_ratio)
        return out[:, :, None] if squeeze else out
