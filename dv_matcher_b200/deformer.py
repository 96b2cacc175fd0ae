"""Deformer head with the reference's parameters and call signature (models/model.py:433-478).

State-dict compatible with the shipped checkpoints (`ckpt/*/ep_deformer_val_best.pth`:
`conv_layer.{weight,bias}`, `deformation_decoder_layer.linear.{0,2,4,6}.{weight,bias}`).
The MLP stays a stock torch module (small cuBLAS GEMMs); what moves into the
library is everything around it: the [B,N,10,128] neighbourhood gathers + 1x1 conv (fused,
dvm_gather_conv_*), `Pi_12 @ feat2` (10-sparse, dvm_sparse_transfer_*), and -- in `forward_fused`
-- the fact that only the K graph-node rows are ever consumed (models/model.py:473-476).
"""
import torch
import torch.nn as nn

from . import ops
from .geometry import gather_conv, index_points_idx
from .maps import SparseSoftMap, _MapBase


class MLP(nn.Module):
    """models/model.py:433-452."""

    def __init__(self, input_dim, output_dim, hidden_dims=(), bias=True, act=None):
        super().__init__()
        act = act or nn.ELU()
        hidden_dims = list(hidden_dims)
        if hidden_dims:
            fc = [nn.Linear(input_dim, hidden_dims[0], bias=bias), act]
            for i in range(len(hidden_dims) - 1):
                fc += [nn.Linear(hidden_dims[i], hidden_dims[i + 1], bias=bias), act]
            fc.append(nn.Linear(hidden_dims[-1], output_dim, bias=bias))
        else:
            fc = [nn.Linear(input_dim, output_dim, bias=bias), act]
        self.linear = nn.Sequential(*fc)

    def forward(self, x):
        return self.linear(x)

    def forward_tc(self, parts):
        """Inference path: the concatenation `parts` (tensors [..., c_i]) goes straight into a staging buffer whose row pitch
        is a multiple of 4 floats, every nn.Linear (+ its ELU) is one dvm_linear_act_fwd launch (tcgen05, 3xTF32)."""
        lead = parts[0].shape[:-1]
        K = sum(p.shape[-1] for p in parts)
        x = torch.empty(*lead, (K + 3) // 4 * 4, dtype=torch.float32, device=parts[0].device)
        c = 0
        for p in parts:
            x[..., c:c + p.shape[-1]] = p
            c += p.shape[-1]
        x = x.reshape(-1, x.shape[-1])
        mods = list(self.linear)
        i, cols = 0, K
        while i < len(mods):
            lin = mods[i]
            elu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ELU) and mods[i + 1].alpha == 1.0
            x = ops.linear_act_fwd(x, lin.weight, lin.bias, "elu" if elu else "none", x_cols=cols)
            cols = None
            i += 2 if elu else 1
            if not elu and i < len(mods) and not isinstance(mods[i], nn.Linear):
                x = mods[i](x)                                   # an activation other than ELU(1): stock torch
                i += 1
        return x.reshape(*lead, x.shape[-1])


class Deformer(nn.Module):
    """models/model.py:454-478."""

    def __init__(self, k):
        super().__init__()
        self.conv_layer = nn.Conv2d(in_channels=k, out_channels=1, kernel_size=(1, 1))
        self.deformation_decoder_layer = MLP(input_dim=128 * 2 + 3 * 2, output_dim=3 + 6, hidden_dims=[512, 256, 128], bias=True, act=nn.ELU())

    def forward(self, feat1_conv, feat2_conv, verts1, verts12, Pi_12, fps1):
        """Reference signature: feat*_conv are the gathered neighbourhoods [B,N,k,128]."""
        feat1 = self.conv_layer(feat1_conv.permute(0, 2, 1, 3)).squeeze(1)
        feat2 = self.conv_layer(feat2_conv.permute(0, 2, 1, 3)).squeeze(1)
        feat2 = torch.matmul(Pi_12, feat2)                       # SparseSoftMap -> 10-sparse gather
        return self._decode(index_points_idx(verts1, fps1), index_points_idx(feat1, fps1),
                            index_points_idx(verts12, fps1), index_points_idx(feat2, fps1))

    def forward_fused(self, feat1, feat2, idx11, idx22, verts1, verts12, Pi_12, fps1):
        """Same result from the un-gathered features feat1 [B,N,128], feat2 [B,M,128] and the xyz k-NN index
        lists: the conv is fused with the gather, and feat1 / Pi@feat2 are evaluated at the K node rows only."""
        w, b = self.conv_layer.weight, self.conv_layer.bias
        B, K = fps1.shape
        k = idx11.shape[-1]
        idx11_nodes = torch.gather(idx11, 1, fps1[..., None].expand(B, K, k))
        st_feat1 = gather_conv(feat1, idx11_nodes, w, b)                       # [B,K,128]
        feat2c = gather_conv(feat2, idx22, w, b)                               # [B,M,128]
        if isinstance(Pi_12, _MapBase):
            sp = Pi_12.sparse()
            kk = sp.idx.shape[-1]
            node_map = SparseSoftMap(torch.gather(sp.idx, 1, fps1[..., None].expand(B, K, kk)).contiguous(),
                                     torch.gather(sp.w, 1, fps1[..., None].expand(B, K, kk)).contiguous(), sp.M)
            st_feat2 = node_map.matmul(feat2c)                                 # [B,K,128]
        else:
            st_feat2 = index_points_idx(torch.matmul(Pi_12, feat2c), fps1)
        return self._decode(index_points_idx(verts1, fps1), st_feat1, index_points_idx(verts12, fps1), st_feat2)

    def _decode(self, st_vts1, st_feat1, st_vts12, st_feat2):
        parts = [st_vts1, st_feat1, st_vts12, st_feat2]
        needs_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in parts) or any(q.requires_grad for q in self.parameters()))
        if parts[0].is_cuda and not needs_grad:
            return self.deformation_decoder_layer.forward_tc(parts)
        return self.deformation_decoder_layer(torch.cat(parts, dim=-1))           # training: autograd through torch's Linear
