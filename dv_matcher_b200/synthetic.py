"""Seeded synthetic shape pairs and features.  Pure torch, device-agnostic.

Geometry: surface samples of an ellipsoid with SCAPE-like semi-axes (0.35, 0.65, 0.14) and a
low-frequency radial bump; the target is an independent sample of the same surface, warped by a
smooth sinusoidal bend (amplitude 0.1) plus N(0, 0.005^2) noise.  Features, two regimes:
  structured   F = g(p) + 0.05 N(0,1), g a fixed seeded MLP 3->256->128 (LeakyReLU 0.2) of the
               canonical (pre-warp) coordinates followed by a BatchNorm-like per-channel
               normalisation and LeakyReLU(0.2), as LG-Net's last block does: peaked soft maps;
  unstructured F = LeakyReLU_0.2(N(0,1)): flat soft maps, worst case for exp-skipping.
"""
import math

import torch
import torch.nn.functional as F

SEMI_AXES = (0.35, 0.65, 0.14)
BASE_SEED = 1234
FEAT_DIM = 128


def _gen(seed, device="cpu"):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def ellipsoid_cloud(n, gen, device="cpu"):
    """Canonical surface samples [n,3] fp32 (area-biased towards uniform by normal-scaling)."""
    u = torch.randn(n, 3, generator=gen, device=device)
    u = u / u.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    bump = 1.0 + 0.08 * torch.sin(3.0 * u[:, 0] + 1.0) * torch.cos(2.0 * u[:, 1] - 0.5)
    ax = torch.tensor(SEMI_AXES, device=device)
    return (u * bump[:, None] * ax).float().contiguous()


def smooth_warp(p):
    """Per-axis sinusoidal bend of amplitude 0.1."""
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    return torch.stack([x + 0.1 * torch.sin(2.2 * y + 0.3),
                        y + 0.1 * torch.sin(3.1 * z * 4.0 + 1.1),
                        z + 0.1 * torch.sin(2.7 * x - 0.7) * 0.4], dim=-1)


class FeatureField:
    """The fixed seeded MLP g: R^3 -> R^C used by the structured regime."""

    def __init__(self, c=FEAT_DIM, seed=4321, device="cpu"):
        g = _gen(seed)
        self.w1 = (torch.randn(3, 256, generator=g) * 4.0).to(device)
        self.b1 = torch.randn(256, generator=g).to(device)
        self.w2 = (torch.randn(256, c, generator=g) / math.sqrt(256.0)).to(device)
        self.b2 = (torch.randn(c, generator=g) * 0.1).to(device)
        # LG-Net ends in conv -> BatchNorm -> LeakyReLU(0.2) (models/model.py:527-529): mimic the
        # per-channel statistics with a fixed normalisation measured on a canonical cloud.
        ref = self._raw(ellipsoid_cloud(4096, g).to(device))
        self.mean = ref.mean(0)
        self.std = ref.std(0).clamp_min(1e-6)

    def _raw(self, p):
        h = F.leaky_relu(p @ self.w1 + self.b1, 0.2)
        return h @ self.w2 + self.b2

    def __call__(self, p):
        return F.leaky_relu((self._raw(p) - self.mean) / self.std, 0.2)


def make_pair(n, m, pair_index=0, regime="structured", c=FEAT_DIM, device="cpu", field=None):
    """One synthetic pair.  Returns dict(xyz1 [n,3], xyz2 [m,3], feat1 [n,c], feat2 [m,c]) fp32."""
    gen = _gen(BASE_SEED + pair_index)          # CPU generator: identical on every device
    p1 = ellipsoid_cloud(n, gen)
    p2 = ellipsoid_cloud(m, gen)
    xyz2 = smooth_warp(p2) + 0.005 * torch.randn(m, 3, generator=gen)
    if regime == "structured":
        field = field or FeatureField(c)
        f1 = field(p1) + 0.05 * torch.randn(n, c, generator=gen)
        f2 = field(p2) + 0.05 * torch.randn(m, c, generator=gen)
    elif regime == "unstructured":
        f1 = F.leaky_relu(torch.randn(n, c, generator=gen), 0.2)
        f2 = F.leaky_relu(torch.randn(m, c, generator=gen), 0.2)
    else:
        raise ValueError(f"unknown regime {regime!r}")
    out = dict(xyz1=p1, xyz2=xyz2.float(), feat1=f1.float(), feat2=f2.float())
    return {k: v.contiguous().to(device) for k, v in out.items()}


def make_batch(b, n, m, first_pair=0, regime="structured", c=FEAT_DIM, device="cpu", pin=False):
    """B stacked pairs: dict of [B,*,*] tensors.  Pair p uses seed BASE_SEED + first_pair + p."""
    field = FeatureField(c) if regime == "structured" else None
    pairs = [make_pair(n, m, first_pair + i, regime, c, "cpu", field) for i in range(b)]
    out = {k: torch.stack([p[k] for p in pairs]).contiguous() for k in pairs[0]}
    if pin and torch.cuda.is_available():
        out = {k: v.pin_memory() for k, v in out.items()}
    return {k: v.to(device) for k, v in out.items()} if device != "cpu" else out


def random_rigid_field(k, gen, rot_scale=0.15, trans_scale=0.02):
    """Seeded Deformer-like output [k,9]: t (3) + 6D residual (6) around the identity."""
    t = trans_scale * torch.randn(k, 3, generator=gen)
    r6 = rot_scale * torch.randn(k, 6, generator=gen)
    return torch.cat([t, r6], dim=-1).float()
