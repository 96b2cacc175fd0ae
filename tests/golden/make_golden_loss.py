"""Golden fixture for the training loss: run the UNMODIFIED reference `GraphDeformLoss_Neural` (full-shape) and
`GraphDeformLoss_Neural_Partial` forward + backward on CPU (build container only) and store the results.

    python tests/golden/make_golden_loss.py

The reference modules are imported through oracle/refimport.py's stub hook.  The one thing that has to be supplied
is `dist_chamfer_3D.chamfer_3DDist` (un-vendored CUDA extension): a differentiable torch restatement
(squared direct-difference distances, min / arg-min, whose autograd gradient 2 g (a - b) is the extension's
backward) is patched in, and `chamfer_loss`'s hard-coded `.cuda()` calls are bypassed.  Real SCAPE geometry
(subsampled), seeded synthetic features, the shipped Deformer checkpoint, Euclidean stand-ins for the geodesic
matrices (float64, like the dataset's).  RNG: torch / numpy / random seeded with 11 before each forward.
"""
import os
import random
import sys
import tempfile

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refimport  # noqa: E402
from dv_matcher_b200 import synthetic  # noqa: E402

QSTEP = 2.0 ** -9
N_FULL, M_PART, C = 1000, 600, 128


class _Chamfer(nn.Module):
    def forward(self, a, b):
        d = ((a[:, :, None, :] - b[:, None, :, :]) ** 2).sum(-1)
        d1, i1 = d.min(2)
        d2, i2 = d.min(1)
        return d1, d2, i1.int(), i2.int()


def _cpu_chamfer_loss_full(self, pos1, pos2):
    d1, d2, _, _ = self.chamfer_dist_3d(pos1, pos2)
    return torch.mean(d1) + torch.mean(d2)


def _cpu_chamfer_loss_partial(self, pos1, pos2):
    d1, d2, _, _ = self.chamfer_dist_3d(pos1, pos2)
    return torch.mean(d1) if d1.shape[1] <= d2.shape[1] else torch.mean(d2)


def quantise(f):
    q = torch.round(f / QSTEP).clamp(-32767, 32767).to(torch.int16)
    return q, q.float() * QSTEP


def seed_all(s):
    torch.manual_seed(s)
    np.random.seed(s)
    random.seed(s)


def run(loss_mod, deformer, feats, verts, dists, alpha):
    f1 = feats[0].clone().requires_grad_(True)
    f2 = feats[1].clone().requires_grad_(True)
    deformer.zero_grad()
    seed_all(11)
    out = loss_mod(f1, f2, dists[0], dists[1], verts[0], verts[1], alpha, deformer)
    out[0].backward()
    vals = [float(o) for o in out]
    grads = {n: p.grad.detach().clone() for n, p in deformer.named_parameters()}
    return vals, f1.grad.detach(), f2.grad.detach(), grads


def main():
    torch.set_num_threads(os.cpu_count())
    ref_loss, _, ref_model, _ = refimport.modules()
    ref_loss.dist_chamfer_3D.chamfer_3DDist = _Chamfer
    root = refimport.REFERENCE_ROOT
    meshes = ["shapes_train/mesh000.off", "shapes_test/mesh053.off", "shapes_train/mesh001.off", "shapes_test/mesh052.off"]
    full = [torch.from_numpy(refimport.load_off_vertices(f"{root}/data/scape_r/{m}")) for m in meshes]
    g = torch.Generator().manual_seed(515)
    sel = [torch.randperm(v.shape[0], generator=g)[:N_FULL].sort().values for v in full]
    v1 = torch.stack([full[0][sel[0]], full[2][sel[2]]]).contiguous()      # B = 2 source shapes
    v2 = torch.stack([full[1][sel[1]], full[3][sel[3]]]).contiguous()
    field = synthetic.FeatureField(C)
    canon = full[0]                                                          # features follow mesh000's coordinates (vertex-ordered)
    f1q, f1 = quantise(torch.stack([field(canon[sel[0] % canon.shape[0]]), field(canon[sel[2] % canon.shape[0]])]) + 0.05 * torch.randn(2, N_FULL, C, generator=g))
    f2q, f2 = quantise(torch.stack([field(canon[sel[1] % canon.shape[0]]), field(canon[sel[3] % canon.shape[0]])]) + 0.05 * torch.randn(2, N_FULL, C, generator=g))
    d1 = torch.cdist(v1.double(), v1.double())
    d2 = torch.cdist(v2.double(), v2.double())

    deformer = ref_model.Deformer(k=10)
    sd = torch.load(f"{root}/ckpt/dvmatcher_scape_r/ep_deformer_val_best.pth", map_location="cpu")
    deformer.load_state_dict(sd, strict=True)

    out = dict(qstep=np.float64(QSTEP), xyz1=v1.numpy(), xyz2=v2.numpy(), feat1_q=f1q.numpy(), feat2_q=f2q.numpy(),
               n_part=np.int64(M_PART))
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)                                                        # the reference dumps OFF files into ./visual_result
        try:
            # ---------------- full-shape loss (config/scape_r.yaml weights, smaller k_dist / N_dist for N = 1000)
            ref_loss.GraphDeformLoss_Neural.chamfer_loss = _cpu_chamfer_loss_full
            crit = ref_loss.GraphDeformLoss_Neural(k_deform=10, w_dist=0.02, w_map=0.005, k_dist=50, N_dist=200, partial=False,
                                                   w_deform=0.5, w_img=0, w_rank=0, w_self_rec=0.5, w_cd=0.1, w_arap=0.01, save_name="golden")
            for alpha in (10.0, 40.0):
                vals, g1, g2, gd = run(crit, deformer, (f1, f2), (v1, v2), (d1, d2), alpha)
                tag = f"full_a{int(alpha)}"
                out[f"{tag}_vals"] = np.asarray(vals, dtype=np.float64)
                out[f"{tag}_gfeat1"] = g1.numpy()[:, ::5].copy()
                out[f"{tag}_gfeat2"] = g2.numpy()[:, ::5].copy()
                out[f"{tag}_gnorms"] = np.asarray([g1.norm().item(), g2.norm().item()], dtype=np.float64)
                for n, t in gd.items():
                    out[f"{tag}_gd_{n}"] = t.numpy() if t.numel() <= 4096 else np.asarray([t.norm().item(), t.flatten()[:64].sum().item()])
                print(tag, vals)
            # ---------------- partial loss (config/scape_partial.yaml weights): full source vs a 600-point subset target
            ref_loss.GraphDeformLoss_Neural_Partial.chamfer_loss = _cpu_chamfer_loss_partial
            critp = ref_loss.GraphDeformLoss_Neural_Partial(k_deform=10, w_dist=0.02, w_map=0.0, k_dist=50, N_dist=200, partial=True,
                                                            w_deform=1000, w_img=0, w_rank=0, w_self_rec=1000, w_cd=0.1, w_arap=0.01, save_name="golden")
            v2p, f2p, d2p = v2[:, :M_PART].contiguous(), f2[:, :M_PART].contiguous(), d2[:, :M_PART, :M_PART].contiguous()
            vals, g1, g2, gd = run(critp, deformer, (f1, f2p), (v1, v2p), (d1, d2p), 40.0)
            out["part_a40_vals"] = np.asarray(vals, dtype=np.float64)
            out["part_a40_gfeat1"] = g1.numpy()[:, ::5].copy()
            out["part_a40_gfeat2"] = g2.numpy()[:, ::5].copy()
            out["part_a40_gnorms"] = np.asarray([g1.norm().item(), g2.norm().item()], dtype=np.float64)
            print("part_a40", vals)
        finally:
            os.chdir(cwd)
    for n, t in sd.items():
        out[f"deformer_{n}"] = t.numpy()
    np.savez_compressed(os.path.join(HERE, "ref_loss.npz"), **out)
    print("ref_loss.npz", sum(v.nbytes for v in out.values() if hasattr(v, "nbytes")) // 1024, "KiB")


if __name__ == "__main__":
    main()
