"""Generate the golden fixtures by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Imports /root/reference's own models/loss.py, lib/deformation_graph_point.py and
lib/deformation_graph.py through oracle/refimport.py's stub hook, runs them on CPU on real SCAPE
geometry (data/scape_r mesh000 -> mesh053, the pair deform.py:159-160 hard-codes) with seeded
synthetic features, and writes small .npz files next to this script.  The inputs are stored in the
fixtures (features quantised to int16 steps of 2^-9 so they are exact in fp32 and portable), so the
tests never need /root/reference.
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refimport  # noqa: E402
from dv_matcher_b200 import synthetic  # noqa: E402

QSTEP = 2.0 ** -9


def quantise(f):
    q = torch.round(f / QSTEP).clamp(-32767, 32767).to(torch.int16)
    return q, q.float() * QSTEP


def main():
    torch.set_num_threads(os.cpu_count())
    ref_loss, ref_dg, ref_model, ref_dg_aa = refimport.modules()
    root = refimport.REFERENCE_ROOT
    v1_full = refimport.load_off_vertices(f"{root}/data/scape_r/shapes_train/mesh000.off")
    v2_full = refimport.load_off_vertices(f"{root}/data/scape_r/shapes_test/mesh053.off")

    # ---------------------------------------------------------------- A. hard / soft maps, N != M
    N, M, C = 1500, 1200, 128
    g = torch.Generator().manual_seed(20260)
    sel1 = torch.randperm(v1_full.shape[0], generator=g)[:N].sort().values
    sel2 = torch.randperm(v2_full.shape[0], generator=g)[:M].sort().values
    xyz1 = torch.from_numpy(v1_full)[sel1].contiguous()
    xyz2 = torch.from_numpy(v2_full)[sel2].contiguous()
    field = synthetic.FeatureField(C)
    # mesh000 / mesh053 are in vertex correspondence only up to remeshing: features are a function
    # of each shape's own coordinates (poses differ), so the maps are peaked but not trivial.
    f1q, f1 = quantise(field(xyz1) + 0.05 * torch.randn(N, C, generator=g))
    f2q, f2 = quantise(field(torch.from_numpy(v1_full)[sel2 % v1_full.shape[0]]) + 0.05 * torch.randn(M, C, generator=g))
    x, y = f1[None], f2[None]
    out = dict(xyz1=xyz1.numpy(), xyz2=xyz2.numpy(), feat1_q=f1q.numpy(), feat2_q=f2q.numpy(), qstep=np.float64(QSTEP))
    with torch.no_grad():
        out["T12"] = ref_loss.knnsearch_t(x, y)[0, :, 0].numpy().astype(np.int32)
        out["T21"] = ref_loss.knnsearch_t(y, x)[0, :, 0].numpy().astype(np.int32)
        out["search_t"] = ref_loss.search_t(x, y)[0, :, 0].numpy().astype(np.int32)
        loss_mod = ref_loss.GraphDeformLoss_Neural.__new__(ref_loss.GraphDeformLoss_Neural)
        idx22 = ref_loss.knn_grad(xyz2[None], xyz2[None], 10)
        nb2 = ref_loss.index_points(xyz2[None], idx22)
        for alpha in (10.0, 50.0, 100.0):
            pi = ref_loss.knnsearch_t_grad(x, y, alpha=alpha)
            pi10 = ref_loss.GraphDeformLoss_Neural.topk_pi(loss_mod, pi)
            vals, idx = torch.topk(pi10, 10, dim=-1)
            tag = f"a{int(alpha)}"
            out[f"pi_vals_{tag}"] = vals[0].numpy()
            out[f"pi_idx_{tag}"] = idx[0].numpy().astype(np.int32)
            out[f"pi_rowsum_{tag}"] = pi10.sum(-1)[0].numpy()
            out[f"verts12_{tag}"] = torch.matmul(pi10, xyz2[None])[0].numpy()
            if alpha == 50.0:
                out["nb_transfer_a50"] = torch.einsum("bij,bjkm->bikm", pi10, nb2)[0].numpy()
                out["idx22"] = idx22[0].numpy().astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "ref_maps.npz"), **out)
    print("ref_maps.npz", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})

    # ---------------------------------------------------------------- B. xyz 10-NN on the full mesh
    verts = torch.from_numpy(v1_full)
    out = dict(xyz=v1_full)
    with torch.no_grad():
        out["knn_grad_k10"] = ref_loss.knn_grad(verts[None], verts[None], 10)[0].numpy().astype(np.int16)

    # ---------------------------------------------------------------- C. graph construction + forward
    torch.manual_seed(7)
    np.random.seed(7)
    random.seed(7)
    num_nodes_all, dg_list = ref_loss.GraphDeformLoss_Neural.deformation_graph_node(loss_mod, verts[None])
    dg = dg_list[0]
    out["nodes_idx"] = np.asarray(dg.nodes_idx).astype(np.int16)
    out["fps_start"] = np.int64(dg.nodes_idx[0])
    out["one_ring"] = np.asarray(dg.one_ring_neigh).astype(np.int16)
    out["influence"] = dg.influence_nodes_idx.numpy().astype(np.int16)
    out["dists"] = dg.dists.numpy()
    out["weights"] = dg.weights.numpy().astype(np.float32)
    out["sigma"] = np.float64(dg.sigma)
    out["num_nodes_all"] = num_nodes_all[0].numpy().astype(np.int16)
    K = len(dg.nodes_idx)
    gen = torch.Generator().manual_seed(99)
    d9 = synthetic.random_rigid_field(K, gen)
    iden = torch.tensor([1, 0, 0, 0, 1, 0], dtype=torch.float32)
    R = ref_loss.rotation_6d_to_matrix(d9[None, :, 3:] + iden)
    T = d9[None, :, :3]
    with torch.no_grad():
        warped, arap, sr = dg(verts, R[0].unsqueeze(0), T[0].unsqueeze(0))
    out["deform9"] = d9.numpy()
    out["R"] = R[0].numpy()
    out["warped"] = warped[0].numpy()
    out["arap"] = np.float32(arap)
    out["sr"] = np.float32(sr)

    # ---------------------------------------------------------------- D. axis-angle variant
    dga = ref_dg_aa.DeformationGraph.__new__(ref_dg_aa.DeformationGraph)
    torch.nn.Module.__init__(dga)
    for name in ("nodes_idx", "influence_nodes_idx", "weights", "one_ring_neigh", "max_neigh_num"):
        setattr(dga, name, getattr(dg, name))
    dga.one_ring_neigh = torch.as_tensor(np.asarray(dg.one_ring_neigh))
    aa = 0.2 * torch.randn(1, K, 3, generator=gen)
    try:
        with torch.no_grad():
            w2, arap2, sr2 = dga(verts, aa, T)
        out["axis_angle"] = aa[0].numpy()
        out["warped_aa"] = w2[0].numpy()
        out["arap_aa"] = np.float32(arap2)
        out["sr_aa"] = np.float32(sr2)
    except Exception as e:  # the axis-angle class is not called by any entry script
        print("axis-angle variant not runnable verbatim:", repr(e))
    np.savez_compressed(os.path.join(HERE, "ref_graph.npz"), **out)
    print("ref_graph.npz", {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
