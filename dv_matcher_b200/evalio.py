"""Wire / on-disk formats either side of the hot path and the geodesic-error evaluator.

Writers produce byte-for-byte what the reference's entry points write, so its MATLAB evaluation (or this module's
GPU evaluator) can read the output of either implementation:

  save_maps_txt      test.py:111-121        T/T_{a}_{b}.txt, one 1-based index per line (np.savetxt fmt '%i')
  save_features_mat  test.py:123-133        feature/usefeature_{name}.mat, key 'uphi'
  save_off_file      deform.py:79-84        vertex-only OFF
  save/load_cache    models/dataset.py:219-228   the (verts_list, used_shapes, fps_list, dist_list) tuple

geodesic_error / evaluate_pairs replace eval/main.m:19-40: `knnsearch(phiT, phiS(vts_src,:))` is the hard map of this
library (dvm_softmap_fwd, mode HARD: index-exact), followed by the lookup `M_T(idx, vts_tar)`.
"""
import os

import numpy as np
import torch

from . import maps
from .losses import save_off_file  # noqa: F401  (re-exported: same writer as deform.py:79-84)


# ------------------------------------------------------------------------------------------------
# writers / readers
# ------------------------------------------------------------------------------------------------
def _np(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def save_maps_txt(save_path, name1, name2, T12, T21):
    """test.py:111-121.  T12, T21: 1-based maps as returned by `search_t` ([1,N,1] / [N,1] / [N])."""
    d = os.path.join(save_path, "T")
    os.makedirs(d, exist_ok=True)
    p12 = os.path.join(d, f"T_{name1}_{name2}.txt")
    p21 = os.path.join(d, f"T_{name2}_{name1}.txt")
    t12, t21 = _np(T12), _np(T21)
    if t12.ndim == 3:
        t12 = t12.squeeze(0)
    if t21.ndim == 3:
        t21 = t21.squeeze(0)
    np.savetxt(p12, t12, fmt="%i")
    np.savetxt(p21, t21, fmt="%i")
    return p12, p21


def load_map_txt(path):
    """1-based map file -> int64 [N] (still 1-based, like the file)."""
    return np.loadtxt(path, dtype=np.int64).reshape(-1)


def save_features_mat(save_path, name, feat):
    """test.py:123-133: feature/usefeature_{name}.mat with key 'uphi' ([N,C])."""
    import scipy.io
    d = os.path.join(save_path, "feature")
    os.makedirs(d, exist_ok=True)
    p = os.path.join(d, f"usefeature_{name}.mat")
    f = _np(feat)
    if f.ndim == 3:
        f = f.squeeze(0)
    scipy.io.savemat(p, {"uphi": f})
    return p


def load_features_mat(path):
    import scipy.io
    return scipy.io.loadmat(path)["uphi"]


def load_off_vertices(path):
    """Vertices of an OFF file (with or without faces; header 'OFF' + 'nv nf ne')."""
    with open(path) as f:
        head = f.readline().strip()
        if head != "OFF":
            if head.startswith("OFF"):                       # 'OFF nv nf ne' on one line
                counts = head[3:].split()
            else:
                raise ValueError(f"{path}: not an OFF file")
        else:
            counts = f.readline().split()
        nv = int(counts[0])
        return np.loadtxt(f, dtype=np.float32, max_rows=nv).reshape(nv, 3)


def load_vts(path):
    """Ground-truth landmark file (`*.vts`): one 1-based vertex index per line -> int64 [L]."""
    return np.loadtxt(path, dtype=np.int64).reshape(-1)


def save_cache(path, verts_list, used_shapes, fps_list, dist_list):
    """models/dataset.py:219-228."""
    torch.save((verts_list, used_shapes, fps_list, dist_list), path)


def load_cache(path):
    verts_list, used_shapes, fps_list, dist_list = torch.load(path, weights_only=False)
    return verts_list, used_shapes, fps_list, dist_list


# ------------------------------------------------------------------------------------------------
# evaluator
# ------------------------------------------------------------------------------------------------
def geodesic_error(phi_src, phi_tar, vts_src, vts_tar, M_tar, device=None, prec=None):
    """eval/main.m:27-38 for one (src, tar) pair.

    phi_src [Ns,C], phi_tar [Nt,C]: per-vertex features ('uphi'); vts_src, vts_tar [L]: 1-based landmark vertex ids in
    ground-truth correspondence; M_tar [Nt,Nt]: the target's (normalised) geodesic distance matrix.
    Returns (errors [L] = M_tar[nn(phi_src[vts_src]), vts_tar], idx [L] 0-based nearest target vertex)."""
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    ps = torch.as_tensor(_np(phi_src), dtype=torch.float32)
    pt = torch.as_tensor(_np(phi_tar), dtype=torch.float32)
    vs = torch.as_tensor(_np(vts_src).astype(np.int64)) - 1
    vt = torch.as_tensor(_np(vts_tar).astype(np.int64)) - 1
    q = ps[vs].to(device).unsqueeze(0).contiguous()
    idx = maps.knnsearch_t(q, pt.to(device).unsqueeze(0).contiguous(), prec=prec).reshape(-1)       # 0-based arg-min, int64
    if torch.is_tensor(M_tar):
        err = M_tar[idx.to(M_tar.device), vt.to(M_tar.device)]
    else:
        err = torch.from_numpy(np.asarray(M_tar)[idx.cpu().numpy(), vt.numpy()])
    return err.reshape(-1), idx


def evaluate_pairs(phis, vts, Ms, device=None, prec=None):
    """eval/main.m:19-42: all ordered pairs (src != tar) of a test set.

    phis[i] [N_i,C], vts[i] [L] (1-based), Ms[i] [N_i,N_i].  Returns (arr [S,S] mean error of src -> tar with zero
    diagonal, errors = concatenation of all per-landmark errors, avg = mean of the off-diagonal entries of arr)."""
    S = len(phis)
    arr = np.zeros((S, S), dtype=np.float64)
    errs = []
    for tar in range(S):
        for src in range(S):
            if src == tar:
                continue
            e, _ = geodesic_error(phis[src], phis[tar], vts[src], vts[tar], Ms[tar], device=device, prec=prec)
            e = e.double().cpu().numpy()
            errs.append(e)
            arr[src, tar] = e.mean()
    off = ~np.eye(S, dtype=bool)
    return arr, (np.concatenate(errs) if errs else np.zeros(0)), float(arr[off].mean()) if S > 1 else 0.0
