"""Drop the B200 hot path into an unmodified DV-Matcher checkout.

    import dv_matcher_b200.install as dvm_install; dvm_install.install()

Rebinds the hot-path names in whichever reference modules are importable / already imported:
`models.loss`, `models.model`, `lib.deformation_graph_point`, `lib.deformation_graph`, and the script modules
that re-define the map functions locally (`test`, `test_partial`, `deform`, `__main__`).  Nothing is patched that
this package does not implement; `uninstall()` restores the originals.
"""
import importlib
import sys

from . import deformation_graph, deformer, geometry, lgnet, losses, maps, secondary

_saved = []

_LOSS_NAMES = dict(
    knnsearch_t=maps.knnsearch_t, search_t=maps.search_t, knnsearch_t_grad=maps.knnsearch_t_grad,
    knn_grad=geometry.knn_grad, knn=geometry.knn, index_points=geometry.index_points, index_points_idx=geometry.index_points_idx,
    rotation_6d_to_matrix=geometry.rotation_6d_to_matrix, dist_chamfer_3D=geometry.dist_chamfer_3D,
    DeformationGraph_geod=deformation_graph.DeformationGraph_geod,
    GraphDeformLoss_Neural=losses.GraphDeformLoss_Neural, GraphDeformLoss_Neural_Partial=losses.GraphDeformLoss_Neural_Partial,
    FrobeniusLoss=losses.FrobeniusLoss,
)
# the entry scripts return 1-based maps from their local copies (test.py:19-28, test_partial.py:170-179, deform.py:86-95)
_SCRIPT_NAMES = dict(
    knnsearch_t=maps.knnsearch_t_1based, search_t=maps.search_t_1based, knnsearch_t_grad=maps.knnsearch_t_grad,
    topk_pi=maps.topk_pi, knn_grad=geometry.knn_grad, index_points=geometry.index_points,
    rotation_6d_to_matrix=geometry.rotation_6d_to_matrix, deformation_graph_node=deformation_graph.deformation_graph_node_list,
    DeformationGraph_geod=deformation_graph.DeformationGraph_geod, Deformer=deformer.Deformer,
    GraphDeformLoss_Neural=losses.GraphDeformLoss_Neural, GraphDeformLoss_Neural_Partial=losses.GraphDeformLoss_Neural_Partial,
    # secondary (cosine / top-40) API of test_partial.py:73-144
    forward_source_target=secondary.forward_source_target, forward_shape=secondary.forward_shape,
    cross_construct=secondary.cross_construct, reconstruction=secondary.reconstruction,
)
_TABLE = {
    "models.loss": _LOSS_NAMES,
    # LG-Net itself stays the reference's torch code, but its seven N x N k-NN calls per forward (knn_new, k = 40,
    # models/model.py:267-278) and their gathers (index_points, :255-265) are the same operators as the loss's
    "models.model": dict(Deformer=deformer.Deformer, knn_new=geometry.knn, index_points=geometry.index_points),
    "lib.deformation_graph_point": dict(DeformationGraph_geod=deformation_graph.DeformationGraph_geod,
                                        farthest_point_sample=deformation_graph.farthest_point_sample),
    "lib.deformation_graph": dict(DeformationGraph=deformation_graph.DeformationGraph),
    "test": _SCRIPT_NAMES, "test_partial": _SCRIPT_NAMES, "deform": _SCRIPT_NAMES, "train": _SCRIPT_NAMES,
    "train_partial": _SCRIPT_NAMES, "__main__": _SCRIPT_NAMES,
}


def _patch(mod, names):
    n = 0
    for name, obj in names.items():
        if hasattr(mod, name):
            _saved.append((mod, name, getattr(mod, name)))
            setattr(mod, name, obj)
            n += 1
    return n


def install(import_missing=True, verbose=False):
    """Returns {module name: number of names rebound}.  `import_missing` tries to import the library modules of the
    reference (`models.loss`, ...) if they are on sys.path but not imported yet; script modules are only patched
    when already loaded (patch again after importing them, or call install() from the script itself)."""
    done = {}
    for modname, names in _TABLE.items():
        mod = sys.modules.get(modname)
        if mod is None and import_missing and "." in modname:
            try:
                mod = importlib.import_module(modname)
            except Exception:
                mod = None
        if mod is None or mod.__name__.startswith("dv_matcher_b200"):
            continue
        c = _patch(mod, names)
        if modname == "models.model" and hasattr(mod, "SA_Layer"):        # LG-Net's global attention without the N x N matrix
            c += _patch(mod.SA_Layer, dict(forward=lgnet.sa_layer_forward))
        if c:
            done[modname] = c
    if verbose:
        print("dv_matcher_b200.install:", done)
    return done


def uninstall():
    while _saved:
        mod, name, obj = _saved.pop()
        setattr(mod, name, obj)
