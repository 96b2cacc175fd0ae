"""CPU-only tests of the host-side logic: install() name rebinding, pair sharding and the one-bucket gradient
all-reduce over gloo with world_size = 2 (the N > 1 path of SURVEY 8e), sparse-map tensor protocol."""
import os
import sys
import types

import pytest
import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_install_rebinds_and_restores():
    from dv_matcher_b200 import install as inst, maps, geometry, losses
    fake = types.ModuleType("models.loss")
    for n in ("knnsearch_t", "knnsearch_t_grad", "knn_grad", "index_points", "GraphDeformLoss_Neural", "dist_chamfer_3D"):
        setattr(fake, n, object())
    fake.unrelated = 42
    script = types.ModuleType("deform")
    script.knnsearch_t = script.topk_pi = object()
    pkg = types.ModuleType("models")
    saved = {k: sys.modules.get(k) for k in ("models", "models.loss", "deform")}
    sys.modules.update({"models": pkg, "models.loss": fake, "deform": script})
    try:
        done = inst.install(import_missing=False)
        assert done["models.loss"] == 6 and done["deform"] == 2
        assert fake.knnsearch_t is maps.knnsearch_t and fake.knn_grad is geometry.knn_grad
        assert fake.GraphDeformLoss_Neural is losses.GraphDeformLoss_Neural and fake.unrelated == 42
        assert script.knnsearch_t is maps.knnsearch_t_1based and script.topk_pi is maps.topk_pi
        inst.uninstall()
        assert not (fake.knnsearch_t is maps.knnsearch_t)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_shard_pairs_partition():
    from dv_matcher_b200.distributed import shard_pairs
    for n in (0, 1, 7, 16):
        for world in (1, 2, 8):
            got = sorted(i for r in range(world) for i in shard_pairs(n, r, world))
            assert got == list(range(n))


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from dv_matcher_b200 import distributed as dd
    r, w, dev = dd.init(backend="gloo")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ELU(), torch.nn.Linear(5, 3))
    data = torch.arange(4 * 6, dtype=torch.float32).reshape(4, 6) / 10.0
    mine = data[dd.shard_pairs(4, r, w)]
    net(mine).pow(2).sum().backward()
    frozen = list(net.parameters())[-1]
    if r == 1:
        frozen.grad = None                           # a rank without a gradient for one tensor contributes zeros
    bucket = dd.allreduce_gradients(list(net.parameters()), world=w)
    res = [p.grad.clone() for p in net.parameters()]
    gathered = dd.gather_pair_results({"rank": r, "pairs": dd.shard_pairs(4, r, w)}, world=w)
    if r == 0:
        torch.save(dict(grads=res, n=bucket.numel(), gathered=gathered), out)
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_gloo_world2(tmp_path):
    out = str(tmp_path / "res.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    # single-process reference: mean over the two shards' gradient sums
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ELU(), torch.nn.Linear(5, 3))
    data = torch.arange(4 * 6, dtype=torch.float32).reshape(4, 6) / 10.0
    grads = []
    for r in range(2):
        net.zero_grad()
        net(data[r::2]).pow(2).sum().backward()
        gs = [p.grad.clone() for p in net.parameters()]
        if r == 1:
            gs[-1] = torch.zeros_like(gs[-1])
        grads.append(gs)
    for a, b0, b1 in zip(got["grads"], grads[0], grads[1]):
        assert torch.allclose(a, (b0 + b1) / 2, rtol=1e-6, atol=1e-7)
    assert got["n"] == sum(p.numel() for p in net.parameters())
    assert [g["pairs"] for g in got["gathered"]] == [[0, 2], [1, 3]]


def test_shard_rows_partition():
    from dv_matcher_b200.distributed import shard_rows
    for n in (1, 7, 200000, 199999):
        for world in (1, 2, 3, 8):
            spans = [shard_rows(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _giant_worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from dv_matcher_b200 import distributed as dd
    from dv_matcher_b200.maps import SparseSoftMap
    from oracle import maps as om
    dd.init(backend="gloo")

    def cpu_soft_map(x, y, alpha, v=None, prec=None):          # the oracle stands in for the CUDA kernel in this host-logic test
        s = om.softmap_sparse(x, y, alpha, v=v, dtype=torch.float32)
        return SparseSoftMap(s["idx"].int(), s["w"], y.shape[1], s["argmin"]), s["piv"]

    g = torch.Generator().manual_seed(0)
    f1, f2, v2 = torch.randn(1, 203, 16, generator=g), torch.randn(1, 150, 16, generator=g), torch.randn(1, 150, 3, generator=g)
    res = dd.match_rows_sharded(f1, f2, v2, alpha=5.0, gather=True, soft_map_fn=cpu_soft_map)
    if rank == 0:
        torch.save({k: v for k, v in res.items()}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_giant_pair_row_sharding_gloo_world2(tmp_path):
    """Rows shard, columns replicate: the gathered result of 2 ranks equals the single-process result (no data-path exchange)."""
    from oracle import maps as om
    out = str(tmp_path / "giant.pt")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_giant_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    g = torch.Generator().manual_seed(0)
    f1, f2, v2 = torch.randn(1, 203, 16, generator=g), torch.randn(1, 150, 16, generator=g), torch.randn(1, 150, 3, generator=g)
    s = om.softmap_sparse(f1, f2, 5.0, v=v2, dtype=torch.float32)
    assert got["rows"] == (0, 203)
    assert torch.equal(got["argmin"], s["argmin"]) and torch.equal(got["top_idx"].long(), s["idx"])
    assert torch.allclose(got["top_w"], s["w"]) and torch.allclose(got["verts_t"], s["piv"])


def test_sparse_soft_map_protocol_cpu():
    """SparseSoftMap's dense view and torch-function interception (no kernels: dense fallback for foreign shapes)."""
    from dv_matcher_b200.maps import SparseSoftMap, topk_pi
    a = torch.rand(2, 5, 12)
    sm = topk_pi(a, k=3)
    assert isinstance(sm, SparseSoftMap) and tuple(sm.shape) == (2, 5, 12)
    dense = sm.to_dense()
    vals, idx = torch.topk(a, 3, dim=-1)
    assert torch.equal(dense, torch.zeros_like(a).scatter(-1, idx, vals))
    assert torch.equal(sm.transpose(1, 2), dense.transpose(1, 2)) and sm.dim() == 3 and sm.size(2) == 12


# ------------------------------------------------------------------------------------------------
# wire / on-disk formats (row f4)
# ------------------------------------------------------------------------------------------------
def test_map_txt_and_feature_mat_formats_match_the_reference_writers(tmp_path):
    """Same bytes as test.py:111-133 (np.savetxt fmt '%i' of the 1-based [N,1] map; savemat key 'uphi')."""
    import scipy.io
    from dv_matcher_b200 import evalio
    g = np.random.default_rng(0)
    t12 = torch.from_numpy(g.integers(1, 500, size=(1, 500, 1)))
    t21 = torch.from_numpy(g.integers(1, 500, size=(1, 480, 1)))
    p12, p21 = evalio.save_maps_txt(str(tmp_path), "mesh052", "mesh053", t12, t21)
    assert p12.endswith("T/T_mesh052_mesh053.txt") and p21.endswith("T/T_mesh053_mesh052.txt")
    ref = tmp_path / "ref.txt"
    np.savetxt(ref, t12.detach().cpu().squeeze(0).numpy(), fmt="%i")            # the reference's own three lines
    assert open(p12, "rb").read() == open(ref, "rb").read()
    assert np.array_equal(evalio.load_map_txt(p21), t21.reshape(-1).numpy())
    feat = torch.randn(1, 500, 128)
    pf = evalio.save_features_mat(str(tmp_path), "mesh052", feat)
    assert pf.endswith("feature/usefeature_mesh052.mat")
    m = scipy.io.loadmat(pf)
    assert m["uphi"].shape == (500, 128) and np.array_equal(m["uphi"], feat[0].numpy())
    assert np.array_equal(evalio.load_features_mat(pf), feat[0].numpy())


def test_off_vts_and_cache_round_trip(tmp_path):
    from dv_matcher_b200 import evalio
    pts = np.random.default_rng(1).standard_normal((37, 3)).astype(np.float32)
    p = tmp_path / "a.off"
    evalio.save_off_file(str(p), pts)
    lines = open(p).read().split("\n")
    assert lines[0] == "OFF" and lines[1] == "37 0 0" and lines[2] == f"{pts[0][0]} {pts[0][1]} {pts[0][2]}"    # deform.py:79-84
    assert np.allclose(evalio.load_off_vertices(str(p)), pts, rtol=0, atol=1e-6)
    (tmp_path / "m.vts").write_text("3\n1\n37\n")
    assert evalio.load_vts(str(tmp_path / "m.vts")).tolist() == [3, 1, 37]
    c = tmp_path / "cache.pt"
    evalio.save_cache(str(c), [torch.from_numpy(pts)], ["a"], [torch.arange(5)], [torch.zeros(37, 37, dtype=torch.float64)])
    v, names, fps, dist = evalio.load_cache(str(c))
    assert names == ["a"] and torch.equal(v[0], torch.from_numpy(pts)) and fps[0].tolist() == [0, 1, 2, 3, 4] and dist[0].dtype == torch.float64


# ------------------------------------------------------------------------------------------------
# bench.py contract pieces that run without a GPU
# ------------------------------------------------------------------------------------------------
def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` = the oracle port timed on the host cores: one JSON line with the reference-arm keys."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--n", "600", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("shape pairs/sec") and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_bench_product_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: without a CUDA device the product arm prints why and measures nothing."""
    import subprocess
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=root)
    assert "no CPU fallback" in (out.stdout + out.stderr)
    assert not any(l.startswith("{") for l in out.stdout.splitlines())


def test_install_patches_lgnet_attention_and_cpu_path_matches_reference():
    """install() rebinds LG-Net's N x N pieces: `models.model.knn_new`, `index_points` and `SA_Layer.forward`.
    On CPU tensors the patched forward runs the reference's dense formula: its output on the reference layer's weights equals the
    unmodified reference's output (tests/golden/ref_lgnet.npz)."""
    from dv_matcher_b200 import install as inst, geometry, lgnet

    class SA_Layer(torch.nn.Module):                                   # the reference layer's modules (models/model.py:98-111)
        def __init__(self, ch=128):
            super().__init__()
            self.q_conv = torch.nn.Conv1d(ch, ch // 4, 1, bias=False)
            self.k_conv = torch.nn.Conv1d(ch, ch // 4, 1, bias=False)
            self.v_conv = torch.nn.Conv1d(ch, ch, 1)
            self.trans_conv = torch.nn.Conv1d(ch, ch, 1)
            self.after_norm = torch.nn.BatchNorm1d(ch)
            self.act = torch.nn.ReLU()
            self.softmax = torch.nn.Softmax(dim=-1)

        def forward(self, x):
            raise AssertionError("the original forward must have been replaced")

    fake = types.ModuleType("models.model")
    fake.SA_Layer, fake.knn_new, fake.index_points, fake.Deformer = SA_Layer, object(), object(), object()
    pkg = types.ModuleType("models")
    saved = {k: sys.modules.get(k) for k in ("models", "models.model")}
    sys.modules.update({"models": pkg, "models.model": fake})
    try:
        done = inst.install(import_missing=False)
        assert done["models.model"] == 4
        assert fake.knn_new is geometry.knn and fake.index_points is geometry.index_points
        assert SA_Layer.forward is lgnet.sa_layer_forward
        z = np.load(os.path.join(ROOT, "tests", "golden", "ref_lgnet.npz"))
        sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd_") and not k.startswith(("sd_bn1", "sd_conv1"))}
        layer = SA_Layer()
        layer.load_state_dict(sd, strict=True)
        layer.eval()
        with torch.no_grad():
            y = layer(torch.from_numpy(z["x"]))
        assert torch.allclose(y, torch.from_numpy(z["y"]), rtol=0, atol=1e-6)
        inst.uninstall()
        assert SA_Layer.forward is not lgnet.sa_layer_forward
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
