"""GPU parity at the BENCHMARKED configurations (BASELINE.json configs 2 and 5) and of the packed deformation path.

The oracle cannot sweep 50k x 50k on the CPU, so every problem is checked on a seeded 512-row sample against the fp64
arbiter restricted to those rows (`oracle.maps.exact_d2_rows_blocked` + `softmap_from_d2`: the semantics of
models/loss.py:91-114, 1339-1347, 1404-1409): arg-min and ordered top-10 `torch.equal` on every row whose deciding gaps
are resolvable in fp32, distances to 2e-6, weights / row sums / transferred coordinates to the stated f16 bound, and the
deform half of `pipeline.match_deform` (deformed coordinates, ARAP, both Chamfer terms) against `oracle.graph.dg_forward`
/ `oracle.geometry.chamfer_3d`.  Measured errors go to gpurun_out/parity_report.jsonl.
"""
import numpy as np
import pytest
import torch

from oracle import geometry as og
from oracle import graph as ogr
from oracle import maps as om

pytestmark = pytest.mark.gpu

# Stated bound on 16-bit-path soft-map weights / row sums / Pi.V.  The candidates (indices, distances,
# the 16 largest terms of every row sum) are exact fp32; what the f16 operands touch is the softmax mass OUTSIDE the 16 exact
# candidates: each of those terms carries the f16 rounding of its distance (relative error alpha * delta_d), so the bound is
# reached where few non-candidate terms dominate the remaining mass.  Measured maxima over the row samples (parity_report):
# 2.1e-3 at 50k alpha=10, 2.0e-3 at 200k alpha=100, 3.3e-4 at 50k alpha=100, <= 1.5e-3 at 20k.
F16_W_BOUND = 3e-3
RTOL = 1e-4               # north_star tolerance for fp32 quantities


def _report(name, **kv):
    import json, os
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_report.jsonl", "a") as f:
        f.write(json.dumps(dict(test=name, **kv)) + "\n")


def _sample_rows(n, count, seed):
    return torch.randperm(n, generator=torch.Generator().manual_seed(seed))[:count].sort().values


def check_softmap_rows(out, p, x, y, v, alphas, rows, w_bound, tag, **info):
    """Problem p of a batched SoftMapOut-like (attributes argmin, top_idx, top_w, top_d, row_sum, piv; `alphas` lists the
    alpha of each output in `out` (a list) -- one distance evaluation serves all of them)."""
    d2 = om.exact_d2_rows_blocked(x[rows], y, torch.float64)
    refs = om.softmap_from_d2(d2, list(alphas), v=v)
    worst = {}
    for o, alpha, s in zip(out, alphas, refs):
        d = s["d"]
        res1 = ~om.near_tie_rows(s["gap1"], d[..., 0], rel=2e-6)
        adj = torch.cat([d[..., 1:] - d[..., :-1], s["gap"][..., None]], -1).min(-1).values
        res = ~om.near_tie_rows(adj, d[..., -1], rel=2e-6)
        assert res.float().mean().item() > 0.95, "too many unresolvable rows in the sample"
        am = o.argmin[p].cpu()[rows]
        assert torch.equal(am[res1], s["argmin"][res1]), f"{tag}: arg-min differs from the fp64 arbiter"
        ti = o.top_idx[p].cpu()[rows].long()
        assert torch.equal(ti[res], s["idx"][res]), f"{tag}: ordered top-10 differs from the fp64 arbiter"
        td = o.top_d[p].cpu()[rows].double()
        derr = ((td - d).abs() / d.clamp_min(1e-30))[res].max().item()
        assert derr <= 2e-6, (tag, derr)
        M = y.shape[0]
        w_ref = torch.zeros(len(rows), M, dtype=torch.float64).scatter_(1, s["idx"], s["w"])
        w_got = torch.zeros(len(rows), M, dtype=torch.float64).scatter_(1, ti, o.top_w[p].cpu()[rows].double())
        sig = (w_ref > 1e-6) & res[:, None]
        werr = ((w_got - w_ref).abs() / w_ref.clamp_min(1e-12))[sig].max().item() if sig.any() else 0.0
        rerr = ((o.row_sum[p].cpu()[rows].double() - s["row_sum"]).abs() / s["row_sum"]).max().item()
        perr = 0.0
        if v is not None and o.piv is not None:
            perr = (o.piv[p].cpu()[rows].double() - s["piv"]).abs().max().item() / v.abs().max().item()
        _report(tag, problem=p, alpha=alpha, rows=len(rows), unresolvable_rows=int((~res).sum()), argmin_mismatch=0, top10_mismatch=0,
                d_rel_err=derr, w_rel_err=werr, rowsum_rel_err=rerr, piv_err=perr, **info)
        assert werr <= w_bound and rerr <= w_bound and perr <= w_bound, (tag, alpha, werr, rerr, perr)
        worst[alpha] = (werr, rerr, perr)
    return worst


@pytest.mark.parametrize("n,pairs,regime", [(20000, 1, "structured"), (50000, 2, "structured"), (50000, 1, "unstructured")])
def test_softmap_f16_at_benchmarked_sizes(n, pairs, regime):
    """ops.softmap_fwd(prec="f16") on the stacked 2B problems of `pairs` pairs, exactly as bench.py launches it."""
    from dv_matcher_b200 import ops, synthetic
    d = synthetic.make_batch(pairs, n, n, regime=regime)
    X = torch.cat([d["feat1"], d["feat2"]])
    Y = torch.cat([d["feat2"], d["feat1"]])
    V = torch.cat([d["xyz2"], d["xyz1"]])
    Xg, Yg, Vg = X.cuda(), Y.cuda(), V.cuda()
    alphas = (100.0, 10.0)
    outs = [ops.softmap_fwd(Xg, Yg, Vg, alpha=a, prec="f16", want_stats=True) for a in alphas]
    torch.cuda.synchronize()
    stats = [o.stats.cpu().tolist() for o in outs]
    for p in range(2 * pairs):
        rows = _sample_rows(n, 512 if p == 0 else 192, 100 + p)
        check_softmap_rows(outs, p, X[p], Y[p], V[p], alphas, rows, F16_W_BOUND, "softmap_f16_scale", n=n, regime=regime,
                           uncertified_rows=[s[0] for s in stats], fp32_pass_rows=[s[2] for s in stats])
    # hard map of the same problems: identical arg-min, bit for bit
    hard = ops.softmap_fwd(Xg, Yg, None, topk=1, soft=False, prec="f16")
    assert torch.equal(hard.argmin, outs[0].argmin)


def test_softmap_f16_exact_row_fallback_at_50k():
    """A batch in which one row survives both the certificate and the rescue scan and takes the exact per-row kernel (a cluster
    of CTAs per row): every row of the 16-bit path must agree with the fp32 path -- same ten indices, weights within the bound --
    and the fallback must cost microseconds, not a single-SM sweep of all 50 000 columns."""
    from dv_matcher_b200 import ops, synthetic
    d = synthetic.make_batch(2, 50000, 50000, first_pair=28)
    X = torch.cat([d["feat1"], d["feat2"]]).cuda()
    Y = torch.cat([d["feat2"], d["feat1"]]).cuda()
    V = torch.cat([d["xyz2"], d["xyz1"]]).cuda()
    out = ops.softmap_fwd(X, Y, V, alpha=100.0, prec="f16", want_stats=True)
    ref = ops.softmap_fwd(X, Y, V, alpha=100.0, prec="fp32")
    torch.cuda.synchronize()
    st = out.stats.cpu().tolist()
    assert st[2] >= 1, f"this batch no longer reaches the exact-row kernel (stats {st}): pick another one"
    assert torch.equal(out.argmin, ref.argmin) and torch.equal(out.top_idx, ref.top_idx)
    sig = ref.top_w > 1e-6
    werr = ((out.top_w - ref.top_w).abs() / ref.top_w.clamp_min(1e-12))[sig].max().item()
    perr = (out.piv - ref.piv).abs().max().item() / V.abs().max().item()
    for _ in range(2):
        ops.softmap_fwd(X, Y, V, alpha=100.0, prec="f16")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(5):
        ops.softmap_fwd(X, Y, V, alpha=100.0, prec="f16")
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 5
    _report("softmap_f16_exact_row_fallback", stats=st, w_rel_err_vs_fp32=werr, piv_err_vs_fp32=perr, ms_per_call=ms)
    assert werr <= F16_W_BOUND and perr <= F16_W_BOUND
    assert ms < 4.5, ms                     # 4.0 ms without a fallback row; 6.0 ms when one CTA swept the row


def test_softmap_f16_200k_one_problem():
    """Config 5's largest size: one 200k x 200k problem (the reference cannot hold its N x M matrix at all)."""
    from dv_matcher_b200 import ops, synthetic
    n = 200000
    d = synthetic.make_batch(1, n, n, regime="structured")
    X, Y, V = d["feat1"], d["feat2"], d["xyz2"]
    out = ops.softmap_fwd(X.cuda(), Y.cuda(), V.cuda(), alpha=100.0, prec="f16", want_stats=True)
    torch.cuda.synchronize()
    rows = _sample_rows(n, 512, 7)
    check_softmap_rows([out], 0, X[0], Y[0], V[0], (100.0,), rows, F16_W_BOUND, "softmap_f16_200k", n=n,
                       uncertified_rows=out.stats.cpu().tolist()[0])


def test_match_partial_shape():
    """Config 2 shape class at training size: N = 4995 source rows against M = 2200 target columns, both directions."""
    from dv_matcher_b200 import pipeline, synthetic
    d = synthetic.make_batch(2, 4995, 2200)
    g = {k: v.cuda() for k, v in d.items()}
    (sm12, sm21), (v12, v21) = pipeline.match(g["feat1"], g["feat2"], g["xyz1"], g["xyz2"], alpha=50.0)
    torch.cuda.synchronize()
    assert sm12.idx.shape == (2, 4995, 10) and sm21.idx.shape == (2, 2200, 10) and v12.shape == (2, 4995, 3) and v21.shape == (2, 2200, 3)

    class O:      # SoftMapOut-like views
        def __init__(self, sm, piv):
            self.argmin, self.top_idx, self.top_w, self.top_d, self.row_sum, self.piv = sm.argmin, sm.idx, sm.w, sm.top_d, sm.row_sum, piv

    for p in range(2):
        check_softmap_rows([O(sm12, v12)], p, d["feat1"][p], d["feat2"][p], d["xyz2"][p], (50.0,), _sample_rows(4995, 384, p),
                           F16_W_BOUND, "match_partial_12")
        check_softmap_rows([O(sm21, v21)], p, d["feat2"][p], d["feat1"][p], d["xyz1"][p], (50.0,), _sample_rows(2200, 384, 9 + p),
                           F16_W_BOUND, "match_partial_21")


def _deform_oracle(src, graphs, d9, p):
    """oracle.graph.dg_forward for problem p from the GPU-built graph and the Deformer output d9 [K,9] (models/loss.py:1257-1273)."""
    iden = torch.tensor([1.0, 0, 0, 0, 1, 0])
    R = og.rotation_6d_to_matrix(d9[..., 3:] + iden)[None]
    t = d9[None, :, :3]
    return ogr.dg_forward(src, graphs.nodes_idx[p].cpu(), graphs.influence[p].cpu(), graphs.weights[p].cpu(), graphs.ring[p].cpu(), R, t)


@pytest.mark.parametrize("n,pairs", [(20000, 1), (50000, 2)])
def test_match_deform_at_benchmarked_sizes(n, pairs):
    """pipeline.match_deform (what bench.py times): match half on row samples vs the fp64 arbiter, deform half vs the oracle."""
    from dv_matcher_b200 import pipeline, synthetic
    from dv_matcher_b200.deformation_graph import build_graphs
    from dv_matcher_b200.deformer import Deformer
    d = synthetic.make_batch(pairs, n, n)
    g = {k: v.cuda() for k, v in d.items()}
    torch.manual_seed(0)
    deformer = Deformer(10).cuda().eval()
    graphs = build_graphs(torch.cat([g["xyz1"], g["xyz2"]]), torch.arange(2 * pairs) % n)
    with torch.no_grad():
        out = pipeline.match_deform(g["feat1"], g["feat2"], g["xyz1"], g["xyz2"], graphs, deformer, alpha=100.0)
    torch.cuda.synchronize()
    src = torch.cat([d["xyz1"], d["xyz2"]])
    tgt = torch.cat([d["xyz2"], d["xyz1"]])
    fsrc = torch.cat([d["feat1"], d["feat2"]])
    ftgt = torch.cat([d["feat2"], d["feat1"]])

    # match half: indices exact, weights / Pi.V within the f16 bound (top_d / row_sum are not part of the step's results)
    for p in range(2 * pairs):
        rows = _sample_rows(n, 256, 40 + p)
        d2 = om.exact_d2_rows_blocked(fsrc[p][rows], ftgt[p], torch.float64)
        s = om.softmap_from_d2(d2, 100.0, v=tgt[p])
        dd = s["d"]
        res1 = ~om.near_tie_rows(s["gap1"], dd[..., 0], rel=2e-6)
        adj = torch.cat([dd[..., 1:] - dd[..., :-1], s["gap"][..., None]], -1).min(-1).values
        res = ~om.near_tie_rows(adj, dd[..., -1], rel=2e-6)
        assert torch.equal(out["T"][p].cpu()[rows][res1], s["argmin"][res1])
        assert torch.equal(out["top_idx"][p].cpu()[rows].long()[res], s["idx"][res])
        perr = (out["verts_t"][p].cpu()[rows].double() - s["piv"]).abs().max().item() / tgt[p].abs().max().item()
        assert perr <= F16_W_BOUND, perr
        _report("match_deform_scale", n=n, problem=p, piv_err=perr, unresolvable_rows=int((~res).sum()))
    # deform half
    d9 = graphs.pack.to_old_order(out["deformations"]).cpu()             # rows follow the packed (Morton) node order
    deformed = out["deformed"].cpu()
    for p in range(2 * pairs):
        warped, arap, _ = _deform_oracle(src[p], graphs, d9[p], p)
        scale = src[p].abs().max().item()
        derr = (deformed[p] - warped[0]).abs().max().item() / scale
        aerr = abs(out["arap"][p].item() - arap.item()) / max(abs(arap.item()), 1e-12)
        assert derr <= RTOL and aerr <= RTOL, (p, derr, aerr)
        # Chamfer terms of the step (parity unpinned vs upstream chamfer3D): the per-point outputs of the same
        # kernel on the same point sets reproduce the step's means, and a row sample of them equals the exact-form oracle
        if p == 0:
            from dv_matcher_b200 import ops
            tg = tgt[p:p + 1].cuda()
            errs = {}
            for name, pts in (("cd_deform", out["deformed"][p:p + 1]), ("cd_self", out["verts_t"][p:p + 1])):
                g1, g2, _, _ = ops.chamfer_fwd(pts, tg)
                assert torch.allclose((g1.mean(1) + g2.mean(1))[0], out[name][p], rtol=1e-6)      # same kernel; the mean is reduced in another batch shape
                rows = _sample_rows(n, 1024, 77)
                c1, _, i1, _ = og.chamfer_3d(pts.cpu()[:, rows], tgt[p:p + 1])
                _, c2, _, _ = og.chamfer_3d(pts.cpu(), tgt[p:p + 1][:, rows]) if n <= 20000 else (None, None, None, None)
                np.testing.assert_allclose(g1.cpu()[0, rows].numpy(), c1[0].numpy(), rtol=RTOL, atol=1e-12)
                if c2 is not None:
                    np.testing.assert_allclose(g2.cpu()[0, rows].numpy(), c2[0].numpy(), rtol=RTOL, atol=1e-12)
                errs[name] = float(((g1.cpu()[0, rows] - c1[0]).abs() / c1[0].clamp_min(1e-20)).max())
            _report("match_deform_scale_deform", n=n, problem=p, deformed_err=derr, arap_rel_err=aerr, chamfer_rows_rel_err=errs)
    # xyz 10-NN of the step against the exact-form oracle on a row sample
    rows = _sample_rows(n, 512, 3)
    ref_idx, _ = og.knn_exact(src[:1, rows], src[:1], 10)
    assert torch.equal(out["knn_self"][0].cpu()[rows], ref_idx[0])


def _same_step(res, ref, i):
    """Integer results are bit-identical run to run; the 16-bit-path softmax mass is summed in queue order, so float results
    that depend on it are reproducible to ~1e-6 relative."""
    for k in ("T", "top_idx"):
        assert torch.equal(res[k], ref[k]), (i, k)
    for k in ("top_w", "verts_t", "deformed", "arap", "cd_deform", "cd_self"):
        assert torch.allclose(res[k], ref[k], rtol=2e-5, atol=1e-7), (i, k, (res[k] - ref[k]).abs().max().item())


def test_engine_step_equals_match_deform_and_cuda_graph_replay():
    """MatchDeformEngine (host buffers in / out, CUDA-graph replay from the second use of a slot) returns exactly what
    pipeline.match_deform computes, step after step, with per-key graph caching and with per-step graph rebuilds."""
    from dv_matcher_b200 import pipeline, synthetic
    from dv_matcher_b200.deformation_graph import build_graphs
    from dv_matcher_b200.deformer import Deformer
    n, B = 3000, 2
    torch.manual_seed(0)
    deformer = Deformer(10).cuda().eval()
    eng = pipeline.MatchDeformEngine(deformer, alpha=100.0)
    batches = [synthetic.make_batch(B, n, n, first_pair=10 * q, pin=True) for q in range(3)]
    refs = []
    for q, h in enumerate(batches):
        g = {k: v.cuda() for k, v in h.items()}
        graphs = build_graphs(torch.cat([g["xyz1"], g["xyz2"]]), torch.arange(2 * B) + q)
        eng.put_graphs(q, graphs)
        with torch.no_grad():
            o = pipeline.match_deform(g["feat1"], g["feat2"], g["xyz1"], g["xyz2"], graphs, deformer, alpha=100.0)
        refs.append({k: o[k].cpu() for k in pipeline.RESULT_NAMES})
    tickets = []
    for i in range(9):                                   # 9 steps over 3 batches: both slots reach the captured-graph path
        q = i % 3
        h = batches[q]
        tickets.append((q, eng.submit(h["feat1"], h["feat2"], h["xyz1"], h["xyz2"], graph_key=q)))
        if len(tickets) == 2:
            qq, t = tickets.pop(0)
            _same_step(eng.result(t), refs[qq], i)
    qq, t = tickets.pop(0)
    _same_step(eng.result(t), refs[qq], 9)
    assert eng.launch_mode == "cuda_graph"
    # a key reused for clouds of another size must raise, not index out of bounds
    small = synthetic.make_batch(B, 1500, 1500, pin=True)
    with pytest.raises(RuntimeError):
        eng.step(small["feat1"], small["feat2"], small["xyz1"], small["xyz2"], graph_key=0)
    # default: graphs rebuilt from the step's own vertices
    r2 = eng.step(small["feat1"], small["feat2"], small["xyz1"], small["xyz2"])
    assert r2["deformed"].shape == (2 * B, 1500, 3) and torch.isfinite(r2["deformed"]).all()


def test_packed_deform_kernels_match_reference_layout_kernels(golden_graph):
    """dvm_node_table / dvm_node_table_from_d9 / dvm_skin_fwd_packed / dvm_arap_fwd_packed / dvm_skin_bwd_csr against the
    reference-layout kernels (dvm_skin_fwd, dvm_arap_fwd, dvm_skin_bwd, dvm_rot6d_fwd) and the oracle, batch of 3 clouds."""
    from dv_matcher_b200 import ops, synthetic
    from dv_matcher_b200.deformation_graph import build_graphs, deform_from_d9
    gen = torch.Generator().manual_seed(5)
    B, n = 3, 4100
    verts = torch.stack([synthetic.ellipsoid_cloud(n, gen) for _ in range(B)])
    vg = verts.cuda()
    graphs = build_graphs(vg, torch.tensor([1, 2, 3]))
    K = n // 2
    d9 = torch.stack([synthetic.random_rigid_field(K, gen) for _ in range(B)])
    d9g = d9.cuda()
    iden = torch.tensor([1.0, 0, 0, 0, 1, 0])
    R = ops.rot6d_fwd((d9g[..., 3:] + iden.cuda()).contiguous())
    t = d9g[..., :3].contiguous()
    # tables (R, t, d9 are in the reference's node order here: read through node_perm): fused == unfused, bit for bit
    pk = graphs.pack
    tb1 = ops.node_table(R, t, pk.nodes_xyz, node_perm=pk.node_perm)
    tb2, R2, t2 = ops.node_table_from_d9(d9g, pk.nodes_xyz, want_rt=True, node_perm=pk.node_perm)
    no = pk.node_perm.long()
    R_new = torch.gather(R, 1, no[..., None, None].expand(B, K, 3, 3))
    t_new = torch.gather(t, 1, no[..., None].expand(B, K, 3))
    assert torch.equal(tb1, tb2) and torch.equal(R_new, R2) and torch.equal(t_new, t2)
    assert torch.equal(tb1[..., :9].reshape(B, K, 3, 3), R_new) and torch.equal(tb1[..., 9:12], t_new) and torch.equal(tb1[..., 12:15], pk.nodes_xyz)
    assert torch.equal(pk.to_old_order(R2), R) and torch.equal(torch.gather(graphs.nodes_idx, 1, no), pk.nodes_idx_m)
    # inputs already in the packed order (what the pipeline does): same table
    tb3 = ops.node_table_from_d9(torch.gather(d9g, 1, no[..., None].expand(B, K, 9)).contiguous(), pk.nodes_xyz)
    assert torch.equal(tb3, tb1)
    # forward: packed vs reference-layout kernels vs oracle
    w_old = ops.skin_fwd(vg, graphs.nodes_idx, graphs.influence, graphs.weights, R, t)
    a_old, s_old = ops.arap_fwd(vg, graphs.nodes_idx, graphs.ring, R, t)
    w_new = ops.skin_fwd_packed(vg, graphs.pack, tb1)
    a_new, s_new = ops.arap_fwd_packed(graphs.pack, tb1, want_sr=True)
    a_nosr, none = ops.arap_fwd_packed(graphs.pack, tb1, want_sr=False)
    assert none is None and torch.equal(a_nosr, a_new)
    assert torch.equal(w_new, w_old)                       # same per-vertex arithmetic, only the visiting order differs
    np.testing.assert_allclose(a_new.cpu().numpy(), a_old.cpu().numpy(), rtol=2e-6)
    np.testing.assert_allclose(s_new.cpu().numpy(), s_old.cpu().numpy(), rtol=2e-6)
    for p in range(B):
        warped, arap, sr = _deform_oracle(verts[p], graphs, d9[p], p)
        assert (w_new[p].cpu() - warped[0]).abs().max().item() <= RTOL * verts[p].abs().max().item()
        assert abs(a_new[p].item() - arap.item()) <= RTOL * abs(arap.item())
        assert abs(s_new[p].item() - sr.item()) <= RTOL * abs(sr.item())
    wf, af, _ = deform_from_d9(vg, graphs, d9g, packed_order=False)
    assert torch.equal(wf, w_new) and torch.equal(af, a_new)
    # backward: node-major CSR kernel vs the atomic kernel (sum order differs) vs fp64
    go = torch.randn(B, n, 3, generator=gen)
    dR_old, dt_old = ops.skin_bwd(vg, graphs.nodes_idx, graphs.influence, graphs.weights, go.cuda())
    dR_new, dt_new = ops.skin_bwd_csr(vg, graphs.pack, go.cuda())
    dR_again, dt_again = ops.skin_bwd_csr(vg, graphs.pack, go.cuda())
    assert torch.equal(dR_new, dR_again) and torch.equal(dt_new, dt_again)          # deterministic
    dR_new, dt_new = pk.to_old_order(dR_new), pk.to_old_order(dt_new)               # packed -> reference node order
    infl = graphs.influence.cpu()
    wts = graphs.weights.cpu().double()
    nodes_xyz = pk.to_old_order(pk.nodes_xyz).cpu().double()
    for p in range(B):
        ref_t = torch.zeros(K, 3, dtype=torch.float64)
        ref_R = torch.zeros(K, 9, dtype=torch.float64)
        for k in range(3):
            gw = wts[p, :, k, None] * go[p].double()
            x = verts[p].double() - nodes_xyz[p][infl[p, :, k]]
            ref_t.index_add_(0, infl[p, :, k], gw)
            ref_R.index_add_(0, infl[p, :, k], (gw[:, :, None] * x[:, None, :]).reshape(n, 9))
        sc_t, sc_R = ref_t.abs().max().item(), ref_R.abs().max().item()
        assert (dt_new[p].cpu().double() - ref_t).abs().max().item() <= 1e-5 * sc_t
        assert (dR_new[p].cpu().double().reshape(K, 9) - ref_R).abs().max().item() <= 1e-5 * sc_R
        assert (dt_old[p].cpu().double() - ref_t).abs().max().item() <= 1e-5 * sc_t


def test_giant_pair_row_sharding_equals_unsharded():
    """SURVEY 8e, one giant pair: the row slabs four ranks would compute (Y replicated, no exchange) concatenate to the
    unsharded result -- indices bit for bit, weights within the stated 16-bit bound (nearly all to the summation order)."""
    from dv_matcher_b200 import distributed as dd, maps, synthetic
    n = 20000
    d = synthetic.make_batch(1, n, n)
    f1, f2, v2 = d["feat1"].cuda(), d["feat2"].cuda(), d["xyz2"].cuda()
    full, vt = maps.soft_map(f1, f2, 100.0, v=v2)
    parts = [dd.match_rows_sharded(f1, f2, v2, alpha=100.0, rank=r, world=4) for r in range(4)]
    assert [p["rows"] for p in parts] == [(0, 5000), (5000, 10000), (10000, 15000), (15000, 20000)]
    assert torch.equal(torch.cat([p["argmin"] for p in parts], 1), full.argmin)
    assert torch.equal(torch.cat([p["top_idx"] for p in parts], 1), full.idx)
    # weights: a row whose certificate fails is re-scored with the exact fp32 mass, and WHICH rows fail depends on the tiling
    # (a 5000-row slab is swept in a different order than 20000 rows), so a few rows may differ by the 16-bit bound itself;
    # all the others agree to the summation order of the mass
    w_cat = torch.cat([p["top_w"] for p in parts], 1)
    rel = ((w_cat - full.w).abs() / full.w.clamp_min(1e-12))[full.w > 1e-6]
    loose = (rel > 2e-5).float().mean().item()
    _report("giant_pair_sharding", w_rel_max=rel.max().item(), frac_rows_beyond_2e5=loose)
    assert rel.max().item() <= F16_W_BOUND and loose <= 0.01, (rel.max().item(), loose)
    assert torch.allclose(torch.cat([p["verts_t"] for p in parts], 1), vt, rtol=F16_W_BOUND, atol=F16_W_BOUND * v2.abs().max().item())


def test_captured_train_step_equals_eager():
    """training.CapturedTrainStep (forward + backward replayed as one CUDA graph from static buffers) against the eager step on
    two different batches, twice each: same loss tuple, same gradients, same host-RNG draws."""
    import gc
    import random
    from dv_matcher_b200 import synthetic, training
    from dv_matcher_b200.deformer import Deformer
    from dv_matcher_b200.deformation_graph import build_graphs
    from dv_matcher_b200.losses import GraphDeformLoss_Neural
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    n = 2000
    head = torch.nn.Linear(128, 128).to(dev)
    deformer = Deformer(10).to(dev)
    params = list(head.parameters()) + list(deformer.parameters())

    def mk():
        return GraphDeformLoss_Neural(k_deform=10, w_dist=0.02, w_map=0.005, k_dist=100, N_dist=200, partial=False, w_deform=0.5,
                                      w_img=0, w_rank=0, w_self_rec=0.5, w_cd=0.1, w_arap=0.01, save_name="t")

    batches, graphs = [], []
    for q in range(2):
        d = {k: v.to(dev) for k, v in synthetic.make_batch(2, n, n, first_pair=2 * q).items()}
        d["dist1"], d["dist2"] = torch.cdist(d["xyz1"], d["xyz1"]), torch.cdist(d["xyz2"], d["xyz2"])
        batches.append({k: d[k] for k in ("feat1", "feat2", "dist1", "dist2", "xyz1", "xyz2")})
        z = torch.zeros(2, dtype=torch.int64, device=dev)
        graphs.append((build_graphs(d["xyz1"], z), build_graphs(d["xyz2"], z)))
    crit = mk()
    eager = []
    for q in range(2):
        random.seed(5 + q)
        for p in params:
            p.grad = None
        crit.static_graphs = [(g.nodes_idx.float(), g) for g in graphs[q]]
        d = batches[q]
        out = crit(head(d["feat1"]), head(d["feat2"]), d["dist1"], d["dist2"], d["xyz1"], d["xyz2"], 100.0, deformer)
        out[0].backward()
        eager.append((torch.stack([o.detach() for o in out]).clone(), torch.cat([p.grad.reshape(-1) for p in params]).clone()))
    del out, crit
    gc.collect()                       # no eager autograd graph may stay alive (its AccumulateGrad nodes pin the default stream)
    step = training.CapturedTrainStep(mk(), lambda a, b: (head(a), head(b)), deformer, params, 100.0)
    for rep in range(2):
        for q in range(2):
            random.seed(5 + q)
            out = step(batches[q], graphs[q])
            torch.cuda.synchronize()
            got = torch.stack(list(out))
            g = torch.cat([p.grad.reshape(-1) for p in params])
            le, ge = eager[q]
            assert torch.allclose(got, le, rtol=1e-5), (got, le)
            rel = float((g - ge).abs().max() / ge.abs().max())
            _report("captured_train_step", rep=rep, batch=q, grad_rel_diff=rel, launches_per_step=step.launches_per_step)
            assert rel <= 1e-4, rel
    assert step.launches_per_step > 50
