"""The oracle restatement (oracle/preshape_oracle.py) against the golden vectors
captured from the unmodified reference (tests/golden/make_golden.py), plus — when
/root/reference is present (build container) — a live cross-check against the
reference itself.  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import preshape_oracle as po
from oracle import ref_shim
from tests.golden_cases import LARGE_CASES, SMALL_CASES, load_case

FLOAT_TOL = 2e-5   # oracle vs golden floats: same ops, possibly different CPU/oneDNN kernels


def _run(name):
    """Single-threaded like the golden generation: oneDNN's conv accumulates in a
    thread-count-dependent order, and 1e-6 centre differences flip FPS arg-max
    decisions on degenerate (collapsed-grid) scenes — SURVEY.md §7 H2."""
    cfg, sd, pts, text_dict, img, g = load_case(name)
    trace = {}
    nt = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        out = _fwd(cfg, sd, pts, text_dict, img, trace)
    finally:
        torch.set_num_threads(nt)
    return cfg, pts, out, trace, g


def _fwd(cfg, sd, pts, text_dict, img, trace):
    return po.forward(sd, pts, text_dict, img, grid_size=cfg.grid_size, dynamic_drop_radio=cfg.dynamic_drop_radio,
                     text_blocks=cfg.text_blocks, img_blocks=cfg.img_blocks, num_sub=cfg.num_sub,
                     num_heads=cfg.num_heads, trace=trace)


@pytest.mark.parametrize("name", SMALL_CASES + LARGE_CASES)
def test_oracle_matches_reference_golden(name):
    cfg, pts, out, tr, g = _run(name)
    # indices: bit-exact
    np.testing.assert_array_equal(tr["idx2"].numpy(), g["idx2"])
    np.testing.assert_array_equal(tr["kept_idx"].numpy(), g["kept_idx"])
    np.testing.assert_array_equal(tr["drop_idx"].numpy(), g["drop_idx"])
    np.testing.assert_array_equal(np.array([o.shape[0] for o in out]), g["out_counts"])
    for k in ("centres", "kept_centres", "point_proxy", "img_proxy", "translate", "transform"):
        np.testing.assert_allclose(tr[k].numpy(), g[k], rtol=0, atol=FLOAT_TOL, err_msg=k)
    P = torch.stack(pts, 0)
    for b, o in enumerate(out):
        ch = (tr["scattered"][b] != P[b]).any(-1).nonzero(as_tuple=True)[0].numpy()
        np.testing.assert_array_equal(ch, g[f"changed_rows_{b}"])
        np.testing.assert_allclose(tr["scattered"][b][ch].numpy(), g[f"changed_vals_{b}"], rtol=0, atol=1e-4)
        if f"out_{b}" in g:
            np.testing.assert_allclose(o.numpy(), g[f"out_{b}"], rtol=0, atol=1e-4)
        else:
            np.testing.assert_allclose(o[:2048].numpy(), g[f"out_head_{b}"], rtol=0, atol=1e-4)
            np.testing.assert_allclose(o[::97].numpy(), g[f"out_stride_{b}"], rtol=0, atol=1e-4)
        np.testing.assert_allclose(o.double().sum(0).numpy(), g["out_sum"][b], rtol=1e-7, atol=1e-2)


def test_golden_cases_cover_edge_conditions():
    """The fixtures must actually contain the edge cases they are named for."""
    g = load_case("c1_sparse")[5]
    assert (g["kept_idx"] == -1).any() and (g["drop_idx"] == -1).any()
    g = load_case("c1_very_sparse")[5]
    assert ((g["kept_idx"] == -1).all(-1)).any(), "expected at least one all-padding kept cluster"
    g = load_case("c1_dups")[5]
    c = g["kept_centres"][0]
    assert len(np.unique(c, axis=0)) < len(c), "expected duplicated clamped centres"
    cfg, sd, pts, *_ = load_case("c1_collapsed")
    ext = pts[0].max(0)[0] - pts[0].min(0)[0]
    assert (ext < 8).all()
    pts = load_case("c1_origin")[2]
    assert (pts[0] == 0).all(-1).any()


def test_ball_query_c_vs_torch_restatement():
    g = torch.Generator().manual_seed(5)
    for trial in range(4):
        p2 = torch.rand(2, 3000, 3, generator=g) * 10
        p1 = torch.rand(2, 40, 3, generator=g) * 10
        r = 1.0 + trial
        i1, k1 = po.ball_query(p1, p2, 30, r)
        i2, k2 = po.ball_query_torch(p1, p2, 30, r)
        assert torch.equal(i1, i2) and torch.equal(k1, k2)


def test_ball_query_empty_and_full():
    p2 = torch.zeros(1, 50, 3)
    p2[0, :, 0] = torch.arange(50) * 0.01
    idx, knn = po.ball_query(torch.tensor([[[100.0, 0, 0], [0.0, 0, 0]]]), p2, 30, 3.0)
    assert (idx[0, 0] == -1).all() and (knn[0, 0] == 0).all()
    assert idx[0, 1].tolist() == list(range(30))
    # strictness of d2 < r^2
    idx, _ = po.ball_query(torch.tensor([[[3.0, 0, 0]]]), torch.tensor([[[0.0, 0, 0], [0.5, 0, 0]]]), 4, 3.0)
    assert idx[0, 0].tolist() == [1, -1, -1, -1]


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted (GPU box)")
def test_fps_matches_reference_in_tree_naive_copy():
    ref = ref_shim.load()
    g = torch.Generator().manual_seed(3)
    pts = torch.rand(3, 200, 3, generator=g)
    pts[1, 50:] = pts[1, 10]          # duplicates -> zero distances -> repeated index 0
    pts[2] = pts[2, 0]                # fully degenerate
    want = ref.sample_farthest_points_naive(pts, K=60)[1]
    got = po.farthest_point_indices(pts, 60)
    assert torch.equal(want, got)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted (GPU box)")
def test_state_dict_spec_matches_reference_module():
    from proxytransformation_b200 import synthetic as syn
    for cfg in (syn.C1, syn.C3):
        net = ref_shim.build_module(cfg.module_kwargs(), None)
        ref_sd = net.state_dict()
        spec = syn.state_dict_spec(cfg)
        assert [k for k, _, _ in spec] == list(ref_sd.keys())
        for k, shape, _ in spec:
            assert tuple(ref_sd[k].shape) == tuple(shape), k


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted (GPU box)")
def test_oracle_live_against_reference_faithful_cost():
    """faithful_cost (all blocks, all 226 queries — what the CPU baseline times) gives the same output."""
    cfg, sd, pts, text_dict, img, g = load_case("c1_blocks3")
    out = po.forward(sd, pts, text_dict, img, grid_size=cfg.grid_size, dynamic_drop_radio=cfg.dynamic_drop_radio,
                     text_blocks=cfg.text_blocks, img_blocks=cfg.img_blocks, num_sub=cfg.num_sub, faithful_cost=True)
    net = ref_shim.build_module(cfg.module_kwargs(), sd)
    want = ref_shim.run_reference(net, pts, text_dict, img)
    for a, b in zip(out, want):
        assert a.shape == b.shape
        assert (a - b).abs().max().item() < 1e-5


def test_sparse_collate_restatement_matches_the_torch_expression():
    """N1 oracle pin: the caller evaluates ``p[:, :3] / voxel_size`` with torch and MinkowskiEngine assigns it into an int32
    tensor; the restatement must agree with exactly that expression (CPU semantics), batch column and ordering included."""
    g = torch.Generator().manual_seed(11)
    pts = [(torch.rand(n, 3, generator=g) - 0.4) * 30.0 for n in (7, 0, 129)]
    coords, feats = po.batch_sparse_collate(pts, 0.01)
    s = 0
    for b, p in enumerate(pts):
        ref = torch.zeros(len(p), 4, dtype=torch.int32)
        ref[:, 1:] = p[:, :3] / 0.01
        ref[:, 0] = b
        assert torch.equal(coords[s:s + len(p)], ref)
        assert torch.equal(feats[s:s + len(p)], p)
        s += len(p)
    assert s == len(coords)
    # truncation, not floor, for negative quotients; floor only on request
    c2, _ = po.batch_sparse_collate([torch.tensor([[-0.015, 0.015, -0.0]])], 0.01)
    assert c2[0].tolist() == [0, -1, 1, 0]
    c3, _ = po.batch_sparse_collate([torch.tensor([[-0.015, 0.015, -0.0]])], 0.01, floor=True)
    assert c3[0].tolist() == [0, -2, 1, 0]


def test_aggregate_sample_restatement_is_the_inverse_transform():
    """N3 oracle sanity: solving extrinsic . x = [p;1] moves ego points back to where the global points were."""
    g = torch.Generator().manual_seed(8)
    glob = [torch.rand(50, 3, generator=g) * 8 for _ in range(3)]
    ext = torch.eye(4).repeat(3, 1, 1)
    for v in range(3):
        q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
        ext[v, :3, :3], ext[v, :3, 3] = q, torch.randn(3, generator=g)
    ego = [(ext[v, :3, :3] @ glob[v].T).T + ext[v, :3, 3] for v in range(3)]
    choices = torch.tensor([149, 0, 75, 75, 3])
    got = po.aggregate_sample(ego, ext, choices)
    assert torch.allclose(got, torch.cat(glob)[choices], atol=1e-5)


N3_CASES = ("xyz", "xyzrgb", "replace")


def load_n3_case(name):
    """tests/golden/n3_aggregate.npz (make_golden_n3.py: the UNMODIFIED reference transforms): views, extrinsics, choices, result."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "n3_aggregate.npz"))
    sizes = z[f"{name}.sizes"].tolist()
    views = list(torch.from_numpy(z[f"{name}.views"]).split(sizes))
    return views, torch.from_numpy(z[f"{name}.extrinsics"]), torch.from_numpy(z[f"{name}.choices"]), z[f"{name}.aggregated"], z[f"{name}.sampled"]


@pytest.mark.parametrize("name", N3_CASES)
def test_aggregate_sample_pinned_to_the_reference_transforms(name):
    """N3 pin: oracle.aggregate_sample against what the reference's own AggregateMultiViewPoints.transform
    (datasets/transforms/multiview.py:224-251) + PointSample._points_random_sampling (points.py:373-419) produced on the same
    views / extrinsics / np.random choices (xyz points, xyz+rgb points, sampling with replacement).  Same torch ops in the same
    order: bit-exact."""
    views, ext, choices, aggregated, sampled = load_n3_case(name)
    got = po.aggregate_sample(views, ext, choices)
    assert np.array_equal(got.numpy(), sampled[:, :3])
    everything = po.aggregate_sample(views, ext, torch.arange(aggregated.shape[0]))
    assert np.array_equal(everything.numpy(), aggregated[:, :3])
