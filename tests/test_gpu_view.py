"""GPU parity of the keyframe association pipeline (K6 renderView, K7 searchCorrespondence, K8
checkMapAssociation) through the C ABI against the CPU oracle.  Bars: every integer output (view membership
and order, candidate lists, associated component) identical; fp64 values bit-identical except through
log() in the 2-D Bhattacharyya distance, which only feeds comparisons."""
import numpy as np
import pytest

from gmmloc_b200 import synth

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def _same(a, b):
    """bit-identical where finite, NaN in the same places (the NaN payload / sign is not part of the contract)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(_bits(a[~na]), _bits(b[~nb]))


def _kf(mean, cov, B, N, seed, poses=None):
    fr = synth.make_frames(mean, cov, B, N, seed=seed, poses_wc=poses)
    fr["depth"] = (synth.BF / (fr["obs"][..., 0] - fr["obs"][..., 2])).astype(np.float32)
    return fr


@pytest.fixture(scope="module")
def v1(gpu, oracle, golden_maps):
    g, _ = gpu
    mean, cov = golden_maps["v1"]["mean"], golden_maps["v1"]["cov"]
    m = g.GMM(mean, cov)
    n = m.buildNeighbors(2.5)  # map/neighbor_dist_thresh, v1.yaml:27
    comps = oracle.map_build(mean, cov)
    nbr = oracle.neighbors(comps, 2.5)
    assert n == nbr[0][-1]
    return m, comps, nbr, mean, cov


def test_render_views_v1_03(gpu, oracle, v1, gt_v1_03):
    g, ctx = gpu
    m, comps, _, mean, cov = v1
    rows = np.arange(20, 2149, 97)
    fr = _kf(mean, cov, len(rows), 4, 1, gt_v1_03[rows, 1:8])
    out = ctx.renderView(m, fr["pose0"])
    assert np.all(out["status"] == 0)
    for b, pose in enumerate(fr["pose0"]):
        view = oracle.render_view(comps, pose)
        V = len(view)
        assert out["count"][b] == V, b
        assert np.array_equal(out["ids"][b, :V], view["id"]), b
        assert np.all(out["ids"][b, V:] == -1)
        assert np.array_equal(_bits(out["mean"][b, :V]), _bits(view["mean"]))
        assert np.array_equal(_bits(out["cov"][b, :V]), _bits(view["cov"]))
        assert np.array_equal(_bits(out["depth"][b, :V]), _bits(view["depth"]))
    assert out["count"].min() > 20 and out["count"].max() > 250


def test_render_views_synthetic_dense(gpu, oracle):
    g, ctx = gpu
    mean, cov = synth.make_map(6000, 0.7, seed=21)
    m = g.GMM(mean, cov)
    comps = oracle.map_build(mean, cov)
    fr = _kf(mean, cov, 6, 4, 21)
    out = ctx.renderView(m, fr["pose0"])
    for b, pose in enumerate(fr["pose0"]):
        view = oracle.render_view(comps, pose)
        assert out["count"][b] == len(view) and np.array_equal(out["ids"][b, :len(view)], view["id"]), b
    # a cap smaller than the view truncates and says so
    V = int(out["count"].max())
    small = ctx.renderView(m, fr["pose0"], cap=V // 2)
    b = int(np.argmax(out["count"]))
    assert small["status"][b] & g.capi.VIEW_OVERFLOW and small["count"][b] == V // 2
    # an empty batch is fine
    assert len(ctx.renderView(m, np.zeros((0, 7)))["count"]) == 0


@pytest.mark.parametrize("num", [1, 5, 8])
def test_search_correspondence(gpu, oracle, v1, gt_v1_03, num):
    g, ctx = gpu
    m, comps, _, mean, cov = v1
    rows = np.array([150, 777, 1400, 2000])
    fr = _kf(mean, cov, len(rows), 700, 2, gt_v1_03[rows, 1:8])
    ctx.renderView(m, fr["pose0"])
    cand, pos = ctx.searchCorrespondence(fr["obs"], num, want_pos=True)
    for b, pose in enumerate(fr["pose0"]):
        view = oracle.render_view(comps, pose)
        want = oracle.search_correspondence(view, fr["obs"][b, :, :2].astype(np.float64), num)
        assert np.array_equal(pos[b], want), b
        ids = np.where(want >= 0, view["id"][np.maximum(want, 0)], -1)
        assert np.array_equal(cand[b], ids), b
    assert (cand >= 0).any() and (cand < 0).any()
    with pytest.raises(g.GmmlocError):
        ctx.searchCorrespondence(fr["obs"][:2], num)  # B differs from the rendered views
    with pytest.raises(g.GmmlocError):
        ctx.searchCorrespondence(fr["obs"], 9)


def test_check_map_association(gpu, oracle, v1, gt_v1_03):
    g, ctx = gpu
    m, comps, nbr, mean, cov = v1
    rows = np.array([300, 1500])
    N = 1200
    fr = _kf(mean, cov, len(rows), N, 3, gt_v1_03[rows, 1:8])
    ctx.renderView(m, fr["pose0"])
    cand = ctx.searchCorrespondence(fr["obs"], 5)
    pt0 = np.empty((len(rows), N, 3))
    for b in range(len(rows)):
        for i in range(N):
            pt0[b, i] = oracle.unproject3(fr["pose0"][b], fr["obs"][b, i, :2].astype(np.float64), fr["depth"][b, i])
    assoc, pt = ctx.checkMapAssociation(m, fr["pose0"], fr["obs"], fr["octave"], cand, pt0)
    moved_unassoc = 0
    for b in range(len(rows)):
        for i in range(N):
            k, x = oracle.check_map_association(comps, nbr, fr["pose0"][b], fr["obs"][b, i].astype(np.float64),
                                                fr["octave"][b, i], cand[b, i], pt0[b, i])
            assert assoc[b, i] == k, (b, i)
            assert np.array_equal(_bits(pt[b, i]), _bits(x)), (b, i)
            moved_unassoc += (k < 0) and not np.array_equal(x, pt0[b, i])
    assert (assoc >= 0).mean() > 0.5 and (assoc < 0).sum() > 50
    assert moved_unassoc > 0  # the kd fallback path ran (moves the point, never associates)


def test_check_map_association_needs_neighbours(gpu, golden_maps):
    g, ctx = gpu
    m = g.GMM(golden_maps["v2"]["mean"][:200], golden_maps["v2"]["cov"][:200])
    with pytest.raises(g.GmmlocError) as ei:
        ctx.checkMapAssociation(m, np.array([[0, 0, 0, 1, 0, 0, 0.0]]), np.zeros((1, 1, 3), np.float32),
                                np.zeros((1, 1), np.uint8), np.full((1, 1, 5), -1, np.int32), np.zeros((1, 1, 3)))
    assert ei.value.status == g.capi.EINVAL


def test_associate_map_elements_fused(gpu, oracle, v1, gt_v1_03):
    g, ctx = gpu
    m, comps, nbr, mean, cov = v1
    rows = np.arange(5, 2149, 179)
    N = 1200  # ORBextractor/num_features, v1.yaml:24
    fr = _kf(mean, cov, len(rows), N, 4, gt_v1_03[rows, 1:8])
    depth = fr["depth"].copy()
    depth[:, ::7] = -1.0  # features without a stereo match
    out = ctx.associateMapElements(m, fr["pose0"], fr["obs"], depth, fr["octave"])
    for b in range(len(rows)):
        want = oracle.keyframe_associate(comps, nbr, fr["pose0"][b], fr["obs"][b], depth[b], fr["octave"][b], nthreads=4)
        assert out["nview"][b] == want["nview"]
        assert np.array_equal(out["cand"][b], want["cand"]), b
        assert np.array_equal(out["assoc"][b], want["assoc"]), b
        assert _same(out["pt3d"][b], want["pt3d"]), b
    assert np.all(out["assoc"][:, ::7] == -1) and np.all(np.isnan(out["pt3d"][:, ::7]))
    assert (out["assoc"] >= 0).mean() > 0.4
    # N = 0 and B = 0
    e = ctx.associateMapElements(m, fr["pose0"][:2], np.zeros((2, 0, 3), np.float32), np.zeros((2, 0), np.float32),
                                 np.zeros((2, 0), np.uint8))
    assert e["assoc"].shape == (2, 0) and np.all(e["nview"] > 0)
    e = ctx.associateMapElements(m, np.zeros((0, 7)), np.zeros((0, 4, 3), np.float32), np.zeros((0, 4), np.float32),
                                 np.zeros((0, 4), np.uint8))
    assert e["assoc"].shape == (0, 4)


def test_associate_map_elements_synthetic_map(gpu, oracle):
    g, ctx = gpu
    mean, cov = synth.make_map(4000, 0.9, seed=33)
    mean = mean * 0.6
    m = g.GMM(mean, cov)
    m.buildNeighbors(2.5)
    comps = oracle.map_build(mean, cov)
    nbr = oracle.neighbors(comps, 2.5)
    fr = _kf(mean, cov, 5, 500, 33)
    out = ctx.associateMapElements(m, fr["pose0"], fr["obs"], fr["depth"], fr["octave"], num=3)
    for b in range(5):
        want = oracle.keyframe_associate(comps, nbr, fr["pose0"][b], fr["obs"][b], fr["depth"][b], fr["octave"][b], num=3)
        assert out["nview"][b] == want["nview"]
        assert np.array_equal(out["cand"][b], want["cand"]) and np.array_equal(out["assoc"][b], want["assoc"]), b
        assert _same(out["pt3d"][b], want["pt3d"]), b


def test_query_point_matches_oracle_and_reference_kdtree(gpu, oracle, golden_maps):
    """GMM::queryPoint (gaussian_mixture.cpp:545-576) through gmmloc_map_query_point: the returned index, the 5-NN list and
    the (discarded) Mahalanobis minimum equal the oracle's; the list also equals the reference's own nanoflann kd-tree
    where oracle/_ref is available."""
    g, ctx = gpu
    gm = golden_maps["v1"]
    m = g.GMM(gm["mean"], gm["cov"])
    comps = oracle.map_build(gm["mean"], gm["cov"])
    pts, _ = synth.make_points_near(gm["mean"], gm["cov"], 3000, seed=8, spread=2.0)
    pts = pts.astype(np.float64)
    pts[:50] = gm["mean"][:50]  # exact hits: distance 0
    out = ctx.queryPoint(m, pts)
    for i in range(0, len(pts), 7):
        r, knn, md = oracle.query_point(comps, pts[i], 5)
        assert out["nearest"][i] == r and np.array_equal(out["knn"][i], knn), i
        assert out["min_mdist2"][i] == md
    ref = oracle.ref_knn(gm["mean"], pts, 5)
    if ref is not None:
        assert np.array_equal(out["knn"], ref[0].astype(np.int32))
    # a map smaller than the list: -1 padded
    small = g.GMM(gm["mean"][:3], gm["cov"][:3])
    o = ctx.queryPoint(small, pts[:4])
    assert np.all(o["knn"][:, 3:] == -1) and np.all(o["knn"][:, :3] >= 0)
    # duplicated means: equal distances -> lower index first
    dup = g.GMM(np.repeat(gm["mean"][:4], 2, axis=0), np.repeat(gm["cov"][:4], 2, axis=0))
    o = ctx.queryPoint(dup, gm["mean"][2:3])
    assert list(o["knn"][0, :2]) == [4, 5] and o["nearest"][0] == 4


def test_render_views_keeps_the_callers_cap_as_row_stride(gpu, golden_maps, gt_v1_03):
    g, ctx = gpu
    import ctypes as C
    gm = golden_maps["v1"]
    K = 40
    m = g.GMM(gm["mean"][:K], gm["cov"][:K])
    pose = synth.make_frames(gm["mean"], gm["cov"], 2, 4, seed=1, poses_wc=gt_v1_03[[100, 900], 1:8])["pose0"]
    cap = K + 7
    cnt = np.zeros(2, np.int32)
    ids = np.full((2, cap), -2, np.int32)
    depth = np.full((2, cap), -2.0)
    cam = g.v1_camera()
    p = lambda a: C.c_void_p(a.ctypes.data)
    g.capi.check(g.capi.lib().gmmloc_render_views(ctx._h, m._h, C.byref(cam), p(pose), 2, cap, p(cnt), None, p(ids), None, None, p(depth)))
    ref = ctx.renderView(m, pose)  # the wrapper clamps cap to K itself
    for b in range(2):
        assert cnt[b] == ref["count"][b]
        assert np.array_equal(ids[b, :cnt[b]], ref["ids"][b, :cnt[b]])
        assert np.array_equal(depth[b, :cnt[b]], ref["depth"][b, :cnt[b]])
        assert np.all(ids[b, K:] == -1) and np.all(depth[b, cnt[b]:] == 0.0)
