"""GPU: the tensor-core distance-tile matcher for small reference clouds (dense.cu) returns the
same exact (distance, index) minima as the tree search and the CPU oracle."""
import numpy as np
import pytest

from oracle import binding as ob
from pgslam_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pm(ctx):
    from pgslam_b200 import pm as _pm
    _pm._DEFAULT_CTX = ctx
    return _pm


def _cloud(pts):
    c = np.ones((4, pts.shape[1]), np.float32)
    c[:3] = pts
    return c


def _both(pm, ctx, ref, qry, max_dist=np.inf):
    out = {}
    for dense in (0, 1 << 14):
        ctx.set_option("dense_max_ref", dense)
        ctx.set_option("dense_count_fallbacks", 1)
        m = pm.Matcher("KDTreeMatcher", {"knn": 1, "maxDist": float(max_dist)} if np.isfinite(max_dist) else {"knn": 1}, ctx=ctx)
        m.init(pm.DataPoints(ref, ctx=ctx))
        got = m.findClosests(pm.DataPoints(qry, ctx=ctx))
        out[dense] = (got.ids, got.dists, m.dense_fallbacks)
    ctx.set_option("dense_max_ref", 0)
    ctx.set_option("dense_count_fallbacks", 0)
    return out[0], out[1 << 14]


@pytest.mark.parametrize("n_ref,n_q", [(1, 3), (5, 200), (100, 1000), (256, 129), (257, 128), (2000, 5000), (8192, 8192), (16384, 3000)])
def test_dense_matcher_equals_tree_and_oracle(pm, ctx, n_ref, n_q):
    g = np.random.default_rng(n_ref)
    # a surface-like cloud far from the origin (the block centring must absorb the offset)
    uv = g.uniform(-1, 1, (2, n_ref))
    ref = _cloud(np.stack([30 + 8 * uv[0], -12 + 8 * uv[1], 1.5 + 0.3 * np.sin(3 * uv[0]) + 0.01 * g.normal(size=n_ref)]).astype(np.float32))
    uq = g.uniform(-1.1, 1.1, (2, n_q))
    qry = _cloud(np.stack([30 + 8 * uq[0], -12 + 8 * uq[1], 1.5 + 0.3 * np.sin(3 * uq[0]) + 0.05 * g.normal(size=n_q)]).astype(np.float32))
    tree, dense = _both(pm, ctx, ref, qry)
    assert np.array_equal(tree[0], dense[0])
    assert np.array_equal(tree[1].view(np.uint32), dense[1].view(np.uint32))
    ids, d2 = ob.kdtree_knn(ref, qry, k=1)
    assert np.array_equal(dense[0], ids) and np.array_equal(dense[1].view(np.uint32), d2.view(np.uint32))
    assert dense[2] <= max(2, n_q // 20)  # almost every query is certified by the 4 candidates


def test_dense_matcher_ties_max_dist_and_scan_data(pm, ctx):
    g = np.random.default_rng(3)
    base = g.uniform(-5, 5, size=(3, 300)).astype(np.float32)
    ref = _cloud(np.concatenate([base] * 6, axis=1))  # every point six times: more ties than candidates -> fallback
    qry = _cloud(g.uniform(-5, 5, size=(3, 700)).astype(np.float32))
    tree, dense = _both(pm, ctx, ref, qry)
    assert np.array_equal(tree[0], dense[0]) and np.array_equal(tree[1], dense[1])
    assert np.all(dense[0] < 300) and dense[2] > 0  # lowest index of every tie, through the exact fallback search
    tree, dense = _both(pm, ctx, ref, qry, max_dist=0.4)
    assert np.array_equal(tree[0], dense[0]) and np.array_equal(tree[1], dense[1]) and (dense[0] == -1).any()
    rd, rf, _ = synth.scan_pair(7, beams=16, az_steps=500)  # 8000-point scans of a 60 m room
    tree, dense = _both(pm, ctx, rf, rd)
    assert np.array_equal(tree[0], dense[0]) and np.array_equal(tree[1].view(np.uint32), dense[1].view(np.uint32))


def test_icp_loop_with_dense_matcher_is_bit_identical(pm, ctx):
    """The fused ICP loop with the tensor-core matcher (small references) returns the very same
    registrations as with the tree: single pairs, a ragged batch, ICPSequence."""
    from tests import util
    pairs = [synth.scan_pair(60 + i, beams=8 + 4 * (i % 3), az_steps=200 + 40 * i)[:2] for i in range(9)]

    def run(dense):
        ctx.set_option("dense_max_ref", dense)
        out = []
        for cfg in (util.C1, util.C2):
            icp = pm.ICP(ctx)
            icp.loadFromYaml(util.to_yaml(cfg))
            rec = icp.compute_batch_array([pm.DataPoints(r, ctx=ctx) for r, _ in pairs], [pm.DataPoints(f, ctx=ctx) for _, f in pairs])
            out.append(rec.tobytes())
            icp(pm.DataPoints(pairs[0][0], ctx=ctx), pm.DataPoints(pairs[0][1], ctx=ctx))
            out.append((icp.last["iterations"], icp.last["T"].tobytes(), icp.last["residual"]))
        seq = pm.ICPSequence(ctx)
        seq.loadFromYaml(util.to_yaml(util.C2))
        seq.setMap(pm.DataPoints(pairs[3][1], ctx=ctx))
        for k in range(3):
            out.append(seq(pm.DataPoints(pairs[3][0], ctx=ctx)).tobytes())
        return out
    try:
        tree = run(0)
        dense = run(1 << 14)
    finally:
        ctx.set_option("dense_max_ref", 0)
    assert tree == dense
    want = ob.icp_run(util.C2, ob.Cloud(pairs[0][0]), ob.Cloud(pairs[0][1]))
    assert dense[3][0] == want["iterations"]
