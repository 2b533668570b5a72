"""CPU suite (`-m "not gpu"`): pins the oracle (golden vectors + independent
cross-checks), exercises the host logic of the library (YAML / registrar /
parameter validation, no device needed), checks that the C-ABI library loads and
exports every symbol include/pgslam_b200.h declares, and covers the N > 1
sharding path with a world_size-2 gloo run."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import binding as ob
from pgslam_b200 import build, dist as pdist, pm, synth
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "pair4000.npz"))


@pytest.fixture(scope="module", autouse=True)
def _built():
    build.build()
    ob.build()


# ------------------------------------------------------------ C ABI surface ---
def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pgslam_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pgs_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 55
    out = subprocess.run(["nm", "-D", "--defined-only", pm.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (pgs_[a-z0-9_]+)", out))
    assert declared <= exported, sorted(declared - exported)
    assert declared == set(pm.SIGNATURES), sorted(declared ^ set(pm.SIGNATURES))
    L = pm.load_library()
    assert b"sm_100a" in L.pgs_version()


def test_no_silent_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pm.CudaError, match="no CPU fallback"):
        pm.Context(0)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pgslam_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.replace("the CPU oracle", "").replace("CPU oracle", "") or f == "synth.py", f


# ----------------------------------------------------- host logic: YAML/registrar ---
def test_config_check_accepts_the_baseline_chains():
    for cfg in (util.C1, util.C2, util.C2_COV, util.C5):
        assert pm.check_config(util.to_yaml(cfg)) >= 5
    assert pm.check_config(util.to_yaml(util.INPUT_FILTERS), chain=False) == 4
    assert pm.check_config("", chain=False) == 0
    assert pm.check_config("- BoundingBoxDataPointsFilter: {xMin: -5, xMax: 5, removeInside: 0}\n", chain=False) == 1
    assert pm.check_config(util.to_yaml(dict(util.C2, outlierFilters=[
        {"TrimmedDistOutlierFilter": {"ratio": 0.9}}, {"SurfaceNormalOutlierFilter": {"maxAngle": 0.5}}]))) >= 6


def test_config_check_block_and_flow_yaml_styles():
    block = """
# comment
readingDataPointsFilters:
  - RandomSamplingDataPointsFilter:
      prob: 0.5
referenceDataPointsFilters:
  - SurfaceNormalDataPointsFilter: {knn: 10, epsilon: 0}   # flow map
matcher:
  KDTreeMatcher:
    knn: 1
    epsilon: 0
outlierFilters:
- TrimmedDistOutlierFilter:
    ratio: 0.75
errorMinimizer:
  PointToPlaneErrorMinimizer
transformationCheckers:
  - CounterTransformationChecker: {maxIterationCount: 40}
  - DifferentialTransformationChecker:
      minDiffRotErr: 0.001
      minDiffTransErr: 0.01
      smoothLength: 4
inspector: NullInspector
logger: NullLogger
"""
    assert pm.check_config(block) == 9
    assert pm.check_config("readingStepDataPointsFilters: []\nmatcher: KDTreeMatcher\n") == 4
    # chains the fused loop accepts since round 1b: knn > 1, step filters, force2D / force4DOF
    assert pm.check_config("readingStepDataPointsFilters:\n  - MaxDistDataPointsFilter: {maxDist: 30}\n"
                           "matcher:\n  KDTreeMatcher: {knn: 3}\n"
                           "errorMinimizer:\n  PointToPlaneErrorMinimizer: {force4DOF: 1}\n") == 5


@pytest.mark.parametrize("text,exc", [
    ("matcher:\n  KDTreeMatcher:\n    knn: 0\n", pm.InvalidParameter),
    ("matcher:\n  KDTreeMatcher: {bogus: 1}\n", pm.InvalidParameter),
    ("matcher:\n  KDTreeMatcher: {knn: abc}\n", pm.InvalidParameter),
    ("outlierFilters:\n  - TrimmedDistOutlierFilter: {ratio: 2}\n", pm.InvalidParameter),
    ("matcher: NoSuchMatcher\n", pm.InvalidElement),
    ("somethingElse:\n  - X\n", pm.InvalidModuleType),
    ("matcher:\n  KDTreeMatcher: {knn: 33}\n", pm.InvalidParameter),  # register k-lists: knn <= 32
    ("errorMinimizer:\n  PointToPlaneWithCovErrorMinimizer: {force2D: 1}\n", pm.InvalidParameter),
])
def test_config_check_rejects_like_libpointmatcher(text, exc):
    with pytest.raises(exc):
        pm.check_config(text)


def test_registrar_lists_the_modules_the_path_needs():
    assert {"RandomSamplingDataPointsFilter", "VoxelGridDataPointsFilter", "SurfaceNormalDataPointsFilter",
            "ObservationDirectionDataPointsFilter", "OrientNormalsDataPointsFilter",
            "SimpleSensorNoiseDataPointsFilter"} <= set(pm.registered("DataPointsFilter"))
    assert pm.registered("Matcher") == ["KDTreeMatcher"]
    assert "TrimmedDistOutlierFilter" in pm.registered("OutlierFilter")
    assert {"PointToPlaneErrorMinimizer", "PointToPlaneWithCovErrorMinimizer", "PointToPointErrorMinimizer"} <= set(
        pm.registered("ErrorMinimizer"))
    assert {"CounterTransformationChecker", "DifferentialTransformationChecker", "BoundTransformationChecker"} == set(
        pm.registered("TransformationChecker"))
    assert pm.registered("Transformation") == ["RigidTransformation"]


# ------------------------------------------------------- oracle vs golden vectors ---
def test_oracle_knn_matches_golden():
    rd, rf = GOLD["reading"], GOLD["reference"]
    ids, d2 = ob.kdtree_knn(rf, rd, k=1)
    assert np.array_equal(ids, GOLD["knn1_ids"]) and np.array_equal(d2, GOLD["knn1_d2"])
    ids, d2 = ob.kdtree_knn(rf, rf, k=5)
    assert np.array_equal(ids, GOLD["knn5_ids"]) and np.array_equal(d2, GOLD["knn5_d2"])


def test_oracle_filters_match_golden():
    oc = ob.Cloud(GOLD["reference"])
    ob.apply_filter(oc, "SurfaceNormalDataPointsFilter", knn=10, keepDensities=1, keepEigenValues=1)
    assert np.array_equal(oc.desc("normals"), GOLD["normals"])
    assert np.array_equal(oc.desc("dens"), GOLD["densities"])
    assert np.array_equal(oc.desc("eigval"), GOLD["eigvalues"])
    vox = ob.Cloud(GOLD["reading"])
    ob.apply_filter(vox, "VoxelGridDataPointsFilter", vSizeX=0.5, vSizeY=0.5, vSizeZ=0.5)
    assert np.array_equal(vox.features, GOLD["voxel_features"])
    st, w = ob.outlier_weights([{"TrimmedDistOutlierFilter": {"ratio": 0.85}}], GOLD["knn1_d2"])
    assert st == 0 and np.array_equal(w, GOLD["trimmed_weights"])


@pytest.mark.parametrize("name,cfg", [("c1", util.C1), ("c2", util.C2), ("c2cov", util.C2_COV)])
def test_oracle_icp_matches_golden(name, cfg):
    r = ob.icp_run(cfg, ob.Cloud(GOLD["reading"]), ob.Cloud(GOLD["reference"]))
    assert r["status"] == 0 and r["iterations"] == int(GOLD[name + "_iterations"])
    np.testing.assert_allclose(r["T"], GOLD[name + "_T"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(r["cov"], GOLD[name + "_cov"], rtol=1e-9, atol=1e-20)
    assert r["residual"] == pytest.approx(float(GOLD[name + "_residual"]), rel=1e-10)


# ------------------------------------------ oracle pinned by independent methods ---
def test_oracle_kdtree_equals_brute_force_twin():
    rd, rf, _ = synth.scan_pair(7, beams=16, az_steps=400)
    for k in (1, 4, 10):
        a = ob.kdtree_knn(rf, rd, k=k)
        b = ob.brute_knn(rf, rd, k=k)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    a = ob.kdtree_knn(rf, rd, k=3, max_dist=0.3)
    b = ob.brute_knn(rf, rd, k=3, max_dist=0.3)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert (a[0] == -1).any() and np.isinf(a[1][a[0] == -1]).all()


def test_oracle_knn_agrees_with_scipy_ckdtree():
    from scipy.spatial import cKDTree
    rd, rf, _ = synth.scan_pair(8, beams=16, az_steps=400)
    ids, d2 = ob.kdtree_knn(rf, rd, k=2)
    dd, ii = cKDTree(rf[:3].T.astype(np.float64)).query(rd[:3].T.astype(np.float64), k=2)
    clear = (dd[:, 1] - dd[:, 0]) > 1e-4  # away from fp32 near-ties
    assert clear.mean() > 0.9
    assert np.array_equal(ids[0][clear], ii[:, 0][clear])
    np.testing.assert_allclose(np.sqrt(d2[0][clear]), dd[:, 0][clear], rtol=1e-5, atol=1e-6)


def test_oracle_ties_go_to_the_lower_index():
    base = np.random.default_rng(0).uniform(-3, 3, size=(3, 200)).astype(np.float32)
    ref = np.ones((4, 600), np.float32)
    ref[:3] = np.concatenate([base, base, base], axis=1)
    ids, d2 = ob.kdtree_knn(ref, ref, k=3)
    assert np.array_equal(ids[0], np.arange(600) % 200)
    assert np.array_equal(ids[1], np.arange(600) % 200 + 200)
    assert (d2 == 0).all()


def test_oracle_quantile_is_the_exact_order_statistic():
    g = np.random.default_rng(1)
    d = g.gamma(2.0, 0.01, size=5000).astype(np.float32)
    d[:40] = 0.0
    d[40:50] = np.inf
    vals = np.sort(d[(d > 0) & np.isfinite(d)])
    for q in (0.5, 0.75, 0.85, 0.999, 1.0):
        st, w = ob.outlier_weights([{"TrimmedDistOutlierFilter": {"ratio": q}}], d[None])
        limit = vals[-1] if q == 1.0 else vals[int(len(vals) * q)]
        assert st == 0 and np.array_equal(w[0], (d <= limit).astype(np.float32))
    st, _ = ob.outlier_weights([{"TrimmedDistOutlierFilter": {"ratio": 0.85}}], np.zeros((1, 10), np.float32))
    assert st == ob.CONVERGENCE_ERROR


def test_oracle_var_trimmed_ratio_minimises_frms():
    """VarTrimmedDistOutlierFilter: the oracle's optimised ratio against a direct numpy
    evaluation of the FRMS criterion on the sorted valid distances."""
    import ctypes as C
    g = np.random.default_rng(5)
    d = np.concatenate([g.gamma(2.0, 0.002, 4000), g.uniform(0.5, 4.0, 600)]).astype(np.float32)  # inliers + outliers
    g.shuffle(d)
    d[:7] = 0.0
    d[7:12] = np.inf
    for lo, hi, lam in ((0.05, 0.99, 0.95), (0.3, 0.8, 2.0), (0.5, 1.0, 0.5)):
        ratio = C.c_float(0)
        assert ob.lib().orc_var_trimmed_ratio(ob._f(d), d.size, lo, hi, lam, C.byref(ratio)) == 0
        v = np.sort(d[np.isfinite(d) & (d > 0)]).astype(np.float64)
        n = d.size
        min_el, max_el = int(np.floor(np.float32(lo) * np.float32(n))), min(int(np.floor(np.float32(hi) * np.float32(n))), v.size)
        ids = np.arange(min_el + 1, max_el + 1, dtype=np.float64)
        frms = (1.0 / (ids / n) ** lam) ** 2 / ids * np.cumsum(v)[min_el:max_el]
        want = np.float32(min_el + int(np.argmin(frms))) / np.float32(n)
        assert ratio.value == want
        # the outlier cluster (the last 13 %) is cut off whenever the range allows it
        if lo < 0.5:
            assert ratio.value < 4000 / 4600 + 0.01
    st, w = ob.outlier_weights(["VarTrimmedDistOutlierFilter"], d[None, :])
    assert st == 0 and w[0, 7:12].sum() == 0 and w[0, :7].sum() == 7  # inf -> 0, exact hits -> 1
    st, _ = ob.outlier_weights(["VarTrimmedDistOutlierFilter"], np.zeros((1, 10), np.float32))
    assert st == ob.CONVERGENCE_ERROR


def test_oracle_dense_algebra_against_numpy():
    import ctypes as C
    L = ob.lib()
    g = np.random.default_rng(2)
    for _ in range(50):
        M = g.normal(size=(3, 3))
        A = np.asfortranarray(M @ M.T)
        w, V = np.zeros(3), np.zeros((3, 3), order="F")
        L.orc_eig3_sym(ob._d(A), ob._d(w), ob._d(V))
        np.testing.assert_allclose(np.sort(w), np.linalg.eigvalsh(A), rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(A @ V, V * w, atol=1e-12)
        G = g.normal(size=(6, 6))
        H = np.asfortranarray(G @ G.T + 0.1 * np.eye(6))
        b, x = g.normal(size=6), np.zeros(6)
        assert L.orc_solve6(ob._d(H), ob._d(b), ob._d(x)) == 6
        np.testing.assert_allclose(x, np.linalg.solve(H, b), rtol=1e-9)
        U, S, Vt = np.zeros((3, 3), order="F"), np.zeros(3), np.zeros((3, 3), order="F")
        Mf = np.asfortranarray(M)
        L.orc_svd3(ob._d(Mf), ob._d(U), ob._d(S), ob._d(Vt))
        np.testing.assert_allclose(S, np.linalg.svd(M, compute_uv=False), rtol=1e-7)
        np.testing.assert_allclose(U @ np.diag(S) @ Vt.T, M, atol=1e-8)
    # rank-deficient normal equations take the minimum-norm branch
    H = np.zeros((6, 6), order="F")
    H[:3, :3] = np.eye(3)
    b, x = np.array([1.0, 2, 3, 0, 0, 0]), np.zeros(6)
    assert L.orc_solve6(ob._d(H), ob._d(b), ob._d(x)) == 3
    np.testing.assert_allclose(x, np.linalg.pinv(H) @ b, atol=1e-12)


@pytest.mark.parametrize("minimizer", ["PointToPointErrorMinimizer", "PointToPlaneErrorMinimizer"])
def test_oracle_icp_recovers_a_known_rigid_motion(minimizer):
    # noiseless box room sampled uniformly on its faces (no ray pattern to alias on)
    g = np.random.default_rng(9)
    n = 4000
    faces = []
    for axis, val in ((0, -5.0), (0, 5.0), (1, -4.0), (1, 4.0), (2, 0.0), (2, 3.0)):
        p = np.stack([g.uniform(-5, 5, n), g.uniform(-4, 4, n), g.uniform(0, 3, n)])
        p[axis] = val
        faces.append(p)
    rf = np.ones((4, 6 * n), np.float32)
    rf[:3] = np.concatenate(faces, axis=1).astype(np.float32)
    T = synth.pose_matrix([0.15, -0.1, 0.04], 0.03, 0.01, -0.015)
    rd = (np.linalg.inv(T) @ rf.astype(np.float64)).astype(np.float32)
    cfg = dict(util.C2, errorMinimizer=minimizer,
               outlierFilters=[{"TrimmedDistOutlierFilter": {"ratio": 0.95}}],
               transformationCheckers=[{"CounterTransformationChecker": {"maxIterationCount": 400}},
                                       {"DifferentialTransformationChecker": {"minDiffRotErr": 1e-7,
                                                                              "minDiffTransErr": 1e-7}}])
    r = ob.icp_run(cfg, ob.Cloud(rd), ob.Cloud(rf))
    assert r["status"] == 0 and not r["max_iter_reached"]
    np.testing.assert_allclose(r["T"], T, atol=1e-5)


def _box_room(n, seed):
    g = np.random.default_rng(seed)
    faces = []
    for axis, val in ((0, -5.0), (0, 5.0), (1, -4.0), (1, 4.0), (2, 0.0), (2, 3.0)):
        p = np.stack([g.uniform(-5, 5, n), g.uniform(-4, 4, n), g.uniform(0, 3, n)])
        p[axis] = val
        faces.append(p)
    rf = np.ones((4, 6 * n), np.float32)
    rf[:3] = np.concatenate(faces, axis=1).astype(np.float32)
    return rf


@pytest.mark.parametrize("mode,T", [
    ("force4DOF", synth.pose_matrix([0.12, -0.08, 0.05], 0.04, 0.0, 0.0)),  # yaw + xyz
    ("force2D", synth.pose_matrix([0.12, -0.08, 0.0], 0.04, 0.0, 0.0)),     # yaw + xy
])
def test_oracle_forced_minimizers_recover_a_motion_inside_their_subspace(mode, T):
    """PointToPlaneErrorMinimizer{force2D, force4DOF}: a motion that lies in the restricted
    parameter space is recovered; the 6-DOF solve of the same case agrees with it."""
    rf = _box_room(3000, 4)
    rd = (np.linalg.inv(T) @ rf.astype(np.float64)).astype(np.float32)
    chk = [{"CounterTransformationChecker": {"maxIterationCount": 400}},
           {"DifferentialTransformationChecker": {"minDiffRotErr": 1e-7, "minDiffTransErr": 1e-7}}]
    cfg = dict(util.C2, errorMinimizer={"PointToPlaneErrorMinimizer": {mode: 1}},
               outlierFilters=[{"TrimmedDistOutlierFilter": {"ratio": 0.95}}], transformationCheckers=chk)
    r = ob.icp_run(cfg, ob.Cloud(rd), ob.Cloud(rf))
    assert r["status"] == 0 and not r["max_iter_reached"]
    np.testing.assert_allclose(r["T"], T, atol=2e-5)
    # the estimate never leaves the subspace: rotation about z only (and no z shift in 2-D)
    np.testing.assert_allclose(r["T"][2, :3], [0, 0, 1], atol=1e-12)
    np.testing.assert_allclose(r["T"][:3, 2], [0, 0, 1], atol=1e-12)
    if mode == "force2D":
        assert abs(r["T"][2, 3]) < 1e-6


def test_oracle_forced_minimizer_is_the_restricted_least_squares_solution():
    """One compute() against numpy: the forced solve equals lstsq over the kept columns of F."""
    rd, rf, _ = synth.scan_pair(21, beams=16, az_steps=200)
    oref = ob.Cloud(rf)
    ob.apply_filter(oref, "SurfaceNormalDataPointsFilter", knn=10)
    ids, d2 = ob.kdtree_knn(rf, rd, k=1)
    w = np.ones_like(d2)
    P = rd[:3].astype(np.float64).T
    Q = rf[:3, ids[0]].astype(np.float64).T
    N = oref.desc("normals")[:, ids[0]].astype(np.float64).T
    for mode, cols in ((1, [2, 3, 4]), (2, [2, 3, 4, 5])):
        st, got = ob.minimize(ob.E_POINT_TO_PLANE, ob.Cloud(rd), oref, ids, d2, w, force_mode=mode)
        assert st == 0
        Nn = N.copy()
        if mode == 1:
            Nn[:, 2] = 0.0
        F = np.concatenate([np.cross(P, Nn), Nn], axis=1)[:, cols]
        e = ((P - Q) * Nn).sum(axis=1)
        x = np.linalg.solve(F.T @ F, -F.T @ e)
        c, s_ = np.cos(x[0]), np.sin(x[0])
        want = np.eye(4)
        want[:2, :2] = [[c, -s_], [s_, c]]
        want[:len(cols) - 1, 3] = x[1:]
        np.testing.assert_allclose(got["T"], want, rtol=0, atol=1e-10)
        assert got["residual"] == pytest.approx(float((e * e).sum()), rel=1e-10)


def test_oracle_icp_with_knn_3_and_step_filters():
    """knn > 1 pairs every reading point with its 3 nearest (ratios over k*N); a step filter is
    applied to the reading on every iteration before the step transform."""
    rd, rf, truth = synth.scan_pair(22, beams=16, az_steps=400)
    base = ob.icp_run(util.C2, ob.Cloud(rd), ob.Cloud(rf))
    k3 = ob.icp_run(dict(util.C2, matcher={"KDTreeMatcher": {"knn": 3}}), ob.Cloud(rd), ob.Cloud(rf))
    assert base["status"] == k3["status"] == 0
    assert k3["point_used_ratio"] == pytest.approx(0.85, abs=1e-3)
    assert np.abs(k3["T"][:3, 3] - truth[:3, 3]).max() < 0.05
    step = dict(util.C2, readingStepDataPointsFilters=[{"MaxDistDataPointsFilter": {"maxDist": 20.0}}])
    s1 = ob.icp_run(step, ob.Cloud(rd), ob.Cloud(rf))
    # the same filter as a reading filter acts in the sensor frame instead of the centred map frame
    pre = dict(util.C2, readingDataPointsFilters=[{"MaxDistDataPointsFilter": {"maxDist": 20.0}}])
    s2 = ob.icp_run(pre, ob.Cloud(rd), ob.Cloud(rf))
    assert s1["status"] == s2["status"] == 0
    assert not np.array_equal(s1["T"], s2["T"])
    assert np.abs(s1["T"][:3, 3] - truth[:3, 3]).max() < 0.05


def test_oracle_sampling_surface_normal_cells():
    """SamplingSurfaceNormal: cells of at most knn points whose sizes follow the
    left = count - count/2 rule; one unit normal per cell; bin mode keeps one point per cell
    at the cell mean; on a plane every normal is the plane normal."""
    g = np.random.default_rng(11)
    n, knn = 5000, 7
    pts = np.ones((4, n), np.float32)
    pts[0], pts[1] = g.uniform(-5, 5, n), g.uniform(-3, 3, n)
    pts[2] = (0.2 * pts[0] - 0.1 * pts[1] + 1.0 + g.normal(0, 1e-4, n)).astype(np.float32)

    def cells(c):
        return 1 if c <= knn else cells(c - c // 2) + cells(c // 2)

    b = ob.Cloud(pts)
    assert ob.apply_filter(b, "SamplingSurfaceNormalDataPointsFilter", knn=knn, samplingMethod=1) == 0
    assert b.n == cells(n)
    nrm = b.desc("normals").astype(np.float64)
    want = np.array([0.2, -0.1, -1.0]) / np.linalg.norm([0.2, -0.1, -1.0])
    assert np.abs(np.abs(nrm.T @ want) - 1.0).max() < 2e-3  # 0.1 mm noise over ~0.2 m cells
    # the kept points are cell means: they lie on the plane, inside the cloud's extent
    m = b.features.astype(np.float64)
    assert np.abs(0.2 * m[0] - 0.1 * m[1] + 1.0 - m[2]).max() < 1e-3
    r = ob.Cloud(pts)
    assert ob.apply_filter(r, "SamplingSurfaceNormalDataPointsFilter", knn=knn, ratio=0.5, seed=3) == 0
    assert abs(r.n / n - 0.5) < 0.03
    # random mode keeps original points, in their original order
    f = r.features
    pos = [np.flatnonzero((pts[:3].T == f[:3, j]).all(axis=1))[0] for j in range(0, r.n, 97)]
    assert pos == sorted(pos)
    # maxBoxDim below the cell size drops everything
    d = ob.Cloud(pts)
    assert ob.apply_filter(d, "SamplingSurfaceNormalDataPointsFilter", knn=knn, maxBoxDim=1e-3) == 0 and d.n == 0


def test_oracle_surface_normal_optional_descriptors():
    """keepMatchedIds / keepMeanDist / sortEigen against numpy on the same neighbourhoods."""
    _, rf, _ = synth.scan_pair(23, beams=16, az_steps=120)
    k = 7
    c = ob.Cloud(rf)
    assert ob.apply_filter(c, "SurfaceNormalDataPointsFilter", knn=k, keepEigenValues=1, keepEigenVectors=1,
                           keepMatchedIds=1, keepMeanDist=1, sortEigen=1) == 0
    ids, _ = ob.kdtree_knn(rf, rf, k=k)
    got = c.descriptors()
    assert np.array_equal(got["matchedIds"], ids.astype(np.float32))
    P = rf[:3].astype(np.float64)
    mean = P[:, ids].mean(axis=1)  # 3 x N
    np.testing.assert_allclose(got["meanDists"][0], np.linalg.norm(P - mean, axis=0), rtol=1e-6, atol=1e-7)
    ev = got["eigValues"]
    assert (np.diff(ev, axis=0) >= 0).all()  # ascending
    # the normal is the first (smallest) eigenvector after sorting
    np.testing.assert_array_equal(got["normals"], np.clip(got["eigVectors"][:3], -1, 1))
    # and it is an eigen-decomposition of the neighbourhood covariance
    i = 100
    Q = P[:, ids[:, i]] - mean[:, [i]]
    Cm = Q @ Q.T / k
    V = got["eigVectors"][:, i].reshape(3, 3).T.astype(np.float64)  # columns = eigenvectors
    np.testing.assert_allclose(Cm @ V, V * ev[:, i].astype(np.float64), atol=1e-6)


def test_oracle_icp_is_equivariant_under_a_common_rigid_motion():
    rd, rf, _ = synth.scan_pair(10, beams=16, az_steps=300)
    G = synth.pose_matrix([1.0, -2.0, 0.3], 0.4, 0.0, 0.0)
    base = ob.icp_run(util.C1, ob.Cloud(rd), ob.Cloud(rf))
    rf2 = (G @ rf.astype(np.float64)).astype(np.float32)
    moved = ob.icp_run(util.C1, ob.Cloud(rd), ob.Cloud(rf2), G)
    assert base["status"] == moved["status"] == 0
    np.testing.assert_allclose(moved["T"], G @ base["T"], atol=5e-3)  # fp32 clouds, different roundings


def test_oracle_sequence_matches_full_icp_up_to_filter_order():
    rd, rf, _ = synth.scan_pair(12, beams=16, az_steps=300)
    seq = ob.IcpSequence(util.C1)  # no reference filters: setMap order is irrelevant
    assert seq.set_map(ob.Cloud(rf)) == 0
    a = seq.run(ob.Cloud(rd))
    b = ob.icp_run(util.C1, ob.Cloud(rd), ob.Cloud(rf))
    assert a["iterations"] == b["iterations"]
    np.testing.assert_array_equal(a["T"], b["T"])


def test_oracle_error_paths():
    rd, rf, _ = synth.scan_pair(13, beams=16, az_steps=100)
    r = ob.icp_run(dict(util.C2, referenceDataPointsFilters=[]), ob.Cloud(rd), ob.Cloud(rf))
    assert r["status"] == ob.INVALID_FIELD
    bad = np.eye(4)
    bad[0, 0] = 2
    assert ob.icp_run(util.C1, ob.Cloud(rd), ob.Cloud(rf), bad)["status"] == ob.TRANSFORMATION_ERROR
    c = ob.Cloud(rd)
    assert ob.apply_filter(c, "OrientNormalsDataPointsFilter") == ob.INVALID_FIELD


def test_synth_scans_are_exact_size_and_reproducible():
    a = synth.scan_pair(3, beams=16, az_steps=64)
    b = synth.scan_pair(3, beams=16, az_steps=64)
    assert a[0].shape == (4, 1024) and a[0].dtype == np.float32
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    r = np.linalg.norm(a[0][:3], axis=0)
    assert r.min() >= 0.99 and r.max() <= 80.01 and np.all(a[0][3] == 1.0)


# ---------------------------------------------------------- N > 1: gloo, 2 ranks ---
def test_shard_ranges_partition_the_pairs():
    for P, G in ((4096, 8), (10, 4), (3, 8), (0, 2)):
        seen = [i for r in range(G) for i in pdist.shard_range(P, r, G)]
        assert seen == list(range(P))


def test_result_packing_roundtrip():
    res = [dict(T=np.arange(16.0).reshape(4, 4), covariance=np.eye(6) * i, iterations=i, status=0,
                max_iterations_reached=bool(i % 2), overlap=0.5, residual=1.5, weighted_point_used_ratio=0.85)
           for i in range(3)]
    back = pdist.unpack_results(pdist.pack_results(res))
    for a, b in zip(res, back):
        assert np.array_equal(a["T"], b["T"]) and np.array_equal(a["covariance"], b["covariance"])
        assert a["iterations"] == b["iterations"] and a["max_iterations_reached"] == b["max_iterations_reached"]


_WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from pgslam_b200 import dist as pdist, synth
from oracle import binding as ob
from tests import util
dist.init_process_group("gloo")
pairs = [synth.scan_pair(s, beams=16, az_steps=60)[:2] for s in range(5)]
def run(block):   # the CPU stand-in for ICP.compute_batch in this CPU-only test
    out = []
    for rd, rf in block:
        r = ob.icp_run(util.C1, ob.Cloud(rd), ob.Cloud(rf))
        out.append(dict(T=r["T"], covariance=r["cov"], iterations=r["iterations"], status=r["status"],
                        max_iterations_reached=r["max_iter_reached"], overlap=r["overlap"], residual=r["residual"],
                        weighted_point_used_ratio=r["weighted_ratio"]))
    return out
res = pdist.register_sharded(run, pairs)
np.save(os.path.join(sys.argv[2], f"rank{dist.get_rank()}.npy"), pdist.pack_results(res))
# the array path bench.py uses: structured pgs_icp_result records of this rank's block
from pgslam_b200 import pm
mine = pdist.shard_range(len(pairs), dist.get_rank(), dist.get_world_size())
rec = np.zeros((len(mine),), dtype=np.dtype(pm.IcpResult))
for j, i in enumerate(mine):
    rec[j]["T"] = np.asarray(res[i]["T"]).ravel(order="F")
    rec[j]["covariance"] = np.asarray(res[i]["covariance"]).ravel(order="F")
    rec[j]["iterations"], rec[j]["status"] = res[i]["iterations"], res[i]["status"]
    rec[j]["max_iterations_reached"] = int(res[i]["max_iterations_reached"])
    rec[j]["overlap"], rec[j]["residual"] = res[i]["overlap"], res[i]["residual"]
    rec[j]["weighted_point_used_ratio"] = res[i]["weighted_point_used_ratio"]
np.save(os.path.join(sys.argv[2], f"rows{dist.get_rank()}.npy"), pdist.gather_records(rec, len(pairs)))
dist.destroy_process_group()
'''


def test_two_rank_gloo_sharding_gives_the_unsharded_results(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="2")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                    "--master-addr", "127.0.0.1", "--master-port", "29617", str(script), ROOT, str(tmp_path)],
                   check=True, env=env, timeout=300, capture_output=True)
    a, b = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b) and a.shape == (5, pdist.RESULT_WIDTH)
    # gather_records (structured records -> rows on every rank) gives the same table
    assert np.array_equal(np.load(tmp_path / "rows0.npy"), a) and np.array_equal(np.load(tmp_path / "rows1.npy"), a)
    for i in range(5):
        rd, rf = synth.scan_pair(i, beams=16, az_steps=60)[:2]
        want = ob.icp_run(util.C1, ob.Cloud(rd), ob.Cloud(rf))
        assert np.array_equal(a[i, :16].reshape(4, 4).T, want["T"])  # independent of the sharding
        assert int(a[i, 52]) == want["iterations"]


def test_oracle_remove_nan_fix_step_shadow():
    """pins of the three element-wise filters against plain numpy statements of the same rules"""
    rd, _, _ = synth.scan_pair(5, beams=16, az_steps=500)
    # RemoveNaN
    bad = rd.copy()
    bad[1, ::7] = np.nan
    c = ob.Cloud(bad)
    assert ob.apply_filter(c, "RemoveNaNDataPointsFilter") == 0
    keep = ~np.isnan(bad[:3]).any(axis=0)
    assert np.array_equal(c.features, bad[:, keep])
    # FixStep: every step-th point from a phase < step, order preserved
    c = ob.Cloud(rd)
    assert ob.apply_filter(c, "FixStepSamplingDataPointsFilter", startStep=9, endStep=9, seed=4) == 0
    out = c.features
    phases = [ph for ph in range(9) if np.array_equal(out, rd[:, ph::9])]
    assert len(phases) == 1
    # Shadow: |cos(normal, ray)| > eps in fp32
    c = ob.Cloud(rd)
    assert ob.apply_filter(c, "SurfaceNormalDataPointsFilter", knn=8) == 0
    nrm = c.descriptors()["normals"].astype(np.float32)
    pts = c.features[:3].astype(np.float32)
    assert ob.apply_filter(c, "ShadowDataPointsFilter", eps=0.3) == 0

    def unit(v):
        s = (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]
        return np.where(s > 0, v / np.sqrt(np.where(s > 0, s, 1)).astype(np.float32), v).astype(np.float32)
    a, b = unit(nrm), unit(pts)
    cosv = np.abs((a[0] * b[0] + a[1] * b[1]) + a[2] * b[2])
    want = cosv > np.float32(0.3)
    assert 0 < want.sum() < want.size
    assert np.array_equal(c.features[:3], pts[:, want])
    assert ob.apply_filter(ob.Cloud(rd), "ShadowDataPointsFilter") != 0  # no normals -> InvalidField


def test_config_accepts_the_added_filters():
    assert pm.check_config("- RemoveNaNDataPointsFilter\n- FixStepSamplingDataPointsFilter: {startStep: 5, endStep: 5}\n"
                           "- ShadowDataPointsFilter: {eps: 0.2}\n- IdentityDataPointsFilter\n", chain=False) == 4
    with pytest.raises(pm.InvalidParameter):
        pm.check_config("- ShadowDataPointsFilter: {epsilon: 0.2}\n", chain=False)


@pytest.mark.parametrize("ext", ["csv", "vtk", "ply"])
def test_cloud_files_round_trip(tmp_path, ext):
    """DataPoints::save / load in libpointmatcher's text formats (host-side, §8f F4)"""
    from pgslam_b200 import cloud_io
    g = np.random.default_rng(0)
    n = 257
    feats = np.vstack([g.normal(size=(3, n)).astype(np.float32) * 30, np.ones((1, n), np.float32)])
    desc = {"normals": g.normal(size=(3, n)).astype(np.float32),
            "simpleSensorNoise": g.random((1, n)).astype(np.float32),
            "observationDirections": g.normal(size=(3, n)).astype(np.float32)}
    if ext == "vtk":
        desc["eigVectors"] = g.normal(size=(9, n)).astype(np.float32)
    path = str(tmp_path / f"cloud.{ext}")
    cloud_io.save(path, feats, desc)
    f2, d2 = cloud_io.load(path)
    assert np.array_equal(f2, feats)          # %.9g round-trips float32 exactly
    assert set(d2) == set(desc)
    for k in desc:
        assert np.array_equal(d2[k], desc[k]), k
    # empty cloud
    cloud_io.save(path, np.zeros((4, 0), np.float32))
    f3, d3 = cloud_io.load(path)
    assert f3.shape == (4, 0) and d3 == {}


def test_cloud_files_as_libpointmatcher_writes_them(tmp_path):
    from pgslam_b200 import cloud_io
    p = tmp_path / "a.csv"
    p.write_text("x,y,z,nx,ny,nz,intensity\n1,2,3,0,0,1,0.5\n4,5,6,1,0,0,0.25\n")
    f, d = cloud_io.load(str(p))
    assert np.array_equal(f, np.array([[1, 4], [2, 5], [3, 6], [1, 1]], np.float32))
    assert np.array_equal(d["normals"], np.array([[0, 1], [0, 0], [1, 0]], np.float32))
    assert np.array_equal(d["intensity"], np.array([[0.5, 0.25]], np.float32))
    p = tmp_path / "b.csv"          # no header: x y z, whitespace separated
    p.write_text("1 2 3\n4 5 6\n")
    f, d = cloud_io.load(str(p))
    assert f.shape == (4, 2) and d == {} and f[2, 1] == 6
    p = tmp_path / "c.vtk"
    p.write_text("# vtk DataFile Version 3.0\nFile created by libpointmatcher\nASCII\nDATASET POLYDATA\n"
                 "POINTS 2 float\n1 2 3\n4 5 6\nVERTICES 2 4\n1 0\n1 1\nPOINT_DATA 2\n"
                 "NORMALS normals float\n0 0 1\n1 0 0\nSCALARS densities float 1\nLOOKUP_TABLE default\n7\n8\n")
    f, d = cloud_io.load(str(p))
    assert np.array_equal(f[:3].T, np.array([[1, 2, 3], [4, 5, 6]], np.float32))
    assert np.array_equal(d["normals"].T, np.array([[0, 0, 1], [1, 0, 0]], np.float32))
    assert np.array_equal(d["densities"], np.array([[7, 8]], np.float32))
    p = tmp_path / "d.ply"
    p.write_text("ply\nformat ascii 1.0\nelement vertex 2\nproperty float x\nproperty float y\nproperty float z\n"
                 "property float nx\nproperty float ny\nproperty float nz\nend_header\n1 2 3 0 0 1\n4 5 6 1 0 0\n")
    f, d = cloud_io.load(str(p))
    assert np.array_equal(f[:3, 1], np.array([4, 5, 6], np.float32)) and d["normals"].shape == (3, 2)
    with pytest.raises(ValueError):
        cloud_io.load(str(tmp_path / "x.pcd"))


def test_config_warnings_name_the_departures_from_upstream():
    assert pm.config_warnings(util.to_yaml(util.C2)) == []
    cfg = dict(util.C2, readingStepDataPointsFilters=[{"RandomSamplingDataPointsFilter": {"prob": 0.5}}],
               readingDataPointsFilters=[{"MaxDensityDataPointsFilter": {"maxDensity": 50}}])
    w = pm.config_warnings(util.to_yaml(cfg))
    assert len(w) == 3 and any("ONCE per registration" in x for x in w) and sum("counter-based hash" in x for x in w) == 2
    w = pm.config_warnings("- RandomSamplingDataPointsFilter: {prob: 0.3}\n", chain=False)
    assert len(w) == 1 and "rand()" in w[0]
    assert "rejected" in pm.config_warnings("matcher:\n  NoSuchMatcher\n")[0]


def test_c_abi_cloud_file_parser_on_the_host(tmp_path):
    """pgs_cloud_file_info: the C ABI's csv / vtk / ply parser (pgs_cloud_load without the upload) reads the
    files the Python module writes, and libpointmatcher-style files, with the same point / descriptor counts."""
    from pgslam_b200 import cloud_io
    g = np.random.default_rng(1)
    n = 129
    feats = np.vstack([g.normal(size=(3, n)).astype(np.float32), np.ones((1, n), np.float32)])
    desc = {"normals": g.normal(size=(3, n)).astype(np.float32), "simpleSensorNoise": g.random((1, n)).astype(np.float32),
            "observationDirections": g.normal(size=(3, n)).astype(np.float32), "eigValues": g.random((3, n)).astype(np.float32)}
    for ext in ("csv", "vtk", "ply"):
        path = str(tmp_path / f"c.{ext}")
        cloud_io.save(path, feats, desc)
        assert pm.cloud_file_info(path) == (n, 4)
    p = tmp_path / "lpm.csv"
    p.write_text("x,y,z,nx,ny,nz,intensity\n1,2,3,0,0,1,0.5\n4,5,6,1,0,0,0.25\n")
    assert pm.cloud_file_info(str(p)) == (2, 2)
    p = tmp_path / "bare.csv"
    p.write_text("1 2 3\n4 5 6\n7 8 9\n")
    assert pm.cloud_file_info(str(p)) == (3, 0)
    with pytest.raises(pm.PointMatcherError):
        pm.cloud_file_info(str(tmp_path / "x.pcd"))
    p = tmp_path / "noz.csv"
    p.write_text("x,y\n1,2\n")
    with pytest.raises(pm.PointMatcherError, match="'z' column"):
        pm.cloud_file_info(str(p))


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): exactly one JSON line on
    stdout with the contract's keys; the GPU arm refuses to run without a device instead of falling back."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "icp_registrations_per_s_120k_pt_pairs"
    for key in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0 and "workload" in d["config"]
    import torch
    if not torch.cuda.is_available():
        ours = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                              timeout=600)
        assert ours.returncode != 0 and "no CUDA device" in (ours.stderr + ours.stdout)


# ------------------------------------------- libnabo-faithful mode vs the contract ---
def test_libnabo_faithful_search_equals_the_contract_at_eps_0_and_not_above():
    """oracle/README.md's divergence table, re-measured: with eps = 0 libnabo's own rules (strict <,
    first-visited ties, incremental rd) return the same ids and distances as the contract's exact
    (distance, index) minimum on scan data; they differ only on exact ties.  With eps > 0 libnabo is
    approximate and a large share of the ids differ - which is why the product refuses eps != 0."""
    rd, rf, _ = synth.scan_pair(3, beams=16, az_steps=900)
    for k in (1, 10):
        a = ob.kdtree_knn(rf, rd, k=k)
        b = ob.kdtree_knn(rf, rd, k=k, mode=ob.SEARCH_NABO)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        c = ob.kdtree_knn(rf, rd, k=k, mode=ob.SEARCH_NABO, epsilon=1.0)
        assert (a[0] != c[0]).mean() > 0.05            # approximate search: not the same neighbours
        assert np.all(c[1][0] <= a[1][0] * 4.0 + 1e-12)  # but within (1+eps)^2 of the true nearest distance
    # exact ties (every point three times): same distances, ids may differ (first visited vs lowest index)
    g = np.random.default_rng(0)
    base = g.uniform(-5, 5, size=(3, 400)).astype(np.float32)
    ref = np.ones((4, 1200), np.float32)
    ref[:3] = np.concatenate([base, base, base], axis=1)
    q = np.ones((4, 200), np.float32)
    q[:3] = g.uniform(-5, 5, size=(3, 200)).astype(np.float32)
    a = ob.kdtree_knn(ref, q, k=3)
    b = ob.kdtree_knn(ref, q, k=3, mode=ob.SEARCH_NABO)
    assert np.array_equal(a[1], b[1])
    assert np.array_equal(np.sort(a[0] % 400, axis=0), np.sort(b[0] % 400, axis=0))  # the same points, other copies
    assert np.all(a[0][0] < 400)  # the contract: lowest index of every tie


def test_libnabo_faithful_icp_matches_the_contract_at_eps_0():
    rd, rf, _ = synth.scan_pair(4, beams=16, az_steps=600)
    want = ob.icp_run(util.C2, ob.Cloud(rd), ob.Cloud(rf))
    try:
        ob.set_search_mode(ob.SEARCH_NABO)
        same = ob.icp_run(util.C2, ob.Cloud(rd), ob.Cloud(rf))
        approx = ob.icp_run(dict(util.C2, matcher={"KDTreeMatcher": {"knn": 1, "epsilon": 3.16}}), ob.Cloud(rd), ob.Cloud(rf))
    finally:
        ob.set_search_mode(ob.SEARCH_CONTRACT)
    assert same["iterations"] == want["iterations"] and np.array_equal(same["T"], want["T"])
    assert np.abs(approx["T"] - want["T"]).max() > 1e-5  # eps > 0 moves the pose well beyond the 1e-5 contract


def test_epsilon_is_rejected_not_ignored():
    """KDTreeMatcher / SurfaceNormalDataPointsFilter `epsilon` != 0 asks for libnabo's approximate
    search, whose result depends on libnabo's tree and visit order: refused (PGS_EPSILON_POLICY=exact
    accepts it and runs the exact search instead)."""
    cfg = dict(util.C2, matcher={"KDTreeMatcher": {"knn": 1, "epsilon": 3.16}})
    with pytest.raises(pm.InvalidParameter, match="epsilon"):
        pm.check_config(util.to_yaml(cfg))
    cfg = dict(util.C2, referenceDataPointsFilters=[{"SurfaceNormalDataPointsFilter": {"knn": 10, "epsilon": 1.33}}])
    with pytest.raises(pm.InvalidParameter, match="epsilon"):
        pm.check_config(util.to_yaml(cfg))
    with pytest.raises(pm.InvalidParameter, match="epsilon"):
        pm.check_config(util.to_yaml([{"SurfaceNormalDataPointsFilter": {"epsilon": 0.5}}]), chain=False)
    assert pm.check_config(util.to_yaml(util.C2)) > 0  # epsilon: 0 passes
    env = dict(os.environ, PGS_EPSILON_POLICY="exact")
    code = ("import sys; sys.path.insert(0, %r); from pgslam_b200 import pm; from tests import util; "
            "print(pm.check_config(util.to_yaml(dict(util.C2, matcher={'KDTreeMatcher': {'knn': 1, 'epsilon': 3.16}}))))" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout
    assert int(out.strip()) > 0


def test_tools_and_bench_compile():
    """every script under tools/ (and bench.py, __graft_entry__.py) is at least valid Python"""
    import ast
    import glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = sorted(glob.glob(os.path.join(root, "tools", "*.py"))) + [os.path.join(root, "bench.py"),
                                                                     os.path.join(root, "__graft_entry__.py")]
    assert len(files) > 10
    for f in files:
        with open(f) as fh:
            ast.parse(fh.read(), filename=f)


def test_clock_sampler_without_a_gpu_reports_why():
    import bench
    s = bench.ClockSampler(0)
    s.start()
    out = s.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    if out["sm_mhz"] is None:
        assert out["reasons"]
