"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the
same seeded inputs.  Bit-exact for ids / distances / weights / filter outputs;
poses within 1e-5 m and 1e-5 rad with equal iteration counts (BASELINE.json)."""
import numpy as np
import pytest

from oracle import binding as ob
from pgslam_b200 import synth
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pm(ctx):
    from pgslam_b200 import pm as _pm
    _pm._DEFAULT_CTX = ctx
    return _pm


@pytest.fixture(scope="module")
def pair30k():
    return synth.scan_pair(3, beams=16, az_steps=1875)


@pytest.fixture(scope="module")
def pair120k():
    return synth.scan_pair(5, beams=64, az_steps=1875)


# ---------------------------------------------------------------- kNN (A9) ---
@pytest.mark.parametrize("k", [1, 5, 10, 32])
def test_knn_bit_exact_30k(pm, pair30k, k):
    rd, rf, _ = pair30k
    m = pm.Matcher("KDTreeMatcher", {"knn": k})
    m.init(pm.DataPoints(rf))
    got = m.findClosests(pm.DataPoints(rd))
    ids, d2 = ob.kdtree_knn(rf, rd, k=k)
    assert np.array_equal(got.ids, ids)
    assert np.array_equal(got.dists.view(np.uint32), d2.view(np.uint32))


def test_knn_bit_exact_120k(pm, pair120k):
    rd, rf, _ = pair120k
    m = pm.Matcher("KDTreeMatcher", {"knn": 1})
    m.init(pm.DataPoints(rf))
    got = m.findClosests(pm.DataPoints(rd))
    ids, d2 = ob.kdtree_knn(rf, rd, k=1)
    assert np.array_equal(got.ids, ids)
    assert np.array_equal(got.dists.view(np.uint32), d2.view(np.uint32))


def test_knn_ties_lower_index_and_max_dist(pm):
    g = np.random.default_rng(0)
    base = g.uniform(-5, 5, size=(3, 500)).astype(np.float32)
    ref = np.ones((4, 1500), np.float32)
    ref[:3] = np.concatenate([base, base, base], axis=1)  # every point three times -> exact ties
    q = np.ones((4, 300), np.float32)
    q[:3] = g.uniform(-5, 5, size=(3, 300)).astype(np.float32)
    for k, md in ((1, np.inf), (4, np.inf), (3, 0.7)):
        m = pm.Matcher("KDTreeMatcher", {"knn": k, "maxDist": md})
        m.init(pm.DataPoints(ref))
        got = m.findClosests(pm.DataPoints(q))
        ids, d2 = ob.brute_knn(ref, q, k=k, max_dist=md)
        assert np.array_equal(got.ids, ids)
        assert np.array_equal(got.dists.view(np.uint32), d2.view(np.uint32))
    if True:
        m = pm.Matcher("KDTreeMatcher", {"knn": 1})
        m.init(pm.DataPoints(ref))
        got = m.findClosests(pm.DataPoints(q))
        assert (got.ids < 500).all()  # lower index of each triple wins


@pytest.mark.parametrize("n", [1, 7, 8, 9, 63, 513])
def test_knn_small_and_ragged(pm, n):
    g = np.random.default_rng(n)
    ref = np.ones((4, n), np.float32)
    ref[:3] = g.normal(size=(3, n)).astype(np.float32)
    q = np.ones((4, 37), np.float32)
    q[:3] = g.normal(size=(3, 37)).astype(np.float32)
    k = min(3, n)
    m = pm.Matcher("KDTreeMatcher", {"knn": k})
    m.init(pm.DataPoints(ref))
    got = m.findClosests(pm.DataPoints(q))
    ids, d2 = ob.brute_knn(ref, q, k=k)
    assert np.array_equal(got.ids, ids)
    assert np.array_equal(got.dists.view(np.uint32), d2.view(np.uint32))


def test_knn_more_neighbours_than_points(pm):
    ref = np.ones((4, 2), np.float32)
    ref[:3, 0] = (0, 0, 0)
    ref[:3, 1] = (1, 0, 0)
    q = np.ones((4, 1), np.float32)
    q[:3, 0] = (0.1, 0, 0)
    m = pm.Matcher("KDTreeMatcher", {"knn": 4})
    m.init(pm.DataPoints(ref))
    got = m.findClosests(pm.DataPoints(q))
    assert got.ids[:, 0].tolist() == [0, 1, -1, -1]
    assert np.isinf(got.dists[2:, 0]).all()


# ------------------------------------------------------------ filters (A3-A7) ---
def _cmp_cloud(dp, oc, exact=True):
    assert dp.getNbPoints() == oc.n
    assert np.array_equal(dp.features.view(np.uint32), oc.features.view(np.uint32))
    want = oc.descriptors()
    got = dp.descriptors
    assert set(got) == set(want)
    for lab in want:
        if exact:
            assert np.array_equal(got[lab].view(np.uint32), want[lab].view(np.uint32)), lab
        else:
            np.testing.assert_allclose(got[lab], want[lab], rtol=1e-6, atol=1e-7, err_msg=lab)


@pytest.mark.parametrize("extra", [{}, {"keepMatchedIds": 1, "keepMeanDist": 1, "sortEigen": 1},
                                   {"knn": 5, "maxDist": 0.3, "keepMatchedIds": 1, "keepMeanDist": 1}])
def test_surface_normal_filter_bit_exact(pm, pair30k, extra):
    _, rf, _ = pair30k
    params = {"knn": 10, "keepNormals": 1, "keepDensities": 1, "keepEigenValues": 1, "keepEigenVectors": 1}
    params.update(extra)
    dp = pm.DataPoints(rf)
    f = pm.DataPointsFilters()
    f.append("SurfaceNormalDataPointsFilter", params)
    f.apply(dp)
    oc = ob.Cloud(rf)
    assert ob.apply_filter(oc, "SurfaceNormalDataPointsFilter", **params) == 0
    _cmp_cloud(dp, oc)


def test_input_filter_chain_bit_exact(pm, pair30k):
    rd, _, _ = pair30k
    dp = pm.DataPoints(rd)
    pm.DataPointsFilters(util.to_yaml(util.INPUT_FILTERS)).apply(dp)
    oc = ob.Cloud(rd)
    for it in util.INPUT_FILTERS:
        (name, p), = ob._modlist([it])
        assert ob.apply_filter(oc, name, **p) == 0
    _cmp_cloud(dp, oc)
    assert {l for l, _ in dp.descriptorLabels()} == {"normals", "observationDirections", "simpleSensorNoise"}


@pytest.mark.parametrize("centroid", [1, 0])
def test_voxel_grid_bit_exact(pm, pair120k, centroid):
    rd, _, _ = pair120k
    params = {"vSizeX": 0.2, "vSizeY": 0.25, "vSizeZ": 0.2, "useCentroid": centroid, "averageExistingDescriptors": 1}
    dp = pm.DataPoints(rd, {"simpleSensorNoise": np.linspace(0, 1, rd.shape[1], dtype=np.float32)[None]})
    f = pm.DataPointsFilters()
    f.append("VoxelGridDataPointsFilter", params)
    f.apply(dp)
    oc = ob.Cloud(rd, {"simpleSensorNoise": np.linspace(0, 1, rd.shape[1], dtype=np.float32)[None]})
    assert ob.apply_filter(oc, "VoxelGridDataPointsFilter", **params) == 0
    assert 1000 < oc.n < rd.shape[1]
    _cmp_cloud(dp, oc)


@pytest.mark.parametrize("name,params", [
    ("RandomSamplingDataPointsFilter", {"prob": 0.6, "seed": 7}),
    ("MaxDistDataPointsFilter", {"dim": -1, "maxDist": 20.0}),
    ("MinDistDataPointsFilter", {"dim": 0, "minDist": 1.5}),
    ("BoundingBoxDataPointsFilter", {"xMin": -8, "xMax": 12, "yMin": -6, "yMax": 5, "zMin": -3, "zMax": 0.5, "removeInside": 0}),
    ("BoundingBoxDataPointsFilter", {"xMin": -8, "xMax": 12, "yMin": -6, "yMax": 5, "zMin": -3, "zMax": 0.5, "removeInside": 1}),
    ("FixStepSamplingDataPointsFilter", {"startStep": 7, "endStep": 7, "seed": 3}),
    ("FixStepSamplingDataPointsFilter", {}),
])
def test_subsampling_filters_bit_exact(pm, pair30k, name, params):
    rd, _, _ = pair30k
    dp = pm.DataPoints(rd)
    f = pm.DataPointsFilters()
    f.append(name, params)
    f.apply(dp)
    oc = ob.Cloud(rd)
    assert ob.apply_filter(oc, name, **params) == 0
    assert 0 < oc.n < rd.shape[1]
    _cmp_cloud(dp, oc)


def test_remove_nan_filter_bit_exact(pm, pair30k):
    rd, _, _ = pair30k
    rd = rd.copy()
    g = np.random.default_rng(3)
    bad = g.choice(rd.shape[1], 500, replace=False)
    rd[g.integers(0, 3, bad.size), bad] = np.nan
    rd[:3, bad[:20]] = np.nan  # whole points too
    desc = {"simpleSensorNoise": np.arange(rd.shape[1], dtype=np.float32)[None]}
    dp = pm.DataPoints(rd, desc)
    pm.DataPointsFilters("- RemoveNaNDataPointsFilter\n").apply(dp)
    oc = ob.Cloud(rd, desc)
    assert ob.apply_filter(oc, "RemoveNaNDataPointsFilter") == 0
    assert oc.n == rd.shape[1] - 500
    _cmp_cloud(dp, oc)
    assert not np.isnan(dp.features).any()
    # a clean cloud passes through untouched, an Identity filter always does
    dp2 = pm.DataPoints(pair30k[0])
    pm.DataPointsFilters("- RemoveNaNDataPointsFilter\n- IdentityDataPointsFilter\n").apply(dp2)
    assert np.array_equal(dp2.features, pair30k[0])


@pytest.mark.parametrize("eps", [0.1, 0.35])
def test_shadow_filter_bit_exact(pm, pair30k, eps):
    rd, _, _ = pair30k
    chain = [{"SurfaceNormalDataPointsFilter": {"knn": 8}}, {"ShadowDataPointsFilter": {"eps": eps}}]
    dp = pm.DataPoints(rd)
    pm.DataPointsFilters(util.to_yaml(chain)).apply(dp)
    oc = ob.Cloud(rd)
    for it in chain:
        (name, p), = ob._modlist([it])
        assert ob.apply_filter(oc, name, **p) == 0
    assert 0 < oc.n < rd.shape[1]
    _cmp_cloud(dp, oc)
    # without normals: InvalidField, as upstream
    with pytest.raises(pm.InvalidField):
        pm.DataPointsFilters("- ShadowDataPointsFilter\n").apply(pm.DataPoints(rd))


def test_fix_step_schedule_is_rejected(pm, pair30k):
    with pytest.raises(pm.InvalidParameter):
        pm.DataPointsFilters("- FixStepSamplingDataPointsFilter: {startStep: 10, endStep: 2, stepMult: 0.5}\n").apply(
            pm.DataPoints(pair30k[0]))


@pytest.mark.parametrize("params", [
    {},
    {"knn": 12, "ratio": 0.3, "seed": 5, "keepDensities": 1, "keepEigenValues": 1, "keepEigenVectors": 1},
    {"samplingMethod": 1, "knn": 9, "keepDensities": 1},
    {"samplingMethod": 1, "averageExistingDescriptors": 0, "maxBoxDim": 0.6, "keepNormals": 0, "keepDensities": 1},
])
def test_sampling_surface_normal_filter_bit_exact(pm, pair30k, params):
    rd, _, _ = pair30k
    # existing descriptors travel with the kept points (and are averaged per cell in bin mode)
    pre = ["ObservationDirectionDataPointsFilter", {"SimpleSensorNoiseDataPointsFilter": {"sensorType": 0}}]
    chain = pre + [{"SamplingSurfaceNormalDataPointsFilter": params}]
    dp = pm.DataPoints(rd)
    pm.DataPointsFilters(util.to_yaml(chain)).apply(dp)
    oc = ob.Cloud(rd)
    for it in chain:
        (name, p), = ob._modlist([it])
        assert ob.apply_filter(oc, name, **p) == 0
    assert 0 < oc.n < rd.shape[1]
    _cmp_cloud(dp, oc)


def test_sampling_surface_normal_small_and_ragged(pm):
    g = np.random.default_rng(1)
    for n in (1, 5, 7, 8, 15, 100, 1001):
        pts = np.ones((4, n), np.float32)
        pts[:3] = g.normal(size=(3, n)).astype(np.float32)
        for params in ({"ratio": 1.0}, {"samplingMethod": 1}):
            dp = pm.DataPoints(pts)
            f = pm.DataPointsFilters()
            f.append("SamplingSurfaceNormalDataPointsFilter", params)
            f.apply(dp)
            oc = ob.Cloud(pts)
            assert ob.apply_filter(oc, "SamplingSurfaceNormalDataPointsFilter", **params) == 0
            _cmp_cloud(dp, oc)


def test_icp_default_chain_is_upstreams_set_default(pm, pair30k):
    """ICP() without a YAML = ICPChainBase::setDefault: RandomSampling reading filter,
    SamplingSurfaceNormal reference filter, TrimmedDist 0.85, point-to-plane."""
    rd, rf, truth = pair30k
    icp = pm.ICP()
    T = icp(pm.DataPoints(rd), pm.DataPoints(rf))
    want = ob.icp_run(ob.default_config(), ob.Cloud(rd), ob.Cloud(rf))
    assert want["status"] == 0
    assert icp.last["iterations"] == want["iterations"]
    util.assert_pose_close(T, want["T"])
    assert icp.last["n_reference"] < rf.shape[1] and icp.last["n_reading"] < rd.shape[1]
    assert np.abs(T[:3, 3] - truth[:3, 3]).max() < 0.05


def test_max_density_filter_bit_exact(pm, pair30k):
    rd, _, _ = pair30k
    chain = [{"SurfaceNormalDataPointsFilter": {"knn": 8, "keepDensities": 1}},
             {"MaxDensityDataPointsFilter": {"maxDensity": 50.0, "seed": 3}}]
    dp = pm.DataPoints(rd)
    pm.DataPointsFilters(util.to_yaml(chain)).apply(dp)
    oc = ob.Cloud(rd)
    for it in chain:
        (name, p), = ob._modlist([it])
        assert ob.apply_filter(oc, name, **p) == 0
    assert 0 < oc.n < rd.shape[1]
    _cmp_cloud(dp, oc)
    # no densities -> InvalidField, as upstream
    with pytest.raises(pm.InvalidField):
        pm.DataPointsFilters(util.to_yaml(chain[1:])).apply(pm.DataPoints(rd))
    assert ob.apply_filter(ob.Cloud(rd), "MaxDensityDataPointsFilter") == ob.INVALID_FIELD


def test_rigid_transformation_bit_exact_and_rigidity_check(pm, pair30k):
    rd, _, truth = pair30k
    dp = pm.DataPoints(rd)
    pm.DataPointsFilters(util.to_yaml(util.INPUT_FILTERS[:2])).apply(dp)
    out = pm.RigidTransformation().compute(dp, truth)
    oc = ob.Cloud(rd)
    for it in util.INPUT_FILTERS[:2]:
        (name, p), = ob._modlist([it])
        ob.apply_filter(oc, name, **p)
    assert ob.rigid_transform(oc, truth) == 0
    _cmp_cloud(out, oc)
    bad = truth.copy()
    bad[:3, :3] *= 1.1
    with pytest.raises(pm.TransformationError):
        pm.RigidTransformation().compute(dp, bad)


def test_concatenate_keeps_common_descriptors(pm, pair30k):
    rd, rf, _ = pair30k
    a = pm.DataPoints(rd[:, :1000], {"normals": np.zeros((3, 1000), np.float32), "densities": np.ones((1, 1000), np.float32)})
    b = pm.DataPoints(rf[:, :500], {"normals": np.ones((3, 500), np.float32)})
    a.concatenate(b)
    assert a.getNbPoints() == 1500
    assert [l for l, _ in a.descriptorLabels()] == ["normals"]
    assert np.array_equal(a.features[:, 1000:], rf[:, :500])
    assert a.getDescriptorByName("normals")[:, 1000:].min() == 1.0


# ------------------------------------------------- outliers + minimizers (A10-A13) ---
@pytest.mark.parametrize("filters", [
    [{"TrimmedDistOutlierFilter": {"ratio": 0.85}}],
    [{"TrimmedDistOutlierFilter": {"ratio": 1.0}}],
    [{"MedianDistOutlierFilter": {"factor": 3}}, {"MaxDistOutlierFilter": {"maxDist": 0.5}}],
    [{"MinDistOutlierFilter": {"minDist": 0.01}}, {"TrimmedDistOutlierFilter": {"ratio": 0.5}}],
    [],
])
def test_outlier_weights_bit_exact(pm, pair30k, filters):
    rd, rf, _ = pair30k
    ids, d2 = ob.kdtree_knn(rf, rd, k=1)
    d2 = d2.copy()
    d2[0, :50] = 0.0       # exact hits are excluded from the quantile (A.3)
    d2[0, 50:60] = np.inf  # unfound matches
    o = pm.OutlierFilters()
    for it in filters:
        (name, p), = ob._modlist([it])
        o.append(name, p)
    got = o.compute(pm.DataPoints(rd), pm.DataPoints(rf), pm.Matches(ids, d2))
    st, want = ob.outlier_weights(filters, d2)
    assert st == 0
    assert np.array_equal(got, want)


def test_var_trimmed_dist_outlier_filter(pm, pair30k):
    rd, rf, _ = pair30k
    ids, d2 = ob.kdtree_knn(rf, rd, k=2)
    d2[0, :50] = 0.0       # exact hits and unfound matches are not "valid" distances
    d2[1, 100:130] = np.inf
    for params in ({}, {"minRatio": 0.3, "maxRatio": 0.9, "lambda": 2.0}):
        filt = [{"VarTrimmedDistOutlierFilter": params}]
        st, want = ob.outlier_weights(filt, d2)
        assert st == 0
        o = pm.OutlierFilters()
        o.append("VarTrimmedDistOutlierFilter", params)
        got = o.compute(pm.DataPoints(rd), pm.DataPoints(rf), pm.Matches(ids, d2))
        assert np.array_equal(got, want)
        assert 0.04 < want.mean() < 1.0
    # fused loop: the ratio is re-optimised every iteration
    cfg = dict(util.C2, outlierFilters=[{"VarTrimmedDistOutlierFilter": {"minRatio": 0.4, "maxRatio": 0.95}}])
    icp, T, want = _run_both(pm, cfg, rd, rf)
    assert want["status"] == 0
    assert icp.last["iterations"] == want["iterations"]
    util.assert_pose_close(T, want["T"])
    assert icp.last["weighted_point_used_ratio"] == pytest.approx(want["weighted_ratio"], rel=1e-12)
    # combined with a fixed limit, and on a batch with a pair that has nothing to filter
    cfg2 = dict(util.C1, outlierFilters=[{"MaxDistOutlierFilter": {"maxDist": 1.0}}, "VarTrimmedDistOutlierFilter"])
    icp, T, want = _run_both(pm, cfg2, rd, rf)
    assert want["status"] == 0 and icp.last["iterations"] == want["iterations"]
    util.assert_pose_close(T, want["T"])
    with pytest.raises(pm.ConvergenceError):
        icp(pm.DataPoints(rf), pm.DataPoints(rf))  # every distance is 0
    assert ob.icp_run(cfg2, ob.Cloud(rf), ob.Cloud(rf))["status"] == ob.CONVERGENCE_ERROR


def test_outlier_no_valid_distance_is_convergence_error(pm):
    q = np.ones((4, 10), np.float32)
    d2 = np.zeros((1, 10), np.float32)
    o = pm.OutlierFilters()
    o.append("TrimmedDistOutlierFilter", {"ratio": 0.85})
    with pytest.raises(pm.ConvergenceError):
        o.compute(pm.DataPoints(q), pm.DataPoints(q), pm.Matches(np.zeros((1, 10), np.int32), d2))


@pytest.mark.parametrize("name,kind", [("PointToPlaneErrorMinimizer", ob.E_POINT_TO_PLANE),
                                       ("PointToPlaneWithCovErrorMinimizer", ob.E_POINT_TO_PLANE_WITH_COV),
                                       ("PointToPointErrorMinimizer", ob.E_POINT_TO_POINT)])
def test_error_minimizer_matches_oracle(pm, pair30k, name, kind):
    rd, rf, _ = pair30k
    oref = ob.Cloud(rf)
    ob.apply_filter(oref, "SurfaceNormalDataPointsFilter", knn=10)
    ord_ = ob.Cloud(rd)
    ids, d2 = ob.kdtree_knn(rf, rd, k=1)
    st, w = ob.outlier_weights([{"TrimmedDistOutlierFilter": {"ratio": 0.85}}], d2)
    st, want = ob.minimize(kind, ord_, oref, ids, d2, w)
    assert st == 0
    ref = pm.DataPoints(rf, {"normals": oref.desc("normals")})
    e = pm.ErrorMinimizer(name)
    T = e.compute(pm.DataPoints(rd), ref, w, pm.Matches(ids, d2))
    np.testing.assert_allclose(T, want["T"], rtol=0, atol=1e-11)
    assert e._last.kept == want["kept"]
    assert e._last.weighted_point_used_ratio == pytest.approx(want["weighted_point_used_ratio"], rel=1e-15)
    assert e._last.residual == pytest.approx(want["residual"], rel=1e-11)
    if kind == ob.E_POINT_TO_PLANE_WITH_COV:
        np.testing.assert_allclose(e.getCovariance(), want["cov"], rtol=1e-8, atol=1e-18)


# ------------------------------------------------------------------ ICP (A15-A16) ---
def _run_both(pm, cfg, rd, rf, T0=None, rd_desc=None):
    icp = pm.ICP()
    icp.loadFromYaml(util.to_yaml(cfg))
    T = icp(pm.DataPoints(rd, rd_desc), pm.DataPoints(rf), T0)
    want = ob.icp_run(cfg, ob.Cloud(rd, rd_desc), ob.Cloud(rf), T0)
    return icp, T, want


def test_icp_c1_point_to_point_30k(pm, pair30k):
    rd, rf, _ = pair30k
    icp, T, want = _run_both(pm, util.C1, rd, rf)
    assert want["status"] == 0
    assert icp.last["iterations"] == want["iterations"]
    util.assert_pose_close(T, want["T"])
    assert icp.errorMinimizer.getOverlap() == pytest.approx(want["overlap"], rel=1e-12)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_icp_c2_point_to_plane_120k(pm, seed):
    rd, rf, truth = synth.scan_pair(seed, beams=64, az_steps=1875)
    icp, T, want = _run_both(pm, util.C2, rd, rf)
    assert want["status"] == 0
    assert icp.last["iterations"] == want["iterations"]
    assert icp.getMaxNumIterationsReached() == want["max_iter_reached"]
    util.assert_pose_close(T, want["T"])
    assert icp.last["residual"] == pytest.approx(want["residual"], rel=1e-7)
    assert icp.last["weighted_point_used_ratio"] == pytest.approx(want["weighted_ratio"], rel=1e-12)
    # and ICP did its job: within a few mm of the generating pose
    assert np.abs(T[:3, 3] - truth[:3, 3]).max() < 0.02


def test_icp_with_cov_and_initial_guess(pm, pair30k):
    rd, rf, truth = pair30k
    T0 = synth.pose_matrix(truth[:3, 3] + [0.05, -0.03, 0.01], 0.01, 0.0, 0.0)
    T0[:3, :3] = truth[:3, :3] @ T0[:3, :3]
    icp, T, want = _run_both(pm, util.C2_COV, rd, rf, T0)
    assert want["status"] == 0
    assert icp.last["iterations"] == want["iterations"]
    util.assert_pose_close(T, want["T"])
    cov = icp.errorMinimizer.getCovariance()
    np.testing.assert_allclose(cov, want["cov"], rtol=1e-6, atol=1e-16)
    assert np.all(np.diag(cov) > 0)


def test_icp_overlap_with_sensor_noise_descriptors(pm, pair30k):
    rd, rf, _ = pair30k
    oc = ob.Cloud(rd)
    for it in util.INPUT_FILTERS:
        (name, p), = ob._modlist([it])
        ob.apply_filter(oc, name, **p)
    desc = oc.descriptors()
    icp, T, want = _run_both(pm, util.C2, rd, rf, rd_desc=desc)
    assert icp.last["iterations"] == want["iterations"]
    util.assert_pose_close(T, want["T"])
    assert 0.0 < want["overlap"] < 1.0
    assert icp.errorMinimizer.getOverlap() == pytest.approx(want["overlap"], abs=1e-12)


def test_icp_c5_voxel_reading_trimmed_075(pm, pair120k):
    rd, rf, _ = pair120k
    icp, T, want = _run_both(pm, util.C5, rd, rf)
    assert want["status"] == 0
    assert icp.last["iterations"] == want["iterations"]
    assert icp.last["n_reading"] < rd.shape[1]
    util.assert_pose_close(T, want["T"])


def test_icp_identical_clouds_gives_identity(pm, pair30k):
    _, rf, _ = pair30k
    # every match distance is 0 -> "no outlier to filter" with a trimmed filter
    icp = pm.ICP()
    icp.loadFromYaml(util.to_yaml(util.C2))
    with pytest.raises(pm.ConvergenceError):
        icp(pm.DataPoints(rf), pm.DataPoints(rf))
    assert ob.icp_run(util.C2, ob.Cloud(rf), ob.Cloud(rf))["status"] == ob.CONVERGENCE_ERROR
    # without outlier filters the x = 0 solution takes the NaN -> identity branch (A.5)
    cfg = dict(util.C2, outlierFilters=[])
    icp, T, want = _run_both(pm, cfg, rf, rf)
    assert want["status"] == 0 and icp.last["iterations"] == want["iterations"]
    np.testing.assert_allclose(T, np.eye(4), atol=1e-6)


def test_icp_errors(pm, pair30k):
    rd, rf, _ = pair30k
    icp = pm.ICP()
    icp.loadFromYaml(util.to_yaml(dict(util.C2, referenceDataPointsFilters=[])))
    with pytest.raises(pm.InvalidField):  # point-to-plane without normals
        icp(pm.DataPoints(rd), pm.DataPoints(rf))
    icp.loadFromYaml(util.to_yaml(util.C1))
    bad = np.eye(4)
    bad[0, 0] = 2.0
    with pytest.raises(pm.TransformationError):
        icp(pm.DataPoints(rd), pm.DataPoints(rf), bad)
    bound = dict(util.C1, transformationCheckers=util.CHECKERS + [
        {"BoundTransformationChecker": {"maxRotationNorm": 0.001, "maxTranslationNorm": 0.001}}])
    icp.loadFromYaml(util.to_yaml(bound))
    with pytest.raises(pm.ConvergenceError):
        icp(pm.DataPoints(rd), pm.DataPoints(rf))
    assert ob.icp_run(bound, ob.Cloud(rd), ob.Cloud(rf))["status"] == ob.CONVERGENCE_ERROR


def test_icp_max_iterations_reached_flag(pm, pair30k):
    rd, rf, _ = pair30k
    cfg = dict(util.C1, transformationCheckers=[{"CounterTransformationChecker": {"maxIterationCount": 3}}])
    icp, T, want = _run_both(pm, cfg, rd, rf)
    assert icp.last["iterations"] == want["iterations"] == 3
    assert icp.getMaxNumIterationsReached() and want["max_iter_reached"]
    util.assert_pose_close(T, want["T"])


def test_icp_sequence_set_map(pm, pair30k):
    rd, rf, truth = pair30k
    seq = pm.ICPSequence()
    seq.loadFromYaml(util.to_yaml(util.C2))
    assert not seq.hasMap()
    seq.setMap(pm.DataPoints(rf))
    assert seq.hasMap()
    oseq = ob.IcpSequence(util.C2)
    assert oseq.set_map(ob.Cloud(rf)) == 0
    T_prev = np.eye(4)
    for step in range(2):
        T = seq(pm.DataPoints(rd), T_prev)
        want = oseq.run(ob.Cloud(rd), T_prev)
        assert seq.last["iterations"] == want["iterations"]
        util.assert_pose_close(T, want["T"])
        T_prev = want["T"]


def test_batch_equals_single_runs(pm):
    pairs = [synth.scan_pair(s, beams=16, az_steps=900 + 100 * s) for s in range(5)]  # ragged sizes
    icp = pm.ICP()
    icp.loadFromYaml(util.to_yaml(util.C2))
    rds = [pm.DataPoints(p[0]) for p in pairs]
    rfs = [pm.DataPoints(p[1]) for p in pairs]
    batch = icp.compute_batch(rds, rfs)
    for i, p in enumerate(pairs):
        T = icp(rds[i], rfs[i])
        assert np.array_equal(T, batch[i]["T"])  # independent of batching, bit for bit
        assert icp.last["iterations"] == batch[i]["iterations"]
        want = ob.icp_run(util.C2, ob.Cloud(p[0]), ob.Cloud(p[1]))
        assert batch[i]["iterations"] == want["iterations"]
        util.assert_pose_close(batch[i]["T"], want["T"])


def test_probes_match_module_by_module_calls(pm, pair30k):
    rd, rf, truth = pair30k
    oref = ob.Cloud(rf)
    ob.apply_filter(oref, "SurfaceNormalDataPointsFilter", knn=10)
    ref = pm.DataPoints(rf, {"normals": oref.desc("normals")})
    icp = pm.ICP()
    icp.loadFromYaml(util.to_yaml(util.C2))
    # Localizer::ComputeOverlapWith
    got = icp.probe_overlap(pm.DataPoints(rd), pm.DataPoints(rf), truth)
    st, want = ob.probe_overlap(util.C2, ob.Cloud(rd), ob.Cloud(rf), truth)
    assert st == 0 and got == pytest.approx(want, rel=1e-15)
    # the same, spelled the way Localizer.hpp:309-347 spells it
    tmp = pm.ICP()
    tmp.loadFromYaml(util.to_yaml(util.C2))
    reference = pm.DataPoints(rf)
    tmp.referenceDataPointsFilters.init()
    tmp.referenceDataPointsFilters.apply(reference)
    tmp.matcher.init(reference)
    reading = pm.DataPoints(rd)
    tmp.readingDataPointsFilters.init()
    tmp.readingDataPointsFilters.apply(reading)
    reading = pm.RigidTransformation().compute(reading, truth)
    tmp.readingStepDataPointsFilters.init()
    tmp.readingStepDataPointsFilters.apply(reading)
    matches = tmp.matcher.findClosests(reading)
    weights = tmp.outlierFilters.compute(reading, reference, matches)
    ee = tmp.errorMinimizer.errorElements(reading, reference, weights, matches)
    assert ee.weightedPointUsedRatio == pytest.approx(want, rel=1e-15)
    # LoopCloser::ComputeResidualError
    got = icp.probe_residual(pm.DataPoints(rd), ref, truth)
    st, want = ob.probe_residual(util.C2, ob.Cloud(rd), oref, truth)
    assert st == 0 and got == pytest.approx(want, rel=1e-10)


def test_assemble_local_map(pm, pair30k):
    rd, rf, truth = pair30k
    a, b = pm.DataPoints(rf), pm.DataPoints(rd)
    out = pm.assemble_local_map([a, b], [np.eye(4), truth])
    oc = ob.Cloud(rd)
    ob.rigid_transform(oc, truth)
    assert out.getNbPoints() == rf.shape[1] + rd.shape[1]
    assert np.array_equal(out.features[:, :rf.shape[1]], rf)
    assert np.array_equal(out.features[:, rf.shape[1]:].view(np.uint32), oc.features.view(np.uint32))


# ------------------------------------------------ committed golden vectors (tests/golden) ---
def test_gpu_against_committed_golden_vectors(pm):
    import os
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "pair4000.npz"))
    rd, rf = G["reading"], G["reference"]
    m = pm.Matcher("KDTreeMatcher", {"knn": 1})
    m.init(pm.DataPoints(rf))
    got = m.findClosests(pm.DataPoints(rd))
    assert np.array_equal(got.ids, G["knn1_ids"]) and np.array_equal(got.dists, G["knn1_d2"])
    m5 = pm.Matcher("KDTreeMatcher", {"knn": 5})
    m5.init(pm.DataPoints(rf))
    got = m5.findClosests(pm.DataPoints(rf))
    assert np.array_equal(got.ids, G["knn5_ids"]) and np.array_equal(got.dists, G["knn5_d2"])
    dp = pm.DataPoints(rf)
    f = pm.DataPointsFilters()
    f.append("SurfaceNormalDataPointsFilter", {"knn": 10, "keepDensities": 1, "keepEigenValues": 1})
    f.apply(dp)
    assert np.array_equal(dp.getDescriptorByName("normals"), G["normals"])
    assert np.array_equal(dp.getDescriptorByName("densities"), G["densities"])
    assert np.array_equal(dp.getDescriptorByName("eigValues"), G["eigvalues"])
    vox = pm.DataPoints(rd)
    f = pm.DataPointsFilters()
    f.append("VoxelGridDataPointsFilter", {"vSizeX": 0.5, "vSizeY": 0.5, "vSizeZ": 0.5})
    f.apply(vox)
    assert np.array_equal(vox.features, G["voxel_features"])
    o = pm.OutlierFilters()
    o.append("TrimmedDistOutlierFilter", {"ratio": 0.85})
    w = o.compute(pm.DataPoints(rd), pm.DataPoints(rf), pm.Matches(G["knn1_ids"], G["knn1_d2"]))
    assert np.array_equal(w, G["trimmed_weights"])
    for name, cfg in (("c1", util.C1), ("c2", util.C2), ("c2cov", util.C2_COV)):
        icp = pm.ICP()
        icp.loadFromYaml(util.to_yaml(cfg))
        T = icp(pm.DataPoints(rd), pm.DataPoints(rf))
        assert icp.last["iterations"] == int(G[name + "_iterations"])
        util.assert_pose_close(T, G[name + "_T"])
        if name == "c2cov":
            np.testing.assert_allclose(icp.errorMinimizer.getCovariance(), G[name + "_cov"], rtol=1e-6, atol=1e-18)


def test_full_size_properties_120k(pm, pair120k):
    """Size-independent properties at BASELINE.json's full size."""
    rd, rf, truth = pair120k
    # (1) self-kNN: every point is its own nearest neighbour at distance 0 and the list is sorted
    m = pm.Matcher("KDTreeMatcher", {"knn": 4})
    ref = pm.DataPoints(rf)
    m.init(ref)
    got = m.findClosests(ref)
    assert np.array_equal(got.ids[0], np.arange(rf.shape[1])) and (got.dists[0] == 0).all()
    assert (np.diff(got.dists, axis=0) >= 0).all()
    # (2) kNN is invariant to a permutation of the reference (ids map through the permutation)
    perm = np.random.default_rng(0).permutation(rf.shape[1])
    m1 = pm.Matcher("KDTreeMatcher", {"knn": 1})
    m1.init(ref)
    a = m1.findClosests(pm.DataPoints(rd))
    m1.init(pm.DataPoints(rf[:, perm]))
    b = m1.findClosests(pm.DataPoints(rd))
    assert np.array_equal(a.dists, b.dists)
    same = perm[b.ids[0]] == a.ids[0]
    assert same.mean() > 0.9999  # only exact-distance ties may resolve to another index
    # (3) trimmed weights keep exactly the requested fraction of the valid matches
    o = pm.OutlierFilters()
    o.append("TrimmedDistOutlierFilter", {"ratio": 0.85})
    w = o.compute(pm.DataPoints(rd), ref, a)
    valid = (a.dists > 0) & np.isfinite(a.dists)
    assert abs(w[valid].mean() - 0.85) < 2e-5
    # (4) registering a cloud against a rigidly moved copy of itself returns the motion
    T = synth.pose_matrix([0.1, -0.05, 0.02], 0.01, 0.002, -0.003)
    moved = (np.linalg.inv(T) @ rf.astype(np.float64)).astype(np.float32)
    icp = pm.ICP()
    icp.loadFromYaml(util.to_yaml(dict(util.C2, transformationCheckers=[
        {"CounterTransformationChecker": {"maxIterationCount": 60}},
        {"DifferentialTransformationChecker": {"minDiffRotErr": 1e-6, "minDiffTransErr": 1e-6}}])))
    out = icp(pm.DataPoints(moved), ref)
    np.testing.assert_allclose(out, T, atol=2e-5)


# ------------------------------------------------------ BASELINE configs C3 / C5, threading ---
def test_c5_one_million_point_scan_to_map(pm):
    """configs[4]: ~1M-point reading vs 1M-point map, voxel subsampling, trimmed 0.75."""
    scene = synth.make_scene(77)
    T_ref = synth.pose_matrix([0.0, 0.0, synth.SENSOR_HEIGHT])
    T_rd = synth.pose_matrix([0.25, -0.15, synth.SENSOR_HEIGHT + 0.03], np.deg2rad(1.5), 0.004, -0.006)
    rf = synth.velodyne_scan(77, 0, T_ref, beams=128, az_steps=7812, scene=scene)
    rd = synth.velodyne_scan(77, 1, T_rd, beams=128, az_steps=7812, scene=scene)
    assert rd.shape[1] == 999936
    icp, T, want = _run_both(pm, util.C5, rd, rf)
    assert want["status"] == 0
    assert icp.last["iterations"] == want["iterations"]
    assert icp.last["n_reference"] == rf.shape[1] and icp.last["n_reading"] < 0.5 * rd.shape[1]
    util.assert_pose_close(T, want["T"])
    truth = np.linalg.inv(T_ref) @ T_rd
    assert np.abs(T[:3, 3] - truth[:3, 3]).max() < 0.05  # the default checkers stop within a few cm


def test_c3_sequential_scan_to_map_odometry(pm):
    """configs[2] in miniature: ICPSequence against a 3-keyframe local map assembled on the
    device the way LocalMap::BuildCloudFromData does, each scan seeded by the previous pose
    (Localizer.hpp:119-126); the oracle walks the same chain."""
    scene = synth.make_scene(5)
    poses = synth.trajectory(8, step=0.5, turn_deg=2.0)
    scans = [synth.velodyne_scan(5, i, poses[i], beams=32, az_steps=900, scene=scene) for i in range(8)]
    kf = [0, 1, 2]
    T_ref_kf = [np.linalg.inv(poses[kf[-1]]) @ poses[k] for k in kf]  # keyframes in the reference-kf frame
    filt = pm.DataPointsFilters(util.to_yaml(util.INPUT_FILTERS[:3]))
    dps = []
    ocs = []
    for k in kf:
        dp = pm.DataPoints(scans[k])
        filt.apply(dp)
        dps.append(dp)
        oc = ob.Cloud(scans[k])
        for it in util.INPUT_FILTERS[:3]:
            (name, p), = ob._modlist([it])
            ob.apply_filter(oc, name, **p)
        ocs.append(oc)
    # reference keyframe first, the others transformed into its frame and concatenated
    local_map = pm.assemble_local_map([dps[2], dps[1], dps[0]], [np.eye(4), T_ref_kf[1], T_ref_kf[0]])
    omap = ocs[2].copy()
    for oc, T in ((ocs[1], T_ref_kf[1]), (ocs[0], T_ref_kf[0])):
        t = oc.copy()
        assert ob.rigid_transform(t, T) == 0
        ob.lib().orc_cloud_concatenate(omap.ptr, t.ptr)
    assert local_map.getNbPoints() == omap.n == 3 * scans[0].shape[1]
    assert np.array_equal(local_map.features.view(np.uint32), omap.features.view(np.uint32))
    assert np.array_equal(local_map.getDescriptorByName("normals").view(np.uint32), omap.desc("normals").view(np.uint32))
    cfg = dict(util.C2, referenceDataPointsFilters=[])  # the local map already carries normals
    seq = pm.ICPSequence()
    seq.loadFromYaml(util.to_yaml(cfg))
    seq.setMap(local_map)
    oseq = ob.IcpSequence(cfg)
    assert oseq.set_map(omap) == 0
    T_prev = np.eye(4)
    for i in range(3, 8):
        guess = T_prev @ (np.linalg.inv(poses[i - 1]) @ poses[i]) if i > 3 else np.linalg.inv(poses[2]) @ poses[3]
        T = seq(pm.DataPoints(scans[i]), guess)
        want = oseq.run(ob.Cloud(scans[i]), guess)
        assert want["status"] == 0 and seq.last["iterations"] == want["iterations"]
        util.assert_pose_close(T, want["T"])
        truth = np.linalg.inv(poses[2]) @ poses[i]
        assert np.abs(T[:3, 3] - truth[:3, 3]).max() < 0.1  # sanity only; parity is against the oracle
        T_prev = want["T"]


def test_two_host_threads_two_contexts(pm):
    """pgslam-MT: localizer and loop closer register concurrently from two host threads,
    each on its own objects (LocalizerMT.hpp:47, LoopCloserMT.hpp:41)."""
    import threading
    pairs = [synth.scan_pair(30 + s, beams=16, az_steps=700) for s in range(2)]
    serial = []
    for rd, rf, _ in pairs:
        icp = pm.ICP()
        icp.loadFromYaml(util.to_yaml(util.C2))
        serial.append((icp(pm.DataPoints(rd), pm.DataPoints(rf)), icp.last["iterations"]))
    out = [None, None]

    def worker(i):
        c = pm.Context(0)
        icp = pm.ICP(c)
        icp.loadFromYaml(util.to_yaml(util.C2))
        rd, rf, _ = pairs[i]
        for _ in range(5):
            T = icp(pm.DataPoints(rd, ctx=c), pm.DataPoints(rf, ctx=c))
        out[i] = (T, icp.last["iterations"])

    th = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    for i in range(2):
        assert out[i] is not None and out[i][1] == serial[i][1]
        assert np.array_equal(out[i][0], serial[i][0])


def test_pinned_uploads_from_a_second_thread_while_a_batch_runs(pm):
    """The end-to-end pipeline of bench.py: the next batch is uploaded from pinned host memory
    (pgs_cloud_create mode 2, side stream) by a helper thread while the current batch computes on
    the same context; results must equal those of plain synchronous uploads, batch after batch."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    ctx = pm.Context(0)
    icp = pm.ICP(ctx)
    icp.loadFromYaml(util.to_yaml(util.C2))
    pairs = [synth.scan_pair(60 + s, beams=16, az_steps=600)[:2] for s in range(8)]
    want = icp.compute_batch([pm.DataPoints(rd, ctx=ctx) for rd, _ in pairs], [pm.DataPoints(rf, ctx=ctx) for _, rf in pairs])
    host = [(torch.from_numpy(np.ascontiguousarray(rd.T)).pin_memory(), torch.from_numpy(np.ascontiguousarray(rf.T)).pin_memory())
            for rd, rf in pairs]

    def upload():
        return ([pm.DataPoints(ctx=ctx, pinned_host_ptr=a.data_ptr(), n=a.shape[0]) for a, _ in host],
                [pm.DataPoints(ctx=ctx, pinned_host_ptr=b.data_ptr(), n=b.shape[0]) for _, b in host])

    with ThreadPoolExecutor(max_workers=1) as pool:
        nxt = pool.submit(upload)
        for _ in range(6):
            rds, rfs = nxt.result()
            nxt = pool.submit(upload)
            got = icp.compute_batch(rds, rfs)
            for g, w in zip(got, want):
                assert g["status"] == 0 and g["iterations"] == w["iterations"]
                assert np.array_equal(g["T"], w["T"])
        nxt.result()


# ------------------------------------------------ SurfaceNormalOutlierFilter (breadth row F4) ---
def test_surface_normal_outlier_filter_module_and_fused(pm, pair30k):
    rd, rf, _ = pair30k
    ord_, orf = ob.Cloud(rd), ob.Cloud(rf)
    for oc in (ord_, orf):
        for it in util.INPUT_FILTERS[:3]:
            (name, p), = ob._modlist([it])
            ob.apply_filter(oc, name, **p)
    rdd, rfd = ord_.descriptors(), orf.descriptors()
    ids, d2 = ob.kdtree_knn(rf, rd, k=1)
    chain = [{"TrimmedDistOutlierFilter": {"ratio": 0.9}}, {"SurfaceNormalOutlierFilter": {"maxAngle": 0.35}}]
    o = pm.OutlierFilters()
    for it in chain:
        (name, p), = ob._modlist([it])
        o.append(name, p)
    got = o.compute(pm.DataPoints(rd, rdd), pm.DataPoints(rf, rfd), pm.Matches(ids, d2))
    st, want = ob.outlier_weights_full(chain, ord_, orf, ids, d2)
    assert st == 0 and np.array_equal(got, want)
    assert 0.3 < want.mean() < 0.9  # the normal test rejects a real share of the matches
    # only the normal filter, and clouds without normals (filter is skipped, weights stay 1)
    o2 = pm.OutlierFilters()
    o2.append("SurfaceNormalOutlierFilter", {"maxAngle": 0.35})
    got2 = o2.compute(pm.DataPoints(rd, rdd), pm.DataPoints(rf, rfd), pm.Matches(ids, d2))
    st, want2 = ob.outlier_weights_full(chain[1:], ord_, orf, ids, d2)
    assert np.array_equal(got2, want2)
    got3 = o2.compute(pm.DataPoints(rd), pm.DataPoints(rf), pm.Matches(ids, d2))
    assert (got3 == 1).all()
    # inside the fused loop
    cfg = dict(util.C2, referenceDataPointsFilters=[], outlierFilters=chain)
    icp = pm.ICP()
    icp.loadFromYaml(util.to_yaml(cfg))
    T = icp(pm.DataPoints(rd, rdd), pm.DataPoints(rf, rfd))
    want = ob.icp_run(cfg, ord_, orf)
    assert want["status"] == 0 and icp.last["iterations"] == want["iterations"]
    util.assert_pose_close(T, want["T"])
    assert icp.last["weighted_point_used_ratio"] == pytest.approx(want["weighted_ratio"], rel=1e-12)


# ------------------------------------------------------------------ edge cases ---
def _status_of(pm, fn):
    try:
        fn()
        return 0
    except pm.PointMatcherError as e:
        return e.status


@pytest.mark.parametrize("n_ref,n_rd", [(1, 5), (3, 1), (7, 7), (9, 40), (100, 3)])
def test_icp_tiny_clouds_match_the_oracle(pm, n_ref, n_rd):
    g = np.random.default_rng(n_ref * 100 + n_rd)
    rf = np.ones((4, n_ref), np.float32)
    rf[:3] = g.normal(size=(3, n_ref)).astype(np.float32)
    rd = np.ones((4, n_rd), np.float32)
    rd[:3] = (g.normal(size=(3, n_rd)) + 0.05).astype(np.float32)
    cfg = dict(util.C1, outlierFilters=[])
    want = ob.icp_run(cfg, ob.Cloud(rd), ob.Cloud(rf))
    icp = pm.ICP()
    icp.loadFromYaml(util.to_yaml(cfg))
    st = _status_of(pm, lambda: icp(pm.DataPoints(rd), pm.DataPoints(rf)))
    assert st == want["status"]
    assert icp.last["iterations"] == want["iterations"]
    if st == 0:
        np.testing.assert_allclose(icp.last["T"], want["T"], atol=1e-9)


def test_icp_empty_clouds_are_convergence_errors(pm, pair30k):
    rd, rf, _ = pair30k
    empty = np.ones((4, 0), np.float32)
    icp = pm.ICP()
    icp.loadFromYaml(util.to_yaml(util.C1))
    with pytest.raises(pm.ConvergenceError):
        icp(pm.DataPoints(empty), pm.DataPoints(rf))
    with pytest.raises(pm.ConvergenceError):
        icp(pm.DataPoints(rd), pm.DataPoints(empty))
    # filters and the matcher accept empty clouds
    dp = pm.DataPoints(empty)
    pm.DataPointsFilters(util.to_yaml(util.INPUT_FILTERS[:1])).apply(dp)
    assert dp.getNbPoints() == 0
    m = pm.Matcher("KDTreeMatcher", {"knn": 2})
    m.init(pm.DataPoints(rf))
    assert m.findClosests(pm.DataPoints(empty)).ids.shape == (2, 0)


@pytest.mark.parametrize("name", ["C1", "C2"])
def test_icp_with_nan_points_follows_the_oracle(pm, pair30k, name):
    """No RemoveNaN filter in the chain: NaN reading points never match and the registration goes on;
    NaN reference points poison the reference mean, which is a convergence error on both sides."""
    rd, rf, _ = pair30k
    g = np.random.default_rng(1)
    bad = g.choice(rd.shape[1], 40, replace=False)
    rdn, rfn = rd.copy(), rf.copy()
    rdn[g.integers(0, 3, bad.size), bad] = np.nan
    rfn[g.integers(0, 3, bad.size), bad] = np.nan
    cfgd = getattr(util, name)
    want = ob.icp_run(ob.config_from_dict(cfgd), ob.Cloud(rdn), ob.Cloud(rf))
    assert want["status"] == 0
    icp = pm.ICP()
    icp.loadFromYaml(util.to_yaml(cfgd))
    T = icp(pm.DataPoints(rdn), pm.DataPoints(rf))
    assert icp.last["iterations"] == want["iterations"]
    np.testing.assert_allclose(T, want["T"], atol=1e-9)
    assert ob.icp_run(ob.config_from_dict(cfgd), ob.Cloud(rd), ob.Cloud(rfn))["status"] != 0
    with pytest.raises(pm.ConvergenceError):
        icp(pm.DataPoints(rd), pm.DataPoints(rfn))


def test_icp_with_matcher_max_dist_and_filter_chain(pm, pair30k):
    """maxDist on the matcher leaves unmatched points (id -1, dist inf) that every later
    stage has to skip the same way; several outlier filters multiply."""
    rd, rf, _ = pair30k
    cfg = dict(util.C2, matcher={"KDTreeMatcher": {"knn": 1, "maxDist": 0.35}},
               outlierFilters=[{"MedianDistOutlierFilter": {"factor": 4}}, {"MaxDistOutlierFilter": {"maxDist": 0.3}},
                               {"MinDistOutlierFilter": {"minDist": 0.001}}])
    icp, T, want = _run_both(pm, cfg, rd, rf)
    assert want["status"] == 0 and icp.last["iterations"] == want["iterations"]
    util.assert_pose_close(T, want["T"])
    assert icp.last["point_used_ratio"] == pytest.approx(want["point_used_ratio"], rel=1e-12)
    assert icp.last["point_used_ratio"] < 0.9
    # a radius so small that nothing matches: "no outlier to filter"
    cfg2 = dict(util.C2, matcher={"KDTreeMatcher": {"knn": 1, "maxDist": 1e-7}})
    icp.loadFromYaml(util.to_yaml(cfg2))
    with pytest.raises(pm.ConvergenceError):
        icp(pm.DataPoints(rd), pm.DataPoints(rf))
    assert ob.icp_run(cfg2, ob.Cloud(rd), ob.Cloud(rf))["status"] == ob.CONVERGENCE_ERROR
    # ... and without outlier filters: "no point to minimize"
    cfg3 = dict(cfg2, outlierFilters=[])
    icp.loadFromYaml(util.to_yaml(cfg3))
    with pytest.raises(pm.ConvergenceError):
        icp(pm.DataPoints(rd), pm.DataPoints(rf))
    assert ob.icp_run(cfg3, ob.Cloud(rd), ob.Cloud(rf))["status"] == ob.CONVERGENCE_ERROR


def test_batch_with_failing_and_ragged_pairs(pm, pair30k):
    rd, rf, _ = pair30k
    small = synth.scan_pair(9, beams=16, az_steps=64)
    icp = pm.ICP()
    icp.loadFromYaml(util.to_yaml(util.C2))
    rds = [pm.DataPoints(rd), pm.DataPoints(rf), pm.DataPoints(small[0])]
    rfs = [pm.DataPoints(rf), pm.DataPoints(rf), pm.DataPoints(small[1])]  # pair 1: identical clouds
    res = icp.compute_batch(rds, rfs)
    assert [r["status"] for r in res] == [0, pm.CONVERGENCE_ERROR, 0]
    for i in (0, 2):
        want = ob.icp_run(util.C2, ob.Cloud([rd, None, small[0]][i]), ob.Cloud([rf, None, small[1]][i]))
        assert res[i]["iterations"] == want["iterations"]
        util.assert_pose_close(res[i]["T"], want["T"])


def test_planar_scene_takes_the_rank_deficient_solve(pm):
    """A single plane constrains 3 of 6 DOF: the normal equations are singular and both
    sides must fall back to the minimum-norm solution (A.5)."""
    g = np.random.default_rng(4)
    n = 5000
    rf = np.ones((4, n), np.float32)
    rf[0] = g.uniform(-5, 5, n)
    rf[1] = g.uniform(-5, 5, n)
    rf[2] = 0.0
    T = synth.pose_matrix([0.0, 0.0, 0.07], 0.0, 0.01, -0.008)
    rd = (np.linalg.inv(T) @ rf.astype(np.float64)).astype(np.float32)
    nrm = np.zeros((3, n), np.float32)
    nrm[2] = 1.0
    cfg = dict(util.C2, referenceDataPointsFilters=[], outlierFilters=[])
    icp = pm.ICP()
    icp.loadFromYaml(util.to_yaml(cfg))
    Tg = icp(pm.DataPoints(rd), pm.DataPoints(rf, {"normals": nrm}))
    want = ob.icp_run(cfg, ob.Cloud(rd), ob.Cloud(rf, {"normals": nrm}))
    assert want["status"] == 0 and icp.last["iterations"] == want["iterations"]
    util.assert_pose_close(Tg, want["T"], 1e-7, 1e-7)
    assert abs(Tg[2, 3] - 0.07) < 1e-3  # the constrained DOF is recovered


# ------------------------------------------- chains the fused loop gained in round 1b ---
@pytest.mark.parametrize("k", [2, 3, 7])
def test_icp_fused_loop_with_knn_greater_than_one(pm, pair30k, k):
    rd, rf, _ = pair30k
    cfg = dict(util.C2, matcher={"KDTreeMatcher": {"knn": k, "epsilon": 0}})
    icp, T, want = _run_both(pm, cfg, rd, rf)
    assert want["status"] == 0
    assert icp.last["iterations"] == want["iterations"]
    util.assert_pose_close(T, want["T"])
    assert icp.last["point_used_ratio"] == pytest.approx(want["point_used_ratio"], rel=1e-12)
    assert icp.last["weighted_point_used_ratio"] == pytest.approx(want["weighted_ratio"], rel=1e-12)
    assert icp.last["residual"] == pytest.approx(want["residual"], rel=1e-7)


def test_icp_knn_with_max_dist_cov_and_point_to_point(pm, pair30k):
    rd, rf, _ = pair30k
    for cfg in (dict(util.C2_COV, matcher={"KDTreeMatcher": {"knn": 4, "maxDist": 0.5}}),
                dict(util.C1, matcher={"KDTreeMatcher": {"knn": 2}})):
        icp, T, want = _run_both(pm, cfg, rd, rf)
        assert want["status"] == 0
        assert icp.last["iterations"] == want["iterations"]
        util.assert_pose_close(T, want["T"])
        assert icp.last["weighted_point_used_ratio"] == pytest.approx(want["weighted_ratio"], rel=1e-12)
        if "WithCov" in str(cfg["errorMinimizer"]):
            np.testing.assert_allclose(icp.errorMinimizer.getCovariance(), want["cov"], rtol=1e-6, atol=1e-16)


def test_icp_reading_step_filters(pm, pair30k):
    rd, rf, _ = pair30k
    cfg = dict(util.C2, readingStepDataPointsFilters=[
        {"MaxDistDataPointsFilter": {"maxDist": 20.0}},
        {"RandomSamplingDataPointsFilter": {"prob": 0.6, "seed": 11}}])
    icp, T, want = _run_both(pm, cfg, rd, rf)
    assert want["status"] == 0
    assert icp.last["iterations"] == want["iterations"]
    assert 0 < icp.last["n_reading"] < rd.shape[1]
    util.assert_pose_close(T, want["T"])
    assert icp.last["weighted_point_used_ratio"] == pytest.approx(want["weighted_ratio"], rel=1e-12)
    # a step filter that removes everything: nothing to match -> ConvergenceError on both sides
    none = dict(util.C2, readingStepDataPointsFilters=[{"MaxDistDataPointsFilter": {"maxDist": 1e-3}}])
    icp = pm.ICP()
    icp.loadFromYaml(util.to_yaml(none))
    with pytest.raises(pm.ConvergenceError):
        icp(pm.DataPoints(rd), pm.DataPoints(rf))
    assert ob.icp_run(none, ob.Cloud(rd), ob.Cloud(rf))["status"] == ob.CONVERGENCE_ERROR


@pytest.mark.parametrize("mode", ["force2D", "force4DOF"])
def test_icp_point_to_plane_forced_modes(pm, pair30k, mode):
    rd, rf, _ = pair30k
    cfg = dict(util.C2, errorMinimizer={"PointToPlaneErrorMinimizer": {mode: 1}})
    icp, T, want = _run_both(pm, cfg, rd, rf)
    assert want["status"] == 0
    assert icp.last["iterations"] == want["iterations"]
    util.assert_pose_close(T, want["T"])
    assert icp.last["residual"] == pytest.approx(want["residual"], rel=1e-7)
    np.testing.assert_allclose(T[2, :3], [0, 0, 1], atol=1e-12)
    # module-level compute() of the same minimizer
    oref = ob.Cloud(rf)
    ob.apply_filter(oref, "SurfaceNormalDataPointsFilter", knn=10)
    ids, d2 = ob.kdtree_knn(rf, rd, k=1)
    st, w = ob.outlier_weights([{"TrimmedDistOutlierFilter": {"ratio": 0.85}}], d2)
    st, ref_out = ob.minimize(ob.E_POINT_TO_PLANE, ob.Cloud(rd), oref, ids, d2, w, force_mode=1 if mode == "force2D" else 2)
    e = pm.ErrorMinimizer("PointToPlaneErrorMinimizer", {mode: 1})
    Tm = e.compute(pm.DataPoints(rd), pm.DataPoints(rf, {"normals": oref.desc("normals")}), w, pm.Matches(ids, d2))
    np.testing.assert_allclose(Tm, ref_out["T"], rtol=0, atol=1e-11)
    assert e._last.residual == pytest.approx(ref_out["residual"], rel=1e-11)


def test_icp_point_to_plane_with_cov_force4dof(pm, pair30k):
    """force4DOF with the covariance estimate (F4 remainder): pose, iterations and the 6x6 Censi
    covariance against the oracle; force2D stays rejected with the covariance."""
    rd, rf, _ = pair30k
    cfg = dict(util.C2, errorMinimizer={"PointToPlaneWithCovErrorMinimizer": {"force4DOF": 1, "sensorStdDev": 0.02}})
    icp, T, want = _run_both(pm, cfg, rd, rf)
    assert want["status"] == 0 and icp.last["iterations"] == want["iterations"]
    util.assert_pose_close(T, want["T"])
    np.testing.assert_allclose(T[2, :3], [0, 0, 1], atol=1e-12)
    cov = icp.errorMinimizer.getCovariance()
    np.testing.assert_allclose(cov, want["cov"], rtol=1e-6, atol=1e-16)
    assert np.all(np.diag(cov) > 0)
    bad = pm.ICP()
    bad.loadFromYaml(util.to_yaml(dict(util.C2, errorMinimizer={"PointToPlaneWithCovErrorMinimizer": {"force2D": 1}})))
    with pytest.raises(pm.InvalidParameter):  # refused when the chain is instantiated for its first registration
        bad(pm.DataPoints(rd), pm.DataPoints(rf))


def test_knn_k10_bit_exact_120k(pm, pair120k):
    """k = 10 at full size (120k was k = 1 only): ids and distances against the oracle."""
    rd, rf, _ = pair120k
    m = pm.Matcher("KDTreeMatcher", {"knn": 10})
    m.init(pm.DataPoints(rf))
    got = m.findClosests(pm.DataPoints(rd))
    ids, d2 = ob.kdtree_knn(rf, rd, k=10)
    assert np.array_equal(got.ids, ids)
    assert np.array_equal(got.dists.view(np.uint32), d2.view(np.uint32))


# ------------------------------------------------ matcher scheduling variants ---
def _bits(res):
    keys = ("T", "covariance", "iterations", "status", "overlap", "weighted_point_used_ratio", "point_used_ratio",
            "residual", "n_reading", "n_reference")
    return [np.asarray(res[k]).tobytes() for k in keys]


@pytest.mark.parametrize("mode", [1, 2, 3, 4, 5])
def test_match_modes_give_bit_identical_registrations(pm, ctx, pair30k, mode):
    """The cell-guided and persistent matchers (pgs_ctx_set_option "match_mode") must return the
    same exact nearest neighbours as the default one: whole results compared bit for bit."""
    cases = [(util.C1, pair30k[0], pair30k[1])]
    rd, rf, _ = synth.scan_pair(2, beams=64, az_steps=1875)
    cases.append((util.C2, rd, rf))
    g = np.random.default_rng(7)
    for n_ref, n_rd in ((1, 5), (7, 7), (9, 40), (17, 33), (100, 3), (1000, 900)):
        a = np.ones((4, n_ref), np.float32); a[:3] = g.uniform(-2, 2, (3, n_ref))
        b = np.ones((4, n_rd), np.float32); b[:3] = g.uniform(-2, 2, (3, n_rd))
        cases.append((util.C1, b, a))
    try:
        for cfg, rd, rf in cases:
            out = {}
            for m in (0, mode):
                ctx.set_option("match_mode", m)
                icp = pm.ICP()
                icp.loadFromYaml(util.to_yaml(cfg))
                try:
                    icp(pm.DataPoints(rd), pm.DataPoints(rf))
                except pm.PointMatcherError:
                    pass
                out[m] = icp.last
            assert _bits(out[0]) == _bits(out[mode]), (mode, rd.shape, rf.shape, out[0]["iterations"], out[mode]["iterations"])
    finally:
        ctx.set_option("match_mode", 0)


@pytest.mark.parametrize("mode", [1, 3, 4, 5])
def test_match_modes_batch_with_ragged_pairs(pm, ctx, mode):
    """A batch of pairs of different sizes through the persistent matcher: ranges of inactive /
    short pairs are skipped, results equal the default matcher's bit for bit."""
    pairs = [synth.scan_pair(20 + i, beams=8 + 4 * i, az_steps=300 + 50 * i)[:2] for i in range(6)]
    try:
        out = {}
        for m in (0, mode):
            ctx.set_option("match_mode", m)
            icp = pm.ICP()
            icp.loadFromYaml(util.to_yaml(util.C2))
            out[m] = icp.compute_batch([pm.DataPoints(r) for r, _ in pairs], [pm.DataPoints(f) for _, f in pairs])
        for a, b in zip(out[0], out[mode]):
            assert _bits(a) == _bits(b)
    finally:
        ctx.set_option("match_mode", 0)


@pytest.mark.parametrize("k", [33, 64, 100])
def test_knn_above_32_matcher_and_surface_normals(pm, pair30k, k):
    """knn > 32 (F4 remainder): the run-time K-list path, against the oracle bit for bit."""
    rd, rf, _ = pair30k
    rd, rf = rd[:, :6000], rf[:, :9000]
    m = pm.Matcher("KDTreeMatcher", {"knn": k})
    m.init(pm.DataPoints(rf))
    got = m.findClosests(pm.DataPoints(rd))
    ids, d2 = ob.kdtree_knn(rf, rd, k=k)
    assert np.array_equal(got.ids, ids)
    assert np.array_equal(got.dists.view(np.uint32), d2.view(np.uint32))
    if k <= 64:
        dp = pm.DataPoints(rf)
        f = pm.DataPointsFilters()
        f.append("SurfaceNormalDataPointsFilter", {"knn": k, "keepDensities": 1})
        f.apply(dp)
        oc = ob.Cloud(rf)
        ob.apply_filter(oc, "SurfaceNormalDataPointsFilter", knn=k, keepDensities=1)
        _cmp_cloud(dp, oc)


@pytest.mark.parametrize("ext", ["csv", "vtk", "ply"])
def test_cloud_files_through_the_c_abi(pm, pair30k, tmp_path, ext):
    """DataPoints::load / save in the C ABI (pgs_cloud_load / pgs_cloud_save, F4 remainder): a filtered cloud
    written by the library is read back bit for bit by the library AND by the independent Python parser,
    and a file written by the Python module loads to the same cloud."""
    from pgslam_b200 import cloud_io
    rd, _, _ = pair30k
    dp = pm.DataPoints(rd[:, :3000])
    f = pm.DataPointsFilters(util.to_yaml(util.INPUT_FILTERS))
    f.apply(dp)
    path = str(tmp_path / f"a.{ext}")
    dp.save(path)
    feats, desc = cloud_io.load(path)
    assert np.array_equal(feats, dp.features) and set(desc) == set(dp.descriptors)
    for k, v in dp.descriptors.items():
        assert np.array_equal(desc[k], v), k
    back = pm.DataPoints.load(path)
    assert np.array_equal(back.features, dp.features)
    assert {k: v.tobytes() for k, v in back.descriptors.items()} == {k: v.tobytes() for k, v in dp.descriptors.items()}
    path2 = str(tmp_path / f"b.{ext}")
    cloud_io.save(path2, dp.features, dp.descriptors)
    again = pm.DataPoints.load(path2)
    assert np.array_equal(again.features, dp.features) and again.getNbPoints() == 3000
    if ext == "ply":  # binary little-endian PLY, as other tools write it
        n = 100
        rec = np.zeros(n, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("intensity", "<u1")])
        rec["x"], rec["y"], rec["z"], rec["intensity"] = np.arange(n), 2 * np.arange(n), -np.arange(n), np.arange(n) % 250
        p3 = tmp_path / "bin.ply"
        with open(p3, "wb") as fh:
            fh.write(b"ply\nformat binary_little_endian 1.0\nelement vertex 100\nproperty float x\nproperty float y\n"
                     b"property float z\nproperty uchar intensity\nend_header\n")
            fh.write(rec.tobytes())
        c = pm.DataPoints.load(str(p3))
        assert c.getNbPoints() == n and np.array_equal(c.features[1], 2 * np.arange(n, dtype=np.float32))
        assert np.array_equal(c.getDescriptorByName("intensity")[0], (np.arange(n) % 250).astype(np.float32))


def test_sort_based_kd_levels_give_the_same_neighbours(pm, pair120k):
    """PGS_KD_SORT_LEVELS=1 (the comparison build of the index: sorted global levels) is read once per
    process, so it runs in a child: any valid tree must return the same exact neighbours."""
    import hashlib
    import os
    import subprocess
    import sys
    rd, rf, _ = pair120k
    m = pm.Matcher("KDTreeMatcher", {"knn": 3})
    m.init(pm.DataPoints(rf))
    got = m.findClosests(pm.DataPoints(rd[:, :20000]))
    want = hashlib.sha256(np.ascontiguousarray(got.ids).tobytes() + np.ascontiguousarray(got.dists).tobytes()).hexdigest()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import hashlib, numpy as np\n"
        "from pgslam_b200 import pm, synth\n"
        "rd, rf, _ = synth.scan_pair(5, beams=64, az_steps=1875)\n"
        "m = pm.Matcher('KDTreeMatcher', {'knn': 3})\n"
        "m.init(pm.DataPoints(rf))\n"
        "g = m.findClosests(pm.DataPoints(rd[:, :20000]))\n"
        "print(hashlib.sha256(np.ascontiguousarray(g.ids).tobytes() + np.ascontiguousarray(g.dists).tobytes()).hexdigest())\n")
    env = dict(os.environ, PGS_KD_SORT_LEVELS="1", PYTHONPATH=root)
    out = subprocess.run([sys.executable, "-c", code], env=env, cwd=root, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().splitlines()[-1] == want
