"""ctypes binding of the CPU oracle (oracle/oracle.h).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(pgslam_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

OK, CONVERGENCE_ERROR, TRANSFORMATION_ERROR, INVALID_PARAMETER, INVALID_FIELD = range(5)

F_RANDOM_SAMPLING, F_VOXEL_GRID, F_SURFACE_NORMAL, F_OBSERVATION_DIRECTION = 1, 2, 3, 4
F_ORIENT_NORMALS, F_SIMPLE_SENSOR_NOISE, F_MAX_DIST, F_MIN_DIST, F_BOUNDING_BOX = 5, 6, 7, 8, 9
F_MAX_DENSITY, F_SAMPLING_SURFACE_NORMAL = 10, 11
F_REMOVE_NAN, F_FIX_STEP_SAMPLING, F_SHADOW, F_IDENTITY = 12, 13, 14, 15
O_TRIMMED_DIST, O_MAX_DIST, O_MIN_DIST, O_MEDIAN_DIST, O_SURFACE_NORMAL = 1, 2, 3, 4, 5
O_VAR_TRIMMED_DIST = 6
E_POINT_TO_PLANE, E_POINT_TO_PLANE_WITH_COV, E_POINT_TO_POINT = 1, 2, 3
MAX_MODS = 8

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)
_dp = C.POINTER(C.c_double)


class CCloud(C.Structure):
    _fields_ = [("n", C.c_int64), ("feat", _fp), ("normals", _fp), ("obsdir", _fp),
                ("noise", _fp), ("dens", _fp), ("eigval", _fp), ("eigvec", _fp),
                ("meandist", _fp), ("matched", _fp), ("matched_span", C.c_int)]


class CFilter(C.Structure):
    _fields_ = [("type", C.c_int), ("p0", C.c_double), ("p1", C.c_double), ("p2", C.c_double),
                ("i0", C.c_int64), ("i1", C.c_int64), ("box", C.c_double * 6)]


class COutlier(C.Structure):
    _fields_ = [("type", C.c_int), ("p0", C.c_double), ("p1", C.c_double), ("p2", C.c_double)]


class CMinOut(C.Structure):
    _fields_ = [("T", C.c_double * 16), ("cov", C.c_double * 36), ("A", C.c_double * 36),
                ("b", C.c_double * 6), ("point_used_ratio", C.c_double),
                ("weighted_point_used_ratio", C.c_double), ("residual", C.c_double),
                ("kept", C.c_int64)]


class CIcpConfig(C.Structure):
    _fields_ = [("reading_filters", CFilter * MAX_MODS), ("n_reading_filters", C.c_int),
                ("reading_step_filters", CFilter * MAX_MODS), ("n_reading_step_filters", C.c_int),
                ("reference_filters", CFilter * MAX_MODS), ("n_reference_filters", C.c_int),
                ("knn", C.c_int), ("epsilon", C.c_double), ("max_dist", C.c_double),
                ("outliers", COutlier * MAX_MODS), ("n_outliers", C.c_int),
                ("minimizer", C.c_int), ("sensor_std_dev", C.c_double),
                ("max_iterations", C.c_int),
                ("has_differential", C.c_int), ("min_diff_rot", C.c_double),
                ("min_diff_trans", C.c_double), ("smooth_length", C.c_int),
                ("has_bound", C.c_int), ("max_rot_norm", C.c_double), ("max_trans_norm", C.c_double),
                ("force_mode", C.c_int)]


class CIcpResult(C.Structure):
    _fields_ = [("T", C.c_double * 16), ("cov", C.c_double * 36), ("iterations", C.c_int),
                ("max_iter_reached", C.c_int), ("status", C.c_int),
                ("overlap", C.c_double), ("weighted_ratio", C.c_double),
                ("point_used_ratio", C.c_double), ("residual", C.c_double),
                ("last_T_iter", C.c_double * 16),
                ("time_filters_s", C.c_double), ("time_index_s", C.c_double),
                ("time_loop_s", C.c_double), ("time_knn_s", C.c_double),
                ("visits", C.c_uint64)]


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with the committed Makefile (gcc only)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle.h", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    cp = C.POINTER(CCloud)
    L.orc_cloud_new.restype = cp
    L.orc_cloud_new.argtypes = [C.c_int64]
    L.orc_cloud_copy.restype = cp
    L.orc_cloud_copy.argtypes = [cp]
    L.orc_cloud_free.argtypes = [cp]
    L.orc_cloud_concatenate.argtypes = [cp, cp]
    L.orc_kdtree_build.restype = C.c_void_p
    L.orc_kdtree_build.argtypes = [_fp, C.c_int64]
    L.orc_kdtree_free.argtypes = [C.c_void_p]
    L.orc_kdtree_knn.restype = C.c_uint64
    L.orc_kdtree_knn.argtypes = [C.c_void_p, _fp, C.c_int64, C.c_int, C.c_float, C.c_int, _ip, _fp]
    L.orc_kdtree_knn_ex.restype = C.c_uint64
    L.orc_kdtree_knn_ex.argtypes = [C.c_void_p, _fp, C.c_int64, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int, _ip, _fp]
    L.orc_set_search_mode.argtypes = [C.c_int]
    L.orc_knn_brute.argtypes = [_fp, C.c_int64, _fp, C.c_int64, C.c_int, C.c_float, _ip, _fp]
    L.orc_filter_apply.argtypes = [C.POINTER(CFilter), cp]
    L.orc_filters_apply.argtypes = [C.POINTER(CFilter), C.c_int, cp]
    L.orc_rigid_transform.argtypes = [cp, _dp]
    L.orc_outlier_weights.argtypes = [C.POINTER(COutlier), C.c_int, _fp, C.c_int64, _fp]
    L.orc_dists_quantile.argtypes = [_fp, C.c_int64, C.c_double, _fp]
    L.orc_var_trimmed_ratio.argtypes = [_fp, C.c_int64, C.c_double, C.c_double, C.c_double, _fp]
    L.orc_outlier_weights_full.argtypes = [C.POINTER(COutlier), C.c_int, cp, cp, _ip, _fp, C.c_int, _fp]
    L.orc_minimize.argtypes = [C.c_int, C.c_double, cp, cp, _ip, _fp, _fp, C.c_int, C.POINTER(CMinOut)]
    L.orc_minimize_ex.argtypes = [C.c_int, C.c_int, C.c_double, cp, cp, _ip, _fp, _fp, C.c_int, C.POINTER(CMinOut)]
    L.orc_overlap.restype = C.c_double
    L.orc_overlap.argtypes = [C.c_int, cp, cp, _ip, _fp, _fp, C.c_int]
    L.orc_eig3_sym.argtypes = [_dp, _dp, _dp]
    L.orc_solve6.argtypes = [_dp, _dp, _dp]
    L.orc_svd3.argtypes = [_dp, _dp, _dp, _dp]
    L.orc_icp_config_default.argtypes = [C.POINTER(CIcpConfig)]
    L.orc_icp_run.argtypes = [C.POINTER(CIcpConfig), cp, cp, _dp, C.POINTER(CIcpResult)]
    L.orc_icp_seq_new.restype = C.c_void_p
    L.orc_icp_seq_new.argtypes = [C.POINTER(CIcpConfig)]
    L.orc_icp_seq_free.argtypes = [C.c_void_p]
    L.orc_icp_seq_set_map.argtypes = [C.c_void_p, cp]
    L.orc_icp_seq_run.argtypes = [C.c_void_p, cp, _dp, C.POINTER(CIcpResult)]
    L.orc_icp_seq_map.restype = cp
    L.orc_icp_seq_map.argtypes = [C.c_void_p]
    L.orc_probe_overlap.argtypes = [C.POINTER(CIcpConfig), cp, cp, _dp, _dp]
    L.orc_probe_residual.argtypes = [C.POINTER(CIcpConfig), cp, cp, _dp, _dp]
    L.orc_num_threads.restype = C.c_int
    L.orc_set_num_threads.argtypes = [C.c_int]
    _LIB = L
    return L


def _f(a):
    return a.ctypes.data_as(_fp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _d(a):
    return a.ctypes.data_as(_dp)


# span None = variable (stored beside the pointer as <field>_span)
_DESC = (("normals", 3), ("obsdir", 3), ("noise", 1), ("dens", 1), ("eigval", 3), ("eigvec", 9),
         ("meandist", 1), ("matched", None))
# PM label <-> oracle field
LABEL_OF = {"normals": "normals", "obsdir": "observationDirections", "noise": "simpleSensorNoise",
            "dens": "densities", "eigval": "eigValues", "eigvec": "eigVectors",
            "meandist": "meanDists", "matched": "matchedIds"}
FIELD_OF = {v: k for k, v in LABEL_OF.items()}


class Cloud:
    """Owns an orc_cloud*.  features: 4xN float32 (column per point)."""

    def __init__(self, features=None, descriptors: dict | None = None, _ptr=None):
        L = lib()
        if _ptr is not None:
            self.ptr = _ptr
            return
        features = np.asarray(features, dtype=np.float32)
        assert features.shape[0] == 4
        n = features.shape[1]
        self.ptr = L.orc_cloud_new(n)
        flat = np.ascontiguousarray(features.T).ravel()
        C.memmove(self.ptr.contents.feat, flat.ctypes.data, flat.nbytes)
        for lab, arr in (descriptors or {}).items():
            self.set_desc(FIELD_OF.get(lab, lab), arr)

    def set_desc(self, field, arr):
        span = dict(_DESC)[field]
        if span is None:
            span = np.asarray(arr).shape[0]
            setattr(self.ptr.contents, field + "_span", span)
        arr = np.ascontiguousarray(np.asarray(arr, dtype=np.float32).reshape(span, -1).T).ravel()
        libc = C.CDLL(None)
        libc.malloc.restype = C.c_void_p
        libc.malloc.argtypes = [C.c_size_t]
        buf = libc.malloc(max(arr.nbytes, 4))
        C.memmove(buf, arr.ctypes.data, arr.nbytes)
        setattr(self.ptr.contents, field, C.cast(buf, _fp))

    def copy(self):
        return Cloud(_ptr=lib().orc_cloud_copy(self.ptr))

    @property
    def n(self):
        return int(self.ptr.contents.n)

    @property
    def features(self):
        n = self.n
        a = np.ctypeslib.as_array(self.ptr.contents.feat, shape=(n, 4)).copy()
        return np.asfortranarray(a.T)

    def desc(self, field):
        span = dict(_DESC)[field]
        if span is None:
            span = getattr(self.ptr.contents, field + "_span")
        p = getattr(self.ptr.contents, field)
        if not p:
            return None
        a = np.ctypeslib.as_array(p, shape=(self.n, span)).copy()
        return np.asfortranarray(a.T)

    def descriptors(self):
        return {LABEL_OF[f]: self.desc(f) for f, _ in _DESC if self.desc(f) is not None}

    def __del__(self):
        try:
            lib().orc_cloud_free(self.ptr)
        except Exception:
            pass


def _pts(a):
    """4xN (any order) float32 -> flat N*4 contiguous."""
    a = np.asarray(a, dtype=np.float32)
    assert a.shape[0] == 4
    return np.ascontiguousarray(a.T)


SEARCH_CONTRACT, SEARCH_NABO = 0, 1


def set_search_mode(mode: int):
    """matcher semantics inside icp_run / icp_seq / the probes: SEARCH_CONTRACT (default) or
    SEARCH_NABO (libnabo verbatim: strict <, first-visited ties, incremental rd, epsilon honoured)"""
    lib().orc_set_search_mode(int(mode))


def kdtree_knn(ref, query, k=1, max_dist=np.inf, allow_self=True, return_visits=False, mode=SEARCH_CONTRACT,
               epsilon=0.0):
    L = lib()
    r, q = _pts(ref), _pts(query)
    nq = q.shape[0]
    ids = np.empty((nq, k), np.int32)
    d2 = np.empty((nq, k), np.float32)
    t = L.orc_kdtree_build(_f(r), r.shape[0])
    v = L.orc_kdtree_knn_ex(t, _f(q), nq, k, float(max_dist), int(allow_self), float(epsilon), int(mode), _i(ids), _f(d2))
    L.orc_kdtree_free(t)
    return (ids.T, d2.T, v) if return_visits else (ids.T, d2.T)


def brute_knn(ref, query, k=1, max_dist=np.inf):
    L = lib()
    r, q = _pts(ref), _pts(query)
    nq = q.shape[0]
    ids = np.empty((nq, k), np.int32)
    d2 = np.empty((nq, k), np.float32)
    L.orc_knn_brute(_f(r), r.shape[0], _f(q), nq, k, float(max_dist), _i(ids), _f(d2))
    return ids.T, d2.T


def make_filter(name: str, **p) -> CFilter:
    f = CFilter()
    if name == "RandomSamplingDataPointsFilter":
        f.type, f.p0, f.i0 = F_RANDOM_SAMPLING, float(p.get("prob", 0.75)), int(p.get("seed", 0))
    elif name == "VoxelGridDataPointsFilter":
        f.type = F_VOXEL_GRID
        f.p0, f.p1, f.p2 = (float(p.get(k, 1.0)) for k in ("vSizeX", "vSizeY", "vSizeZ"))
        f.i0, f.i1 = int(p.get("useCentroid", 1)), int(p.get("averageExistingDescriptors", 1))
    elif name == "SurfaceNormalDataPointsFilter":
        f.type, f.i0 = F_SURFACE_NORMAL, int(p.get("knn", 5))
        f.p0 = float(p.get("maxDist", np.inf))
        f.i1 = (int(p.get("keepNormals", 1)) | int(p.get("keepDensities", 0)) << 1 |
                int(p.get("keepEigenValues", 0)) << 2 | int(p.get("keepEigenVectors", 0)) << 3 |
                int(p.get("keepMatchedIds", 0)) << 4 | int(p.get("keepMeanDist", 0)) << 5 |
                int(p.get("sortEigen", 0)) << 6)
    elif name == "ObservationDirectionDataPointsFilter":
        f.type = F_OBSERVATION_DIRECTION
        f.p0, f.p1, f.p2 = (float(p.get(k, 0.0)) for k in ("x", "y", "z"))
    elif name == "OrientNormalsDataPointsFilter":
        f.type, f.i0 = F_ORIENT_NORMALS, int(p.get("towardCenter", 1))
    elif name == "SimpleSensorNoiseDataPointsFilter":
        f.type, f.i0, f.p0 = F_SIMPLE_SENSOR_NOISE, int(p.get("sensorType", 0)), float(p.get("gain", 1.0))
    elif name == "IdentityDataPointsFilter":
        f.type = F_IDENTITY
    elif name == "RemoveNaNDataPointsFilter":
        f.type = F_REMOVE_NAN
    elif name == "FixStepSamplingDataPointsFilter":
        if float(p.get("stepMult", 1)) != 1 or int(p.get("endStep", p.get("startStep", 10))) != int(p.get("startStep", 10)):
            raise KeyError("FixStepSampling: step schedules are not supported")
        f.type, f.i0, f.i1 = F_FIX_STEP_SAMPLING, int(p.get("startStep", 10)), int(p.get("seed", 0))
    elif name == "ShadowDataPointsFilter":
        f.type, f.p0 = F_SHADOW, float(p.get("eps", 0.1))
    elif name == "MaxDistDataPointsFilter":
        f.type, f.i0, f.p0 = F_MAX_DIST, int(p.get("dim", -1)), float(p.get("maxDist", 1.0))
    elif name == "MinDistDataPointsFilter":
        f.type, f.i0, f.p0 = F_MIN_DIST, int(p.get("dim", -1)), float(p.get("minDist", 1.0))
    elif name == "SamplingSurfaceNormalDataPointsFilter":
        f.type = F_SAMPLING_SURFACE_NORMAL
        f.p0, f.p1, f.p2 = float(p.get("ratio", 0.5)), float(p.get("maxBoxDim", np.inf)), float(p.get("seed", 0))
        f.i0 = int(p.get("knn", 7))
        f.i1 = (int(p.get("keepNormals", 1)) | int(p.get("keepDensities", 0)) << 1 |
                int(p.get("keepEigenValues", 0)) << 2 | int(p.get("keepEigenVectors", 0)) << 3 |
                int(p.get("samplingMethod", 0)) << 4 | int(p.get("averageExistingDescriptors", 1)) << 5)
    elif name == "MaxDensityDataPointsFilter":
        f.type, f.p0, f.i0 = F_MAX_DENSITY, float(p.get("maxDensity", 10.0)), int(p.get("seed", 0))
    elif name == "BoundingBoxDataPointsFilter":
        f.type, f.i0 = F_BOUNDING_BOX, int(p.get("removeInside", 1))
        for j, (key, dflt) in enumerate((("xMin", -1), ("xMax", 1), ("yMin", -1), ("yMax", 1), ("zMin", -1), ("zMax", 1))):
            f.box[j] = float(p.get(key, dflt))
    else:
        raise KeyError(name)
    return f


def make_outlier(name: str, **p) -> COutlier:
    o = COutlier()
    if name == "TrimmedDistOutlierFilter":
        o.type, o.p0 = O_TRIMMED_DIST, float(p.get("ratio", 0.85))
    elif name == "MaxDistOutlierFilter":
        o.type, o.p0 = O_MAX_DIST, float(p.get("maxDist", 1.0))
    elif name == "MinDistOutlierFilter":
        o.type, o.p0 = O_MIN_DIST, float(p.get("minDist", 1.0))
    elif name == "MedianDistOutlierFilter":
        o.type, o.p0 = O_MEDIAN_DIST, float(p.get("factor", 3.0))
    elif name == "SurfaceNormalOutlierFilter":
        o.type, o.p0 = O_SURFACE_NORMAL, float(p.get("maxAngle", 1.57))
    elif name == "VarTrimmedDistOutlierFilter":
        o.type = O_VAR_TRIMMED_DIST
        o.p0, o.p1, o.p2 = float(p.get("minRatio", 0.05)), float(p.get("maxRatio", 0.99)), float(p.get("lambda", 0.95))
    else:
        raise KeyError(name)
    return o


def apply_filter(cloud: Cloud, name: str, **p) -> int:
    f = make_filter(name, **p)
    return lib().orc_filter_apply(C.byref(f), cloud.ptr)


def rigid_transform(cloud: Cloud, T) -> int:
    T = np.asfortranarray(np.asarray(T, dtype=np.float64))
    return lib().orc_rigid_transform(cloud.ptr, _d(T))


def _modlist(items):
    """[('Name', {params}) | 'Name' | {'Name': {params}}] -> [(name, params)]"""
    out = []
    for it in items or []:
        if isinstance(it, str):
            out.append((it, {}))
        elif isinstance(it, dict):
            (k, v), = it.items()
            out.append((k, v or {}))
        else:
            out.append((it[0], it[1] or {}))
    return out


def default_config() -> CIcpConfig:
    """ICPChainBase::setDefault."""
    c = CIcpConfig()
    lib().orc_icp_config_default(C.byref(c))
    return c


def config_from_dict(cfg: dict) -> CIcpConfig:
    """Same dict that yaml.safe_load gives for a libpointmatcher ICP YAML (A.10)."""
    c = CIcpConfig()
    lib().orc_icp_config_default(C.byref(c))
    for key, arr, cnt in (("readingDataPointsFilters", c.reading_filters, "n_reading_filters"),
                          ("readingStepDataPointsFilters", c.reading_step_filters, "n_reading_step_filters"),
                          ("referenceDataPointsFilters", c.reference_filters, "n_reference_filters")):
        mods = _modlist(cfg.get(key))
        for j, (name, p) in enumerate(mods):
            arr[j] = make_filter(name, **p)
        setattr(c, cnt, len(mods))
    if "matcher" in cfg:
        (name, p), = _modlist([cfg["matcher"]])
        assert name == "KDTreeMatcher"
        c.knn = int(p.get("knn", 1))
        c.epsilon = float(p.get("epsilon", 0))
        c.max_dist = float(p.get("maxDist", np.inf))
    if "outlierFilters" in cfg:
        mods = _modlist(cfg["outlierFilters"])
        for j, (name, p) in enumerate(mods):
            c.outliers[j] = make_outlier(name, **p)
        c.n_outliers = len(mods)
    if "errorMinimizer" in cfg:
        (name, p), = _modlist([cfg["errorMinimizer"]])
        c.minimizer = {"PointToPlaneErrorMinimizer": E_POINT_TO_PLANE,
                       "PointToPlaneWithCovErrorMinimizer": E_POINT_TO_PLANE_WITH_COV,
                       "PointToPointErrorMinimizer": E_POINT_TO_POINT}[name]
        c.sensor_std_dev = float(p.get("sensorStdDev", 0.01))
        c.force_mode = 1 if int(p.get("force2D", 0)) else (2 if int(p.get("force4DOF", 0)) else 0)
    if "transformationCheckers" in cfg:
        c.max_iterations, c.has_differential, c.has_bound = 0, 0, 0
        for name, p in _modlist(cfg["transformationCheckers"]):
            if name == "CounterTransformationChecker":
                c.max_iterations = int(p.get("maxIterationCount", 40))
            elif name == "DifferentialTransformationChecker":
                c.has_differential = 1
                c.min_diff_rot = float(p.get("minDiffRotErr", 0.001))
                c.min_diff_trans = float(p.get("minDiffTransErr", 0.001))
                c.smooth_length = int(p.get("smoothLength", 3))
            elif name == "BoundTransformationChecker":
                c.has_bound = 1
                c.max_rot_norm = float(p.get("maxRotationNorm", 1))
                c.max_trans_norm = float(p.get("maxTranslationNorm", 1))
            else:
                raise KeyError(name)
    return c


def _result(r: CIcpResult) -> dict:
    return dict(T=np.array(r.T).reshape(4, 4).T.copy(), cov=np.array(r.cov).reshape(6, 6).T.copy(),
                iterations=r.iterations, max_iter_reached=bool(r.max_iter_reached), status=r.status,
                overlap=r.overlap, weighted_ratio=r.weighted_ratio, point_used_ratio=r.point_used_ratio,
                residual=r.residual, last_T_iter=np.array(r.last_T_iter).reshape(4, 4).T.copy(),
                time_filters_s=r.time_filters_s, time_index_s=r.time_index_s,
                time_loop_s=r.time_loop_s, time_knn_s=r.time_knn_s, visits=int(r.visits))


def icp_run(cfg, reading: Cloud, reference: Cloud, T_init=None) -> dict:
    c = cfg if isinstance(cfg, CIcpConfig) else config_from_dict(cfg)
    T = np.asfortranarray(np.eye(4) if T_init is None else np.asarray(T_init, dtype=np.float64))
    r = CIcpResult()
    lib().orc_icp_run(C.byref(c), reading.ptr, reference.ptr, _d(T), C.byref(r))
    return _result(r)


class IcpSequence:
    def __init__(self, cfg):
        self.cfg = cfg if isinstance(cfg, CIcpConfig) else config_from_dict(cfg)
        self.h = lib().orc_icp_seq_new(C.byref(self.cfg))

    def set_map(self, cloud: Cloud) -> int:
        return lib().orc_icp_seq_set_map(self.h, cloud.ptr)

    def run(self, reading: Cloud, T_init=None) -> dict:
        T = np.asfortranarray(np.eye(4) if T_init is None else np.asarray(T_init, dtype=np.float64))
        r = CIcpResult()
        lib().orc_icp_seq_run(self.h, reading.ptr, _d(T), C.byref(r))
        return _result(r)

    def map(self) -> Cloud:
        return Cloud(_ptr=lib().orc_cloud_copy(lib().orc_icp_seq_map(self.h)))

    def __del__(self):
        try:
            lib().orc_icp_seq_free(self.h)
        except Exception:
            pass


def outlier_weights(outliers, d2):
    mods = _modlist(outliers)
    arr = (COutlier * max(1, len(mods)))()
    for j, (name, p) in enumerate(mods):
        arr[j] = make_outlier(name, **p)
    d = np.ascontiguousarray(np.asarray(d2, np.float32).T).ravel()
    w = np.empty_like(d)
    st = lib().orc_outlier_weights(arr, len(mods), _f(d), d.size, _f(w))
    return st, w.reshape(np.asarray(d2).T.shape).T


def outlier_weights_full(outliers, reading: Cloud, reference: Cloud, ids, d2):
    mods = _modlist(outliers)
    arr = (COutlier * max(1, len(mods)))()
    for j, (name, p) in enumerate(mods):
        arr[j] = make_outlier(name, **p)
    k = np.asarray(d2).shape[0]
    I = np.ascontiguousarray(np.asarray(ids, np.int32).T).ravel()
    d = np.ascontiguousarray(np.asarray(d2, np.float32).T).ravel()
    w = np.empty_like(d)
    st = lib().orc_outlier_weights_full(arr, len(mods), reading.ptr, reference.ptr, _i(I), _f(d), k, _f(w))
    return st, w.reshape(np.asarray(d2).T.shape).T


def minimize(kind, reading: Cloud, reference: Cloud, ids, d2, w, sensor_std_dev=0.01, force_mode=0):
    k = ids.shape[0]
    I = np.ascontiguousarray(np.asarray(ids, np.int32).T).ravel()
    D = np.ascontiguousarray(np.asarray(d2, np.float32).T).ravel()
    W = np.ascontiguousarray(np.asarray(w, np.float32).T).ravel()
    o = CMinOut()
    st = lib().orc_minimize_ex(kind, force_mode, sensor_std_dev, reading.ptr, reference.ptr, _i(I), _f(D), _f(W), k, C.byref(o))
    return st, dict(T=np.array(o.T).reshape(4, 4).T.copy(), cov=np.array(o.cov).reshape(6, 6).T.copy(),
                    A=np.array(o.A).reshape(6, 6).T.copy(), b=np.array(o.b), kept=o.kept,
                    point_used_ratio=o.point_used_ratio,
                    weighted_point_used_ratio=o.weighted_point_used_ratio, residual=o.residual)


def probe_overlap(cfg, reading: Cloud, reference: Cloud, T):
    c = cfg if isinstance(cfg, CIcpConfig) else config_from_dict(cfg)
    T = np.asfortranarray(np.asarray(T, dtype=np.float64))
    out = C.c_double(0)
    st = lib().orc_probe_overlap(C.byref(c), reading.ptr, reference.ptr, _d(T), C.byref(out))
    return st, out.value


def probe_residual(cfg, reading: Cloud, reference: Cloud, T):
    c = cfg if isinstance(cfg, CIcpConfig) else config_from_dict(cfg)
    T = np.asfortranarray(np.asarray(T, dtype=np.float64))
    out = C.c_double(0)
    st = lib().orc_probe_residual(C.byref(c), reading.ptr, reference.ptr, _d(T), C.byref(out))
    return st, out.value
