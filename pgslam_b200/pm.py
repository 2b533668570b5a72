"""ctypes host layer over libpgslam_b200.so, shaped like the libpointmatcher
objects pgslam uses (types.h:19-27): DataPoints, DataPointsFilters, ICP,
ICPSequence, and the ICPChainBase members pgslam touches module by module
(Localizer.hpp:309-347, LoopCloser.hpp:346-362).

This is the Python counterpart of include/pgslam_b200/pm_adapter.hpp; both sit
on the same C ABI (include/pgslam_b200.h).  There is NO CPU fallback: if the
CUDA library is missing or no GPU is present, construction fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpgslam_b200.so")
_LIB = None

OK, CONVERGENCE_ERROR, TRANSFORMATION_ERROR, INVALID_PARAMETER, INVALID_FIELD, \
    INVALID_MODULE_TYPE, INVALID_ELEMENT, CUDA_ERROR, INVALID_ARGUMENT = range(9)


class PointMatcherError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"[{self.__class__.__name__}] {msg}")
        self.status = status


class ConvergenceError(PointMatcherError): pass
class TransformationError(PointMatcherError): pass
class InvalidParameter(PointMatcherError): pass
class InvalidField(PointMatcherError): pass
class InvalidModuleType(PointMatcherError): pass
class InvalidElement(PointMatcherError): pass
class CudaError(PointMatcherError): pass


_EXC = {CONVERGENCE_ERROR: ConvergenceError, TRANSFORMATION_ERROR: TransformationError,
        INVALID_PARAMETER: InvalidParameter, INVALID_FIELD: InvalidField,
        INVALID_MODULE_TYPE: InvalidModuleType, INVALID_ELEMENT: InvalidElement,
        CUDA_ERROR: CudaError, INVALID_ARGUMENT: InvalidParameter}

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)
_dp = C.POINTER(C.c_double)
_vp = C.c_void_p


class MinResult(C.Structure):
    _fields_ = [("T", C.c_double * 16), ("covariance", C.c_double * 36),
                ("point_used_ratio", C.c_double), ("weighted_point_used_ratio", C.c_double),
                ("residual", C.c_double), ("overlap", C.c_double), ("kept", C.c_int64)]


class IcpResult(C.Structure):
    _fields_ = [("T", C.c_double * 16), ("covariance", C.c_double * 36),
                ("iterations", C.c_int32), ("max_iterations_reached", C.c_int32),
                ("status", C.c_int32), ("reserved", C.c_int32),
                ("overlap", C.c_double), ("weighted_point_used_ratio", C.c_double),
                ("point_used_ratio", C.c_double), ("residual", C.c_double),
                ("n_reading", C.c_int64), ("n_reference", C.c_int64)]


class HostCloud(C.Structure):
    """pgs_host_cloud: a host-resident DataPoints (features 4 x N column-major + descriptor blocks)."""
    _fields_ = [("features4xN", C.c_void_p), ("n", C.c_int64), ("n_descriptors", C.c_int),
                ("labels", C.POINTER(C.c_char_p)), ("spans", C.POINTER(C.c_int)), ("data", C.POINTER(C.c_void_p))]


class StageTimes(C.Structure):
    _fields_ = [("filters_ms", C.c_float), ("index_ms", C.c_float), ("loop_ms", C.c_float),
                ("total_ms", C.c_float), ("match_ms", C.c_float), ("select_ms", C.c_float),
                ("accumulate_ms", C.c_float), ("iterations_launched", C.c_int32)]


# name -> (restype, argtypes); every symbol include/pgslam_b200.h declares
SIGNATURES = {
    "pgs_ctx_create": (C.c_int, [C.c_int, _vp, C.POINTER(_vp)]),
    "pgs_ctx_destroy": (None, [_vp]),
    "pgs_last_error": (C.c_char_p, [_vp]),
    "pgs_ctx_synchronize": (C.c_int, [_vp]),
    "pgs_version": (C.c_char_p, []),
    "pgs_cloud_create": (C.c_int, [_vp, _vp, C.c_int64, C.c_int, C.POINTER(_vp)]),
    "pgs_cloud_set_descriptor": (C.c_int, [_vp, C.c_char_p, C.c_int, _vp, C.c_int]),
    "pgs_cloud_remove_descriptor": (C.c_int, [_vp, C.c_char_p]),
    "pgs_cloud_num_points": (C.c_int64, [_vp]),
    "pgs_cloud_num_descriptors": (C.c_int, [_vp]),
    "pgs_cloud_descriptor_info": (C.c_int, [_vp, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int)]),
    "pgs_cloud_get_features": (C.c_int, [_vp, _vp, C.c_int]),
    "pgs_cloud_get_descriptor": (C.c_int, [_vp, C.c_char_p, _vp, C.c_int]),
    "pgs_cloud_copy": (C.c_int, [_vp, C.POINTER(_vp)]),
    "pgs_cloud_concatenate": (C.c_int, [_vp, _vp]),
    "pgs_cloud_destroy": (None, [_vp]),
    "pgs_cloud_load": (C.c_int, [_vp, C.c_char_p, C.POINTER(_vp)]),
    "pgs_cloud_save": (C.c_int, [_vp, C.c_char_p]),
    "pgs_cloud_file_info": (C.c_int, [C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int), C.c_char_p, C.c_int]),
    "pgs_rigid_transform": (C.c_int, [_vp, _dp]),
    "pgs_cloud_assemble": (C.c_int, [_vp, C.c_int, C.POINTER(_vp), _dp, C.POINTER(_vp)]),
    "pgs_filters_create_from_yaml": (C.c_int, [_vp, C.c_char_p, C.c_size_t, C.POINTER(_vp)]),
    "pgs_filters_create": (C.c_int, [_vp, C.POINTER(_vp)]),
    "pgs_filters_append": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_char_p), C.c_int]),
    "pgs_filters_count": (C.c_int, [_vp]),
    "pgs_filters_apply": (C.c_int, [_vp, _vp]),
    "pgs_filters_destroy": (None, [_vp]),
    "pgs_matcher_create": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.POINTER(_vp)]),
    "pgs_matcher_init": (C.c_int, [_vp, _vp]),
    "pgs_matcher_knn": (C.c_int, [_vp]),
    "pgs_matcher_find": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int]),
    "pgs_matcher_dense_fallbacks": (C.c_uint, [_vp]),
    "pgs_matcher_destroy": (None, [_vp]),
    "pgs_outliers_create": (C.c_int, [_vp, C.POINTER(_vp)]),
    "pgs_outliers_append": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_char_p), C.c_int]),
    "pgs_outliers_compute": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int]),
    "pgs_outliers_destroy": (None, [_vp]),
    "pgs_minimizer_create": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.POINTER(_vp)]),
    "pgs_minimizer_compute": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.POINTER(MinResult)]),
    "pgs_minimizer_destroy": (None, [_vp]),
    "pgs_icp_create_from_yaml": (C.c_int, [_vp, C.c_char_p, C.c_size_t, C.POINTER(_vp)]),
    "pgs_icp_create_default": (C.c_int, [_vp, C.POINTER(_vp)]),
    "pgs_icp_destroy": (None, [_vp]),
    "pgs_icp_reading_filters": (_vp, [_vp]),
    "pgs_icp_reading_step_filters": (_vp, [_vp]),
    "pgs_icp_reference_filters": (_vp, [_vp]),
    "pgs_icp_matcher": (_vp, [_vp]),
    "pgs_icp_outliers": (_vp, [_vp]),
    "pgs_icp_minimizer": (_vp, [_vp]),
    "pgs_icp_run": (C.c_int, [_vp, _vp, _vp, _dp, C.POINTER(IcpResult)]),
    "pgs_icp_set_map": (C.c_int, [_vp, _vp]),
    "pgs_icp_has_map": (C.c_int, [_vp]),
    "pgs_icp_run_sequence": (C.c_int, [_vp, _vp, _dp, C.POINTER(IcpResult)]),
    "pgs_icp_run_batch": (C.c_int, [_vp, C.c_int, C.POINTER(_vp), C.POINTER(_vp), _dp, C.POINTER(IcpResult)]),
    "pgs_icp_run_batch_multi": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.POINTER(HostCloud), C.POINTER(HostCloud),
                                          _dp, C.c_int, C.POINTER(IcpResult)]),
    "pgs_icp_probe_overlap": (C.c_int, [_vp, _vp, _vp, _dp, _dp]),
    "pgs_icp_probe_residual": (C.c_int, [_vp, _vp, _vp, _dp, _dp]),
    "pgs_config_check": (C.c_int, [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.c_int), C.c_char_p, C.c_int]),
    "pgs_config_warnings": (C.c_int, [C.c_char_p, C.c_size_t, C.c_int, C.c_char_p, C.c_int]),
    "pgs_registrar_count": (C.c_int, [C.c_int]),
    "pgs_registrar_name": (C.c_char_p, [C.c_int, C.c_int]),
    "pgs_registrar_param_count": (C.c_int, [C.c_int, C.c_char_p]),
    "pgs_registrar_param": (C.c_int, [C.c_int, C.c_char_p, C.c_int] + [C.POINTER(C.c_char_p)] * 5 + [C.c_char_p]),
    "pgs_module_validate": (C.c_int, [C.c_int, C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_int]),
    "pgs_ctx_launch_count": (C.c_uint64, [_vp]),
    "pgs_ctx_set_profiling": (C.c_int, [_vp, C.c_int]),
    "pgs_ctx_set_batch_streams": (C.c_int, [_vp, C.c_int]),
    "pgs_ctx_last_stage_times": (C.c_int, [_vp, C.POINTER(StageTimes)]),
    "pgs_ctx_set_option": (C.c_int, [_vp, C.c_char_p, C.c_double]),
}


def load_library():
    """dlopen libpgslam_b200.so and type every entry point.  No fallback."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m pgslam_b200.build` "
                "(__graft_entry__.build()).  pgslam_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


KINDS = ("DataPointsFilter", "Matcher", "OutlierFilter", "ErrorMinimizer", "TransformationChecker", "Inspector",
         "Logger", "Transformation")


def check_config(yaml_text: str, chain: bool = True) -> int:
    """Parse + validate a libpointmatcher YAML on the host (no GPU needed).
    Returns the number of modules; raises the PointMatcher-style exception otherwise."""
    L = load_library()
    b = yaml_text.encode()
    n = C.c_int(0)
    err = C.create_string_buffer(512)
    st = L.pgs_config_check(b, len(b), int(chain), C.byref(n), err, 512)
    if st != OK:
        raise _EXC.get(st, PointMatcherError)(st, err.value.decode())
    return n.value


def config_warnings(yaml_text: str, chain: bool = True) -> list[str]:
    """Where a valid configuration still departs from libpointmatcher's behaviour (pgs_config_warnings)."""
    L = load_library()
    b = yaml_text.encode()
    buf = C.create_string_buffer(4096)
    L.pgs_config_warnings(b, len(b), int(chain), buf, 4096)
    return [ln for ln in buf.value.decode().splitlines() if ln]


def registered(kind: str) -> list[str]:
    """PM::get().REG(kind): the names of the registered modules."""
    L = load_library()
    k = KINDS.index(kind)
    return [L.pgs_registrar_name(k, i).decode() for i in range(L.pgs_registrar_count(k))]


def available_parameters(kind: str, name: str) -> list[dict]:
    """Parametrizable::availableParameters() of a registered module."""
    L = load_library()
    k = KINDS.index(kind)
    n = L.pgs_registrar_param_count(k, name.encode())
    if n < 0:
        raise InvalidElement(INVALID_ELEMENT, f"Trying to instanciate unknown element {name}")
    out = []
    for i in range(n):
        f = [C.c_char_p() for _ in range(5)]
        t = C.create_string_buffer(2)
        L.pgs_registrar_param(k, name.encode(), i, *[C.byref(x) for x in f], t)
        out.append(dict(name=f[0].value.decode(), doc=f[1].value.decode(), default=f[2].value.decode(),
                        min=f[3].value.decode(), max=f[4].value.decode(), type=t.value.decode()))
    return out


def cloud_file_info(path: str) -> tuple[int, int]:
    """(points, descriptors) of a csv / vtk / ply file, parsed on the host (no device needed)."""
    L = load_library()
    n, nd = C.c_int64(0), C.c_int(0)
    err = C.create_string_buffer(512)
    st = L.pgs_cloud_file_info(os.fsencode(path), C.byref(n), C.byref(nd), err, 512)
    if st != OK:
        raise _EXC.get(st, PointMatcherError)(st, err.value.decode())
    return int(n.value), int(nd.value)


def _mat(T):
    T = np.eye(4) if T is None else np.asarray(T, dtype=np.float64)
    assert T.shape == (4, 4)
    return np.asfortranarray(T)


def _kv(params: dict | None):
    items = []
    for k, v in (params or {}).items():
        items += [str(k).encode(), (repr(float(v)) if isinstance(v, float) else str(v)).encode()]
    arr = (C.c_char_p * max(1, len(items)))(*items)
    return arr, len(items) // 2


class Context:
    """One device + one stream.  `stream` may be a raw cudaStream_t (int), e.g.
    torch.cuda.current_stream().cuda_stream, so torch events time our kernels."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = load_library()
        h = _vp()
        st = self.lib.pgs_ctx_create(device, _vp(stream) if stream else None, C.byref(h))
        if st != OK:
            raise _EXC.get(st, PointMatcherError)(st, self.lib.pgs_last_error(None).decode())
        self.h = h
        self.device = device

    def check(self, st):
        if st != OK:
            raise _EXC.get(st, PointMatcherError)(st, self.lib.pgs_last_error(self.h).decode())

    def synchronize(self):
        self.check(self.lib.pgs_ctx_synchronize(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.pgs_ctx_launch_count(self.h))

    def set_profiling(self, on: bool):
        self.lib.pgs_ctx_set_profiling(self.h, int(on))

    def set_batch_streams(self, n: int):
        """Worker streams a batch of independent pairs is split over (default 4)."""
        self.check(self.lib.pgs_ctx_set_batch_streams(self.h, int(n)))

    def set_option(self, key: str, value):
        """Scheduling knobs of the hot kernels (pgs_ctx_set_option); results never depend on them."""
        self.check(self.lib.pgs_ctx_set_option(self.h, key.encode(), float(value)))

    def stage_times(self) -> dict:
        t = StageTimes()
        self.lib.pgs_ctx_last_stage_times(self.h, C.byref(t))
        return {f: getattr(t, f) for f, _ in StageTimes._fields_}

    def close(self):
        if getattr(self, "h", None):
            self.lib.pgs_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_DEFAULT_CTX = None


def default_context() -> Context:
    global _DEFAULT_CTX
    if _DEFAULT_CTX is None:
        _DEFAULT_CTX = Context(0)
    return _DEFAULT_CTX


class DataPoints:
    """PM::DataPoints (types.h:20), resident on the device.

    features: 4 x N float32 (column per point, last row 1);
    descriptors: {label: span x N float32}."""

    def __init__(self, features=None, descriptors: dict | None = None, ctx: Context | None = None,
                 _handle=None, device_ptr: int | None = None, n: int | None = None,
                 pinned_host_ptr: int | None = None):
        self.ctx = ctx or default_context()
        L = self.ctx.lib
        if _handle is not None:
            self.h = _handle
            return
        h = _vp()
        if pinned_host_ptr is not None:
            # N x 4 float32 in PINNED host memory, uploaded asynchronously on the side
            # stream (mode 2 of the ABI): the caller keeps the buffer alive until the
            # next call that returns results
            self.ctx.check(L.pgs_cloud_create(self.ctx.h, _vp(pinned_host_ptr), n, 2, C.byref(h)))
        elif device_ptr is not None:
            self.ctx.check(L.pgs_cloud_create(self.ctx.h, _vp(device_ptr), n, 1, C.byref(h)))
        else:
            f = np.asarray(features, dtype=np.float32)
            if f.ndim != 2 or f.shape[0] != 4:
                raise InvalidField(INVALID_FIELD, "features must be 4 x N (x, y, z, pad)")
            flat = np.ascontiguousarray(f.T)
            self.ctx.check(L.pgs_cloud_create(self.ctx.h, flat.ctypes.data_as(_vp), flat.shape[0], 0, C.byref(h)))
        self.h = h
        for label, d in (descriptors or {}).items():
            self.addDescriptor(label, d)

    # --- PM-style accessors -------------------------------------------------
    def getNbPoints(self) -> int:
        return int(self.ctx.lib.pgs_cloud_num_points(self.h))

    @property
    def features(self) -> np.ndarray:
        n = self.getNbPoints()
        out = np.empty((n, 4), np.float32)
        self.ctx.check(self.ctx.lib.pgs_cloud_get_features(self.h, out.ctypes.data_as(_vp), 0))
        return np.asfortranarray(out.T)

    def descriptorLabels(self) -> list[tuple[str, int]]:
        L = self.ctx.lib
        out = []
        for i in range(L.pgs_cloud_num_descriptors(self.h)):
            buf = C.create_string_buffer(128)
            span = C.c_int(0)
            self.ctx.check(L.pgs_cloud_descriptor_info(self.h, i, buf, 128, C.byref(span)))
            out.append((buf.value.decode(), span.value))
        return out

    def descriptorExists(self, label: str) -> bool:
        return any(l == label for l, _ in self.descriptorLabels())

    def getDescriptorByName(self, label: str) -> np.ndarray:
        span = dict(self.descriptorLabels()).get(label)
        if span is None:
            raise InvalidField(INVALID_FIELD, f"Cannot find descriptor {label}")
        out = np.empty((self.getNbPoints(), span), np.float32)
        self.ctx.check(self.ctx.lib.pgs_cloud_get_descriptor(self.h, label.encode(), out.ctypes.data_as(_vp), 0))
        return np.asfortranarray(out.T)

    def addDescriptor(self, label: str, data):
        d = np.asarray(data, dtype=np.float32)
        if d.ndim == 1:
            d = d[None, :]
        flat = np.ascontiguousarray(d.T)
        self.ctx.check(self.ctx.lib.pgs_cloud_set_descriptor(self.h, label.encode(), d.shape[0],
                                                             flat.ctypes.data_as(_vp), 0))

    def removeDescriptor(self, label: str):
        self.ctx.check(self.ctx.lib.pgs_cloud_remove_descriptor(self.h, label.encode()))

    @property
    def descriptors(self) -> dict:
        return {l: self.getDescriptorByName(l) for l, _ in self.descriptorLabels()}

    @classmethod
    def load(cls, path: str, ctx: "Context | None" = None) -> "DataPoints":
        """DataPoints::load (csv / vtk / ply, by extension) through the C ABI (pgs_cloud_load)."""
        ctx = ctx or default_context()
        h = _vp()
        ctx.check(ctx.lib.pgs_cloud_load(ctx.h, os.fsencode(path), C.byref(h)))
        return cls(ctx=ctx, _handle=h)

    def save(self, path: str):
        """DataPoints::save (csv / vtk / ply, by extension) through the C ABI (pgs_cloud_save)."""
        self.ctx.check(self.ctx.lib.pgs_cloud_save(self.h, os.fsencode(path)))

    def copy(self) -> "DataPoints":
        h = _vp()
        self.ctx.check(self.ctx.lib.pgs_cloud_copy(self.h, C.byref(h)))
        return DataPoints(ctx=self.ctx, _handle=h)

    def concatenate(self, other: "DataPoints"):
        self.ctx.check(self.ctx.lib.pgs_cloud_concatenate(self.h, other.h))

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.pgs_cloud_destroy(self.h)
        except Exception:
            pass


class RigidTransformation:
    """PM::get().REG(Transformation).create("RigidTransformation") (Localizer.hpp:20)."""

    def __init__(self, ctx: Context | None = None):
        self.ctx = ctx or default_context()

    def compute(self, cloud: DataPoints, T) -> DataPoints:
        out = cloud.copy()
        T = _mat(T)
        self.ctx.check(self.ctx.lib.pgs_rigid_transform(out.h, T.ctypes.data_as(_dp)))
        return out


class Transformations:
    """ICPChainBase::transformations: apply() is in place (LoopCloser.hpp:352)."""

    def __init__(self, ctx):
        self.ctx = ctx

    def apply(self, cloud: DataPoints, T):
        T = _mat(T)
        self.ctx.check(self.ctx.lib.pgs_rigid_transform(cloud.h, T.ctypes.data_as(_dp)))


def assemble_local_map(clouds: list[DataPoints], transforms: list) -> DataPoints:
    """LocalMap::BuildCloudFromData (LocalMap.hpp:209-224) on the device."""
    ctx = clouds[0].ctx
    arr = (_vp * len(clouds))(*[c.h for c in clouds])
    T = np.concatenate([_mat(t).ravel(order="F") for t in transforms])
    h = _vp()
    ctx.check(ctx.lib.pgs_cloud_assemble(ctx.h, len(clouds), arr, T.ctypes.data_as(_dp), C.byref(h)))
    return DataPoints(ctx=ctx, _handle=h)


class DataPointsFilters:
    """PM::DataPointsFilters (types.h:27): DataPointsFilters(yaml), init(), apply()."""

    def __init__(self, yaml_text: str | None = None, ctx: Context | None = None, _borrowed=None):
        self.ctx = ctx or default_context()
        self._own = _borrowed is None
        if _borrowed is not None:
            self.h = _borrowed
            return
        h = _vp()
        if yaml_text is None:
            self.ctx.check(self.ctx.lib.pgs_filters_create(self.ctx.h, C.byref(h)))
        else:
            b = yaml_text.encode()
            self.ctx.check(self.ctx.lib.pgs_filters_create_from_yaml(self.ctx.h, b, len(b), C.byref(h)))
        self.h = h

    def append(self, name: str, params: dict | None = None):
        kv, n = _kv(params)
        self.ctx.check(self.ctx.lib.pgs_filters_append(self.h, name.encode(), kv, n))

    def __len__(self):
        return self.ctx.lib.pgs_filters_count(self.h)

    def init(self):
        pass

    def apply(self, cloud: DataPoints):
        self.ctx.check(self.ctx.lib.pgs_filters_apply(self.h, cloud.h))

    def __del__(self):
        try:
            if self._own and self.h and self.ctx.h:
                self.ctx.lib.pgs_filters_destroy(self.h)
        except Exception:
            pass


class Matches:
    """PM::Matches: ids (k x N int32), dists (k x N float32, SQUARED)."""

    def __init__(self, ids, dists):
        self.ids = ids
        self.dists = dists


class Matcher:
    """PM::Matcher / KDTreeMatcher: init(reference), findClosests(reading)."""

    def __init__(self, name="KDTreeMatcher", params: dict | None = None, ctx: Context | None = None, _borrowed=None):
        self.ctx = ctx or default_context()
        self._own = _borrowed is None
        if _borrowed is not None:
            self.h = _borrowed
            return
        kv, n = _kv(params)
        h = _vp()
        self.ctx.check(self.ctx.lib.pgs_matcher_create(self.ctx.h, name.encode(), kv, n, C.byref(h)))
        self.h = h

    @property
    def knn(self):
        return self.ctx.lib.pgs_matcher_knn(self.h)

    def init(self, reference: DataPoints):
        self.ctx.check(self.ctx.lib.pgs_matcher_init(self.h, reference.h))

    @property
    def dense_fallbacks(self) -> int:
        return int(self.ctx.lib.pgs_matcher_dense_fallbacks(self.h))

    def findClosests(self, reading: DataPoints) -> Matches:
        n, k = reading.getNbPoints(), self.knn
        ids = np.empty((n, k), np.int32)
        d2 = np.empty((n, k), np.float32)
        self.ctx.check(self.ctx.lib.pgs_matcher_find(self.h, reading.h, ids.ctypes.data_as(_vp), d2.ctypes.data_as(_vp), 0))
        return Matches(np.asfortranarray(ids.T), np.asfortranarray(d2.T))

    def __del__(self):
        try:
            if self._own and self.h and self.ctx.h:
                self.ctx.lib.pgs_matcher_destroy(self.h)
        except Exception:
            pass


class OutlierFilters:
    """PM::OutlierFilters: compute(reading, reference, matches) -> OutlierWeights."""

    def __init__(self, ctx: Context | None = None, _borrowed=None):
        self.ctx = ctx or default_context()
        self._own = _borrowed is None
        if _borrowed is not None:
            self.h = _borrowed
            return
        h = _vp()
        self.ctx.check(self.ctx.lib.pgs_outliers_create(self.ctx.h, C.byref(h)))
        self.h = h

    def append(self, name: str, params: dict | None = None):
        kv, n = _kv(params)
        self.ctx.check(self.ctx.lib.pgs_outliers_append(self.h, name.encode(), kv, n))

    def compute(self, reading: DataPoints, reference: DataPoints, matches: Matches) -> np.ndarray:
        k, n = matches.dists.shape
        ids = np.ascontiguousarray(matches.ids.T, dtype=np.int32)
        d2 = np.ascontiguousarray(matches.dists.T, dtype=np.float32)
        w = np.empty((n, k), np.float32)
        self.ctx.check(self.ctx.lib.pgs_outliers_compute(self.h, reading.h, reference.h, ids.ctypes.data_as(_vp),
                                                         d2.ctypes.data_as(_vp), k, w.ctypes.data_as(_vp), 0))
        return np.asfortranarray(w.T)

    def __del__(self):
        try:
            if self._own and self.h and self.ctx.h:
                self.ctx.lib.pgs_outliers_destroy(self.h)
        except Exception:
            pass


class ErrorElements:
    """PM::ErrorMinimizer::ErrorElements(reading, reference, weights, matches)
    (Localizer.hpp:332): only the ratios pgslam reads are exposed."""

    def __init__(self, res: MinResult):
        self.pointUsedRatio = res.point_used_ratio
        self.weightedPointUsedRatio = res.weighted_point_used_ratio
        self.nbRejectedPoints = None
        self.kept = res.kept


class ErrorMinimizer:
    """PM::ErrorMinimizer: compute(), getOverlap(), getCovariance(), getResidualError()."""

    def __init__(self, name="PointToPlaneErrorMinimizer", params: dict | None = None,
                 ctx: Context | None = None, _borrowed=None):
        self.ctx = ctx or default_context()
        self._own = _borrowed is None
        self._last = None
        if _borrowed is not None:
            self.h = _borrowed
            return
        kv, n = _kv(params)
        h = _vp()
        self.ctx.check(self.ctx.lib.pgs_minimizer_create(self.ctx.h, name.encode(), kv, n, C.byref(h)))
        self.h = h

    def _run(self, reading, reference, weights, matches) -> MinResult:
        k, n = matches.dists.shape
        ids = np.ascontiguousarray(matches.ids.T, dtype=np.int32)
        d2 = np.ascontiguousarray(matches.dists.T, dtype=np.float32)
        w = np.ascontiguousarray(np.asarray(weights).T, dtype=np.float32)
        res = MinResult()
        self.ctx.check(self.ctx.lib.pgs_minimizer_compute(self.h, reading.h, reference.h, ids.ctypes.data_as(_vp),
                                                          d2.ctypes.data_as(_vp), w.ctypes.data_as(_vp), k, 0,
                                                          C.byref(res)))
        return res

    def compute(self, reading, reference, weights, matches) -> np.ndarray:
        self._last = self._run(reading, reference, weights, matches)
        return np.array(self._last.T).reshape(4, 4).T.copy()

    def errorElements(self, reading, reference, weights, matches) -> ErrorElements:
        return ErrorElements(self._run(reading, reference, weights, matches))

    def getResidualError(self, reading, reference, weights, matches) -> float:
        return self._run(reading, reference, weights, matches).residual

    def getOverlap(self) -> float:
        if self._last is None:
            raise ConvergenceError(CONVERGENCE_ERROR, "getOverlap() before any compute()")
        return self._last.overlap

    def getCovariance(self) -> np.ndarray:
        if self._last is None:
            return np.zeros((6, 6))
        return np.array(self._last.covariance).reshape(6, 6).T.copy()

    def _set_from_icp(self, r: IcpResult):
        m = MinResult()
        m.overlap = r.overlap
        m.weighted_point_used_ratio = r.weighted_point_used_ratio
        m.point_used_ratio = r.point_used_ratio
        m.residual = r.residual
        for i in range(36):
            m.covariance[i] = r.covariance[i]
        self._last = m

    def __del__(self):
        try:
            if self._own and self.h and self.ctx.h:
                self.ctx.lib.pgs_minimizer_destroy(self.h)
        except Exception:
            pass


def _icp_result(r: IcpResult) -> dict:
    return dict(T=np.array(r.T).reshape(4, 4).T.copy(), covariance=np.array(r.covariance).reshape(6, 6).T.copy(),
                iterations=int(r.iterations), max_iterations_reached=bool(r.max_iterations_reached),
                status=int(r.status), overlap=r.overlap, weighted_point_used_ratio=r.weighted_point_used_ratio,
                point_used_ratio=r.point_used_ratio, residual=r.residual, n_reading=int(r.n_reading),
                n_reference=int(r.n_reference))


class ICP:
    """PM::ICP (types.h:24): loadFromYaml / setDefault, operator(), and the
    public ICPChainBase members pgslam drives one by one."""

    def __init__(self, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        self.h = None
        self.last = None
        self.setDefault()

    def _wire(self):
        L = self.ctx.lib
        c = self.ctx
        self.readingDataPointsFilters = DataPointsFilters(ctx=c, _borrowed=_vp(L.pgs_icp_reading_filters(self.h)))
        self.readingStepDataPointsFilters = DataPointsFilters(ctx=c, _borrowed=_vp(L.pgs_icp_reading_step_filters(self.h)))
        self.referenceDataPointsFilters = DataPointsFilters(ctx=c, _borrowed=_vp(L.pgs_icp_reference_filters(self.h)))
        self.matcher = Matcher(ctx=c, _borrowed=_vp(L.pgs_icp_matcher(self.h)))
        self.outlierFilters = OutlierFilters(ctx=c, _borrowed=_vp(L.pgs_icp_outliers(self.h)))
        self.errorMinimizer = ErrorMinimizer(ctx=c, _borrowed=_vp(L.pgs_icp_minimizer(self.h)))
        self.transformations = Transformations(c)

    def _replace(self, h):
        if self.h:
            self.ctx.lib.pgs_icp_destroy(self.h)
        self.h = h
        self._wire()

    def setDefault(self):
        h = _vp()
        self.ctx.check(self.ctx.lib.pgs_icp_create_default(self.ctx.h, C.byref(h)))
        self._replace(h)

    def loadFromYaml(self, text: str):
        b = text.encode()
        h = _vp()
        self.ctx.check(self.ctx.lib.pgs_icp_create_from_yaml(self.ctx.h, b, len(b), C.byref(h)))
        self._replace(h)

    def _finish(self, st, r: IcpResult):
        self.last = _icp_result(r)
        self.errorMinimizer._set_from_icp(r)
        self.ctx.check(st)
        return self.last["T"]

    def __call__(self, reading: DataPoints, reference: DataPoints, T_init=None) -> np.ndarray:
        T = _mat(T_init)
        r = IcpResult()
        st = self.ctx.lib.pgs_icp_run(self.h, reading.h, reference.h, T.ctypes.data_as(_dp), C.byref(r))
        return self._finish(st, r)

    def getMaxNumIterationsReached(self) -> bool:
        """The accessor pgslam's author patched in (LoopCloser.hpp:310-318)."""
        return bool(self.last and self.last["max_iterations_reached"])

    def compute_batch(self, readings: list[DataPoints], references: list[DataPoints], T_inits=None) -> list[dict]:
        """P independent registrations run concurrently on this context's GPU."""
        P = len(readings)
        ra = (_vp * P)(*[c.h for c in readings])
        fa = (_vp * P)(*[c.h for c in references])
        Tp = None
        if T_inits is not None:
            Tflat = np.concatenate([_mat(t).ravel(order="F") for t in T_inits])
            Tp = Tflat.ctypes.data_as(_dp)
        res = (IcpResult * P)()
        st = self.ctx.lib.pgs_icp_run_batch(self.h, P, ra, fa, Tp, res)
        out = [_icp_result(r) for r in res]
        self.ctx.check(st)
        return out

    def compute_batch_array(self, readings, references, T_inits=None, handles=None) -> np.ndarray:
        """compute_batch without per-pair Python objects: returns the pgs_icp_result records as one
        structured numpy array (fields T [col-major 16], covariance, iterations, status, ...).
        `handles` = (reading handle array, reference handle array) from batch_handles(), reusable."""
        ra, fa = handles if handles is not None else batch_handles(readings, references)
        P = len(ra)
        Tp = None
        if T_inits is not None:
            Tflat = np.ascontiguousarray(np.concatenate([_mat(t).ravel(order="F") for t in T_inits]))
            Tp = Tflat.ctypes.data_as(_dp)
        res = (IcpResult * P)()
        st = self.ctx.lib.pgs_icp_run_batch(self.h, P, ra, fa, Tp, res)
        out = np.frombuffer(res, dtype=np.dtype(IcpResult)).copy()
        self.ctx.check(st)
        return out

    def probe_overlap(self, reading, reference, T) -> float:
        """Localizer::ComputeOverlapWith (Localizer.hpp:282-348), one fused call."""
        T = _mat(T)
        out = C.c_double(0)
        self.ctx.check(self.ctx.lib.pgs_icp_probe_overlap(self.h, reading.h, reference.h, T.ctypes.data_as(_dp), C.byref(out)))
        return out.value

    def probe_residual(self, reading, reference, T) -> float:
        """LoopCloser::ComputeResidualError (LoopCloser.hpp:343-365), one fused call."""
        T = _mat(T)
        out = C.c_double(0)
        self.ctx.check(self.ctx.lib.pgs_icp_probe_residual(self.h, reading.h, reference.h, T.ctypes.data_as(_dp), C.byref(out)))
        return out.value

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.pgs_icp_destroy(self.h)
        except Exception:
            pass


def empty_records() -> np.ndarray:
    return np.zeros((0,), dtype=np.dtype(IcpResult))


def batch_handles(readings, references):
    """ctypes handle arrays for compute_batch_array (build once, reuse every call)."""
    P = len(readings)
    return (_vp * P)(*[c.h for c in readings]), (_vp * P)(*[c.h for c in references])


def host_clouds(arrays, descriptors=None):
    """pgs_host_cloud array over host buffers.  `arrays`: (N, 4) float32 C-contiguous numpy arrays
    ({x,y,z,1} per point: PM's column-major 4 x N), or (pointer, n) tuples for raw (e.g. pinned)
    memory.  `descriptors`: optional list of {label: (N, span) float32 array}.  The returned
    object keeps every buffer alive."""
    P = len(arrays)
    out = (HostCloud * P)()
    keep = [out]
    for i, a in enumerate(arrays):
        if isinstance(a, tuple):
            ptr, n = a
        else:
            if a.dtype != np.float32 or a.ndim != 2 or a.shape[1] != 4 or not a.flags["C_CONTIGUOUS"]:
                raise InvalidField(INVALID_FIELD, "host clouds must be (N, 4) float32 C-contiguous")
            keep.append(a)
            ptr, n = a.ctypes.data, a.shape[0]
        out[i].features4xN = ptr
        out[i].n = n
        d = (descriptors[i] if descriptors else None) or {}
        out[i].n_descriptors = len(d)
        if d:
            labels = (C.c_char_p * len(d))(*[k.encode() for k in d])
            blocks = [np.ascontiguousarray(v, dtype=np.float32) for v in d.values()]
            spans = (C.c_int * len(d))(*[(b.shape[1] if b.ndim == 2 else 1) for b in blocks])
            data = (C.c_void_p * len(d))(*[b.ctypes.data for b in blocks])
            out[i].labels, out[i].spans, out[i].data = labels, spans, data
            keep += [labels, blocks, spans, data]
    out._keep = keep
    return out


def compute_batch_multi(icps, readings, references, T_inits=None, pinned=False) -> np.ndarray:
    """pgs_icp_run_batch_multi: host-resident pairs registered on the contexts of `icps` (one ICP
    object per GPU, same chain), contiguous-block sharded; returns the pgs_icp_result records as
    a structured numpy array in pair order.  `readings` / `references`: host_clouds(...) arrays."""
    P = len(readings)
    hs = (_vp * len(icps))(*[i.h for i in icps])
    Tp = None
    if T_inits is not None:
        Tflat = np.ascontiguousarray(np.concatenate([_mat(t).ravel(order="F") for t in T_inits]))
        Tp = Tflat.ctypes.data_as(_dp)
    res = (IcpResult * max(P, 1))()
    st = icps[0].ctx.lib.pgs_icp_run_batch_multi(hs, len(icps), P, readings, references, Tp, int(bool(pinned)), res)
    out = np.frombuffer(res, dtype=np.dtype(IcpResult)).copy()[:P]
    icps[0].ctx.check(st)
    return out


class ICPSequence(ICP):
    """PM::ICPSequence (types.h:25): setMap / hasMap / operator()(reading, T)."""

    def setMap(self, cloud: DataPoints):
        self.ctx.check(self.ctx.lib.pgs_icp_set_map(self.h, cloud.h))

    def hasMap(self) -> bool:
        return bool(self.ctx.lib.pgs_icp_has_map(self.h))

    def __call__(self, reading: DataPoints, T_init=None, _unused=None) -> np.ndarray:
        T = _mat(T_init)
        r = IcpResult()
        st = self.ctx.lib.pgs_icp_run_sequence(self.h, reading.h, T.ctypes.data_as(_dp), C.byref(r))
        return self._finish(st, r)
