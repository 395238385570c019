"""ctypes binding of oracle/libicp_oracle.so -- the CPU checker (test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

from norlab_icp_mapper_b200._abi import Config, Result

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "libicp_oracle.so")
_lib = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build():
    subprocess.run(["make", "-s", "-C", os.path.join(_ROOT, "oracle")], check=True)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(_ROOT, "oracle", "icp_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        build()
    L = C.CDLL(_SO)
    L.orc_kdtree_build.restype = C.c_void_p
    L.orc_kdtree_build.argtypes = [f32p, C.c_int32, C.c_int64, C.c_int32]
    L.orc_kdtree_free.argtypes = [C.c_void_p]
    L.orc_kdtree_knn.argtypes = [C.c_void_p, f32p, C.c_int32, C.c_int64, C.c_int32, C.c_float,
                                 i32p, f32p, C.c_int32]
    L.orc_kdtree_knn_ex.argtypes = [C.c_void_p, f32p, C.c_int32, C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_int32,
                                    i32p, f32p, C.c_int32]
    L.orc_icp_set_reading_max_search_dist.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    L.orc_icp_create.restype = C.c_void_p
    L.orc_icp_create.argtypes = [C.POINTER(Config)]
    L.orc_icp_destroy.argtypes = [C.c_void_p]
    L.orc_icp_last_error.restype = C.c_char_p
    L.orc_icp_last_error.argtypes = [C.c_void_p]
    L.orc_icp_set_map.argtypes = [C.c_void_p, f32p, C.c_int32, C.c_void_p, C.c_int64]
    L.orc_icp_get_mean.argtypes = [C.c_void_p, f32p]
    L.orc_icp_register.argtypes = [C.c_void_p, f32p, C.c_int32, C.c_int64, C.c_void_p, f32p,
                                   C.POINTER(Result), C.c_void_p, C.c_void_p, C.c_int32]
    L.orc_icp_match.argtypes = [C.c_void_p, f32p, C.c_int32, C.c_int64, i32p, f32p, C.c_int32]
    L.orc_transform.argtypes = [f32p, C.c_int32, C.c_void_p, C.c_int64, f32p]
    L.orc_point_distance_keep.restype = C.c_int64
    L.orc_point_distance_keep.argtypes = [f32p, C.c_int32, C.c_int64, f32p, C.c_int64, C.c_float,
                                          u8p, C.c_int32]
    L.orc_surface_normals.argtypes = [f32p, C.c_int32, C.c_int64, C.c_int32, f32p, C.c_int32]
    L.orc_num_threads.restype = C.c_int32
    _lib = L
    return L


def _cloud(a):
    """(N, rows) C-contiguous fp32 == the reference's column-major rows x N Eigen matrix."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2
    return a


def knn(ref, queries, k, dim=None, max_radius=np.inf, nthreads=0, max_radii=None, strict=False):
    """max_radii: one search radius per query (libnabo's vector-of-radii overload); strict: dist2 < r^2 instead of <=."""
    ref, queries = _cloud(ref), _cloud(queries)
    dim = dim or min(ref.shape[1], 3)
    L = lib()
    t = L.orc_kdtree_build(ref, ref.shape[1], ref.shape[0], dim)
    assert t
    ids = np.empty((queries.shape[0], k), np.int32)
    d2 = np.empty((queries.shape[0], k), np.float32)
    rptr = None
    if max_radii is not None:
        max_radii = np.ascontiguousarray(max_radii, np.float32)
        assert max_radii.shape == (queries.shape[0],)
        rptr = max_radii.ctypes.data_as(C.c_void_p)
    L.orc_kdtree_knn_ex(t, queries, queries.shape[1], queries.shape[0], k, max_radius, rptr, int(strict), ids, d2, nthreads)
    L.orc_kdtree_free(t)
    return ids, d2


class OracleICP:
    """PM::ICPSequence restated on the CPU."""

    def __init__(self, cfg: Config):
        self.cfg = cfg
        self.n = cfg.dim + 1
        self._h = lib().orc_icp_create(C.byref(cfg))
        if not self._h:
            raise ValueError("bad config")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_icp_destroy(self._h)
            self._h = None

    def last_error(self):
        return lib().orc_icp_last_error(self._h).decode()

    def last_robust_scale(self):
        L = lib()
        L.orc_icp_last_robust_scale.restype = C.c_float
        L.orc_icp_last_robust_scale.argtypes = [C.c_void_p]
        return float(L.orc_icp_last_robust_scale(self._h))

    def set_map(self, features, normals=None):
        features = _cloud(features)
        nptr = None
        if normals is not None:
            normals = _cloud(normals)
            nptr = normals.ctypes.data_as(C.c_void_p)
        return lib().orc_icp_set_map(self._h, features, features.shape[1], nptr, features.shape[0])

    def mean(self):
        m = np.zeros(3, np.float32)
        lib().orc_icp_get_mean(self._h, m)
        return m

    def match(self, queries, nthreads=0):
        queries = _cloud(queries)
        k = self.cfg.knn
        ids = np.empty((queries.shape[0], k), np.int32)
        d2 = np.empty((queries.shape[0], k), np.float32)
        rc = lib().orc_icp_match(self._h, queries, queries.shape[1], queries.shape[0], ids, d2, nthreads)
        return rc, ids, d2

    def set_reading_max_search_dist(self, radii):
        """The reading's `maxSearchDist` descriptor for the following register / match calls (None clears it)."""
        if radii is None:
            lib().orc_icp_set_reading_max_search_dist(self._h, None, 0)
        else:
            r = np.ascontiguousarray(radii, np.float32)
            lib().orc_icp_set_reading_max_search_dist(self._h, r.ctypes.data_as(C.c_void_p), len(r))

    def register(self, reading, T_init=None, nthreads=0, want_trace=False, reading_normals=None):
        """Returns (status, T (n x n, row-major numpy view of the math matrix), Result, trace, secs)."""
        reading = _cloud(reading)
        L = lib()
        L.orc_icp_set_reading_normals.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        if reading_normals is not None:
            rn = np.ascontiguousarray(reading_normals, np.float32)
            assert rn.shape == (reading.shape[0], self.n - 1)
            L.orc_icp_set_reading_normals(self._h, rn.ctypes.data_as(C.c_void_p), rn.shape[0])
        else:
            L.orc_icp_set_reading_normals(self._h, None, 0)
        n = self.n
        T_out = np.zeros(n * n, np.float32)
        res = Result()
        tptr = None
        if T_init is not None:
            T_cm = np.ascontiguousarray(np.asarray(T_init, np.float32).T).ravel()  # column-major
            tptr = T_cm.ctypes.data_as(C.c_void_p)
        trace = None
        trptr = None
        if want_trace:
            trace = np.zeros((max(self.cfg.max_iteration_count, 1) + 1, n * n), np.float32)
            trptr = trace.ctypes.data_as(C.c_void_p)
        secs = np.zeros(4, np.float64)
        rc = lib().orc_icp_register(self._h, reading, reading.shape[1], reading.shape[0], tptr, T_out,
                                    C.byref(res), trptr, secs.ctypes.data_as(C.c_void_p), nthreads)
        T = T_out.reshape(n, n).T.copy()
        if trace is not None:
            trace = trace[:res.iterations].reshape(-1, n, n).transpose(0, 2, 1).copy()
        return rc, T, res, trace, secs


def transform(features, T, normals=None):
    features = _cloud(features).copy()
    n = features.shape[1]
    T_cm = np.ascontiguousarray(np.asarray(T, np.float32).T).ravel()
    nptr = None
    if normals is not None:
        normals = _cloud(normals).copy()
        nptr = normals.ctypes.data_as(C.c_void_p)
    rc = lib().orc_transform(features, n, nptr, features.shape[0], T_cm)
    return rc, features, normals


def point_distance_keep(map_feat, input_feat, min_dist, nthreads=0):
    map_feat, input_feat = _cloud(map_feat), _cloud(input_feat)
    keep = np.zeros(input_feat.shape[0], np.uint8)
    kept = lib().orc_point_distance_keep(map_feat, map_feat.shape[1], map_feat.shape[0], input_feat,
                                         input_feat.shape[0], min_dist, keep, nthreads)
    return kept, keep.astype(bool)


def surface_normals(feat, knn_, nthreads=0):
    feat = _cloud(feat)
    dim = feat.shape[1] - 1
    out = np.zeros((feat.shape[0], dim), np.float32)
    rc = lib().orc_surface_normals(feat, feat.shape[1], feat.shape[0], knn_, out, nthreads)
    return rc, out
