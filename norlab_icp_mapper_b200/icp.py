"""Python face of the B200 ICP core, shaped like the reference's use of `PM::ICPSequence`
(/root/reference/norlab_icp_mapper/Mapper.h:23): `setMap(cloud)` (Map.cpp:111,178,528,581),
`icp(input)` (Mapper.cpp:213) and `errorMinimizer->getOverlap()` (Mapper.cpp:219).

Clouds are numpy arrays of shape (N, dim+1), C-contiguous fp32 -- byte-identical to the reference's
column-major (dim+1) x N `DataPoints::features`.  Transforms are returned as ordinary (dim+1, dim+1)
numpy matrices (row-major view of the math matrix).  Everything runs in libb200icp.so on the GPU;
errors that libpointmatcher raises as C++ exceptions surface as B200ICPError with the C-ABI status.
"""
import ctypes as C

import numpy as np

from . import _abi
from ._abi import Config, Result, Timing, make_config  # noqa: F401  (re-exported)
from ._lib import B200ICPError, load


def _cloud(a, rows=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or (rows is not None and a.shape[1] != rows):
        raise ValueError(f"expected an (N, {rows}) array, got {a.shape}")
    return a


def _T_to_colmajor(T, n):
    T = np.asarray(T, dtype=np.float32)
    if T.shape != (n, n):
        raise ValueError(f"expected a ({n}, {n}) transform")
    return np.ascontiguousarray(T.T).ravel()


class ICP:
    """One `PM::ICPSequence`: a context bound to one GPU and one stream."""

    def __init__(self, cfg: Config, device: int = 0):
        self._L = load()
        self.cfg = cfg
        self.dim = cfg.dim
        self.n = cfg.dim + 1
        h = C.c_void_p()
        rc = self._L.b200icp_create(C.byref(cfg), device, C.byref(h))
        if rc != _abi.OK:
            raise B200ICPError(rc, self._L.b200icp_last_error(None).decode())
        self._h = h
        self.last_result = Result()

    def close(self):
        if getattr(self, "_h", None):
            self._L.b200icp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != _abi.OK:
            raise B200ICPError(rc, self._L.b200icp_last_error(self._h).decode())

    # -- icp.setMap(cloud) -------------------------------------------------------------------
    def set_map(self, features, normals=None):
        features = _cloud(features, self.n)
        nptr = None
        if normals is not None:
            normals = _cloud(normals, self.dim)
            if len(normals) != len(features):
                raise ValueError("normals and features differ in length")
            nptr = normals.ctypes.data
        self._check(self._L.b200icp_set_map(self._h, features.ctypes.data, self.n, nptr, len(features)))
        return len(features) > 0

    def set_map_device(self, d_features_ptr, n, d_normals_ptr=None):
        self._check(self._L.b200icp_set_map_device(self._h, d_features_ptr, self.n, d_normals_ptr, n))

    def has_map(self):
        return self._L.b200icp_map_size(self._h) > 0

    def map_mean(self):
        m = np.zeros(3, np.float32)
        self._check(self._L.b200icp_get_map_mean(self._h, m.ctypes.data))
        return m

    def grid_info(self):
        h = C.c_float()
        dims = np.zeros(3, np.int32)
        self._check(self._L.b200icp_get_grid_info(self._h, C.byref(h), dims.ctypes.data))
        return h.value, tuple(int(x) for x in dims)

    # -- correction = icp(input) ---------------------------------------------------------------
    def __call__(self, reading, T_init=None, reading_normals=None, reading_max_search_dist=None):
        """reading_normals: the reading's `normals` descriptor (N x dim), needed by SurfaceNormalOutlierFilter only.
        reading_max_search_dist: the reading's `maxSearchDist` descriptor (N radii): per-point search radius instead of maxDist."""
        reading = _cloud(reading, self.n)
        tptr = None
        if T_init is not None:
            T_cm = _T_to_colmajor(T_init, self.n)
            tptr = T_cm.ctypes.data
        T_out = np.zeros(self.n * self.n, np.float32)
        res = Result()
        if reading_max_search_dist is not None:
            rm = np.ascontiguousarray(reading_max_search_dist, np.float32)
            assert rm.shape == (len(reading),)
            rn = None if reading_normals is None else _cloud(reading_normals, self.dim)
            rc = self._L.b200icp_register_descriptors(self._h, reading.ctypes.data, self.n, len(reading), None if rn is None else rn.ctypes.data,
                                                      rm.ctypes.data, tptr, T_out.ctypes.data, C.byref(res))
        elif reading_normals is not None:
            rn = _cloud(reading_normals, self.dim)
            assert len(rn) == len(reading)
            rc = self._L.b200icp_register_normals(self._h, reading.ctypes.data, self.n, len(reading), rn.ctypes.data, tptr, T_out.ctypes.data,
                                                  C.byref(res))
        else:
            rc = self._L.b200icp_register(self._h, reading.ctypes.data, self.n, len(reading), tptr, T_out.ctypes.data, C.byref(res))
        self.last_result = res
        self._check(rc)
        return T_out.reshape(self.n, self.n).T.copy()

    def register_device(self, d_reading_ptr, nq, T_init=None):
        tptr = None
        if T_init is not None:
            T_cm = _T_to_colmajor(T_init, self.n)
            tptr = T_cm.ctypes.data
        T_out = np.zeros(self.n * self.n, np.float32)
        res = Result()
        rc = self._L.b200icp_register_device(self._h, d_reading_ptr, self.n, nq, tptr, T_out.ctypes.data, C.byref(res))
        self.last_result = res
        self._check(rc)
        return T_out.reshape(self.n, self.n).T.copy()

    def get_overlap(self):
        """errorMinimizer->getOverlap() of the last registration (Mapper.cpp:219)."""
        return float(self.last_result.overlap)

    # -- matcher->findClosests / Nabo::NNS::knn ------------------------------------------------
    def match(self, queries):
        queries = _cloud(queries, self.n)
        k = self.cfg.knn
        ids = np.empty((len(queries), k), np.int32)
        d2 = np.empty((len(queries), k), np.float32)
        self._check(self._L.b200icp_match(self._h, queries.ctypes.data, self.n, len(queries), ids.ctypes.data, d2.ctypes.data))
        return ids, d2

    def knn(self, ref, queries, k, dim=None, max_radius=float("inf")):
        ref, queries = _cloud(ref), _cloud(queries)
        dim = dim or self.dim
        ids = np.empty((len(queries), k), np.int32)
        d2 = np.empty((len(queries), k), np.float32)
        self._check(self._L.b200icp_knn(self._h, ref.ctypes.data, ref.shape[1], len(ref), queries.ctypes.data,
                                        queries.shape[1], len(queries), dim, k, max_radius, ids.ctypes.data, d2.ctypes.data))
        return ids, d2

    # -- RigidTransformation::compute ------------------------------------------------------------
    def transform(self, features, T, normals=None):
        features = _cloud(features).copy()
        rows = features.shape[1]
        T_cm = _T_to_colmajor(T, rows)
        nptr = None
        if normals is not None:
            normals = _cloud(normals, rows - 1).copy()
            nptr = normals.ctypes.data
        self._check(self._L.b200icp_transform(self._h, features.ctypes.data, rows, nptr, len(features), T_cm.ctypes.data))
        return features, normals

    # -- device-resident local map (Map::updateLocalPointCloud / updatePose pieces) -----------------
    def map_insert_point_distance(self, input_features, min_dist_new_point, input_normals=None, want_keep=False):
        """PointDistanceMapperModule::inPlaceUpdateMap on the device map.  Returns (n_added, keep mask or None)."""
        inp = _cloud(input_features, self.n)
        nptr = None
        if input_normals is not None:
            input_normals = _cloud(input_normals, self.dim)
            nptr = input_normals.ctypes.data
        added = C.c_int64()
        keep = np.zeros(len(inp), np.uint8) if want_keep else None
        self._check(self._L.b200icp_map_insert_point_distance(self._h, inp.ctypes.data, self.n, len(inp), nptr, min_dist_new_point,
                                                              C.byref(added), keep.ctypes.data if want_keep else None))
        return added.value, (keep.astype(bool) if want_keep else None)

    def map_surface_normals(self, knn):
        self._check(self._L.b200icp_map_surface_normals(self._h, knn))

    def debug_selfknn_redone(self):
        """Queries of the last staged self k-NN (selfknn.cu) that the shell walk had to redo; -1 = staged kernel not used."""
        self._L.b200icp_debug_selfknn_redone.argtypes = [C.c_void_p]
        self._L.b200icp_debug_selfknn_redone.restype = C.c_int64
        return self._L.b200icp_debug_selfknn_redone(self._h)

    def cloud_surface_normals(self, features, knn):
        """SurfaceNormalDataPointsFilter{knn} on a host cloud (N x (dim + 1)); returns N x dim normals."""
        pts = _cloud(features, self.n)
        out = np.zeros((len(pts), self.dim), np.float32)
        self._check(self._L.b200icp_cloud_surface_normals(self._h, pts.ctypes.data, self.n, len(pts), knn, out.ctypes.data))
        return out

    def map_window(self, load, slab):
        slab = np.ascontiguousarray(slab, np.int32)
        assert slab.shape == (6,)
        changed = C.c_int64()
        self._check(self._L.b200icp_map_window(self._h, int(load), slab.ctypes.data, C.byref(changed)))
        return changed.value

    def map_commit(self):
        self._check(self._L.b200icp_map_commit(self._h))

    def map_counts(self):
        a, b = C.c_int64(), C.c_int64()
        self._check(self._L.b200icp_map_counts(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def map_download(self, global_map=False):
        n = C.c_int64()
        self._check(self._L.b200icp_map_download(self._h, int(global_map), None, None, 0, C.byref(n)))
        feat = np.zeros((n.value, self.n), np.float32)
        has_n = bool(self._L.b200icp_map_has_normals(self._h))
        nrm = np.zeros((n.value, self.dim), np.float32) if has_n else None
        if n.value:
            self._check(self._L.b200icp_map_download(self._h, int(global_map), feat.ctypes.data, nrm.ctypes.data if has_n else None,
                                                     n.value, C.byref(n)))
        return feat[:n.value], (nrm[:n.value] if has_n else None)

    def map_set_prob(self, prob=None, constant=0.6):
        ptr = None
        if prob is not None:
            prob = np.ascontiguousarray(prob, np.float32)
            ptr = prob.ctypes.data
        self._check(self._L.b200icp_map_set_prob(self._h, ptr, constant))

    def map_download_prob(self, global_map=False):
        n = self.map_counts()[1 if global_map else 0]
        out = np.zeros(n, np.float32)
        if n:
            self._check(self._L.b200icp_map_download_prob(self._h, int(global_map), out.ctypes.data, n))
        return out

    def map_octree(self, input_features, max_size_by_node, sampling_method=0, input_normals=None, input_prob=None, max_point_by_node=1):
        """OctreeMapperModule::inPlaceUpdateMap.  Returns the local map size afterwards."""
        inp = _cloud(input_features, self.n)
        nptr = pptr = None
        if input_normals is not None:
            input_normals = _cloud(input_normals, self.dim)
            nptr = input_normals.ctypes.data
        if input_prob is not None:
            input_prob = np.ascontiguousarray(input_prob, np.float32)
            pptr = input_prob.ctypes.data
        n_after = C.c_int64()
        self._check(self._L.b200icp_map_octree(self._h, inp.ctypes.data, self.n, len(inp), nptr, pptr, max_size_by_node,
                                               max_point_by_node, sampling_method, C.byref(n_after)))
        return n_after.value

    def filter_cloud(self, features, filters):
        """DataPointsFilters::apply for the `input:` chain (Mapper.cpp:187-191); filters: mapper.InputFilter entries.
        Returns the surviving points in input order."""
        pts = _cloud(features, self.n).copy()
        n = C.c_int64(len(pts))
        arr = (type(filters[0]) * len(filters))(*filters) if len(filters) else None
        self._check(self._L.b200icp_filter_cloud(self._h, pts.ctypes.data, self.n, C.byref(n), arr, len(filters)))
        return pts[:n.value]

    def map_cut_at_threshold(self, threshold, use_larger_than=True):
        n = C.c_int64()
        self._check(self._L.b200icp_map_cut_at_threshold(self._h, threshold, int(use_larger_than), C.byref(n)))
        return n.value

    def map_dynamic_points(self, input_features, input_prob, pose, params=None):
        """DynamicPointsMapperModule::inPlaceUpdateMap (input in the map frame)."""
        inp = _cloud(input_features, self.n)
        pptr = None
        if input_prob is not None:
            input_prob = np.ascontiguousarray(input_prob, np.float32)
            pptr = input_prob.ctypes.data
        T = _T_to_colmajor(pose, self.n)
        params = params or _abi.DynamicParams()
        self._check(self._L.b200icp_map_dynamic_points(self._h, inp.ctypes.data, self.n, len(inp), pptr, T.ctypes.data, C.byref(params)))

    # -- the map's other descriptors (intensity, t, ring ...) ------------------------------------------
    def map_set_extra(self, extra):
        """extra: N x rows (one row of the array per map point), or None to drop them."""
        if extra is None:
            self._check(self._L.b200icp_map_set_extra(self._h, None, 0))
            return
        extra = np.ascontiguousarray(extra, np.float32)
        self._check(self._L.b200icp_map_set_extra(self._h, extra.ctypes.data, extra.shape[1]))

    def map_extra_rows(self):
        return self._L.b200icp_map_extra_rows(self._h)

    def map_select_extra(self, rows):
        rows = np.ascontiguousarray(rows, np.int32)
        self._check(self._L.b200icp_map_select_extra(self._h, rows.ctypes.data, len(rows)))

    def map_download_extra(self, global_map=False):
        n = self.map_counts()[1 if global_map else 0]
        out = np.zeros((n, self.map_extra_rows()), np.float32)
        if n and out.shape[1]:
            self._check(self._L.b200icp_map_download_extra(self._h, int(global_map), out.ctypes.data, n))
        return out

    def map_replace_local(self, features, normals=None, prob=None, extra=None):
        """localPointCloud = cloud (what a host-signature MapperModule returned)."""
        f = _cloud(features, self.n)
        nrm = None if normals is None else _cloud(normals, self.dim)
        pr = None if prob is None else np.ascontiguousarray(prob, np.float32)
        ex = None if extra is None else np.ascontiguousarray(extra, np.float32)
        self._check(self._L.b200icp_map_replace_local(self._h, f.ctypes.data, self.n, len(f), None if nrm is None else nrm.ctypes.data,
                                                      None if pr is None else pr.ctypes.data, None if ex is None else ex.ctypes.data,
                                                      0 if ex is None else ex.shape[1]))

    def map_insert_point_distance_prob(self, input_features, min_dist_new_point, input_normals=None, input_prob=None):
        inp = _cloud(input_features, self.n)
        nrm = None if input_normals is None else _cloud(input_normals, self.dim)
        pr = None if input_prob is None else np.ascontiguousarray(input_prob, np.float32)
        added = C.c_int64()
        self._check(self._L.b200icp_map_insert_point_distance_prob(self._h, inp.ctypes.data, self.n, len(inp), None if nrm is None else nrm.ctypes.data,
                                                                   None if pr is None else pr.ctypes.data, min_dist_new_point, C.byref(added), None))
        return added.value

    # -- the device-resident scan slot (one upload per scan; b200icp_scan_*) ---------------------------
    def scan_upload(self, features, normals=None, prob=None, extra=None, rotating_rows=()):
        f = _cloud(features, self.n)
        self._check(self._L.b200icp_scan_upload(self._h, f.ctypes.data, self.n, len(f)))
        if normals is not None or prob is not None or extra is not None:
            nrm = None if normals is None else _cloud(normals, self.dim)
            pr = None if prob is None else np.ascontiguousarray(prob, np.float32)
            ex = None if extra is None else np.ascontiguousarray(extra, np.float32)
            rot = np.ascontiguousarray(rotating_rows, np.int32)
            self._check(self._L.b200icp_scan_set_descriptors(self._h, None if nrm is None else nrm.ctypes.data, None if pr is None else pr.ctypes.data,
                                                             None if ex is None else ex.ctypes.data, 0 if ex is None else ex.shape[1],
                                                             rot.ctypes.data if len(rot) else None, len(rot)))

    def scan_info(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        self._check(self._L.b200icp_scan_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"n": self._L.b200icp_scan_size(self._h), "normals": bool(a.value), "prob": bool(b.value), "extra_rows": c.value}

    def scan_filter(self, filters):
        arr = (type(filters[0]) * len(filters))(*filters) if len(filters) else None
        n = C.c_int64()
        self._check(self._L.b200icp_scan_filter(self._h, arr, len(filters), C.byref(n)))
        return n.value

    def scan_add_prob(self, constant):
        self._check(self._L.b200icp_scan_add_prob(self._h, constant))

    def scan_surface_normals(self, knn):
        self._check(self._L.b200icp_scan_surface_normals(self._h, knn))

    def scan_select_extra(self, rows):
        rows = np.ascontiguousarray(rows, np.int32)
        self._check(self._L.b200icp_scan_select_extra(self._h, rows.ctypes.data, len(rows)))

    def scan_transform(self, T):
        T_cm = _T_to_colmajor(T, self.n)
        self._check(self._L.b200icp_scan_transform(self._h, T_cm.ctypes.data))

    def scan_register(self, T_init=None):
        T_out = np.zeros(self.n * self.n, np.float32)
        Ti = None if T_init is None else _T_to_colmajor(T_init, self.n)
        res = Result()
        self._check(self._L.b200icp_scan_register(self._h, None if Ti is None else Ti.ctypes.data, T_out.ctypes.data, C.byref(res)))
        self.last_result = res
        return T_out.reshape(self.n, self.n).T.copy()

    def scan_insert_point_distance(self, min_dist_new_point):
        added = C.c_int64()
        self._check(self._L.b200icp_scan_insert_point_distance(self._h, min_dist_new_point, C.byref(added)))
        return added.value

    def scan_append(self):
        added = C.c_int64()
        self._check(self._L.b200icp_scan_append(self._h, C.byref(added)))
        return added.value

    def scan_octree(self, max_size_by_node, sampling_method=0, max_point_by_node=1):
        n_after = C.c_int64()
        self._check(self._L.b200icp_scan_octree(self._h, max_size_by_node, max_point_by_node, sampling_method, C.byref(n_after)))
        return n_after.value

    def scan_dynamic_points(self, pose, params=None):
        T = _T_to_colmajor(pose, self.n)
        params = params or _abi.DynamicParams()
        self._check(self._L.b200icp_scan_dynamic_points(self._h, T.ctypes.data, C.byref(params)))

    def scan_download(self):
        """Returns (features, normals or None, prob or None, extra or None)."""
        info = self.scan_info()
        n = info["n"]
        f = np.zeros((n, self.n), np.float32)
        got = C.c_int64()
        self._check(self._L.b200icp_scan_download(self._h, f.ctypes.data, n, C.byref(got)))
        nrm = np.zeros((n, self.dim), np.float32) if info["normals"] else None
        pr = np.zeros(n, np.float32) if info["prob"] else None
        ex = np.zeros((n, info["extra_rows"]), np.float32) if info["extra_rows"] else None
        if nrm is not None or pr is not None or ex is not None:
            self._check(self._L.b200icp_scan_download_descriptors(self._h, None if nrm is None else nrm.ctypes.data, None if pr is None else pr.ctypes.data,
                                                                  None if ex is None else ex.ctypes.data, n))
        return f, nrm, pr, ex

    # -- instrumentation ---------------------------------------------------------------------------
    def stream(self):
        return self._L.b200icp_stream(self._h)

    def set_sm_share(self, n_sms):
        """SMs the registration loop may occupy (0 = all): contexts working side by side on one GPU each take a share."""
        self._check(self._L.b200icp_set_sm_share(self._h, int(n_sms)))

    def set_profiling(self, on=True):
        self._check(self._L.b200icp_set_profiling(self._h, int(on)))

    def set_trace(self, on=True):
        self._check(self._L.b200icp_set_trace(self._h, int(on)))

    def trace(self):
        have = self._L.b200icp_get_trace(self._h, None, 0)
        if have <= 0:
            return np.zeros((0, self.n, self.n), np.float32)
        buf = np.zeros((have, self.n * self.n), np.float32)
        self._L.b200icp_get_trace(self._h, buf.ctypes.data, have)
        return buf.reshape(have, self.n, self.n).transpose(0, 2, 1).copy()

    def timing(self):
        t = Timing()
        self._check(self._L.b200icp_get_timing(self._h, C.byref(t)))
        return t
