"""Python face of the host-side mirror of `norlab_icp_mapper::Mapper` (libb200mapper.so, C++ in
norlab_icp_mapper_b200/host/).  Method names follow the reference's pybind11 module
(/root/reference/python/src/mapper.cpp:11-25): Mapper(config, is3D, isOnline, isMapping,
saveMapCellsOnHardDrive), applyInputFilters, processInput, getMap, setMap, getNewLocalMap, getPose,
getIsMapping, setIsMapping, getTrajectory."""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._lib import B200ICPError, load as _load_icp

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libb200mapper.so")
SYMBOLS = ["b200mapper_create", "b200mapper_destroy", "b200mapper_last_error", "b200mapper_apply_input_filters",
           "b200mapper_process_input", "b200mapper_get_pose", "b200mapper_get_map", "b200mapper_get_new_local_map",
           "b200mapper_set_map", "b200mapper_get_is_mapping", "b200mapper_set_is_mapping", "b200mapper_trajectory_size",
           "b200mapper_get_trajectory", "b200mapper_get_stats", "b200mapper_get_window_updates", "b200mapper_process_raw_input",
           "b200mapper_vtk_save", "b200mapper_vtk_load",
           "b200mapper_set_map_descriptors", "b200mapper_get_map_prob", "b200mapper_create_from_yaml", "b200mapper_yaml_summary", "b200mapper_map_update_in_flight", "b200mapper_wait_for_map_update", "b200mapper_get_local_map"]


class InputFilter(C.Structure):
    """b200icp_filter: kind 1 = BoundingBox{lo, hi, removeInside}, 2 = DistanceLimit{dim, dist, removeInside},
    3 = RandomSampling{prob = dist, seed = dim}."""
    _fields_ = [("kind", C.c_int32), ("lo", C.c_float * 3), ("hi", C.c_float * 3), ("dim", C.c_int32), ("dist", C.c_float),
                ("remove_inside", C.c_int32)]


def bounding_box(lo, hi, removeInside=True):
    f = InputFilter()
    f.kind = 1
    for i in range(3):
        f.lo[i], f.hi[i] = lo[i], hi[i]
    f.remove_inside = int(removeInside)
    return f


def distance_limit(dist, dim=-1, removeInside=False):
    f = InputFilter()
    f.kind, f.dim, f.dist, f.remove_inside = 2, dim, dist, int(removeInside)
    return f


def random_sampling(prob, seed=0):
    """RandomSamplingDataPointsFilter{prob}: keep each point with probability prob (reproducible counter-based generator)."""
    f = InputFilter()
    f.kind, f.dim, f.dist = 3, int(seed), float(prob)
    return f


class MapperConfig(C.Structure):
    _fields_ = [("icp", _abi.Config), ("update_condition", C.c_int32), ("update_value", C.c_float),
                ("sensor_max_range", C.c_float), ("min_dist_new_point", C.c_float), ("surface_normal_knn", C.c_int32),
                ("is_3d", C.c_int32), ("is_online", C.c_int32), ("is_mapping", C.c_int32),
                ("use_dynamic_points", C.c_int32), ("dynamic_points", _abi.DynamicParams),
                ("use_octree", C.c_int32), ("octree_max_size_by_node", C.c_float), ("octree_sampling_method", C.c_int32),
                ("use_cut_at_threshold", C.c_int32), ("cut_threshold", C.c_float),
                ("n_input_filters", C.c_int32), ("input_filters", InputFilter * 6),
                ("add_probability_dynamic", C.c_int32), ("probability_dynamic_value", C.c_float),
                ("reserve_points", C.c_int32), ("input_surface_normal_knn", C.c_int32), ("cell_spill", C.c_int32), ("reserved", C.c_int32 * 1),
                ("cell_folder", C.c_char * 128)]


class MapperStats(C.Structure):
    _fields_ = [("overlap", C.c_float), ("iterations", C.c_int32), ("map_updated", C.c_int32), ("n_window_updates", C.c_int32),
                ("n_local", C.c_int64), ("n_global", C.c_int64)]


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    _load_icp()  # libb200icp.so first (also resolved through the rpath)
    if not os.path.exists(SO_PATH):
        raise ImportError(f"{SO_PATH} is missing: build it with `make -C norlab_icp_mapper_b200/host`")
    L = C.CDLL(SO_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.b200mapper_create.argtypes = [C.POINTER(MapperConfig), i32, C.POINTER(vp)]
    L.b200mapper_create_from_yaml.argtypes = [C.c_char_p, i32, i32, i32, i32, i32, i32, C.POINTER(vp)]
    L.b200mapper_destroy.argtypes = [vp]
    L.b200mapper_destroy.restype = None
    L.b200mapper_last_error.argtypes = [vp]
    L.b200mapper_last_error.restype = C.c_char_p
    L.b200mapper_apply_input_filters.argtypes = [vp, vp, i32, C.POINTER(i64)]
    L.b200mapper_process_input.argtypes = [vp, vp, i32, i64, vp, C.c_double]
    L.b200mapper_process_raw_input.argtypes = [vp, vp, i32, i64, vp, C.c_double, C.POINTER(i64)]
    L.b200mapper_set_map_descriptors.argtypes = [vp, vp, i32, vp, vp, i64]
    L.b200mapper_get_map_prob.argtypes = [vp, vp, i64, C.POINTER(i64)]
    L.b200mapper_get_local_map.argtypes = [vp, vp, vp, i64, C.POINTER(i64)]
    L.b200mapper_map_update_in_flight.argtypes = [vp]
    L.b200mapper_wait_for_map_update.argtypes = [vp]
    L.b200mapper_get_pose.argtypes = [vp, vp]
    L.b200mapper_get_map.argtypes = [vp, vp, vp, i64, C.POINTER(i64)]
    L.b200mapper_get_new_local_map.argtypes = [vp, vp, vp, i64, C.POINTER(i64), C.POINTER(i32)]
    L.b200mapper_set_map.argtypes = [vp, vp, i32, vp, i64]
    L.b200mapper_get_is_mapping.argtypes = [vp]
    L.b200mapper_set_is_mapping.argtypes = [vp, i32]
    L.b200mapper_trajectory_size.argtypes = [vp]
    L.b200mapper_trajectory_size.restype = i64
    L.b200mapper_get_trajectory.argtypes = [vp, vp, vp, i64]
    L.b200mapper_get_stats.argtypes = [vp, C.POINTER(MapperStats)]
    L.b200mapper_get_window_updates.argtypes = [vp, vp, i32]
    L.b200mapper_vtk_save.argtypes = [C.c_char_p, vp, i32, i64, vp, vp, i32]
    L.b200mapper_vtk_load.argtypes = [C.c_char_p, i32, vp, vp, vp, i64, C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)]
    _lib = L
    return L


def yaml_summary(path, is3D=True):
    """What the C++ reader (host/YamlConfig.h) makes of a Mapper configuration file: dict of key -> value strings.  No GPU needed."""
    L = load()
    L.b200mapper_yaml_summary.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32]
    buf = C.create_string_buffer(16384)
    rc = L.b200mapper_yaml_summary(os.fspath(path).encode(), int(is3D), buf, len(buf))
    if rc != _abi.OK:
        raise B200ICPError(rc, L.b200mapper_last_error(None).decode())
    return dict(line.split("=", 1) for line in buf.value.decode().split("\n") if line)


class Mapper:
    CONDITIONS = {"distance": 0, "delay": 1, "overlap": 2}

    def __init__(self, icp_config, is3D=True, isOnline=False, isMapping=True, saveMapCellsOnHardDrive=False, *,
                 updateCondition=("distance", 1.0), sensorMaxRange=200.0, minDistNewPoint=0.15, surfaceNormalKnn=0,
                 dynamicPoints=None, octree=None, cutAtThreshold=None, inputFilters=(), addProbabilityDynamic=None, reservePoints=0,
                 inputSurfaceNormalKnn=0, cellSpill="device", cellFolder="", device=0):
        """dynamicPoints: _abi.DynamicParams or None; octree: (maxSizeByNode, samplingMethod) or None (then PointDistance);
        cutAtThreshold: threshold or None; inputFilters: InputFilter list; addProbabilityDynamic: value or None."""
        self._L = load()
        self.dim = 3 if is3D else 2
        self.n = self.dim + 1
        if isinstance(icp_config, (str, bytes, os.PathLike)):
            # the reference's signature: Mapper(configFilePath, is3D, isOnline, isMapping, saveMapCellsOnHardDrive) with the YAML
            # file of Mapper::loadYamlConfig (python/src/mapper.cpp:23); the keyword arguments below do not apply
            h = C.c_void_p()
            rc = self._L.b200mapper_create_from_yaml(os.fspath(icp_config).encode() if not isinstance(icp_config, bytes) else icp_config,
                                                     int(is3D), int(isOnline), int(isMapping), int(saveMapCellsOnHardDrive), device,
                                                     int(reservePoints), C.byref(h))
            if rc != _abi.OK:
                raise B200ICPError(rc, self._L.b200mapper_last_error(None).decode())
            self._h = h
            return
        cfg = MapperConfig()
        cfg.icp = icp_config
        cfg.update_condition = self.CONDITIONS.get(updateCondition[0], 99)
        cfg.update_value = updateCondition[1]
        cfg.sensor_max_range = sensorMaxRange
        cfg.min_dist_new_point = minDistNewPoint
        cfg.surface_normal_knn = surfaceNormalKnn
        cfg.is_3d, cfg.is_online, cfg.is_mapping = int(is3D), int(isOnline), int(isMapping)
        if dynamicPoints is not None:
            cfg.use_dynamic_points, cfg.dynamic_points = 1, dynamicPoints
        if octree is not None:
            cfg.use_octree, cfg.octree_max_size_by_node, cfg.octree_sampling_method = 1, octree[0], octree[1]
        if cutAtThreshold is not None:
            cfg.use_cut_at_threshold, cfg.cut_threshold = 1, cutAtThreshold
        cfg.n_input_filters = len(inputFilters)
        for i, f in enumerate(inputFilters):
            cfg.input_filters[i] = f
        if addProbabilityDynamic is not None:
            cfg.add_probability_dynamic, cfg.probability_dynamic_value = 1, addProbabilityDynamic
        cfg.reserve_points = int(reservePoints)
        cfg.input_surface_normal_knn = int(inputSurfaceNormalKnn)  # input: SurfaceNormalDataPointsFilter{knn} on the reading
        # cells the window leaves: stay in HBM ("device"), RAMCellManager ("ram"), HardDriveCellManager ("disk" / saveMapCellsOnHardDrive)
        cfg.cell_spill = 2 if saveMapCellsOnHardDrive else {"device": 0, "ram": 1, "disk": 2}[cellSpill]
        cfg.cell_folder = os.fspath(cellFolder).encode()
        self.dim = 3 if is3D else 2
        self.n = self.dim + 1
        h = C.c_void_p()
        rc = self._L.b200mapper_create(C.byref(cfg), device, C.byref(h))
        if rc != _abi.OK:
            raise B200ICPError(rc, self._L.b200mapper_last_error(None).decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.b200mapper_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != _abi.OK:
            raise B200ICPError(rc, self._L.b200mapper_last_error(self._h).decode())

    def applyInputFilters(self, cloud):
        cloud = np.ascontiguousarray(cloud, np.float32).copy()
        n = C.c_int64(len(cloud))
        self._check(self._L.b200mapper_apply_input_filters(self._h, cloud.ctypes.data, self.n, C.byref(n)))
        return cloud[:n.value]

    def processInput(self, cloud_sensor_frame, estimatedPose, timeStamp):
        cloud = np.ascontiguousarray(cloud_sensor_frame, np.float32)
        T = np.ascontiguousarray(np.asarray(estimatedPose, np.float32).T).ravel()
        self._check(self._L.b200mapper_process_input(self._h, cloud.ctypes.data, self.n, len(cloud), T.ctypes.data, float(timeStamp)))

    def processRawInput(self, raw_cloud_sensor_frame, estimatedPose, timeStamp):
        """applyInputFilters + processInput with ONE host-to-device copy (the filter chain runs on the device-resident scan).
        Returns the point count after the `input:` chain."""
        cloud = np.ascontiguousarray(raw_cloud_sensor_frame, np.float32)
        T = np.ascontiguousarray(np.asarray(estimatedPose, np.float32).T).ravel()
        n = C.c_int64()
        self._check(self._L.b200mapper_process_raw_input(self._h, cloud.ctypes.data, self.n, len(cloud), T.ctypes.data, float(timeStamp), C.byref(n)))
        return n.value

    def mapUpdateInFlight(self):
        """isOnline: the asynchronous map update dispatched by an earlier processInput is still running."""
        return bool(self._L.b200mapper_map_update_in_flight(self._h))

    def waitForMapUpdate(self):
        """isOnline: block until the map update in flight and the queued cell-window updates are done."""
        self._check(self._L.b200mapper_wait_for_map_update(self._h))

    def getPose(self):
        T = np.zeros(self.n * self.n, np.float32)
        self._check(self._L.b200mapper_get_pose(self._h, T.ctypes.data))
        return T.reshape(self.n, self.n).T.copy()

    def getMap(self):
        n = C.c_int64()
        self._check(self._L.b200mapper_get_map(self._h, None, None, 0, C.byref(n)))
        feat = np.zeros((n.value, self.n), np.float32)
        nrm = np.full((n.value, self.dim), np.nan, np.float32)
        if n.value:
            self._check(self._L.b200mapper_get_map(self._h, feat.ctypes.data, nrm.ctypes.data, n.value, C.byref(n)))
        return feat, (None if np.isnan(nrm).all() else nrm)

    def getLocalMap(self):
        """Map::getLocalPointCloud as arrays: (features, normals or None)."""
        st = self.stats()
        n = C.c_int64()
        feat = np.zeros((st.n_local, self.n), np.float32)
        nrm = np.full((st.n_local, self.dim), np.nan, np.float32)
        self._check(self._L.b200mapper_get_local_map(self._h, feat.ctypes.data, nrm.ctypes.data, st.n_local, C.byref(n)))
        return feat[:n.value], (None if np.isnan(nrm).all() else nrm[:n.value])

    def getNewLocalMap(self):
        """Mapper::getNewLocalMap (python/src/mapper.cpp:16): (True, features, normals or None) when the local map changed since
        the last call (the flag is consumed, as upstream), else (False, None, None)."""
        cap = int(self.stats().n_local) + 1024
        n, avail = C.c_int64(), C.c_int32()
        feat = np.zeros((cap, self.n), np.float32)
        nrm = np.full((cap, self.dim), np.nan, np.float32)
        self._check(self._L.b200mapper_get_new_local_map(self._h, feat.ctypes.data, nrm.ctypes.data, cap, C.byref(n), C.byref(avail)))
        if not avail.value:
            return False, None, None
        nrm = nrm[:n.value]
        return True, feat[:n.value].copy(), (None if np.isnan(nrm).all() else nrm.copy())

    def getMapProbabilityDynamic(self):
        """The probabilityDynamic descriptor of getMap(), or None when the map does not carry it."""
        n = C.c_int64()
        self._check(self._L.b200mapper_get_map_prob(self._h, None, 0, C.byref(n)))
        if n.value == 0:
            return None
        out = np.zeros(n.value, np.float32)
        self._check(self._L.b200mapper_get_map_prob(self._h, out.ctypes.data, n.value, C.byref(n)))
        return out

    def setMap(self, features, normals=None, probabilityDynamic=None):
        features = np.ascontiguousarray(features, np.float32)
        nptr = pptr = None
        if normals is not None:
            normals = np.ascontiguousarray(normals, np.float32)
            nptr = normals.ctypes.data
        if probabilityDynamic is not None:
            probabilityDynamic = np.ascontiguousarray(probabilityDynamic, np.float32)
            pptr = probabilityDynamic.ctypes.data
        self._check(self._L.b200mapper_set_map_descriptors(self._h, features.ctypes.data, self.n, nptr, pptr, len(features)))

    def getIsMapping(self):
        return bool(self._L.b200mapper_get_is_mapping(self._h))

    def setIsMapping(self, v):
        self._check(self._L.b200mapper_set_is_mapping(self._h, int(v)))

    def getTrajectory(self):
        n = self._L.b200mapper_trajectory_size(self._h)
        poses = np.zeros((n, self.n * self.n), np.float32)
        stamps = np.zeros(n, np.float64)
        if n:
            self._check(self._L.b200mapper_get_trajectory(self._h, poses.ctypes.data, stamps.ctypes.data, n))
        return poses.reshape(n, self.n, self.n).transpose(0, 2, 1).copy(), stamps

    def stats(self):
        s = MapperStats()
        self._check(self._L.b200mapper_get_stats(self._h, C.byref(s)))
        return s

    def windowUpdates(self):
        buf = np.zeros((64, 7), np.int32)
        n = self._L.b200mapper_get_window_updates(self._h, buf.ctypes.data, 64)
        return buf[:n].copy()
