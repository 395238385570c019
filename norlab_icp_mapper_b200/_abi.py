"""ctypes mirror of include/b200icp.h (struct layouts and enums).  Pure declarations: shared by the
product binding (icp.py) and by the tests' oracle binding so both speak the same config."""
import ctypes as C

ABI_VERSION = 6
MAX_OUTLIER_FILTERS = 4

# b200icp_status
OK, ERR_INVALID_ARG, ERR_CUDA, ERR_NO_MAP, ERR_CONVERGENCE, ERR_BOUND, ERR_NAN, ERR_TRANSFORM, \
    ERR_INVALID_FIELD, ERR_NOT_IMPLEMENTED = range(10)
# b200icp_outlier_kind
OUTLIER_TRIMMED_DIST, OUTLIER_MAX_DIST, OUTLIER_MIN_DIST, OUTLIER_MEDIAN_DIST, OUTLIER_VAR_TRIMMED_DIST, OUTLIER_SURFACE_NORMAL, \
    OUTLIER_ROBUST = 1, 2, 3, 4, 5, 6, 7
# b200icp_robust_fct / b200icp_robust_scale / b200icp_robust_dist (libpointmatcher parameter values)
ROBUST_FCTS = {"cauchy": 0, "welsch": 1, "sc": 2, "gm": 3, "tukey": 4, "huber": 5, "L1": 6, "student": 7}
ROBUST_SCALES = {"none": 0, "mad": 1, "berg": 2, "std": 3}
ROBUST_DISTS = {"point2point": 0, "point2plane": 1}
# b200icp_minimizer_kind
MIN_POINT_TO_PLANE, MIN_POINT_TO_POINT, MIN_IDENTITY = 0, 1, 2


class Config(C.Structure):
    _fields_ = [
        ("dim", C.c_int32),
        ("knn", C.c_int32),
        ("max_dist", C.c_float),
        ("epsilon", C.c_float),
        ("n_outlier", C.c_int32),
        ("outlier_kind", C.c_int32 * MAX_OUTLIER_FILTERS),
        ("outlier_param", C.c_float * MAX_OUTLIER_FILTERS),
        ("outlier_param2", C.c_float * MAX_OUTLIER_FILTERS),
        ("outlier_param3", C.c_float * MAX_OUTLIER_FILTERS),
        ("minimizer", C.c_int32),
        ("max_iteration_count", C.c_int32),
        ("use_differential", C.c_int32),
        ("min_diff_rot_err", C.c_float),
        ("min_diff_trans_err", C.c_float),
        ("smooth_length", C.c_int32),
        ("use_bound", C.c_int32),
        ("max_rotation_norm", C.c_float),
        ("max_translation_norm", C.c_float),
        ("sort_reading", C.c_int32),
        ("use_graph", C.c_int32),
        ("nn_variant", C.c_int32),
        ("outlier_mode", C.c_int32 * MAX_OUTLIER_FILTERS),
        ("checker_order", C.c_int32),
        ("minimizer_flags", C.c_int32),
        ("conventions", C.c_int32),
        ("reserved2", C.c_int32 * 5),
    ]


class Result(C.Structure):
    _fields_ = [
        ("overlap", C.c_float),
        ("point_used_ratio", C.c_float),
        ("iterations", C.c_int32),
        ("max_iter_reached", C.c_int32),
        ("pairs_last_iter", C.c_int64),
    ]


class Timing(C.Structure):
    _fields_ = [
        ("total_ms", C.c_float),
        ("nn_ms_sum", C.c_float),
        ("nn_launches", C.c_int32),
        ("kernel_launches", C.c_int32),
        ("setmap_ms", C.c_float),
        ("select_ms_sum", C.c_float),
        ("acc_ms_sum", C.c_float),
        ("loop_iterations", C.c_int32),
        ("loop_search_ms_sum", C.c_float),
        ("loop_total_ms", C.c_float),
        ("loop_fast_iterations", C.c_int32),
        ("loop_searched_queries", C.c_int32),
        ("loop_two_barrier_iterations", C.c_int32),
        ("loop_kernel_ms", C.c_float),
    ]


class Pair(C.Structure):
    """b200icp_pair: one scan <-> submap alignment of b200icp_register_batch (host pointers)."""
    _fields_ = [("map_features", C.c_void_p), ("map_normals", C.c_void_p), ("n_map", C.c_int64),
                ("reading", C.c_void_p), ("n_reading", C.c_int64), ("T_init", C.c_void_p)]


class PairResult(C.Structure):
    _fields_ = [("T", C.c_float * 16), ("result", Result), ("status", C.c_int32), ("setmap_ms", C.c_float),
                ("register_ms", C.c_float)]


def make_config(dim=3, knn=1, max_dist=float("inf"), epsilon=0.0, outliers=(("trimmed", 0.85),),
                minimizer="point_to_plane", max_iteration_count=40, differential=None, bound=None,
                sort_reading=1, use_graph=1, nn_variant=0, checker_order=0, force2D=False, force4DOF=False, conventions=0):
    """Build a Config from the names used in the reference's `icp:` YAML node
    (docs/MapperConfiguration.md:172-189)."""
    kinds = {"trimmed": OUTLIER_TRIMMED_DIST, "max_dist": OUTLIER_MAX_DIST,
             "min_dist": OUTLIER_MIN_DIST, "median": OUTLIER_MEDIAN_DIST, "var_trimmed": OUTLIER_VAR_TRIMMED_DIST,
             "surface_normal": OUTLIER_SURFACE_NORMAL, "robust": OUTLIER_ROBUST}
    mins = {"point_to_plane": MIN_POINT_TO_PLANE, "point_to_point": MIN_POINT_TO_POINT,
            "identity": MIN_IDENTITY}
    c = Config()
    c.dim, c.knn, c.max_dist, c.epsilon = dim, knn, max_dist, epsilon
    if len(outliers) > MAX_OUTLIER_FILTERS:
        raise ValueError("too many outlier filters")
    c.n_outlier = len(outliers)
    for i, (name, *params) in enumerate(outliers):  # ("var_trimmed", minRatio, maxRatio, lambda); one parameter otherwise
        c.outlier_kind[i] = kinds[name]
        if name == "robust":  # ("robust", {robustFct, tuning, scaleEstimator, nbIterationForScale, distanceType, approximation})
            rp = dict(robustFct="cauchy", tuning=1.0, scaleEstimator="mad", nbIterationForScale=0, distanceType="point2point",
                      approximation=float("inf"))  # libpointmatcher's defaults
            rp.update(params[0] if params else {})
            c.outlier_param[i], c.outlier_param2[i], c.outlier_param3[i] = rp["tuning"], rp["approximation"], 0.0
            c.outlier_mode[i] = (ROBUST_FCTS[rp["robustFct"]] | (ROBUST_SCALES[rp["scaleEstimator"]] << 8) |
                                 (ROBUST_DISTS[rp["distanceType"]] << 12) | (int(rp["nbIterationForScale"]) << 16))
            continue
        c.outlier_param[i] = params[0]
        c.outlier_param2[i] = params[1] if len(params) > 1 else 0.0
        c.outlier_param3[i] = params[2] if len(params) > 2 else 0.0
    c.minimizer = mins[minimizer]
    c.max_iteration_count = max_iteration_count
    if differential is not None:
        c.use_differential = 1
        c.min_diff_rot_err, c.min_diff_trans_err, c.smooth_length = differential
    else:
        c.use_differential, c.min_diff_rot_err, c.min_diff_trans_err, c.smooth_length = 0, 1e-3, 1e-3, 3
    if bound is not None:
        c.use_bound = 1
        c.max_rotation_norm, c.max_translation_norm = bound
    else:
        c.use_bound, c.max_rotation_norm, c.max_translation_norm = 0, 1.0, 1.0
    c.sort_reading, c.use_graph, c.nn_variant = sort_reading, use_graph, nn_variant
    c.checker_order = checker_order
    c.minimizer_flags = (1 if force2D else 0) | (2 if force4DOF else 0)
    c.conventions = conventions
    return c


class DynamicParams(C.Structure):
    """DynamicPointsMapperModule parameters with the reference's defaults (DynamicPointsMapperModule.h:33-44)."""
    _fields_ = [("threshold_dynamic", C.c_float), ("alpha", C.c_float), ("beta", C.c_float), ("beam_half_angle", C.c_float),
                ("epsilon_a", C.c_float), ("epsilon_d", C.c_float), ("sensor_max_range", C.c_float)]

    def __init__(self, thresholdDynamic=0.6, alpha=0.8, beta=0.99, beamHalfAngle=0.01, epsilonA=0.01, epsilonD=0.01, sensorMaxRange=200.0):
        super().__init__(thresholdDynamic, alpha, beta, beamHalfAngle, epsilonA, epsilonD, sensorMaxRange)
