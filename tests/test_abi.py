"""The C-ABI shared library: it loads, exports every symbol include/b200icp.h declares, the ctypes
struct mirrors match the C layout, and there is no CPU fallback.  No GPU needed."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from norlab_icp_mapper_b200 import _abi, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200icp.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.SO_PATH):
        subprocess.run(["make", "-s", "-j4", "-C", os.path.join(ROOT, "norlab_icp_mapper_b200", "csrc")], check=True)
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    text = open(HEADER).read()
    declared = sorted(set(re.findall(r"\b(b200icp_[a-z0-9_]+)\s*\(", text)))
    assert len(declared) >= 20
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_lib.SYMBOLS) == declared


def test_struct_layout_matches_c(lib):
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "b200icp.h"
int main(void) {
    printf("%zu %zu %zu %zu %zu %zu %zu %zu %d %d\n", sizeof(b200icp_config), offsetof(b200icp_config, outlier_param),
           offsetof(b200icp_config, minimizer), offsetof(b200icp_config, sort_reading), sizeof(b200icp_result),
           offsetof(b200icp_result, pairs_last_iter), sizeof(b200icp_timing), offsetof(b200icp_config, outlier_mode),
           (int)B200ICP_OUTLIER_ROBUST, (int)B200ICP_ROBUST_MODE(B200ICP_ROBUST_HUBER, B200ICP_SCALE_BERG, B200ICP_DIST_POINT2PLANE, 3));
    return 0;
}'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "t")
        subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    robust = _abi.make_config(outliers=(("robust", dict(robustFct="huber", scaleEstimator="berg", distanceType="point2plane", nbIterationForScale=3)),))
    want = [C.sizeof(_abi.Config), _abi.Config.outlier_param.offset, _abi.Config.minimizer.offset, _abi.Config.sort_reading.offset,
            C.sizeof(_abi.Result), _abi.Result.pairs_last_iter.offset, C.sizeof(_abi.Timing), _abi.Config.outlier_mode.offset,
            _abi.OUTLIER_ROBUST, robust.outlier_mode[0]]
    assert got == want


def test_config_default_is_lpm_set_default(lib):
    cfg = _abi.Config()
    lib.b200icp_config_default(C.byref(cfg), 3)
    assert (cfg.dim, cfg.knn, cfg.n_outlier, cfg.outlier_kind[0], cfg.minimizer) == (3, 1, 1, _abi.OUTLIER_TRIMMED_DIST, _abi.MIN_POINT_TO_PLANE)
    assert cfg.outlier_param[0] == pytest.approx(0.85) and cfg.max_iteration_count == 40
    assert cfg.use_differential == 1 and cfg.smooth_length == 3 and cfg.max_dist == float("inf")


def test_invalid_configs_are_rejected_before_touching_cuda(lib):
    for kw, code in ((dict(dim=4), _abi.ERR_INVALID_ARG), (dict(knn=0), _abi.ERR_INVALID_ARG), (dict(knn=33), _abi.ERR_INVALID_ARG),
                     (dict(outliers=(("trimmed", 1.5),)), _abi.ERR_INVALID_ARG),
                     (dict(outliers=(("trimmed", 0.8), ("median", 3.0))), _abi.ERR_NOT_IMPLEMENTED)):
        cfg = _abi.make_config(**kw)
        h = C.c_void_p()
        assert lib.b200icp_create(C.byref(cfg), 0, C.byref(h)) == code, kw
        assert not h.value and lib.b200icp_last_error(None)


def test_no_cpu_fallback(lib):
    """Without a GPU, creation fails loudly; with one, this test is trivially satisfied."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = _abi.make_config()
    h = C.c_void_p()
    assert lib.b200icp_create(C.byref(cfg), 0, C.byref(h)) == _abi.ERR_CUDA
    assert b"no CPU fallback" in lib.b200icp_last_error(None)
    from norlab_icp_mapper_b200.icp import ICP, B200ICPError
    with pytest.raises(B200ICPError):
        ICP(cfg)
    assert lib.b200icp_map_size(None) == 0


def test_mapper_library_exports_its_header(lib):
    """libb200mapper.so (host-side C++ mirror of Mapper/Map/MapperModule) exports include/b200mapper.h."""
    from norlab_icp_mapper_b200 import mapper
    if not os.path.exists(mapper.SO_PATH):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "norlab_icp_mapper_b200", "host")], check=True)
    M = mapper.load()
    text = open(os.path.join(ROOT, "include", "b200mapper.h")).read()
    declared = sorted(set(re.findall(r"\b(b200mapper_[a-z0-9_]+)\s*\(", text)))
    assert declared == sorted(mapper.SYMBOLS)
    assert not [s for s in declared if not hasattr(M, s)]
    src = r'''
#include <stdio.h>
#include "b200mapper.h"
int main(void) { printf("%zu %zu\n", sizeof(b200mapper_config), sizeof(b200mapper_stats)); return 0; }'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "t")
        subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [C.sizeof(mapper.MapperConfig), C.sizeof(mapper.MapperStats)]
    import torch
    if not torch.cuda.is_available():  # no fallback: constructing a Mapper without a GPU fails loudly
        from norlab_icp_mapper_b200._lib import B200ICPError
        with pytest.raises(B200ICPError):
            mapper.Mapper(_abi.make_config(), True, False, True, False)
