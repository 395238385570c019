"""Mapper::loadYamlConfig without yaml-cpp (host/YamlConfig.h): the reference's configuration grammar (the layout of
examples/config.yaml and of the snippets in docs/MapperConfiguration.md, restated here), checked against PyYAML's reading of the
same text; unknown names fail the way libpointmatcher's registrar does.  The CPU tests need no GPU; the GPU test builds a Mapper
from the file like the reference's constructor."""
import numpy as np
import pytest
import yaml

EXAMPLE = """\
input:
  - BoundingBoxDataPointsFilter:
      xMin: -1.5
      xMax: 0.5
      yMin: -1
      yMax: 1
      zMin: -1
      zMax: 0.5
      removeInside: 1

  - BoundingBoxDataPointsFilter:
      xMin: -6
      xMax: -1.5
      yMin: -2.5
      yMax: 2.5
      zMin: -1
      zMax: 1
      removeInside: 1

  - AddDescriptorDataPointsFilter:
      descriptorName: probabilityDynamic
      descriptorDimension: 1
      descriptorValues: [0.6] # the initial probability of each point being dynamic

post:
    - SurfaceNormalDataPointsFilter:
        knn: 10

    - CutAtDescriptorThresholdDataPointsFilter:
        descName: probabilityDynamic
        useLargerThan: 1
        threshold: 0.65

mapper:
  updateCondition:
    type: delay
    value: 0.05

  mapperModule:
    - DynamicPointsMapperModule:
        thresholdDynamic: 0.9
        alpha: 0.8
        beta: 0.99
        beamHalfAngle: 0.01
        epsilonA: 0.01
        epsilonD: 0.01

    - OctreeMapperModule:
        buildParallel: 1
        maxSizeByNode: 0.15
        samplingMethod: 1

  sensorMaxRange: 200

icp:
  matcher:
    KDTreeMatcher:
      knn: 6
      maxDist: 2.0
      epsilon: 1

  errorMinimizer:
    IdentityErrorMinimizer:

  transformationCheckers:
    - CounterTransformationChecker:
        maxIterationCount: 10

  inspector: NullInspector
"""

NORLAB_STYLE = """\
icp:
  readingDataPointsFilters:
    - IdentityDataPointsFilter
  matcher:
    KDTreeMatcher:
      knn: 3
      maxDist: 1.5
  outlierFilters:
    - TrimmedDistOutlierFilter:
        ratio: 0.9
    - SurfaceNormalOutlierFilter:
        maxAngle: 0.42
    - RobustOutlierFilter:
        robustFct: "cauchy"
        tuning: 1.5
        scaleEstimator: mad
        distanceType: point2plane
        nbIterationForScale: 2
  errorMinimizer:
    PointToPlaneErrorMinimizer:
      force4DOF: 1
  transformationCheckers:
    - DifferentialTransformationChecker:
        minDiffRotErr: 0.002
        minDiffTransErr: 0.01
        smoothLength: 4
    - CounterTransformationChecker:
        maxIterationCount: 40
    - BoundTransformationChecker:
        maxRotationNorm: 0.8
        maxTranslationNorm: 15
mapper:
  updateCondition: {type: distance, value: 2.5}
"""


def _write(tmp_path, text):
    p = tmp_path / "config.yaml"
    p.write_text(text)
    return p


def test_example_config_is_read_like_pyyaml_reads_it(tmp_path):
    from norlab_icp_mapper_b200.mapper import yaml_summary
    s = yaml_summary(_write(tmp_path, EXAMPLE))
    ref = yaml.safe_load(EXAMPLE)
    km = ref["icp"]["matcher"]["KDTreeMatcher"]
    assert (int(s["icp.knn"]), float(s["icp.maxDist"]), float(s["icp.epsilon"])) == (km["knn"], km["maxDist"], km["epsilon"])
    assert s["icp.minimizer"] == "2" and s["icp.nOutlier"] == "0"  # IdentityErrorMinimizer, no outlierFilters
    assert s["icp.counter"] == str(ref["icp"]["transformationCheckers"][0]["CounterTransformationChecker"]["maxIterationCount"])
    assert s["icp.differential"].startswith("0,") and s["icp.bound"].startswith("0,")
    boxes = [e["BoundingBoxDataPointsFilter"] for e in ref["input"] if "BoundingBoxDataPointsFilter" in e]
    assert s["input.n"] == str(len(boxes)) == "2"
    for i, b in enumerate(boxes):
        got = [float(v) for v in s[f"input{i}"].split(",")]
        assert got[0] == 1 and got[1:7] == [b["xMin"], b["xMax"], b["yMin"], b["yMax"], b["zMin"], b["zMax"]] and got[9] == b["removeInside"]
    assert s["input.addProbabilityDynamic"] == "1,0.6" and s["input.surfaceNormalKnn"] == "0"
    assert s["post.surfaceNormalKnn"] == "10" and s["post.cut"] == "1,1,0.65"
    assert s["mapper.updateCondition"] == "delay,0.05" and float(s["mapper.sensorMaxRange"]) == 200
    assert s["mapper.nModules"] == "2"
    mods = ref["mapper"]["mapperModule"]
    for i, m in enumerate(mods):
        (name, params), = m.items()
        fields = s[f"module{i}"].split(";")
        assert fields[0] == name
        assert {k: float(v) for k, v in (f.split("=") for f in fields[1:])} == {k: float(v) for k, v in params.items()}


def test_icp_chain_in_norlab_style(tmp_path):
    from norlab_icp_mapper_b200 import _abi
    from norlab_icp_mapper_b200.mapper import yaml_summary
    s = yaml_summary(_write(tmp_path, NORLAB_STYLE))
    assert s["icp.knn"] == "3" and float(s["icp.maxDist"]) == 1.5 and s["icp.minimizer"] == "0" and s["icp.minimizerFlags"] == "2"
    assert s["icp.nOutlier"] == "3"
    assert s["icp.outlier0"].split(",")[:2] == ["1", "0.9"] and s["icp.outlier1"].split(",")[:2] == ["6", "0.42"]
    kind, tuning, approx, _, mode = s["icp.outlier2"].split(",")
    assert kind == "7" and float(tuning) == 1.5 and approx == "inf"
    assert int(mode) == (_abi.ROBUST_FCTS["cauchy"] | (_abi.ROBUST_SCALES["mad"] << 8) | (_abi.ROBUST_DISTS["point2plane"] << 12) | (2 << 16))
    assert s["icp.counter"] == "40" and s["icp.differential"] == "1,0.002,0.01,4" and s["icp.bound"] == "1,0.8,15"
    assert s["icp.checkerOrder"] == "1"  # the Differential checker is listed before the Counter, the Bound checker after it
    assert s["mapper.updateCondition"] == "distance,2.5" and s["mapper.nModules"] == "0"  # -> setDefaultMapperModule


def test_defaults_without_nodes(tmp_path):
    from norlab_icp_mapper_b200.mapper import yaml_summary
    s = yaml_summary(_write(tmp_path, "mapper:\n  sensorMaxRange: 80\n"))
    # icp.setDefault(): knn 1, TrimmedDist 0.85, PointToPlane, Counter 40 + Differential
    assert s["icp.knn"] == "1" and s["icp.nOutlier"] == "1" and s["icp.outlier0"].startswith("1,0.85") and s["icp.minimizer"] == "0"
    assert s["icp.counter"] == "40" and s["icp.differential"].startswith("1,")
    assert s["mapper.updateCondition"] == "distance,1" and float(s["mapper.sensorMaxRange"]) == 80


@pytest.mark.parametrize("text, message", [
    ("input:\n  - VoxelGridDataPointsFilter:\n      vSizeX: 0.1\n", "unknown DataPointsFilter VoxelGridDataPointsFilter"),
    ("icp:\n  outlierFilters:\n    - GenericDescriptorOutlierFilter:\n        source: reference\n", "unknown OutlierFilter"),
    ("icp:\n  matcher:\n    KDTreeMatcher:\n      knnn: 3\n", "Parameter knnn for module KDTreeMatcher was set but is not used"),
    ("icp:\n  errorMinimizer:\n    PointToPlaneWithCovErrorMinimizer:\n", "unknown ErrorMinimizer"),
    ("post:\n  - CutAtDescriptorThresholdDataPointsFilter:\n      descName: intensity\n      threshold: 3\n", "only descName probabilityDynamic"),
    ("input:\n\t- IdentityDataPointsFilter\n", "tabs are not allowed"),
    ("surprise: 1\n", "unknown top-level key surprise"),
])
def test_unknown_names_fail_loudly(tmp_path, text, message):
    from norlab_icp_mapper_b200._lib import B200ICPError
    from norlab_icp_mapper_b200.mapper import yaml_summary
    with pytest.raises(B200ICPError) as e:
        yaml_summary(_write(tmp_path, text))
    assert e.value.status == 1 and message in str(e.value)


def test_missing_file(tmp_path):
    from norlab_icp_mapper_b200._lib import B200ICPError
    from norlab_icp_mapper_b200.mapper import yaml_summary
    with pytest.raises(B200ICPError) as e:
        yaml_summary(tmp_path / "nope.yaml")
    assert "Cannot open config file" in str(e.value)


@pytest.mark.gpu
def test_mapper_from_the_example_yaml(tmp_path):
    """Mapper(configFilePath, is3D, isOnline, isMapping, saveMapCellsOnHardDrive) -- the reference's constructor -- on the example
    configuration; a short drive gives the same map as the same configuration spelled out through the keyword arguments."""
    from norlab_icp_mapper_b200 import _abi, synth
    from norlab_icp_mapper_b200._abi import make_config
    from norlab_icp_mapper_b200.mapper import Mapper, bounding_box
    text = EXAMPLE.replace("samplingMethod: 1", "samplingMethod: 0")  # (the random sampler is seeded per call count: use `first`)
    a = Mapper(str(_write(tmp_path, text)), True, False, True, False)
    cfg = make_config(dim=3, knn=6, max_dist=2.0, epsilon=1.0, outliers=(), minimizer="identity", max_iteration_count=10)
    dyn = _abi.DynamicParams(thresholdDynamic=0.9, alpha=0.8, beta=0.99, beamHalfAngle=0.01, epsilonA=0.01, epsilonD=0.01)
    boxes = [((-1.5, -1, -1), (0.5, 1, 0.5)), ((-6, -2.5, -1), (-1.5, 2.5, 1))]
    b = Mapper(cfg, True, False, True, False, updateCondition=("delay", 0.05), sensorMaxRange=200.0, surfaceNormalKnn=10, dynamicPoints=dyn,
               octree=(0.15, 0), cutAtThreshold=0.65, inputFilters=[bounding_box(lo, hi, True) for lo, hi in boxes], addProbabilityDynamic=0.6)
    world = synth.World3D(seed=5, size=(60.0, 60.0), n_boxes=8)
    for i in range(3):
        T = synth.make_T((0.8 * i, 0.2 * i, 1.5), (0, 0, 3.0 * i))
        S, _ = world.sample(15_000, np.random.default_rng(50 + i), noise=0.01, center=T[:3, 3], radius=40.0)
        scan = synth.homog(synth.apply_T(np.linalg.inv(T), S))
        for m in (a, b):
            m.processInput(m.applyInputFilters(scan), T.astype(np.float32), 0.1 * i)
    fa, na = a.getMap()
    fb, nb = b.getMap()
    assert len(fa) > 5000 and np.array_equal(fa, fb) and np.array_equal(na, nb)
    assert np.array_equal(a.getMapProbabilityDynamic(), b.getMapProbabilityDynamic())
    a.close()
    b.close()
