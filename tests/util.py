"""Shared helpers for the parity tests: the BASELINE.json configs as
libpointmatcher YAML and product/oracle cloud
construction from the same arrays."""
from __future__ import annotations

import numpy as np
import yaml

CHECKERS = [{"CounterTransformationChecker": {"maxIterationCount": 40}},
            {"DifferentialTransformationChecker": {"minDiffRotErr": 0.001, "minDiffTransErr": 0.001, "smoothLength": 3}}]

# C1: point-to-point, KDTreeMatcher k=1, trimmed 0.85, no filters
C1 = dict(matcher={"KDTreeMatcher": {"knn": 1, "epsilon": 0}},
          outlierFilters=[{"TrimmedDistOutlierFilter": {"ratio": 0.85}}],
          errorMinimizer="PointToPointErrorMinimizer",
          transformationCheckers=CHECKERS, inspector="NullInspector", logger="NullLogger")

# C2: point-to-plane with the SurfaceNormal reference filter
C2 = dict(referenceDataPointsFilters=[{"SurfaceNormalDataPointsFilter": {"knn": 10, "epsilon": 0, "keepNormals": 1}}],
          matcher={"KDTreeMatcher": {"knn": 1, "epsilon": 0}},
          outlierFilters=[{"TrimmedDistOutlierFilter": {"ratio": 0.85}}],
          errorMinimizer="PointToPlaneErrorMinimizer",
          transformationCheckers=CHECKERS, inspector="NullInspector", logger="NullLogger")

C2_COV = dict(C2, errorMinimizer={"PointToPlaneWithCovErrorMinimizer": {"sensorStdDev": 0.01}})

# C5: voxel-subsampled reading, trimmed 0.75
C5 = dict(readingDataPointsFilters=[{"VoxelGridDataPointsFilter": {"vSizeX": 0.2, "vSizeY": 0.2, "vSizeZ": 0.2,
                                                                 "useCentroid": 1}}],
          referenceDataPointsFilters=[{"SurfaceNormalDataPointsFilter": {"knn": 10}}],
          matcher={"KDTreeMatcher": {"knn": 1}},
          outlierFilters=[{"TrimmedDistOutlierFilter": {"ratio": 0.75}}],
          errorMinimizer="PointToPlaneErrorMinimizer",
          transformationCheckers=CHECKERS)

INPUT_FILTERS = [{"SurfaceNormalDataPointsFilter": {"knn": 10, "epsilon": 0, "keepNormals": 1}},
                 "ObservationDirectionDataPointsFilter",
                 {"OrientNormalsDataPointsFilter": {"towardCenter": 1}},
                 {"SimpleSensorNoiseDataPointsFilter": {"sensorType": 0, "gain": 1}}]


def to_yaml(cfg) -> str:
    return yaml.safe_dump(cfg, default_flow_style=False, sort_keys=False)


def rot_angle(Ra, Rb) -> float:
    # small-angle safe (arccos of the trace is ill-conditioned near 0): the angle
    # from the antisymmetric part, sin(theta) = |vee(R - R^T)| / 2
    R = Ra[:3, :3].T @ Rb[:3, :3]
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2.0
    s = float(np.linalg.norm(v))
    c = (np.trace(R) - 1.0) / 2.0
    return float(np.arctan2(s, c))


def assert_pose_close(Ta, Tb, tol_t=1e-5, tol_r=1e-5):
    dt = float(np.abs(Ta[:3, 3] - Tb[:3, 3]).max())
    dr = rot_angle(Ta, Tb)
    assert dt <= tol_t and dr <= tol_r, f"pose mismatch: dt={dt:.3e} m, dr={dr:.3e} rad"
