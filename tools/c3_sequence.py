"""BASELINE.json configs[2] in the small: sequential scan-to-map odometry with ICPSequence
(Localizer.hpp:119-126): a 3-keyframe local map (3 x 120k points, normals from the input
filters), then scans registered one after the other, each seeded by the previous result.
Inherently sequential (one GPU, latency-bound).  Prints scans/s and the per-scan latency."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pgslam_b200 import pm, synth  # noqa: E402
from tests import util  # noqa: E402

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 40
ctx = pm.Context(0)
scene = synth.make_scene(5)
poses = synth.trajectory(3 + n_scans, step=0.25, turn_deg=1.0)
scans = [synth.velodyne_scan(5, i, poses[i], beams=64, az_steps=1875, scene=scene) for i in range(3 + n_scans)]
filt = pm.DataPointsFilters(util.to_yaml(util.INPUT_FILTERS[:3]), ctx=ctx)
kfs = []
for k in range(3):
    dp = pm.DataPoints(scans[k], ctx=ctx)
    filt.apply(dp)
    kfs.append(dp)
T_ref = [np.linalg.inv(poses[2]) @ poses[k] for k in range(3)]
local_map = pm.assemble_local_map([kfs[2], kfs[1], kfs[0]], [np.eye(4), T_ref[1], T_ref[0]])
cfg = dict(util.C2, referenceDataPointsFilters=[])
seq = pm.ICPSequence(ctx)
seq.loadFromYaml(util.to_yaml(cfg))
t0 = time.perf_counter()
seq.setMap(local_map)
ctx.synchronize()
t_map = time.perf_counter() - t0
T_prev = np.linalg.inv(poses[2]) @ poses[3]
its, lat = [], []
for rep in range(2):  # first pass warms the pools
    T_prev = np.linalg.inv(poses[2]) @ poses[3]
    its, lat = [], []
    t_all = time.perf_counter()
    for i in range(3, 3 + n_scans):
        t1 = time.perf_counter()
        dp = pm.DataPoints(scans[i], ctx=ctx)          # H2D of the scan
        filt.apply(dp)                                   # input filters (normals etc.), Localizer.hpp:103
        T = seq(dp, T_prev)                              # Localizer.hpp:126
        lat.append(time.perf_counter() - t1)
        its.append(seq.last["iterations"])
        step = np.linalg.inv(poses[i]) @ poses[min(i + 1, len(poses) - 1)]
        T_prev = T @ step                                # constant-velocity style seed
    total = time.perf_counter() - t_all
err = np.abs(T[:3, 3] - (np.linalg.inv(poses[2]) @ poses[2 + n_scans])[:3, 3]).max()
print(f"map: {local_map.getNbPoints()} pts, setMap {1e3 * t_map:.1f} ms (first call); "
      f"{n_scans} scans of {scans[0].shape[1]} pts: {n_scans / total:.1f} scans/s, "
      f"median latency {1e3 * np.median(lat):.2f} ms (input filters + ICP, host to host), "
      f"iterations mean {np.mean(its):.1f}, final position error {err:.3f} m")
