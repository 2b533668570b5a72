"""Regenerates tests/golden/*.npz from the CPU oracle.

The reference ships no fixtures (tests/instantiation.cpp:4-19 pins nothing) and
its numerics live in libraries that cannot be built here, so these vectors are
ORACLE outputs on small seeded inputs: they pin the oracle against regressions
and give the GPU tests a /root/reference-free anchor.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402
from pgslam_b200 import synth  # noqa: E402
from tests import util  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    rd, rf, truth = synth.scan_pair(42, beams=16, az_steps=250)  # 4000 points
    ids1, d1 = ob.kdtree_knn(rf, rd, k=1)
    ids5, d5 = ob.kdtree_knn(rf, rf, k=5)
    oc = ob.Cloud(rf)
    ob.apply_filter(oc, "SurfaceNormalDataPointsFilter", knn=10, keepDensities=1, keepEigenValues=1)
    vox = ob.Cloud(rd)
    ob.apply_filter(vox, "VoxelGridDataPointsFilter", vSizeX=0.5, vSizeY=0.5, vSizeZ=0.5)
    st, w = ob.outlier_weights([{"TrimmedDistOutlierFilter": {"ratio": 0.85}}], d1)
    res = {}
    for name, cfg in (("c1", util.C1), ("c2", util.C2), ("c2cov", util.C2_COV)):
        r = ob.icp_run(cfg, ob.Cloud(rd), ob.Cloud(rf))
        assert r["status"] == 0
        res[name + "_T"] = r["T"]
        res[name + "_iterations"] = np.int32(r["iterations"])
        res[name + "_cov"] = r["cov"]
        res[name + "_residual"] = np.float64(r["residual"])
        res[name + "_overlap"] = np.float64(r["overlap"])
    np.savez_compressed(os.path.join(OUT, "pair4000.npz"), reading=rd, reference=rf, truth=truth,
                        knn1_ids=ids1, knn1_d2=d1, knn5_ids=ids5, knn5_d2=d5,
                        normals=oc.desc("normals"), densities=oc.desc("dens"), eigvalues=oc.desc("eigval"),
                        voxel_features=vox.features, trimmed_weights=w, **res)
    print("wrote", os.path.join(OUT, "pair4000.npz"))


if __name__ == "__main__":
    main()
