"""The C++ adapter (include/pgslam_b200/pm_adapter.hpp): pgslam's own call
sequences (tests/cpp/callsites.cpp) compile against it on the CPU box and, on the
GPU box, give the oracle's answers."""
import os
import subprocess

import numpy as np
import pytest

from pgslam_b200 import build, pm, synth
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_callsites(out_dir=None):
    """compile tests/cpp/callsites.cpp against the adapter (also used by bench.py's drop-in leg)"""
    import pathlib
    import tempfile
    d = pathlib.Path(out_dir or tempfile.mkdtemp(prefix="pgs_callsites_"))
    return _compile(d)


def _compile(tmp_path):
    build.build()
    exe = str(tmp_path / "callsites")
    libdir = os.path.dirname(pm.LIB_PATH)
    subprocess.run(["/usr/bin/g++", "-std=c++14", "-Wall", "-Wextra", "-Werror", "-O1", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "callsites.cpp"), "-o", exe, "-L" + libdir, "-lpgslam_b200",
                    "-Wl,-rpath," + libdir, "-pthread"], check=True, capture_output=True, text=True)
    return exe


def test_pgslam_call_sites_compile_against_the_adapter(tmp_path):
    assert os.path.exists(_compile(tmp_path))


def _write_cloud(path, feat):
    with open(path, "wb") as f:
        f.write(np.int32(feat.shape[1]).tobytes())
        f.write(np.ascontiguousarray(feat.T, dtype=np.float32).tobytes())


@pytest.mark.gpu
def test_pgslam_call_sites_run_and_match_the_oracle(tmp_path):
    from oracle import binding as ob
    exe = _compile(tmp_path)
    rd, rf, _ = synth.scan_pair(21, beams=16, az_steps=500)
    _write_cloud(tmp_path / "rd.bin", rd)
    _write_cloud(tmp_path / "rf.bin", rf)
    cfg = dict(util.C2_COV, referenceDataPointsFilters=[])  # the local-map clouds already carry normals
    (tmp_path / "icp.yaml").write_text(util.to_yaml(cfg))
    (tmp_path / "filters.yaml").write_text(util.to_yaml(util.INPUT_FILTERS))
    out = subprocess.run([exe, str(tmp_path / "rd.bin"), str(tmp_path / "rf.bin"), str(tmp_path / "icp.yaml"),
                          str(tmp_path / "filters.yaml")], check=True, capture_output=True, text=True, timeout=120).stdout
    kv = dict(line.split("=", 1) for line in out.strip().splitlines())
    # the same sequence on the oracle
    ord_, orf = ob.Cloud(rd), ob.Cloud(rf)
    for it in util.INPUT_FILTERS:
        (name, p), = ob._modlist([it])
        ob.apply_filter(ord_, name, **p)
        ob.apply_filter(orf, name, **p)
    want = ob.icp_run(cfg, ord_, orf)
    T = np.array([float(x) for x in kv["T"].split(",")]).reshape(4, 4).T
    assert int(kv["descriptors"]) == 3
    assert int(kv["iterations"]) == want["iterations"]
    assert int(kv["max_iter_reached"]) == int(want["max_iter_reached"])
    util.assert_pose_close(T, want["T"], 2e-6, 2e-6)  # T crosses the adapter as float
    assert float(kv["overlap"]) == pytest.approx(want["overlap"], abs=1e-6)
    assert float(kv["cov00"]) == pytest.approx(want["cov"][0, 0], rel=1e-4)
    Tf = T.astype(np.float32).astype(np.float64)
    st, res = ob.probe_residual(cfg, ord_, orf, Tf)
    assert st == 0 and float(kv["residual"]) == pytest.approx(res, rel=1e-5)
    st, ratio = ob.probe_overlap(cfg, ord_, orf, Tf)
    assert st == 0 and float(kv["weighted_ratio"]) == pytest.approx(ratio, abs=1e-6)
    assert int(kv["local_map_points"]) == rd.shape[1] + rf.shape[1]
    assert int(kv["has_map_before"]) == 0 and int(kv["seq_iterations"]) >= 1
    assert int(kv["no_map_identity"]) == 1
    # plugin registrar, YAML-free
    assert kv["reg_filter_class"] == "RandomSamplingDataPointsFilter" and kv["reg_filter_seed_default"] == "0"
    assert int(kv["reg_filter_nparams"]) == 2 and 0.4 * rf.shape[1] < int(kv["reg_filter_points"]) < 0.6 * rf.shape[1]
    ids, _ = ob.kdtree_knn(rf, rd, k=2)
    assert int(kv["reg_matcher_k"]) == 2 and int(kv["reg_matcher_id0"]) == int(ids[0, 0])
    assert float(kv["reg_outlier_ratio"]) == pytest.approx(0.5, abs=1e-3) and int(kv["reg_outlier_chain_equal"]) == 1
    assert np.isfinite(float(kv["reg_minimizer_t03"])) and int(kv["reg_checker_max"]) == 7
    assert int(kv["reg_errors"]) == 7 and int(kv["reg_names"]) >= 15
    # times follow the points through a device filter
    assert int(kv["times_cols"]) == int(kv["times_points"]) > 0 and int(kv["times_bad"]) == 0
    # the candidate loop as one batch: same answers as the single calls, on one or two contexts
    Tb = np.array([float(x) for x in kv["T_batch0"].split(",")]).reshape(4, 4).T
    assert np.array_equal(Tb, T) and int(kv["batch_iterations0"]) == want["iterations"]
    assert int(kv["batch_converged"]) == 1 and int(kv["batch_two_contexts_equal"]) == 1
    assert int(kv["batch_pair2_status"]) == ob.icp_run(cfg, orf, orf)["status"]
    assert int(kv["file_roundtrip"]) == 1
    # T = double: the pose is returned in double, not through float
    Td = np.array([float(x) for x in kv["T_double"].split(",")]).reshape(4, 4).T
    assert np.abs(Td - want["T"]).max() < 1e-9 and not np.array_equal(Td, Td.astype(np.float32).astype(np.float64))


@pytest.mark.gpu
def test_dropin_bench_mode(tmp_path):
    import json
    exe = _compile(tmp_path)
    out = subprocess.run([exe, "--bench", "20000"], check=True, capture_output=True, text=True, timeout=120).stdout
    d = json.loads([ln for ln in out.splitlines() if ln.startswith("{")][0])
    assert d["points"] == 20000 and d["dropin_ms"] > 0 and d["dropin_ms_unchanged_clouds"] <= d["dropin_ms"] * 1.5
