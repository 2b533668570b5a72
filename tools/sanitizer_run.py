"""A small pass over the hot path for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
two ragged 30k-pt C2 registrations as a batch, one 120k pair, the input-filter chain, a matcher with k = 10.
usage: compute-sanitizer --tool memcheck python tools/sanitizer_run.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from pgslam_b200 import pm, synth  # noqa: E402
from tests import util  # noqa: E402

ctx = pm.Context(0)
pm._DEFAULT_CTX = ctx
rd, rf, _ = synth.scan_pair(3, beams=16, az_steps=1875)
icp = pm.ICP(ctx)
icp.loadFromYaml(util.to_yaml(util.C2))
a = [pm.DataPoints(rd, ctx=ctx), pm.DataPoints(rd[:, :20011], ctx=ctx), pm.DataPoints(rd[:, :7], ctx=ctx)]
b = [pm.DataPoints(rf, ctx=ctx), pm.DataPoints(rf[:, :25013], ctx=ctx), pm.DataPoints(rf[:, :9], ctx=ctx)]
res = icp.compute_batch(a, b)
print("batch", [(r["status"], r["iterations"]) for r in res])
if len(sys.argv) > 1 and sys.argv[1] == "big":
    rd2, rf2, _ = synth.scan_pair(5, beams=64, az_steps=1875)
    T = icp(pm.DataPoints(rd2, ctx=ctx), pm.DataPoints(rf2, ctx=ctx))
    print("120k", icp.last["iterations"])
dp = pm.DataPoints(rd, ctx=ctx)
pm.DataPointsFilters(util.to_yaml(util.INPUT_FILTERS)).apply(dp)
print("filters", dp.getNbPoints())
m = pm.Matcher("KDTreeMatcher", {"knn": 10})
m.init(pm.DataPoints(rf, ctx=ctx))
print("knn", m.findClosests(pm.DataPoints(rd[:, :5000], ctx=ctx)).ids.shape)
icp1 = pm.ICP(ctx)
icp1.loadFromYaml(util.to_yaml(util.C1))
icp1(pm.DataPoints(rd, ctx=ctx), pm.DataPoints(rf, ctx=ctx))
print("C1", icp1.last["iterations"])
