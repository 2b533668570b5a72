"""Per-iteration trace of the fused ICP loop on a C2 batch (PGS_TRACE_LOOP=1):
match / select / accumulate milliseconds and the number of seeded queries the leaf
adjacency lists could not prove complete.  Usage: python tools/trace_loop.py [pairs]"""
import os
import sys

os.environ["PGS_TRACE_LOOP"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pgslam_b200 import pm  # noqa: E402
from tests import util  # noqa: E402
import bench  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 48
ctx = pm.Context(0)
icp = pm.ICP(ctx)
icp.loadFromYaml(util.to_yaml(util.C2))
data = bench.gen_pairs(range(pairs))
rd = [pm.DataPoints(r, ctx=ctx) for r, _ in data]
rf = [pm.DataPoints(f, ctx=ctx) for _, f in data]
for _ in range(2):
    icp.compute_batch(rd, rf)
ctx.set_profiling(True)
sys.stderr.write("---- traced batch ----\n")
res = icp.compute_batch(rd, rf)
print("iterations", sorted(r["iterations"] for r in res))
print(ctx.stage_times())
