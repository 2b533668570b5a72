"""Short driver for ncu: a few C2 registrations (120k-pt pairs) through the C ABI.
Usage: python tools/profile_run.py [pairs] [repeats] [batch_streams=1]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pgslam_b200 import pm  # noqa: E402
from tests import util  # noqa: E402
import bench  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
streams = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ctx = pm.Context(0)
ctx.set_batch_streams(streams)  # 1: the whole batch in every launch, as in bench.py's profiled step
icp = pm.ICP(ctx)
icp.loadFromYaml(util.to_yaml(util.C2))
data = bench.gen_pairs(range(pairs))
rd = [pm.DataPoints(r, ctx=ctx) for r, _ in data]
rf = [pm.DataPoints(f, ctx=ctx) for _, f in data]
for _ in range(reps):
    res = icp.compute_batch(rd, rf)
print("iterations", [r["iterations"] for r in res], "launches", ctx.launch_count)
