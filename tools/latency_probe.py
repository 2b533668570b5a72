"""Where one C2 registration's latency goes: wall time per call, stage times from the library's
own events, launches per call.  Usage: python tools/latency_probe.py [repeats=30] [profile=1]
With profile=0 it only runs the calls (the form to put under an ncu launch list)."""
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from pgslam_b200 import pm  # noqa: E402
from tests import util  # noqa: E402
import bench  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
prof = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = pm.Context(0)
icp = pm.ICP(ctx)
icp.loadFromYaml(util.to_yaml(util.C2))
(r, f), = bench.gen_pairs([0])
rd, rf = pm.DataPoints(r, ctx=ctx), pm.DataPoints(f, ctx=ctx)
for _ in range(5):
    rec = icp.compute_batch_array([rd], [rf])
l0 = ctx.launch_count
wall = []
for _ in range(reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rec = icp.compute_batch_array([rd], [rf])
    wall.append((time.perf_counter() - t0) * 1e3)
launches = (ctx.launch_count - l0) / reps
wall.sort()
print("wall ms: min %.3f p50 %.3f p90 %.3f   launches/call %.1f  iterations %d" %
      (wall[0], wall[len(wall) // 2], wall[int(0.9 * len(wall))], launches, int(rec["iterations"][0])))
if prof:
    ctx.set_profiling(True)
    st = []
    for _ in range(5):
        icp.compute_batch_array([rd], [rf])
        st.append(ctx.stage_times())
    ctx.set_profiling(False)
    keys = st[0].keys()
    print("stage ms (median of 5):", {k: round(statistics.median(s[k] for s in st), 4) for k in keys})
