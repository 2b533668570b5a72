"""Time the matcher variants on the bench batch (96 C2 pairs of 120k points, one stream, CUDA
events around every launch).  Usage: python tools/match_modes.py [pairs] ["mode:blocks:refill:pw:lw" ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pgslam_b200 import pm  # noqa: E402
from tests import util  # noqa: E402
import bench  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 96
specs = sys.argv[2:] or ["0", "1", "2", "3"]
ctx = pm.Context(0)
ctx.set_batch_streams(1)
icp = pm.ICP(ctx)
icp.loadFromYaml(util.to_yaml(util.C2))
data = bench.gen_pairs(range(pairs))
rd = [pm.DataPoints(r, ctx=ctx) for r, _ in data]
rf = [pm.DataPoints(f, ctx=ctx) for _, f in data]
ref = None
for spec in specs:
    resort = -1
    if "r" in spec:
        spec, r = spec.split("r")
        resort = int(r)
    ctx.set_option("resort_it", resort)
    f = [int(x) for x in spec.split(":")]
    ctx.set_option("match_mode", f[0])
    keys = ("mq_batches", "mq_blocks") if f[0] == 4 else ("pm_blocks", "pm_refill", "pm_pair_w", "pm_leaf_w")
    for key, val in zip(keys, f[1:]):
        ctx.set_option(key, val)
    for _ in range(2):
        icp.compute_batch(rd, rf)
    ctx.set_profiling(True)
    best = None
    for _ in range(3):
        res = icp.compute_batch(rd, rf)
        st = ctx.stage_times()
        if best is None or st["match_ms"] < best["match_ms"]:
            best = st
    ctx.set_profiling(False)
    sig = [(r["iterations"], np.round(r["T"], 9).tobytes()) for r in res]
    if ref is None:
        ref = sig
    print(f"spec {spec}: match {best['match_ms']:.3f} ms select {best['select_ms']:.3f} acc {best['accumulate_ms']:.3f} "
          f"loop {best['loop_ms']:.3f} total {best['total_ms']:.3f} launches {best['iterations_launched']} identical {sig == ref}",
          flush=True)
