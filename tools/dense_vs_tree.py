"""KDTreeMatcher k = 1 on small clouds: tensor-core distance tiles (dense.cu) against the tree.
usage: python tools/dense_vs_tree.py [sizes ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from pgslam_b200 import pm, synth  # noqa: E402

sizes = [int(x) for x in sys.argv[1:]] or [512, 1024, 2048, 4096, 8192, 16384]
ctx = pm.Context(0)
for n in sizes:
    beams = 16
    az = max(8, n // beams)
    rd, rf, _ = synth.scan_pair(5, beams=beams, az_steps=az)
    ref, qry = pm.DataPoints(rf, ctx=ctx), pm.DataPoints(rd, ctx=ctx)
    import ctypes as C
    nq = rd.shape[1]
    import torch
    ids = torch.empty(nq, dtype=torch.int32, device="cuda")
    d2 = torch.empty(nq, dtype=torch.float32, device="cuda")
    res = {}
    for name, dense in (("tree", 0), ("dense", 1 << 14)):
        ctx.set_option("dense_max_ref", dense)
        m = pm.Matcher("KDTreeMatcher", {"knn": 1}, ctx=ctx)
        m.init(ref)
        call = lambda: ctx.check(ctx.lib.pgs_matcher_find(m.h, qry.h, C.c_void_p(ids.data_ptr()), C.c_void_p(d2.data_ptr()), 1))
        for _ in range(5):
            call()
        ctx.synchronize()
        reps = 200
        t0 = time.perf_counter()
        for _ in range(reps):
            call()
        ctx.synchronize()
        res[name] = (1e6 * (time.perf_counter() - t0) / reps, ids.clone())
    same = bool(torch.equal(res["tree"][1], res["dense"][1]))
    print(f"n_ref = n_query = {rf.shape[1]:6d}: tree {res['tree'][0]:8.1f} us/call, dense {res['dense'][0]:8.1f} us/call "
          f"(back-to-back calls, one stream), identical {same}", flush=True)

# ---- the same inside the fused ICP loop: a batch of small pairs (loop-closure candidates on subsampled clouds)
from tests import util  # noqa: E402
for n in [s for s in sizes if s <= 8192]:
    pairs = [synth.scan_pair(100 + i, beams=16, az_steps=max(8, n // 16))[:2] for i in range(48)]
    rd = [pm.DataPoints(r, ctx=ctx) for r, _ in pairs]
    rf = [pm.DataPoints(f, ctx=ctx) for _, f in pairs]
    icp = pm.ICP(ctx)
    icp.loadFromYaml(util.to_yaml(util.C2))
    line = f"48 pairs of {pairs[0][0].shape[1]:5d}-pt clouds, C2 chain:"
    sig = None
    for name, dense in (("tree", 0), ("dense", 1 << 14)):
        ctx.set_option("dense_max_ref", dense)
        ctx.set_batch_streams(1)
        for _ in range(2):
            icp.compute_batch_array(rd, rf)
        ctx.set_profiling(True)
        rec = icp.compute_batch_array(rd, rf)
        st = ctx.stage_times()
        ctx.set_profiling(False)
        sig = sig or rec.tobytes()
        line += f"  {name}: match {st['match_ms']:.3f} ms of loop {st['loop_ms']:.3f} ms ({st['iterations_launched']} launches)"
        same = rec.tobytes() == sig
    print(line + f"  identical {same}", flush=True)
ctx.set_option("dense_max_ref", 0)
