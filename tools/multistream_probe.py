"""Experiment: one batch of C2 pairs split over S host threads, each with its own context
(stream).  Wall-clock registrations/s for S = 1, 2, 3, 4, 6.  Usage: python tools/multistream_probe.py [pairs]"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from pgslam_b200 import pm  # noqa: E402
from tests import util  # noqa: E402
import bench  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 48
data = bench.gen_pairs(range(pairs))
for S in (1, 2, 3, 4, 6, 8):
    ctxs = [pm.Context(0) for _ in range(S)]
    icps = []
    for c in ctxs:
        icp = pm.ICP(c)
        icp.loadFromYaml(util.to_yaml(util.C2))
        icps.append(icp)
    share = [list(range(s, pairs, S)) for s in range(S)]
    rd = [[pm.DataPoints(data[i][0], ctx=ctxs[s]) for i in share[s]] for s in range(S)]
    rf = [[pm.DataPoints(data[i][1], ctx=ctxs[s]) for i in share[s]] for s in range(S)]

    def work(s, n):
        for _ in range(n):
            icps[s].compute_batch(rd[s], rf[s])

    def run(n):
        th = [threading.Thread(target=work, args=(s, n)) for s in range(S)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    run(3)
    dt = run(8)
    print(f"S={S}: {pairs * 8 / dt:.0f} registrations/s ({1e3 * dt / 8:.2f} ms per {pairs}-pair step)", flush=True)
    del rd, rf, icps, ctxs
