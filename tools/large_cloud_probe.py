"""Large-cloud robustness: an 8M-point reference (kd build, 20 global levels + local), exact k = 1 and k = 4 neighbours
checked against a brute-force torch scan for a sample of queries; SurfaceNormal on 4M points runs through.
usage: python tools/large_cloud_probe.py [points=8000000]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from pgslam_b200 import pm  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
torch.cuda.set_device(0)
ctx = pm.Context(0, torch.cuda.current_stream().cuda_stream)
g = torch.Generator(device="cuda").manual_seed(7)
ref = torch.ones((n, 4), device="cuda", dtype=torch.float32)
ref[:, :3] = (torch.rand((n, 3), generator=g, device="cuda") - 0.5) * torch.tensor([200.0, 200.0, 20.0], device="cuda")
nq = 200_000
qry = torch.ones((nq, 4), device="cuda", dtype=torch.float32)
qry[:, :3] = (torch.rand((nq, 3), generator=g, device="cuda") - 0.5) * torch.tensor([200.0, 200.0, 20.0], device="cuda")
torch.cuda.synchronize()
rf = pm.DataPoints(ctx=ctx, device_ptr=ref.data_ptr(), n=n)
rd = pm.DataPoints(ctx=ctx, device_ptr=qry.data_ptr(), n=nq)
for k in (1, 4):
    m = pm.Matcher("KDTreeMatcher", {"knn": k}, ctx=ctx)
    t0 = time.perf_counter()
    m.init(rf)
    ctx.synchronize()
    t1 = time.perf_counter()
    got = m.findClosests(rd)
    t2 = time.perf_counter()
    ids = torch.from_numpy(np.ascontiguousarray(got.ids)).cuda().long()   # k x nq
    d2 = torch.from_numpy(np.ascontiguousarray(got.dists)).cuda()
    bad = 0
    sample = torch.arange(0, nq, nq // 512, device="cuda")[:512]
    for s0 in range(0, sample.numel(), 64):
        sq = sample[s0:s0 + 64]
        q = qry[sq, :3]
        # brute force in float64 for the check of the SET (distances re-evaluated in fp32 below)
        best = torch.full((q.shape[0], k), float("inf"), device="cuda", dtype=torch.float64)
        for c0 in range(0, n, 2_000_000):
            c = ref[c0:c0 + 2_000_000, :3].double()
            dd = ((q.double()[:, None, :] - c[None, :, :]) ** 2).sum(-1)
            best = torch.cat([best, dd], dim=1).topk(k, dim=1, largest=False).values
        mine = ((q.double()[:, None, :] - ref[ids[:, sq].T, :3].double()) ** 2).sum(-1)  # 64 x k
        bad += int((torch.abs(mine.sort(dim=1).values - best) > 1e-6 * (1 + best)).sum())
    print(f"k={k}: build {1e3 * (t1 - t0):.1f} ms for {n} points, {nq} queries in {1e3 * (t2 - t1):.1f} ms (incl. download), "
          f"mismatches against brute force on 512 queries: {bad}", flush=True)
    assert bad == 0
nn = min(n, 4_000_000)
dp = pm.DataPoints(ctx=ctx, device_ptr=ref.data_ptr(), n=nn)
t0 = time.perf_counter()
pm.DataPointsFilters("- SurfaceNormalDataPointsFilter:\n    knn: 10\n", ctx=ctx).apply(dp)
ctx.synchronize()
print(f"SurfaceNormal(knn=10) on {nn} points: {1e3 * (time.perf_counter() - t0):.1f} ms, descriptors {dp.descriptor_labels() if hasattr(dp, 'descriptor_labels') else ''}")
