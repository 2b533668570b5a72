"""Soak: many registrations in one process, device memory watched (cudaMallocAsync pools must reach a steady state).
usage: python tools/soak.py [single_calls=3000] [batch_calls=40]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from pgslam_b200 import pm, synth_torch  # noqa: E402
from tests import util  # noqa: E402

singles = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
batches = int(sys.argv[2]) if len(sys.argv) > 2 else 40
torch.cuda.set_device(0)
ctx = pm.Context(0, torch.cuda.current_stream().cuda_stream)
icp = pm.ICP(ctx)
icp.loadFromYaml(util.to_yaml(util.C2))
data, _ = synth_torch.scan_pairs(range(96), "cuda")
torch.cuda.synchronize()
rd = [pm.DataPoints(ctx=ctx, device_ptr=r.data_ptr(), n=r.shape[0]) for r, _ in data]
rf = [pm.DataPoints(ctx=ctx, device_ptr=f.data_ptr(), n=f.shape[0]) for _, f in data]
ctx.synchronize()


def used_mb():
    free, total = torch.cuda.mem_get_info()
    return (total - free) / 2**20


first = None
t0 = time.perf_counter()
for i in range(singles):
    rec = icp.compute_batch_array([rd[i % 96]], [rf[i % 96]])
    if first is None:
        first = rec.tobytes()
    if i % 96 == 0:
        assert rec.tobytes() == first, "result of pair 0 changed"
    if i % 500 == 0:
        print(f"single {i}: device memory in use {used_mb():.0f} MB", flush=True)
print(f"{singles} single registrations in {time.perf_counter() - t0:.1f} s, {used_mb():.0f} MB in use", flush=True)
ref = None
for b in range(batches):
    rec = icp.compute_batch_array(rd, rf)
    ref = ref or rec.tobytes()
    assert rec.tobytes() == ref, "batch result changed"
    if b % 10 == 0:
        print(f"batch {b}: device memory in use {used_mb():.0f} MB", flush=True)
print(f"{batches} batches of 96, {used_mb():.0f} MB in use; results identical throughout", flush=True)
