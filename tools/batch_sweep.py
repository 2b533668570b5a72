"""Throughput of one big batch of C2 pairs for several (worker streams, chunk) settings.
usage: python tools/batch_sweep.py [pairs] ["streams:chunk[:match_mode]" ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from pgslam_b200 import pm, synth_torch  # noqa: E402
from tests import util  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 384
specs = sys.argv[2:] or ["4:24", "4:16", "6:16", "8:12", "3:32", "2:48"]
torch.cuda.set_device(0)
ctx = pm.Context(0, torch.cuda.current_stream().cuda_stream)
icp = pm.ICP(ctx)
icp.loadFromYaml(util.to_yaml(util.C2))
t0 = time.perf_counter()
data, _ = synth_torch.scan_pairs(range(pairs), "cuda")
torch.cuda.synchronize()
print(f"generated {pairs} pairs on the device in {time.perf_counter() - t0:.2f} s", flush=True)
rd = [pm.DataPoints(ctx=ctx, device_ptr=r.data_ptr(), n=r.shape[0]) for r, _ in data]
rf = [pm.DataPoints(ctx=ctx, device_ptr=f.data_ptr(), n=f.shape[0]) for _, f in data]
ctx.synchronize()
handles = pm.batch_handles(rd, rf)
ref = None
for spec in specs:
    f = [int(x) for x in spec.split(":")]
    ctx.set_batch_streams(f[0])
    ctx.set_option("batch_chunk", f[1])
    ctx.set_option("match_mode", f[2] if len(f) > 2 else 0)
    icp.compute_batch_array(rd, rf, handles=handles)
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = icp.compute_batch_array(rd, rf, handles=handles)
        best = min(best, time.perf_counter() - t0)
    sig = res.tobytes()
    ref = ref or sig
    print(f"streams:chunk {spec}: {pairs / best:8.1f} reg/s ({1e3 * best:.1f} ms per {pairs} pairs), ok {int((res['status'] == 0).sum())}, "
          f"iterations {res['iterations'].mean():.3f}, identical {sig == ref}", flush=True)

# the same through the host-memory entry (pinned buffers, uploads inside)
host_rd = [r.cpu().pin_memory() for r, _ in data]
host_rf = [f.cpu().pin_memory() for _, f in data]
hrd = pm.host_clouds([(t.data_ptr(), t.shape[0]) for t in host_rd])
hrf = pm.host_clouds([(t.data_ptr(), t.shape[0]) for t in host_rf])
for spec in (specs if os.environ.get('PGS_SWEEP_HOST_ALL') else specs[:3]):
    f = [int(x) for x in spec.split(":")]
    ctx.set_batch_streams(f[0])
    ctx.set_option("batch_chunk", f[1])
    pm.compute_batch_multi([icp], hrd, hrf, pinned=True)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        res = pm.compute_batch_multi([icp], hrd, hrf, pinned=True)
        best = min(best, time.perf_counter() - t0)
    print(f"host entry, streams:chunk {spec}: {pairs / best:8.1f} reg/s end to end, identical {res.tobytes() == ref}", flush=True)
