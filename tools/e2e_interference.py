"""Does host->device DMA traffic alone slow the resident batch?  The resident leg of a 384-pair batch
is timed (a) alone, (b) while a helper thread streams pinned host memory to the device at about the
rate the host-entry leg needs, (c) through the host entry itself.
usage: python tools/e2e_interference.py [pairs=384]"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from pgslam_b200 import pm, synth_torch  # noqa: E402
from tests import util  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 384
torch.cuda.set_device(0)
ctx = pm.Context(0, torch.cuda.current_stream().cuda_stream)
icp = pm.ICP(ctx)
icp.loadFromYaml(util.to_yaml(util.C2))
data, _ = synth_torch.scan_pairs(range(pairs), "cuda")
torch.cuda.synchronize()
rd = [pm.DataPoints(ctx=ctx, device_ptr=r.data_ptr(), n=r.shape[0]) for r, _ in data]
rf = [pm.DataPoints(ctx=ctx, device_ptr=f.data_ptr(), n=f.shape[0]) for _, f in data]
ctx.synchronize()
handles = pm.batch_handles(rd, rf)
ctx.set_batch_streams(8)
ctx.set_option("batch_chunk", 12)


def resident(reps=3):
    icp.compute_batch_array(rd, rf, handles=handles)
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        icp.compute_batch_array(rd, rf, handles=handles)
        best = min(best, time.perf_counter() - t0)
    return pairs / best


print(f"resident alone: {resident():.0f} reg/s", flush=True)
stop = False
moved = [0]
host = torch.empty(46 * 1024 * 1024, dtype=torch.uint8).pin_memory()
dev = torch.empty_like(host, device="cuda")
side = torch.cuda.Stream()


def pump(sleep_s):
    with torch.cuda.stream(side):
        while not stop:
            dev.copy_(host, non_blocking=True)
            side.synchronize()
            moved[0] += host.numel()
            if sleep_s:
                time.sleep(sleep_s)


for sleep_s in (0.0015, 0.0):
    stop = False
    moved[0] = 0
    th = threading.Thread(target=pump, args=(sleep_s,))
    th.start()
    t0 = time.perf_counter()
    r = resident()
    dt = time.perf_counter() - t0
    stop = True
    th.join()
    print(f"resident with {moved[0] / dt / 1e9:.1f} GB/s of H2D beside it: {r:.0f} reg/s", flush=True)

host_rd = [r.cpu().pin_memory() for r, _ in data]
host_rf = [f.cpu().pin_memory() for _, f in data]
hrd = pm.host_clouds([(t.data_ptr(), t.shape[0]) for t in host_rd])
hrf = pm.host_clouds([(t.data_ptr(), t.shape[0]) for t in host_rf])
pm.compute_batch_multi([icp], hrd, hrf, pinned=True)
best = 1e9
for _ in range(3):
    t0 = time.perf_counter()
    pm.compute_batch_multi([icp], hrd, hrf, pinned=True)
    best = min(best, time.perf_counter() - t0)
print(f"host entry: {pairs / best:.0f} reg/s ({pairs * 2 * 1.92e6 / best / 1e9:.1f} GB/s of uploads)", flush=True)
