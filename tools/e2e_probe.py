import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from pgslam_b200 import pm
from tests import util
torch.cuda.set_device(0)
stream = torch.cuda.current_stream()
ctx = pm.Context(0, stream.cuda_stream)
icp = pm.ICP(ctx); icp.loadFromYaml(util.to_yaml(util.C2))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
ctx.set_batch_streams(int(sys.argv[2]) if len(sys.argv) > 2 else 4)
pairs = bench.gen_pairs(range(B))
host = [(torch.from_numpy(np.ascontiguousarray(rd.T)).pin_memory(), torch.from_numpy(np.ascontiguousarray(rf.T)).pin_memory()) for rd, rf in pairs]
def upload():
    rds = [pm.DataPoints(ctx=ctx, pinned_host_ptr=h.data_ptr(), n=h.shape[0]) for h, _ in host]
    rfs = [pm.DataPoints(ctx=ctx, pinned_host_ptr=h.data_ptr(), n=h.shape[0]) for _, h in host]
    return rds, rfs
for mode in ('sync-upload', 'pipelined'):
    tu = tc = 0.0
    pending = [upload()]
    torch.cuda.synchronize()
    t_all = time.perf_counter()
    per = []
    for step in range(int(sys.argv[3]) if len(sys.argv) > 3 else 8):
        t0 = time.perf_counter()
        if mode == 'pipelined':
            cur = pending.pop(); pending.append(upload())
        else:
            cur = upload()
        t1 = time.perf_counter()
        res = icp.compute_batch(*cur)
        t2 = time.perf_counter()
        del cur
        tu += t1 - t0; tc += t2 - t1
        t3 = time.perf_counter()
        per.append((round(1e3*(t1-t0),1), round(1e3*(t2-t1),1), round(1e3*(t3-t2),1)))
    torch.cuda.synchronize()
    tot = time.perf_counter() - t_all
    print('per step (upload-call, compute-call, del):', per)
    print(mode, 'per step ms: upload-call %.2f compute-call %.2f total %.2f -> %.0f reg/s' % (1e3*tu/len(per), 1e3*tc/len(per), 1e3*tot/len(per), B*len(per)/tot))
# resident for comparison
rds, rfs = upload(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(8): icp.compute_batch(rds, rfs)
torch.cuda.synchronize(); tot = time.perf_counter() - t0
print('resident per step ms %.2f -> %.0f reg/s' % (1e3*tot/8, B*8/tot))
# host-side cost of the compute_batch call itself (python + ctypes marshalling)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); icp.compute_batch(rds, rfs); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(8)
