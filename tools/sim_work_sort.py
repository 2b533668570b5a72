"""CPU simulation (numpy + Python loops, no GPU): if the queries of the seeded matcher were re-grouped by the work their
search took in the PREVIOUS iteration, how much shorter would the slowest lane of every warp be?  Four ICP-like poses
converging on the truth, 16k Morton-ordered queries of a synthetic 120k-pt pair.  Result:
2.5 % at a work correlation of 0.65, 11-13 % at 0.85; sorting by the (unknowable) own work would give 40 %.
usage: python tools/sim_work_sort.py [queries=16384]"""
import numpy as np, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pgslam_b200 import synth
rd4, rf4, T_true = synth.scan_pair(3)
ref = np.ascontiguousarray(np.asarray(rf4, dtype=np.float32)[:3].T)
rdp = np.ascontiguousarray(np.asarray(rd4, dtype=np.float32)[:3].T)
n = len(ref); L = 8
nl = (n + L - 1) // L
P = 1
while P < nl: P *= 2
depth = int(np.log2(P))
order = np.arange(n)
sys.setrecursionlimit(100000)
def build(s, e, cap):
    if cap == 1 or e - s <= 0: return
    half = cap // 2 * L
    if e - s > half:
        seg = order[s:e]
        ext = ref[seg].max(0) - ref[seg].min(0)
        ax = int(np.argmax(ext))
        idx = np.argpartition(ref[seg, ax], half - 1)
        order[s:e] = seg[idx]
        build(s, s + half, cap // 2); build(s + half, e, cap // 2)
    else:
        build(s, e, cap // 2)
build(0, n, P)
sp = ref[order]
lo = np.full((2 * P, 3), np.inf, np.float32); hi = np.full((2 * P, 3), -np.inf, np.float32)
for j in range(nl):
    seg = sp[j * L:(j + 1) * L]
    lo[P + j] = seg.min(0); hi[P + j] = seg.max(0)
for i in range(P - 1, 0, -1):
    lo[i] = np.minimum(lo[2 * i], lo[2 * i + 1]); hi[i] = np.maximum(hi[2 * i], hi[2 * i + 1])
lo_l = lo.tolist(); hi_l = hi.tolist(); sp_l = sp.tolist()
def lb(q, i):
    a = lo_l[i]; b = hi_l[i]; s = 0.0
    for d in range(3):
        t = a[d] - q[d]
        if t > 0: s += t * t
        else:
            t = q[d] - b[d]
            if t > 0: s += t * t
    return s
def scan(q, leaf, best):
    base = leaf * L
    bd, bp = best
    for j in range(base, min(base + L, n)):
        p = sp_l[j]
        d = (p[0]-q[0])**2 + (p[1]-q[1])**2 + (p[2]-q[2])**2
        if d < bd: bd = d; bp = j
    return (bd, bp)
def trav(q, x, best, cnt):
    # descend nearest-first with re-test of far child
    if x >= P:
        if x - P < nl:
            cnt[1] += 1
            return scan(q, x - P, best)
        return best
    cnt[0] += 1
    l0 = lb(q, 2 * x); l1 = lb(q, 2 * x + 1)
    if l0 <= l1: near, far, ln, lf = 2 * x, 2 * x + 1, l0, l1
    else: near, far, ln, lf = 2 * x + 1, 2 * x, l1, l0
    if ln <= best[0]: best = trav(q, near, best, cnt)
    if lf <= best[0]:
        cnt[2] += 1
        best = trav(q, far, best, cnt)
    return best
H = 10
def seeded(q, seed_pos):
    cnt = [0, 0, 0]
    if seed_pos < 0:
        best = trav(q, 1, (np.inf, -1), cnt)
        return best, cnt
    p = sp_l[seed_pos]
    best = ((p[0]-q[0])**2 + (p[1]-q[1])**2 + (p[2]-q[2])**2, seed_pos)
    leaf = seed_pos // L
    node = (P + leaf) >> H
    best = trav(q, node, best, cnt)
    d = depth - H
    while d > 0:
        sib = node ^ 1
        if lb(q, sib) <= best[0]: best = trav(q, sib, best, cnt)
        node >>= 1; d -= 1
    return best, cnt
def pose(dx, dyaw):
    c, s = np.cos(dyaw), np.sin(dyaw)
    D = np.eye(4); D[:2, :2] = [[c, -s], [s, c]]; D[0, 3] = dx
    return D @ np.asarray(T_true, dtype=np.float64)
def xform(T):
    return (rdp.astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
# Morton order of the reading (at the first pose)
qa = xform(pose(0.10, np.deg2rad(0.5)))
mn = ref.min(0); ext = ref.max(0) - mn
g = np.clip(((qa - mn) / ext * 255).astype(np.int64), 0, 255)
def spread(v):
    v = (v | (v << 16)) & 0x030000FF; v = (v | (v << 8)) & 0x0300F00F; v = (v | (v << 4)) & 0x030C30C3; v = (v | (v << 2)) & 0x09249249
    return v
mort = spread(g[:, 0]) | (spread(g[:, 1]) << 1) | (spread(g[:, 2]) << 2)
mo = np.argsort(mort, kind='stable')
NQ = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
start = 40000
sel = mo[start:start + NQ]
poses = [pose(0.10, np.deg2rad(0.5)), pose(0.03, np.deg2rad(0.15)), pose(0.01, np.deg2rad(0.05)), pose(0.003, np.deg2rad(0.015))]
seed = np.full(NQ, -1)
works = []
for it, T in enumerate(poses):
    q = xform(T)[sel].tolist()
    w = np.zeros(NQ); newseed = np.zeros(NQ, dtype=np.int64)
    for i in range(NQ):
        best, cnt = seeded(q[i], int(seed[i]))
        w[i] = cnt[0] + 3.0 * cnt[1]   # descend step ~40 instr, leaf scan ~134
        newseed[i] = best[1]
    seed = newseed
    works.append(w)
    print("iteration", it, "mean work", w.mean(), "p50", np.median(w), "p90", np.percentile(w, 90), "p99", np.percentile(w, 99), "max", w.max(), flush=True)
def warp_cost(w, perm):
    ww = w[perm]
    m = ww[:len(ww) // 32 * 32].reshape(-1, 32)
    return m.max(1).sum(), m.sum() / (32 * m.max(1).sum())
ident = np.arange(NQ)
for it in (2, 3):
    w_prev, w_cur = works[it - 1], works[it]
    print("iteration", it, "corr(prev, cur) =", np.corrcoef(w_prev, w_cur)[0, 1])
    c0, e0 = warp_cost(w_cur, ident)
    # buckets from the previous iteration's work, stable inside a bucket
    for nb in (4, 8, 16):
        edges = np.quantile(w_prev, np.linspace(0, 1, nb + 1)[1:-1])
        b = np.searchsorted(edges, w_prev)
        perm = np.argsort(b, kind='stable')
        c1, e1 = warp_cost(w_cur, perm)
        print(f"  {nb} buckets by previous work: warp-max cost {c1 / c0:.3f} of Morton order (lane eff {e0:.3f} -> {e1:.3f})")
    perm = np.argsort(w_cur, kind='stable')
    c2, e2 = warp_cost(w_cur, perm)
    print(f"  oracle sort by own work: {c2 / c0:.3f} (lane eff {e2:.3f})")
