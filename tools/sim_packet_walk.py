"""CPU simulation (numpy + Python loops, no GPU): would a self-kNN walk shared by the 32 tree-ordered queries of a warp
(a node is entered if ANY lane needs it) beat 32 per-thread walks?  Counts node tests and leaf scans of both on a
synthetic 120k-pt scan.  Result: no - the union is 80 steps, the slowest lane needs 40.
usage: python tools/sim_packet_walk.py [k=10]"""
import numpy as np, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pgslam_b200 import synth
K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
rd, rf, T = synth.scan_pair(3)
pts = np.ascontiguousarray(np.asarray(rf, dtype=np.float32)[:3].T)
n = len(pts); L = 8
nl = (n + L - 1) // L
P = 1
while P < nl: P *= 2
depth = int(np.log2(P))
order = np.arange(n)
def build(s, e, cap):  # segment [s,e) of `order`, capacity cap leaves
    if cap == 1 or e - s <= 0: return
    half = cap // 2 * L
    if e - s > half:
        seg = order[s:e]
        ext = pts[seg].max(0) - pts[seg].min(0)
        ax = int(np.argmax(ext))
        k = half
        idx = np.argpartition(pts[seg, ax], k - 1)  # left gets `half` smallest
        # make deterministic-ish
        order[s:e] = seg[idx]
        build(s, s + half, cap // 2); build(s + half, e, cap // 2)
    else:
        build(s, e, cap // 2)
sys.setrecursionlimit(100000)
build(0, n, P)
sp = pts[order]
lo = np.full((2 * P, 3), np.inf, np.float32); hi = np.full((2 * P, 3), -np.inf, np.float32)
for j in range(nl):
    seg = sp[j * L:(j + 1) * L]
    lo[P + j] = seg.min(0); hi[P + j] = seg.max(0)
for i in range(P - 1, 0, -1):
    lo[i] = np.minimum(lo[2 * i], lo[2 * i + 1]); hi[i] = np.maximum(hi[2 * i], hi[2 * i + 1])
def lb(q, i):
    d = np.maximum(0, np.maximum(lo[i] - q, q - hi[i]))
    return float((d * d).sum())
class Acc:
    def __init__(s): s.d = [np.inf] * K
    def bound(s): return s.d[-1]
    def offer(s, dd):
        if dd < s.d[-1]:
            s.d[-1] = dd; s.d.sort()
def scan(q, leaf, acc):
    seg = sp[leaf * L:(leaf + 1) * L]
    for dd in ((seg - q) ** 2).sum(1): acc.offer(float(dd))
# the per-thread walk: nearest child first, the far child re-tested against the tighter bound
def trav(q, x, acc, cnt):
    if x >= P:
        if x - P < nl: scan(q, x - P, acc); cnt[1] += 1
        return
    cnt[0] += 1
    l0, l1 = lb(q, 2 * x), lb(q, 2 * x + 1)
    near, far, ln, lf = (2 * x, 2 * x + 1, l0, l1) if l0 <= l1 else (2 * x + 1, 2 * x, l1, l0)
    if ln <= acc.bound(): trav(q, near, acc, cnt)
    if lf <= acc.bound(): cnt[0] += 0.3; trav(q, far, acc, cnt)
def single2(j, group=1):
    q = sp[j]; acc = Acc(); leaf = j // L
    first = leaf & ~((1 << group) - 1)
    for l in range(first, first + (1 << group)):
        if l < nl: scan(q, l, acc)
    cnt = [0, 0]
    node = (P + leaf) >> group; d = depth - group
    while d > 0:
        sib = node ^ 1
        cnt[0] += 0.3
        if lb(q, sib) <= acc.bound(): trav(q, sib, acc, cnt)
        node >>= 1; d -= 1
    return cnt
def packet(w, group=2):
    js = [j for j in range(32 * w, min(32 * w + 32, n))]
    qs = [sp[j] for j in js]; accs = [Acc() for _ in js]
    first = (js[0] // L) & ~((1 << group) - 1)
    for l in range(first, first + (1 << group)):
        if l < nl:
            for q, a in zip(qs, accs): scan(q, l, a)
    cnt = [0, 0, 0]  # node steps, leaf scans, sum of lanes wanting leaf
    def anyq(x):
        return [lb(q, x) <= a.bound() for q, a in zip(qs, accs)]
    def ptrav(x):
        if x >= P:
            if x - P < nl:
                want = anyq(x)
                cnt[1] += 1; cnt[2] += sum(want)
                for q, a, wnt in zip(qs, accs, want):
                    if wnt: scan(q, x - P, a)
            return
        cnt[0] += 1
        # order: by lane majority / first lane's nearer child
        l0 = min(lb(q, 2 * x) for q in qs); l1 = min(lb(q, 2 * x + 1) for q in qs)
        near, far = (2 * x, 2 * x + 1) if l0 <= l1 else (2 * x + 1, 2 * x)
        if any(anyq(near)): ptrav(near)
        if any(anyq(far)): cnt[0] += 0.3; ptrav(far)
    node = (P + js[0] // L) >> group; d = depth - group
    while d > 0:
        sib = node ^ 1
        cnt[0] += 0.3
        if any(anyq(sib)): ptrav(sib)
        node >>= 1; d -= 1
    return cnt
rng = np.random.default_rng(0)
ws = rng.choice(n // 32, 60, replace=False)
tot_single_max = 0; tot_single_sum = 0; tot_packet = 0; tot_leafwant = 0; tot_pleaf = 0
for w in ws:
    per = [single2(j) for j in range(32 * w, 32 * w + 32)]
    cost = [c[0] + 2.5 * c[1] for c in per]   # leaf scan ~2.5x a node step
    tot_single_max += max(cost); tot_single_sum += sum(cost)
    pc = packet(w)
    tot_packet += pc[0] + 2.5 * pc[1]; tot_leafwant += pc[2]; tot_pleaf += pc[1]
print("K", K, "per-lane mean cost", tot_single_sum / (32 * len(ws)), "warp max cost", tot_single_max / len(ws),
      "lane eff (mean/max)", tot_single_sum / (32 * tot_single_max))
print("packet cost", tot_packet / len(ws), "leaf scans/warp", tot_pleaf / len(ws), "lanes wanting per leaf", tot_leafwant / max(1, tot_pleaf))
