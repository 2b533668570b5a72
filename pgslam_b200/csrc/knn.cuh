// knn.cuh — stackless exact nearest-neighbour traversal of the implicit box tree.
//
// Replaces libnabo's recursive kd-tree descent (KDTreeMatcher::findClosests ->
// NNS::knn).  Result contract: the exact
// lexicographic minimum of (fp32 squared distance, original reference index),
// i.e. eps = 0 with ties broken by the lower index.
//
// Exactness: a subtree is skipped only if the lower bound of its box, computed
// with the same fp32 operation order as the distance itself, is strictly
// greater than the current k-th best distance; IEEE rounding is monotone, so
// the bound never exceeds the distance of any point in the box.  Equal bounds
// are visited because an equal distance with a lower index must still win.
//
// Traversal state is two registers: the 1-based heap index of the current node
// and a bit trail (bit j set = the sibling j levels up is still pending).  No
// stack memory; a sibling's box is re-tested against the (tighter) bound when
// the walk comes back to it.
#pragma once

#include "core.cuh"

namespace pgs {

#ifdef __CUDACC__

struct Best1 {
  float d;
  int id;
  int pos;
  __device__ __forceinline__ void init() {
    d = __int_as_float(0x7f800000);
    id = 0x7fffffff;
    pos = -1;
  }
  __device__ __forceinline__ float bound() const { return d; }
  __device__ __forceinline__ void offer(float dd, int iid, int ppos) {
    if (dd < d || (dd == d && iid < id)) { d = dd; id = iid; pos = ppos; }
  }
};

// ascending (d, id) list of capacity KCAP, of which the first k entries count
template <int KCAP>
struct BestK {
  float d[KCAP];
  int id[KCAP];
  int k;
  __device__ __forceinline__ void init(int kk) {
    k = kk;
#pragma unroll
    for (int j = 0; j < KCAP; ++j) { d[j] = __int_as_float(0x7f800000); id[j] = 0x7fffffff; }
  }
  __device__ __forceinline__ float bound() const {
    float b = d[0];
#pragma unroll
    for (int j = 1; j < KCAP; ++j) b = (j == k - 1) ? d[j] : b;
    return (k == 1) ? d[0] : b;
  }
  __device__ __forceinline__ void offer(float dd, int iid, int) {
    // cheap reject against the k-th entry first
    float bd = bound();
    if (dd > bd) return;
    int bid = id[0];
#pragma unroll
    for (int j = 1; j < KCAP; ++j) bid = (j == k - 1) ? id[j] : bid;
    if (dd == bd && iid >= bid) return;
#pragma unroll
    for (int j = KCAP - 1; j >= 1; --j) {
      if (j < k) {
        bool before_prev = dd < d[j - 1] || (dd == d[j - 1] && iid < id[j - 1]);
        bool before_this = dd < d[j] || (dd == d[j] && iid < id[j]);
        float nd = before_prev ? d[j - 1] : (before_this ? dd : d[j]);
        int ni = before_prev ? id[j - 1] : (before_this ? iid : id[j]);
        d[j] = nd;
        id[j] = ni;
      }
    }
    bool first = dd < d[0] || (dd == d[0] && iid < id[0]);
    if (first) { d[0] = dd; id[0] = iid; }
  }
};

template <class Acc>
__device__ __forceinline__ void knn_traverse(const TreeView& t, float qx, float qy, float qz, float maxr2,
                                             Acc& acc) {
  const float4* __restrict__ nodes4 = reinterpret_cast<const float4*>(t.nodes);
  unsigned node = 1, trail = 0;
  int depth = 0;
  while (true) {
    bool descend = false;
    if (depth == t.depth) {
      const int leaf = (int)node - t.P;
      if (leaf < t.n_leaves) {
        const float4* __restrict__ lp = t.pts + (size_t)leaf * kLeaf;
#pragma unroll
        for (int j = 0; j < kLeaf; ++j) {
          float4 p = lp[j];
          float dd = dist2_rn(qx, qy, qz, p.x, p.y, p.z);
          if (dd <= maxr2) acc.offer(dd, __float_as_int(p.w), leaf * kLeaf + j);
        }
      }
    } else {
      // both children: 12 consecutive floats at node*48 bytes
      const float4* __restrict__ c = nodes4 + (size_t)node * 3;
      float4 a = c[0], b = c[1], e = c[2];
      float lb0 = box_lb_rn(qx, qy, qz, a.x, a.y, a.z, a.w, b.x, b.y);
      float lb1 = box_lb_rn(qx, qy, qz, b.z, b.w, e.x, e.y, e.z, e.w);
      float bound = fminf(acc.bound(), maxr2);
      bool near1 = lb1 < lb0;
      float lbn = near1 ? lb1 : lb0, lbf = near1 ? lb0 : lb1;
      if (lbn <= bound) {
        trail = (trail << 1) | ((lbf <= bound) ? 1u : 0u);
        node = node * 2 + (near1 ? 1u : 0u);
        ++depth;
        descend = true;
      }
    }
    if (descend) continue;
    // walk back to the deepest pending sibling whose box still qualifies
    while (true) {
      if (trail == 0) return;
      int up = __ffs(trail) - 1;
      node >>= up;
      depth -= up;
      trail >>= up;
      node ^= 1u;
      trail ^= 1u;
      const float* __restrict__ nb = t.nodes + (size_t)node * 6;
      float lb = box_lb_rn(qx, qy, qz, nb[0], nb[1], nb[2], nb[3], nb[4], nb[5]);
      if (lb <= fminf(acc.bound(), maxr2)) break;
    }
  }
}

#endif  // __CUDACC__

}  // namespace pgs
