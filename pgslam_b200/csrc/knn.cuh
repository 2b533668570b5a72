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
// Candidates are kept as 64-bit keys (distance bits << 32 | index): squared
// distances are non-negative, so their bit patterns order like the values and
// one unsigned compare implements the (distance, index) lexicographic rule.
//
// Traversal state is two registers: the 1-based heap index of the current node
// and a bit trail (bit j set = the sibling j levels up is still pending).  No
// stack memory; a sibling's box is re-tested against the (tighter) bound when
// the walk comes back to it.
#pragma once

#include "core.cuh"

namespace pgs {

#ifdef __CUDACC__

constexpr unsigned long long kEmptyKey = 0x7f8000007fffffffull;  // (+inf, INT_MAX)

__device__ __forceinline__ unsigned long long make_key(float d, int id) {
  return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)id;
}
__device__ __forceinline__ float key_dist(unsigned long long k) { return __uint_as_float((unsigned)(k >> 32)); }
__device__ __forceinline__ int key_id(unsigned long long k) { return (int)(unsigned)(k & 0xffffffffull); }

// The matcher's maxDist is folded into the initial state: every slot starts as the key
// (maxDist^2, INT_MAX), so a candidate farther than maxDist never wins, one exactly at
// maxDist does (its index is below INT_MAX), and neither the leaf scan nor the box tests need
// a separate radius compare.  A slot still holding INT_MAX at the end means "not found".
struct Best1 {
  unsigned long long key;  // (distance bits << 32) | original index
  int pos;
  __device__ __forceinline__ void init(float maxr2) {
    key = make_key(maxr2, 0x7fffffff);
    pos = -1;
  }
  __device__ __forceinline__ float bound() const { return key_dist(key); }
  __device__ __forceinline__ void offer(float dd, int iid, int ppos) {
    const unsigned long long nk = make_key(dd, iid);
    if (nk < key) { key = nk; pos = ppos; }
  }
  // squared distance of the result; +inf when nothing lies within maxDist
  __device__ __forceinline__ float dist() const { return pos < 0 ? __int_as_float(0x7f800000) : key_dist(key); }
};

// ascending list of exactly K keys, held in registers (all indices static)
template <int K>
struct BestK {
  unsigned long long key[K];
  __device__ __forceinline__ void init(float maxr2) {
#pragma unroll
    for (int j = 0; j < K; ++j) key[j] = make_key(maxr2, 0x7fffffff);
  }
  __device__ __forceinline__ float bound() const { return key_dist(key[K - 1]); }
  __device__ __forceinline__ void offer(float dd, int iid, int) {
    const unsigned long long nk = make_key(dd, iid);
    if (nk >= key[K - 1]) return;  // the common case: one compare
#pragma unroll
    for (int j = K - 1; j >= 1; --j) {
      const unsigned long long prev = key[j - 1];
      key[j] = (nk < prev) ? prev : ((nk < key[j]) ? nk : key[j]);
    }
    if (nk < key[0]) key[0] = nk;
  }
};

// K-list of run-time length (knn > 32): the keys live in local memory and an insertion shifts
// with a loop.  Slower than the register lists above - it exists so that a libpointmatcher
// configuration with a large `knn` (SurfaceNormal on sparse clouds) runs at all.
constexpr int kMaxDynK = 256;
struct BestDyn {
  unsigned long long* key;  // k entries, ascending
  int k;
  __device__ __forceinline__ void init(float maxr2) {
    for (int j = 0; j < k; ++j) key[j] = make_key(maxr2, 0x7fffffff);
  }
  __device__ __forceinline__ float bound() const { return key_dist(key[k - 1]); }
  __device__ __forceinline__ void offer(float dd, int iid, int) {
    const unsigned long long nk = make_key(dd, iid);
    if (nk >= key[k - 1]) return;
    int j = k - 1;
    while (j > 0 && key[j - 1] > nk) {
      key[j] = key[j - 1];
      --j;
    }
    key[j] = nk;
  }
};

// Batcher's odd-even merge sort of 16 keys held in registers (63 compare-exchanges, checked
// with the 0-1 principle by the generator): every index is a literal, so the array never
// leaves the register file.
__device__ __forceinline__ void sort_keys_network16(unsigned long long (&a)[16]) {
#define PGS_CSWAP(i, j)                                  \
  {                                                      \
    const unsigned long long x = a[i], y = a[j];         \
    const bool sw = y < x;                               \
    a[i] = sw ? y : x;                                   \
    a[j] = sw ? x : y;                                   \
  }
  PGS_CSWAP(0, 1); PGS_CSWAP(2, 3); PGS_CSWAP(4, 5); PGS_CSWAP(6, 7); PGS_CSWAP(8, 9); PGS_CSWAP(10, 11);
  PGS_CSWAP(12, 13); PGS_CSWAP(14, 15); PGS_CSWAP(0, 2); PGS_CSWAP(1, 3); PGS_CSWAP(4, 6); PGS_CSWAP(5, 7);
  PGS_CSWAP(8, 10); PGS_CSWAP(9, 11); PGS_CSWAP(12, 14); PGS_CSWAP(13, 15); PGS_CSWAP(1, 2); PGS_CSWAP(5, 6);
  PGS_CSWAP(9, 10); PGS_CSWAP(13, 14); PGS_CSWAP(0, 4); PGS_CSWAP(1, 5); PGS_CSWAP(2, 6); PGS_CSWAP(3, 7);
  PGS_CSWAP(8, 12); PGS_CSWAP(9, 13); PGS_CSWAP(10, 14); PGS_CSWAP(11, 15); PGS_CSWAP(2, 4); PGS_CSWAP(3, 5);
  PGS_CSWAP(10, 12); PGS_CSWAP(11, 13); PGS_CSWAP(1, 2); PGS_CSWAP(3, 4); PGS_CSWAP(5, 6); PGS_CSWAP(9, 10);
  PGS_CSWAP(11, 12); PGS_CSWAP(13, 14); PGS_CSWAP(0, 8); PGS_CSWAP(1, 9); PGS_CSWAP(2, 10); PGS_CSWAP(3, 11);
  PGS_CSWAP(4, 12); PGS_CSWAP(5, 13); PGS_CSWAP(6, 14); PGS_CSWAP(7, 15); PGS_CSWAP(4, 8); PGS_CSWAP(5, 9);
  PGS_CSWAP(6, 10); PGS_CSWAP(7, 11); PGS_CSWAP(2, 4); PGS_CSWAP(3, 5); PGS_CSWAP(6, 8); PGS_CSWAP(7, 9);
  PGS_CSWAP(10, 12); PGS_CSWAP(11, 13); PGS_CSWAP(1, 2); PGS_CSWAP(3, 4); PGS_CSWAP(5, 6); PGS_CSWAP(7, 8);
  PGS_CSWAP(9, 10); PGS_CSWAP(11, 12); PGS_CSWAP(13, 14);
#undef PGS_CSWAP
}

// Seed of a self-search: the aligned group of 2^group leaves (<= kSeedSlots points) around the
// query's own leaf is the same for all the lanes that share it, so instead of offering its
// points one by one (a divergent insertion each) every lane sorts the whole group's keys with
// a branch-free network and keeps the K smallest.  maxr2 caps the list as in BestK::init.
constexpr int kSeedSlots = 16;  // = the network's width
template <int K>
__device__ __forceinline__ void knn_seed_group(const TreeView& t, int leaf, int group, float qx, float qy, float qz,
                                               float maxr2, BestK<K>& acc) {
  static_assert(K <= kSeedSlots, "the seed group must hold at least K points");
  group = min(group, t.depth);
  const int first = (leaf & ~((1 << group) - 1)) * kLeaf;
  const int count = min(kLeaf << group, t.n_leaves * kLeaf - first);
  const unsigned long long qxy = pack_f32x2(qx, qy);
  unsigned long long a[kSeedSlots];
#pragma unroll
  for (int j = 0; j < kSeedSlots; ++j) {
    a[j] = kEmptyKey;
    if (j < count) {
      const float4 p = __ldg(t.pts + first + j);
      a[j] = make_key(dist2_rn_packed(qxy, qz, p.x, p.y, p.z), __float_as_int(p.w));
    }
  }
  sort_keys_network16(a);
  const unsigned long long cap = make_key(maxr2, 0x7fffffff);
#pragma unroll
  for (int j = 0; j < K; ++j) acc.key[j] = a[j] < cap ? a[j] : cap;
}

template <class Acc>
__device__ __forceinline__ void knn_scan_leaf(const TreeView& t, int leaf, float qx, float qy, float qz,
                                              Acc& acc, int skip_lo, int skip_hi) {
  if (leaf >= t.n_leaves) return;
  const float4* __restrict__ lp = t.pts + (size_t)leaf * kLeaf;
  const int base = leaf * kLeaf;
  const unsigned long long qxy = pack_f32x2(qx, qy);
#pragma unroll
  for (int j = 0; j < kLeaf; ++j) {
    float4 p = __ldg(lp + j);
    float dd = dist2_rn_packed(qxy, qz, p.x, p.y, p.z);
    const int pos = base + j;
    if (pos < skip_lo || pos > skip_hi) acc.offer(dd, __float_as_int(p.w), pos);
  }
}

// The same for a K-list.  ncu on knn_kernel<10> (profiles/r2_knn10_blocks.md): 44 % of its warp
// instructions were the eight unrolled 72-instruction insertions of a leaf scan, each running
// with ~4 of 32 lanes - whichever lanes happened to accept THAT point.  Here every lane first
// marks the points that beat its K-th key (8 distances, no insertion), then the lanes insert
// their marked points together, one per trip: the insertion code exists once and runs as many
// times as the lane with the most candidates needs (~3), not once per point slot (~6).
template <int K>
__device__ __forceinline__ void knn_scan_leaf(const TreeView& t, int leaf, float qx, float qy, float qz,
                                              BestK<K>& acc, int skip_lo, int skip_hi) {
  if (leaf >= t.n_leaves) return;
  const float4* __restrict__ lp = t.pts + (size_t)leaf * kLeaf;
  const int base = leaf * kLeaf;
  const unsigned long long qxy = pack_f32x2(qx, qy);
  const unsigned long long worst = acc.key[K - 1];
  unsigned mask = 0u;
#pragma unroll
  for (int j = 0; j < kLeaf; ++j) {
    const float4 p = __ldg(lp + j);
    const unsigned long long nk = make_key(dist2_rn_packed(qxy, qz, p.x, p.y, p.z), __float_as_int(p.w));
    const int pos = base + j;
    if ((pos < skip_lo || pos > skip_hi) && nk < worst) mask |= 1u << j;
  }
#pragma unroll 1
  while (mask) {
    const int j = __ffs(mask) - 1;
    mask &= mask - 1u;
    const float4 p = __ldg(lp + j);
    acc.offer(dist2_rn_packed(qxy, qz, p.x, p.y, p.z), __float_as_int(p.w), base + j);
  }
}

// Exhaustive best-first walk of the subtree under `node` (at `depth`), whose own
// box the caller has already accepted.  skip_lo..skip_hi: sorted positions the
// caller has already offered (pass an empty range (0, -1) otherwise).
template <class Acc>
__device__ __forceinline__ void knn_traverse_from(const TreeView& t, unsigned node, int depth, float qx, float qy,
                                                  float qz, Acc& acc, int skip_lo, int skip_hi) {
  const ulonglong2* __restrict__ nodes16 = reinterpret_cast<const ulonglong2*>(t.nodes);
  const unsigned long long* __restrict__ nodes8 = reinterpret_cast<const unsigned long long*>(t.nodes);
  const QueryPk q = pack_query(qx, qy, qz);
  unsigned trail = 0;
  while (true) {
    // ---- descend while the nearer child qualifies -------------------------
    bool at_leaf = true;
    while (depth < t.depth) {
      const ulonglong2* __restrict__ c = nodes16 + (size_t)node * 3;  // both children: 48 bytes
      const ulonglong2 a = __ldg(c), b = __ldg(c + 1), e = __ldg(c + 2);
      float lb0 = box_lb_packed(q, a.x, a.y, b.x);
      float lb1 = box_lb_packed(q, b.y, e.x, e.y);
      const float bound = acc.bound();
      bool near1 = lb1 < lb0;
      float lbn = near1 ? lb1 : lb0, lbf = near1 ? lb0 : lb1;
      if (!(lbn <= bound)) { at_leaf = false; break; }
      trail = (trail << 1) | ((lbf <= bound) ? 1u : 0u);
      node = node * 2 + (near1 ? 1u : 0u);
      ++depth;
    }
    if (at_leaf) knn_scan_leaf(t, (int)node - t.P, qx, qy, qz, acc, skip_lo, skip_hi);
    // ---- walk back to the deepest pending sibling whose box still qualifies --
    while (true) {
      if (trail == 0) return;
      int up = __ffs(trail) - 1;
      node >>= up;
      depth -= up;
      trail >>= up;
      node ^= 1u;
      trail ^= 1u;
      const unsigned long long* __restrict__ nb = nodes8 + (size_t)node * 3;
      float lb = box_lb_packed(q, __ldg(nb), __ldg(nb + 1), __ldg(nb + 2));
      if (lb <= acc.bound()) break;
    }
  }
}

// The walk of knn_traverse_from, entered at its walk-back stage: (node, depth) is a node that has
// been dealt with and `trail` marks the pending siblings above it (bit g = the sibling of the
// ancestor g levels up still has to be looked at).
template <class Acc>
__device__ __forceinline__ void knn_traverse_resume(const TreeView& t, unsigned node, int depth, unsigned trail, float qx,
                                                    float qy, float qz, Acc& acc) {
  const ulonglong2* __restrict__ nodes16 = reinterpret_cast<const ulonglong2*>(t.nodes);
  const unsigned long long* __restrict__ nodes8 = reinterpret_cast<const unsigned long long*>(t.nodes);
  const QueryPk q = pack_query(qx, qy, qz);
  while (true) {
    // ---- walk back to the deepest pending sibling whose box still qualifies --
    while (true) {
      if (trail == 0) return;
      const int up = __ffs(trail) - 1;
      node >>= up;
      depth -= up;
      trail >>= up;
      node ^= 1u;
      trail ^= 1u;
      const unsigned long long* __restrict__ nb = nodes8 + (size_t)node * 3;
      if (box_lb_packed(q, __ldg(nb), __ldg(nb + 1), __ldg(nb + 2)) <= acc.bound()) break;
    }
    // ---- descend while the nearer child qualifies -------------------------
    bool at_leaf = true;
    while (depth < t.depth) {
      const ulonglong2* __restrict__ c = nodes16 + (size_t)node * 3;
      const ulonglong2 a = __ldg(c), b = __ldg(c + 1), e = __ldg(c + 2);
      const float lb0 = box_lb_packed(q, a.x, a.y, b.x);
      const float lb1 = box_lb_packed(q, b.y, e.x, e.y);
      const float bound = acc.bound();
      const bool near1 = lb1 < lb0;
      const float lbn = near1 ? lb1 : lb0, lbf = near1 ? lb0 : lb1;
      if (!(lbn <= bound)) { at_leaf = false; break; }
      trail = (trail << 1) | ((lbf <= bound) ? 1u : 0u);
      node = node * 2 + (near1 ? 1u : 0u);
      ++depth;
    }
    if (at_leaf) knn_scan_leaf(t, (int)node - t.P, qx, qy, qz, acc, 0, -1);
  }
}

// Seeded search along the seed leaf's own path: the sibling of every ancestor of the seed leaf is
// tested ONCE against the bound the caller already has (one 24-byte box per level, addresses known
// from the leaf index alone, four loads in flight), which yields the pending-sibling trail a
// descent from the root would have built on its way to the seed leaf - without testing the boxes
// ON the path (the seed leaf is scanned whatever they say) and without the dependent loads of a
// descent.  Then the seed leaf is scanned and the walk continues as knn_traverse_from would.
template <class Acc>
__device__ __forceinline__ void knn_seed_path(const TreeView& t, int leaf, float qx, float qy, float qz, Acc& acc) {
  const unsigned long long* __restrict__ nodes8 = reinterpret_cast<const unsigned long long*>(t.nodes);
  const QueryPk q = pack_query(qx, qy, qz);
  const unsigned node = (unsigned)(t.P + leaf);
  const float bound = acc.bound();
  unsigned trail = 0u;
  for (int g = 0; g < t.depth; g += 4) {
    float lbv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      lbv[u] = __int_as_float(0x7f800000);
      if (g + u < t.depth) {
        const unsigned sib = (node >> (g + u)) ^ 1u;
        const unsigned long long* __restrict__ nb = nodes8 + (size_t)sib * 3;
        lbv[u] = box_lb_packed(q, __ldg(nb), __ldg(nb + 1), __ldg(nb + 2));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (g + u < t.depth && lbv[u] <= bound) trail |= 1u << (g + u);
  }
  knn_scan_leaf(t, leaf, qx, qy, qz, acc, 0, -1);
  knn_traverse_resume(t, node, t.depth, trail, qx, qy, qz, acc);
}

// top-down search from the root (no prior knowledge about the query)
template <class Acc>
__device__ __forceinline__ void knn_traverse(const TreeView& t, float qx, float qy, float qz,
                                             Acc& acc, int skip_lo = 0, int skip_hi = -1) {
  knn_traverse_from(t, 1u, 0, qx, qy, qz, acc, skip_lo, skip_hi);
}

// Bottom-up search from a seed leaf (last iteration's match, or the query's own
// leaf for self-kNN): scan the leaf, then climb to the root testing the SIBLING
// subtree at every level.  The sibling boxes sit at addresses known from the
// leaf index alone, so their loads are independent (four levels in flight at a
// time) instead of the root descent's chain of dependent loads; with a tight
// seed almost every sibling fails its bound test and is never entered.  The
// union of the seed leaf and all sibling subtrees is the whole tree, so the
// result is the same exact minimum as the top-down walk.
// `group` > 0 widens the seed from one leaf to the aligned group of 2^group leaves around it
// (scanned by every lane in lock step) and starts the climb `group` levels higher: the
// lowest siblings are the ones a query pokes into most often, and entering them one lane
// at a time costs far more issue slots than scanning them uniformly.
template <class Acc>
__device__ __forceinline__ void knn_climb(const TreeView& t, int leaf, float qx, float qy, float qz,
                                          Acc& acc, int skip_lo = 0, int skip_hi = -1, int group = 0,
                                          bool scan_seed = true) {
  group = min(group, t.depth);
  if (scan_seed) {
    const int first = leaf & ~((1 << group) - 1);
    knn_scan_leaf(t, leaf, qx, qy, qz, acc, skip_lo, skip_hi);
    for (int l = first; l < first + (1 << group); ++l)
      if (l != leaf) knn_scan_leaf(t, l, qx, qy, qz, acc, skip_lo, skip_hi);
  }
  unsigned node = (unsigned)(t.P + leaf) >> group;
  int depth = t.depth - group;
  const float inf = __int_as_float(0x7f800000);
  const unsigned long long* __restrict__ nodes8 = reinterpret_cast<const unsigned long long*>(t.nodes);
  const QueryPk q = pack_query(qx, qy, qz);
  while (depth > 0) {
    float lbv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      lbv[u] = inf;
      if (depth - u > 0) {
        const unsigned sib = (node >> u) ^ 1u;
        const unsigned long long* __restrict__ nb = nodes8 + (size_t)sib * 3;
        lbv[u] = box_lb_packed(q, __ldg(nb), __ldg(nb + 1), __ldg(nb + 2));
      }
    }
#pragma unroll 1
    for (int u = 0; u < 4; ++u) {
      const float lb = u == 0 ? lbv[0] : (u == 1 ? lbv[1] : (u == 2 ? lbv[2] : lbv[3]));
      // an empty box has lb = +inf: the `< inf` test keeps an unbounded first search out of it
      if (lb < inf && lb <= acc.bound())
        knn_traverse_from(t, (node >> u) ^ 1u, depth - u, qx, qy, qz, acc, skip_lo, skip_hi);
    }
    node >>= 4;
    depth -= 4;
  }
}

#endif  // __CUDACC__

}  // namespace pgs
