// kdorder.cu — balanced kd-tree ORDER of a cloud, built on the device.
//
// The implicit box tree (index.cu) is only as good as the order its leaves are
// cut from.  Cutting a Morton-sorted array at count midpoints gives sibling
// boxes that overlap massively (a Z-curve range is not convex): measured on a
// 120k-point scan, a seeded nearest-neighbour search then visits ~35 inner
// nodes and ~3.9 leaves per query.  Ordering the points as a balanced kd-tree
// instead (every node split at its count midpoint along the widest axis of its
// own points, left child filled first — exactly the shape of the implicit
// complete tree) makes siblings disjoint: ~4 inner nodes and ~1.9 leaves per
// query (profiles/r1_tree_quality.md).  Same tree arithmetic, same traversal,
// 5x less work per search.
//
// Build, batched over clouds (blockIdx.y / job index = cloud x segment):
//   global levels  while a segment spans more than kLocal points: per segment
//                  bounding box -> widest axis -> 11-bit key along it -> histogram
//                  select of the median bin -> stable partition (kd_hist /
//                  kd_part_count / kd_part_scatter below).  Only the split matters:
//                  both halves are re-split along their own axis one level down.
//   local levels   one block per kLocal-point segment finishes the remaining
//                  levels in shared memory with a bitonic network restricted
//                  to the (shrinking) sub-segments.
// The order depends only on the points, not on a translation of them, so the
// SurfaceNormal filter's index and the (mean-centred) matcher index share it.
#include "core.cuh"

namespace pgs {

namespace {

constexpr int kLocal = 4096;       // points finished by one block in shared memory
constexpr int kLocalThreads = 1024;
constexpr int kFineLevels = 2;     // global levels sorted on 16-bit keys; the ones below use 8 bits

struct KdCloud {
  const float4* pts;
  int n;
};

// number of real points of segment s (span S) of a cloud with n points
__device__ __forceinline__ int seg_count(int n, int s, int S) {
  long long r = (long long)n - (long long)s * S;
  return r < 0 ? 0 : (r > S ? S : (int)r);
}

__global__ void kd_iota_kernel(const KdCloud* __restrict__ clouds, uint32_t* __restrict__ vals, int span) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < clouds[b].n) vals[(size_t)b * span + i] = (uint32_t)i;
}

__global__ void kd_box_init_kernel(unsigned* __restrict__ bbox, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) bbox[i] = (i % 6 < 3) ? 0xffffffffu : 0u;
}

// A 256-thread block's (lo, hi) merged into a segment's box {lo.xyz, hi.xyz} of ordered-uint coordinates:
// one set of atomics per BLOCK.  (Per warp they were the whole run time of these kernels: the boxes of
// a batch share a few L2 lines, and atomics on one line are served one after the other.)
// Every thread of the block must call it; `slot` is 8 x 6 words of shared memory.
__device__ __forceinline__ void kd_box_commit(unsigned* __restrict__ box, unsigned (&lo)[3], unsigned (&hi)[3],
                                              unsigned (*slot)[6]) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    lo[d] = __reduce_min_sync(0xffffffffu, lo[d]);
    hi[d] = __reduce_max_sync(0xffffffffu, hi[d]);
  }
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { slot[w][d] = lo[d]; slot[w][3 + d] = hi[d]; }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    const bool is_lo = threadIdx.x < 3;
    unsigned v = slot[0][threadIdx.x];
    for (int q = 1; q < 8; ++q) v = is_lo ? min(v, slot[q][threadIdx.x]) : max(v, slot[q][threadIdx.x]);
    // an empty block leaves lo = 0xffffffff / hi = 0: both are no-ops on the box
    if (is_lo) atomicMin(box + threadIdx.x, v);
    else atomicMax(box + threadIdx.x, v);
  }
  __syncthreads();  // the slots may be used again
}

// the identity order and the cloud's bounding box (the only segment of level 0) in one pass
__global__ void __launch_bounds__(256)
kd_iota_box_kernel(const KdCloud* __restrict__ clouds, uint32_t* __restrict__ vals, int span, unsigned* __restrict__ bbox) {
  __shared__ unsigned slot[8][6];
  const int b = blockIdx.y;
  const KdCloud c = clouds[b];
  unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.n; i += gridDim.x * blockDim.x) {
    vals[(size_t)b * span + i] = (uint32_t)i;
    const float4 p = c.pts[i];
    const unsigned u[3] = {f2ord(p.x), f2ord(p.y), f2ord(p.z)};
#pragma unroll
    for (int d = 0; d < 3; ++d) { lo[d] = min(lo[d], u[d]); hi[d] = max(hi[d], u[d]); }
  }
  kd_box_commit(bbox + 6 * b, lo, hi, slot);
}

__global__ void kd_jobs_kernel(const KdCloud* __restrict__ clouds, int n_clouds, int segs, int S,
                               int* __restrict__ job_n, unsigned* __restrict__ bbox) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_clouds * segs) return;
  job_n[j] = seg_count(clouds[j / segs].n, j % segs, S);
  for (int d = 0; d < 3; ++d) { bbox[6 * j + d] = 0xffffffffu; bbox[6 * j + 3 + d] = 0u; }
}

__global__ void __launch_bounds__(256)
kd_bbox_kernel(const KdCloud* __restrict__ clouds, const uint32_t* __restrict__ vals, int span, int segs, int S,
               unsigned* __restrict__ bbox) {
  const int j = blockIdx.y;
  const int b = j / segs, s = j % segs;
  const KdCloud c = clouds[b];
  const int cnt = seg_count(c.n, s, S);
  const uint32_t* v = vals + (size_t)b * span + (size_t)s * S;
  unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
    float4 p = c.pts[v[i]];
    unsigned u[3] = {f2ord(p.x), f2ord(p.y), f2ord(p.z)};
#pragma unroll
    for (int d = 0; d < 3; ++d) { lo[d] = min(lo[d], u[d]); hi[d] = max(hi[d], u[d]); }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    lo[d] = __reduce_min_sync(0xffffffffu, lo[d]);
    hi[d] = __reduce_max_sync(0xffffffffu, hi[d]);
  }
  if ((threadIdx.x & 31) == 0 && lo[0] <= hi[0]) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { atomicMin(bbox + 6 * j + d, lo[d]); atomicMax(bbox + 6 * j + 3 + d, hi[d]); }
  }
}

__device__ __forceinline__ int widest_axis(const unsigned* bb) {
  float e0 = ord2f(bb[3]) - ord2f(bb[0]), e1 = ord2f(bb[4]) - ord2f(bb[1]), e2 = ord2f(bb[5]) - ord2f(bb[2]);
  int a = 0;
  float e = e0;
  if (e1 > e) { e = e1; a = 1; }
  if (e2 > e) { a = 2; }
  return a;
}

__global__ void __launch_bounds__(256)
kd_key_kernel(const KdCloud* __restrict__ clouds, const uint32_t* __restrict__ vals, int span, int segs, int S,
              const unsigned* __restrict__ bbox, uint32_t* __restrict__ keys, float key_max) {
  const int j = blockIdx.y;
  const int b = j / segs, s = j % segs;
  const KdCloud c = clouds[b];
  const int cnt = seg_count(c.n, s, S);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  const unsigned* bb = bbox + 6 * j;
  const int axis = widest_axis(bb);
  const size_t o = (size_t)b * span + (size_t)s * S + i;
  float4 p = c.pts[vals[o]];
  // key: the coordinate quantised over the segment's own extent, 16 bits (two radix passes)
  // at the two top levels, 8 bits (one pass) below.  Any order yields a valid tree (boxes
  // come from the points); points inside the median's bin are split by position, so the
  // sibling boxes may overlap by one bin: 1/65536th of a 60 m extent at the top, 1/256th of
  // a 30 m one (12 cm) at level 2 - a slab that holds well under 1 % of the queries.
  const float lo = ord2f(bb[axis]), hi = ord2f(bb[3 + axis]);
  const float x = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
  const float scale = hi > lo ? key_max / (hi - lo) : 0.f;
  int q = (int)((x - lo) * scale);
  keys[o] = (uint32_t)max(0, min((int)key_max, q));
}

// ---- global levels as a median SPLIT (radix select + stable partition) ----------------
//
// A level only has to put the smaller half of every segment first; both halves are re-split
// along their own axis one level down, so sorting them is wasted work.  Per level:
//   kd_hist     11-bit key of every point along the segment's widest axis (kept for the next
//               two kernels) -> per-block shared-memory histogram -> merged into the segment's
//               2048-bin histogram; the last block of a segment to finish finds the bin that
//               holds the median and how many of its members still go left;
//   kd_part_count   per tile: how many keys are below / in the median bin;
//   kd_part_scatter exclusive prefix over the tiles before it, warp-ballot ranks inside the
//               tile -> every index moves to its slot.  Stable, so the order is deterministic.
// Points inside the median's bin are split by position: sibling boxes may overlap by 1/2047th
// of the segment's extent (3 cm of a 60 m scene at the top level).
constexpr int kSplitBins = 2048;
constexpr int kSplitTile = 2048;  // elements per block: 256 threads x 8

struct SplitMeta {
  int med_bin;  // keys < med_bin go left; kSplitBins = every point goes left
  int n_less;   // points with key < med_bin
  int n_eq;     // points with key == med_bin
  int tie;      // how many of those still go left
};

__global__ void __launch_bounds__(256)
kd_hist_kernel(const KdCloud* __restrict__ clouds, const uint32_t* __restrict__ vals, int span, int segs, int S,
               const unsigned* __restrict__ bbox, unsigned short* __restrict__ keys, unsigned* __restrict__ hist,
               unsigned* __restrict__ tickets, SplitMeta* __restrict__ meta) {
  __shared__ unsigned sh[kSplitBins];
  __shared__ unsigned scan[256];
  __shared__ bool last;
  const int j = blockIdx.y;
  const int b = j / segs, s = j % segs;
  const KdCloud c = clouds[b];
  const int cnt = seg_count(c.n, s, S);
  if (cnt == 0) return;
  const int tid = threadIdx.x, lane = tid & 31;
  for (int d = tid; d < kSplitBins; d += 256) sh[d] = 0;
  __syncthreads();
  const unsigned* bb = bbox + 6 * j;
  const int axis = widest_axis(bb);
  const float lo = ord2f(bb[axis]), hi = ord2f(bb[3 + axis]);
  const float scale = hi > lo ? (float)(kSplitBins - 1) / (hi - lo) : 0.f;
  const size_t base = (size_t)j * S;
  const int t0 = blockIdx.x * kSplitTile;
#pragma unroll
  for (int r = 0; r < kSplitTile / 256; ++r) {
    const int i = t0 + r * 256 + tid;
    const bool valid = i < cnt;
    unsigned key = 0;
    if (valid) {
      const float4 p = c.pts[vals[base + i]];
      const float x = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
      key = (unsigned)max(0, min(kSplitBins - 1, (int)((x - lo) * scale)));
      keys[base + i] = (unsigned short)key;
    }
    const unsigned act = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const unsigned m = __match_any_sync(act, key);
      if (lane == __ffs(m) - 1) atomicAdd(&sh[key], (unsigned)__popc(m));
    }
  }
  __syncthreads();
  unsigned* gh = hist + (size_t)j * kSplitBins;
  for (int d = tid; d < kSplitBins; d += 256)
    if (sh[d]) atomicAdd(&gh[d], sh[d]);
  __threadfence();
  __syncthreads();
  if (tid == 0) last = (atomicAdd(&tickets[j], 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  // the last block owns the segment's merged histogram: read it and leave it cleared for the next level
  const int target = S / 2;  // the left child takes the first S/2 slots
  unsigned mine[kSplitBins / 256], sum = 0;
#pragma unroll
  for (int u = 0; u < kSplitBins / 256; ++u) {
    mine[u] = __ldcg(&gh[tid * (kSplitBins / 256) + u]);
    gh[tid * (kSplitBins / 256) + u] = 0u;
    sum += mine[u];
  }
  scan[tid] = sum;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {  // inclusive scan of the 256 partial sums
    const unsigned v = tid >= o ? scan[tid - o] : 0u;
    __syncthreads();
    scan[tid] += v;
    __syncthreads();
  }
  if (tid == 0) tickets[j] = 0;
  if (cnt <= target) {
    if (tid == 0) meta[j] = SplitMeta{kSplitBins, cnt, 0, 0};
    return;
  }
  const unsigned incl = scan[tid], excl = incl - sum;
  if (excl < (unsigned)target && incl >= (unsigned)target) {  // exactly one thread
    unsigned cum = excl;
    int u = 0;
    for (; u < kSplitBins / 256 - 1; ++u) {
      if (cum + mine[u] >= (unsigned)target) break;
      cum += mine[u];
    }
    meta[j] = SplitMeta{tid * (kSplitBins / 256) + u, (int)cum, (int)mine[u], target - (int)cum};
  }
}

__global__ void __launch_bounds__(256)
kd_part_count_kernel(const KdCloud* __restrict__ clouds, int segs, int S, const unsigned short* __restrict__ keys,
                     const SplitMeta* __restrict__ meta, int2* __restrict__ counts, int ntiles) {
  __shared__ int wl[8], we[8];
  const int j = blockIdx.y;
  const int cnt = seg_count(clouds[j / segs].n, j % segs, S);
  if (cnt == 0) return;
  const int med = meta[j].med_bin;
  const size_t base = (size_t)j * S;
  const int t0 = blockIdx.x * kSplitTile;
  int nl = 0, ne = 0;
#pragma unroll
  for (int r = 0; r < kSplitTile / 256; ++r) {
    const int i = t0 + r * 256 + threadIdx.x;
    if (i < cnt) {
      const int k = keys[base + i];
      nl += k < med;
      ne += k == med;
    }
  }
  nl = __reduce_add_sync(0xffffffffu, nl);
  ne = __reduce_add_sync(0xffffffffu, ne);
  if ((threadIdx.x & 31) == 0) { wl[threadIdx.x >> 5] = nl; we[threadIdx.x >> 5] = ne; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int a = 0, e = 0;
    for (int w = 0; w < 8; ++w) { a += wl[w]; e += we[w]; }
    counts[(size_t)j * ntiles + blockIdx.x] = make_int2(a, e);
  }
}

__global__ void __launch_bounds__(256)
kd_part_scatter_kernel(const KdCloud* __restrict__ clouds, int segs, int S, const uint32_t* __restrict__ vals_in,
                       uint32_t* __restrict__ vals_out, const unsigned short* __restrict__ keys,
                       const SplitMeta* __restrict__ meta, const int2* __restrict__ counts, int ntiles,
                       unsigned* __restrict__ child_bbox) {
  __shared__ int wl[8], we[8];
  __shared__ int tile_l, tile_e;
  __shared__ unsigned slot[8][6];
  const int j = blockIdx.y;
  const int cnt = seg_count(clouds[j / segs].n, j % segs, S);
  if (cnt == 0) return;
  const SplitMeta mt = meta[j];
  const size_t base = (size_t)j * S;
  const int t0 = blockIdx.x * kSplitTile;
  if (t0 >= cnt) return;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  if (tid < 32) {  // points below / in the median bin in the tiles before this one
    int a = 0, e = 0;
    for (int t = lane; t < (int)blockIdx.x; t += 32) {
      const int2 c2 = counts[(size_t)j * ntiles + t];
      a += c2.x;
      e += c2.y;
    }
    a = __reduce_add_sync(0xffffffffu, a);
    e = __reduce_add_sync(0xffffffffu, e);
    if (lane == 0) { tile_l = a; tile_e = e; }
  }
  // a warp owns 256 consecutive slots of the tile (8 rounds of 32): ranks follow the slot order
  const int w0 = t0 + w * 256;
  int side[8], rank[8];  // rank among the warp's own less / equal / greater slots
  uint32_t val[8];
  int run_l = 0, run_e = 0;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i = w0 + r * 32 + lane;
    const bool valid = i < cnt;
    const int key = valid ? (int)keys[base + i] : 0x7fffffff;
    val[r] = valid ? vals_in[base + i] : 0u;
    side[r] = !valid ? 3 : (key < mt.med_bin ? 0 : (key == mt.med_bin ? 1 : 2));
    const unsigned bl = __ballot_sync(0xffffffffu, side[r] == 0);
    const unsigned be = __ballot_sync(0xffffffffu, side[r] == 1);
    const unsigned lt = (1u << lane) - 1u;
    const int less_before = run_l + __popc(bl & lt), eq_before = run_e + __popc(be & lt);
    // valid slots are contiguous from the start of the segment, so every slot before a valid one is valid
    rank[r] = side[r] == 0 ? less_before : (side[r] == 1 ? eq_before : r * 32 + lane - less_before - eq_before);
    run_l += __popc(bl);
    run_e += __popc(be);
  }
  if (lane == 0) { wl[w] = run_l; we[w] = run_e; }
  __syncthreads();
  int off_l = tile_l, off_e = tile_e;  // less / equal points of the segment before this warp's slots
  for (int q = 0; q < w; ++q) { off_l += wl[q]; off_e += we[q]; }
  const int half = S / 2;
  const int off_g = w0 - off_l - off_e;
  // child_bbox (null at the last global level): the boxes of the next level's segments 2j and 2j+1 are
  // reduced here, where every point's side is known, instead of in a pass of their own
  const float4* __restrict__ pts = clouds[j / segs].pts;
  unsigned llo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, lhi[3] = {0u, 0u, 0u};
  unsigned rlo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, rhi[3] = {0u, 0u, 0u};
  float4 pt[8];
  if (child_bbox) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (side[r] != 3) pt[r] = pts[val[r]];
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    if (side[r] == 3) continue;
    int dst;
    if (side[r] == 0) {
      dst = off_l + rank[r];
    } else if (side[r] == 1) {
      const int e = off_e + rank[r];
      dst = e < mt.tie ? mt.n_less + e : half + (e - mt.tie);
    } else {
      dst = half + (mt.n_eq - mt.tie) + off_g + rank[r];
    }
    vals_out[base + dst] = val[r];
    if (child_bbox) {
      const unsigned u[3] = {f2ord(pt[r].x), f2ord(pt[r].y), f2ord(pt[r].z)};
      const bool left = dst < half;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        if (left) { llo[d] = min(llo[d], u[d]); lhi[d] = max(lhi[d], u[d]); }
        else { rlo[d] = min(rlo[d], u[d]); rhi[d] = max(rhi[d], u[d]); }
      }
    }
  }
  if (child_bbox) {
    kd_box_commit(child_bbox + 6 * (size_t)(2 * j), llo, lhi, slot);
    kd_box_commit(child_bbox + 6 * (size_t)(2 * j + 1), rlo, rhi, slot);
  }
}

// ---- local levels: one block owns m0 (<= kLocal) consecutive slots ------------
//
// Sub-segments of kRadixMin points or more are SPLIT, not sorted: only the median
// partition matters (both halves are re-split along their own axis one level
// down).  Per level: bounding box -> widest axis -> 8-bit key -> one 256-bin
// histogram pass finds the median's bin and how many of its members go left
// -> one stable partition pass (per-32-chunk counts + prefix, deterministic).
// Smaller sub-segments are sorted with a bitonic network whose strides stay
// inside a warp's window, so they need almost no block barriers.
constexpr int kRadixMin = 256;
constexpr int kMaxRadixSegs = kLocal / kRadixMin;  // 16
constexpr int kChunks = kLocal / 32;               // 128

__global__ void __launch_bounds__(kLocalThreads)
kd_local_kernel(const KdCloud* __restrict__ clouds, const uint32_t* vals, uint32_t* vals_out, int span, int segs,
                int m0) {
  extern __shared__ unsigned smem[];
  unsigned* sx = smem;                 // ordered-uint coordinates, fixed slots
  unsigned* sy = sx + kLocal;
  unsigned* sz = sy + kLocal;
  unsigned* key = sz + kLocal;         // key per position (16-bit in radix levels, 32-bit in bitonic levels)
  unsigned* bb = key + kLocal;         // 6 x (kLocal / 16) bounding boxes
  unsigned* hist = bb + 6 * (kLocal / 16);            // kMaxRadixSegs x 256
  int* seg_less = reinterpret_cast<int*>(hist + kMaxRadixSegs * 256);  // per radix segment
  int* seg_eq = seg_less + kMaxRadixSegs;
  int* seg_tie = seg_eq + kMaxRadixSegs;
  int* seg_b1 = seg_tie + kMaxRadixSegs;
  int* seg_below = seg_b1 + kMaxRadixSegs;
  int* seg_kp = seg_below + kMaxRadixSegs;
  unsigned short* perm_a = reinterpret_cast<unsigned short*>(seg_kp + kMaxRadixSegs);  // slot per position
  unsigned short* perm_b = perm_a + kLocal;
  unsigned short* ch_l = perm_b + kLocal;  // per 32-chunk counts, then exclusive offsets
  unsigned short* ch_e = ch_l + kChunks;
  unsigned short* ch_ol = ch_e + kChunks;
  unsigned short* ch_oe = ch_ol + kChunks;
  unsigned char* axis = reinterpret_cast<unsigned char*>(ch_oe + kChunks);

  const int b = blockIdx.y, s = blockIdx.x;
  const KdCloud c = clouds[b];
  const int cnt = seg_count(c.n, s, m0);
  if (cnt == 0) return;
  const uint32_t* v = vals + (size_t)b * span + (size_t)s * m0;
  uint32_t* vo = vals_out + (size_t)b * span + (size_t)s * m0;  // may alias v: gather, barrier, write
  const int tid = threadIdx.x, lane = tid & 31;
  unsigned short* perm = perm_a;
  unsigned short* perm_next = perm_b;
  for (int i = tid; i < m0; i += kLocalThreads) {
    // slot i is a real point iff i < cnt (the original indices stay in global memory)
    if (i < cnt) {
      float4 p = c.pts[v[i]];
      sx[i] = f2ord(p.x); sy[i] = f2ord(p.y); sz[i] = f2ord(p.z);
    } else {
      sx[i] = sy[i] = sz[i] = 0xffffffffu;  // padding sorts last on every axis
    }
    perm[i] = (unsigned short)i;
  }
  __syncthreads();
  for (int m = m0; m > kLeaf; m >>= 1) {
    // m0 and m are powers of two: every i / m below is a shift
    const int lm = 31 - __clz(m);
    const int nseg = m0 >> lm;
    // ---- bounding box of every sub-segment (real points only) -> widest axis ----
    for (int q = tid; q < nseg * 6; q += kLocalThreads) bb[q] = (q % 6 < 3) ? 0xffffffffu : 0u;  // constant divisor
    __syncthreads();
    for (int i = tid; i < m0; i += kLocalThreads) {
      const int slot = perm[i];
      const bool real = slot < cnt;
      unsigned lo[3] = {real ? sx[slot] : 0xffffffffu, real ? sy[slot] : 0xffffffffu, real ? sz[slot] : 0xffffffffu};
      unsigned hi[3] = {real ? sx[slot] : 0u, real ? sy[slot] : 0u, real ? sz[slot] : 0u};
      // lanes of a warp (of a half-warp when m == 16) share a sub-segment: reduce before the atomics
      const unsigned mask = m >= 32 ? 0xffffffffu : ((tid & 16) ? 0xffff0000u : 0x0000ffffu);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        lo[d] = __reduce_min_sync(mask, lo[d]);
        hi[d] = __reduce_max_sync(mask, hi[d]);
      }
      const bool leader = m >= 32 ? (tid & 31) == 0 : (tid & 15) == 0;
      if (leader && lo[0] <= hi[0]) {
        unsigned* q = bb + 6 * (i >> lm);
#pragma unroll
        for (int d = 0; d < 3; ++d) { atomicMin(q + d, lo[d]); atomicMax(q + 3 + d, hi[d]); }
      }
    }
    __syncthreads();
    for (int q = tid; q < nseg; q += kLocalThreads) axis[q] = (unsigned char)widest_axis(bb + 6 * q);
    __syncthreads();

    if (m >= kRadixMin) {
      // ================= median split by radix select + stable partition =================
      // 8-bit key: the coordinate quantised over the sub-segment's own extent (padding = 255).
      // Points that share the median's bin are split by position, so sibling boxes can overlap
      // by 1/255 of the extent - millimetres at these levels - which no query notices, and
      // one histogram pass finds the split.
      for (int i = tid; i < m0; i += kLocalThreads) {
        const int slot = perm[i];
        const int sg = i >> lm;
        const int a = axis[sg];
        unsigned k8 = 255u;
        if (slot < cnt) {
          const unsigned* q = bb + 6 * sg;
          const float lo = ord2f(q[a]), hi = ord2f(q[3 + a]);
          const float x = ord2f(a == 0 ? sx[slot] : (a == 1 ? sy[slot] : sz[slot]));
          const float scale = hi > lo ? 254.0f / (hi - lo) : 0.f;
          k8 = (unsigned)max(0, min(254, (int)((x - lo) * scale)));
        }
        key[i] = k8;
      }
      const int target = m / 2;  // the left child takes the `target` smallest keys
      for (int q = tid; q < nseg * 256; q += kLocalThreads) hist[q] = 0;
      __syncthreads();
      for (int i = tid; i < m0; i += kLocalThreads) {
        const unsigned bin = key[i];
        const unsigned peers = __match_any_sync(0xffffffffu, bin | ((unsigned)(i >> lm) << 8));
        if (lane == __ffs(peers) - 1) atomicAdd(&hist[(i >> lm) * 256 + bin], (unsigned)__popc(peers));
      }
      __syncthreads();
      {
        // one warp per sub-segment: smallest bin whose cumulative count reaches the target
        const int w = tid >> 5;
        if (w < nseg) {
          const unsigned* h = hist + w * 256;
          int mine[8], sum = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) { mine[j] = (int)h[lane * 8 + j]; sum += mine[j]; }
          int incl = sum;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
          }
          const unsigned reach = __ballot_sync(0xffffffffu, incl >= target);
          const int sel = __ffs(reach) - 1;  // always found: the segment holds m >= target slots
          if (lane == sel) {
            int cum = incl - sum, j = 0;
            for (; j < 7; ++j) {
              if (cum + mine[j] >= target) break;
              cum += mine[j];
            }
            seg_kp[w] = lane * 8 + j;
            seg_less[w] = cum;
            seg_eq[w] = mine[j];
            seg_tie[w] = target - cum;  // ties that still go left (>= 1)
          }
        }
        __syncthreads();
      }
      // stable partition: per-chunk counts -> exclusive offsets inside the sub-segment
      const int rounds = (m0 + kLocalThreads - 1) / kLocalThreads;
      int my_rank[kLocal / kLocalThreads];
      signed char my_side[kLocal / kLocalThreads];  // 0 less, 1 equal, 2 greater
#pragma unroll
      for (int r = 0; r < kLocal / kLocalThreads; ++r) {
        const int i = r * kLocalThreads + tid;
        my_side[r] = -1;
        my_rank[r] = 0;
        if (r < rounds && i < m0) {
          const int sg = i >> lm;
          const int k16 = (int)key[i], kp = seg_kp[sg];
          const int side = k16 < kp ? 0 : (k16 == kp ? 1 : 2);
          const unsigned bl = __ballot_sync(0xffffffffu, side == 0);
          const unsigned be = __ballot_sync(0xffffffffu, side == 1);
          const unsigned bg = ~(bl | be);
          const unsigned lt = (1u << lane) - 1u;
          my_side[r] = (signed char)side;
          my_rank[r] = __popc((side == 0 ? bl : (side == 1 ? be : bg)) & lt);
          if (lane == 0) { ch_l[i >> 5] = (unsigned short)__popc(bl); ch_e[i >> 5] = (unsigned short)__popc(be); }
        }
      }
      __syncthreads();
      if (tid < m0 / 32) {
        const int first = (tid >> (lm - 5)) << (lm - 5);
        int sl = 0, se = 0;
        for (int c2 = first; c2 < tid; ++c2) { sl += ch_l[c2]; se += ch_e[c2]; }
        ch_ol[tid] = (unsigned short)sl;
        ch_oe[tid] = (unsigned short)se;
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < kLocal / kLocalThreads; ++r) {
        const int i = r * kLocalThreads + tid;
        if (my_side[r] >= 0) {
          const int sg = i >> lm, ch = i >> 5, base = sg << lm;
          const int ol = ch_ol[ch], oe = ch_oe[ch];
          int pos;
          if (my_side[r] == 0) {
            pos = base + ol + my_rank[r];
          } else if (my_side[r] == 1) {
            const int e = oe + my_rank[r], tie = seg_tie[sg];
            pos = e < tie ? base + seg_less[sg] + e : base + target + (e - tie);
          } else {
            const int og = (ch - (sg << (lm - 5))) * 32 - ol - oe;
            pos = base + target + (seg_eq[sg] - seg_tie[sg]) + og + my_rank[r];
          }
          perm_next[pos] = perm[i];
        }
      }
      __syncthreads();
      unsigned short* t = perm; perm = perm_next; perm_next = t;
    } else {
      // ================= small sub-segments: bitonic sort of (key, perm) ==================
      for (int i = tid; i < m0; i += kLocalThreads) {
        const int slot = perm[i];
        const int a = axis[i >> lm];
        key[i] = a == 0 ? sx[slot] : (a == 1 ? sy[slot] : sz[slot]);
      }
      // Pair t of a stage with stride jj is (i, i|jj), i = t with a zero bit inserted
      // at log2(jj): for jj <= 32 the 32 pairs of a warp stay inside the warp's own
      // 64-element window.  A stage with jj >= 32 follows one that wrote across
      // warps or opens a new merge, so it starts with a block barrier; stages with
      // jj <= 16 follow warp-local stages and only need a warp barrier.
      __syncthreads();
      for (int k = 2; k <= m; k <<= 1) {
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
          if (jj >= 32) __syncthreads(); else __syncwarp();
          for (int t = tid; t < m0 / 2; t += kLocalThreads) {
            const int i = ((t & ~(jj - 1)) << 1) | (t & (jj - 1));
            const int l = i | jj;
            const bool up = ((i & (m - 1)) & k) == 0;
            const unsigned ki = key[i], kl = key[l];
            if ((ki > kl) == up && ki != kl) {
              key[i] = kl; key[l] = ki;
              const unsigned short pi = perm[i];
              perm[i] = perm[l]; perm[l] = pi;
            }
          }
        }
      }
      __syncthreads();
    }
  }
  // apply the permutation to the original indices in place: gather, barrier, write
  uint32_t out[kLocal / kLocalThreads];
#pragma unroll
  for (int r = 0; r < kLocal / kLocalThreads; ++r) {
    const int i = r * kLocalThreads + tid;
    out[r] = i < cnt ? v[perm[i]] : 0u;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kLocal / kLocalThreads; ++r) {
    const int i = r * kLocalThreads + tid;
    if (i < cnt) vo[i] = out[r];
  }
}

}  // namespace

// d_vals_out: n_clouds x span uint32 (span = pow2 leaves x kLeaf >= every n).
// On return d_vals_out[b*span + j] is the original index of the j-th point of
// cloud b in kd order.
void kd_order_batched(Ctx* ctx, const std::vector<const float4*>& pts, const std::vector<int>& n, int span,
                      uint32_t* d_vals_out) {
  const int B = (int)pts.size();
  if (B == 0) return;
  cudaStream_t st = ctx->stream;
  std::vector<KdCloud> hc(B);
  int max_n = 0;
  for (int b = 0; b < B; ++b) { hc[b] = KdCloud{pts[b], n[b]}; max_n = std::max(max_n, n[b]); }
  if (max_n == 0) return;
  DBuf<KdCloud> clouds(ctx, B);
  ctx->upload_small(clouds.p, hc.data(), sizeof(KdCloud) * B);
  const size_t total = (size_t)B * span;
  DBuf<uint32_t> keys_a(ctx, total), keys_b(ctx, total), vals_b(ctx, total);
  uint32_t* cur = d_vals_out;   // the buffer that holds the current order
  uint32_t* other = vals_b.p;
  // PGS_KD_SORT_LEVELS=1 brings back the earlier global levels (full segmented radix sorts on
  // 16 / 8-bit keys) for comparison: 2.2 ms per 96 clouds against 1.1 ms for the median split
  static const bool sort_levels = std::getenv("PGS_KD_SORT_LEVELS") && std::atoi(std::getenv("PGS_KD_SORT_LEVELS")) != 0;
  int level = 0;
  int n_levels = 0;
  while ((span >> n_levels) > kLocal) ++n_levels;
  if (!sort_levels && n_levels > 0) {
    // median split: per level histogram -> tile counts -> scatter; the scatter also reduces the boxes of
    // the two children, the last block of a histogram leaves it cleared (3 launches per level)
    const int max_jobs = B << (n_levels - 1);
    const int max_tiles = ceil_div(std::min(span, max_n), kSplitTile);
    DBuf<unsigned> hist(ctx, (size_t)max_jobs * kSplitBins);
    hist.zero();
    DBuf<unsigned> tickets(ctx, (size_t)max_jobs);
    tickets.zero();
    DBuf<SplitMeta> meta(ctx, (size_t)max_jobs);
    DBuf<int2> counts(ctx, (size_t)B * max_tiles * 2 + max_jobs);  // tiles per job halve as jobs double
    const int all_boxes = 6 * B * ((1 << n_levels) - 1);          // level l starts at 6 * B * (2^l - 1)
    DBuf<unsigned> bbox(ctx, (size_t)all_boxes);
    kd_box_init_kernel<<<ceil_div(all_boxes, 256), 256, 0, st>>>(bbox.p, all_boxes);
    kd_iota_box_kernel<<<dim3(std::max(1, std::min(ceil_div(max_n, 1024), 128)), B), 256, 0, st>>>(clouds.p, cur, span, bbox.p);
    ctx_count_launches(ctx, 2);
    unsigned short* keys16 = reinterpret_cast<unsigned short*>(keys_a.p);
    for (; level < n_levels; ++level) {
      const int segs = 1 << level, S = span >> level, jobs = B * segs;
      const int ntiles = ceil_div(std::min(S, max_n), kSplitTile);
      unsigned* boxes = bbox.p + (size_t)6 * B * (segs - 1);
      unsigned* child_boxes = level + 1 < n_levels ? bbox.p + (size_t)6 * B * (2 * segs - 1) : nullptr;
      const dim3 grid(ntiles, jobs);
      kd_hist_kernel<<<grid, 256, 0, st>>>(clouds.p, cur, span, segs, S, boxes, keys16, hist.p, tickets.p, meta.p);
      kd_part_count_kernel<<<grid, 256, 0, st>>>(clouds.p, segs, S, keys16, meta.p, counts.p, ntiles);
      kd_part_scatter_kernel<<<grid, 256, 0, st>>>(clouds.p, segs, S, cur, other, keys16, meta.p, counts.p, ntiles, child_boxes);
      ctx_count_launches(ctx, 3);
      std::swap(cur, other);
    }
  } else {
    kd_iota_kernel<<<dim3(ceil_div(max_n, 256), B), 256, 0, st>>>(clouds.p, cur, span);
    ctx_count_launches(ctx, 1);
  }
  for (; level < n_levels; ++level) {  // the sort-based levels (comparison build only)
    const int segs = 1 << level, S = span >> level, jobs = B * segs;
    DBuf<int> job_n(ctx, jobs);
    DBuf<unsigned> bbox(ctx, (size_t)6 * jobs);
    kd_jobs_kernel<<<ceil_div(jobs, 128), 128, 0, st>>>(clouds.p, B, segs, S, job_n.p, bbox.p);
    const int seg_max = std::min(S, max_n);
    kd_bbox_kernel<<<dim3(std::max(1, std::min(ceil_div(seg_max, 2048), 64)), jobs), 256, 0, st>>>(clouds.p, cur, span,
                                                                                                 segs, S, bbox.p);
    const int key_bits = level < kFineLevels ? 16 : 8;
    kd_key_kernel<<<dim3(ceil_div(seg_max, 256), jobs), 256, 0, st>>>(clouds.p, cur, span, segs, S, bbox.p, keys_a.p,
                                                                      (float)((1 << key_bits) - 1));
    ctx_count_launches(ctx, 3);
    // jobs are laid out back to back with stride S: cloud b, segment s starts at (b*segs + s) * S
    const bool in_b = radix_sort_pairs<uint32_t>(ctx, keys_a.p, keys_b.p, cur, other, job_n.p, jobs, S, seg_max, key_bits);
    if (in_b) std::swap(cur, other);  // a one-pass sort leaves the order in the other buffer
  }
  const int m0 = std::min(span, kLocal);
  const int segs = span / m0;
  const size_t smem = (size_t)kLocal * 4 * 4 + 6 * (kLocal / 16) * 4 + (size_t)kMaxRadixSegs * 256 * 4 +
                      6 * kMaxRadixSegs * 4 + (size_t)kLocal * 2 * 2 + 4 * kChunks * 2 + kLocal / 16 + 64;
  PGS_CUDA(cudaFuncSetAttribute(kd_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kd_local_kernel<<<dim3(segs, B), kLocalThreads, smem, st>>>(clouds.p, cur, d_vals_out, span, segs, m0);
  ctx_count_launches(ctx, 1);
  PGS_LAUNCH_CHECK();
}

}  // namespace pgs
