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
//                  bounding box -> widest axis -> key = that coordinate, 16 bits ->
//                  segmented 2-pass radix sort (sort.cu, one job per segment).
//                  Positional halving of a sorted segment IS the median split.
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
              const unsigned* __restrict__ bbox, uint32_t* __restrict__ keys) {
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
  // 16-bit key: the coordinate quantised over the segment's own extent.  Any
  // order yields a valid tree (boxes come from the points); a split that is off
  // by one 1/65536th of the extent costs nothing measurable and halves the
  // number of radix passes per level.
  const float lo = ord2f(bb[axis]), hi = ord2f(bb[3 + axis]);
  const float x = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
  const float scale = hi > lo ? 65535.0f / (hi - lo) : 0.f;
  int q = (int)((x - lo) * scale);
  keys[o] = (uint32_t)max(0, min(65535, q));
}

// ---- local levels: one block owns m0 (<= kLocal) consecutive slots ------------
__global__ void __launch_bounds__(kLocalThreads)
kd_local_kernel(const KdCloud* __restrict__ clouds, uint32_t* __restrict__ vals, int span, int segs, int m0) {
  extern __shared__ unsigned smem[];
  unsigned* sx = smem;                 // ordered-uint coordinates, fixed slots
  unsigned* sy = sx + kLocal;
  unsigned* sz = sy + kLocal;
  unsigned* sid = sz + kLocal;         // original index per slot
  unsigned* key = sid + kLocal;        // sort key per position
  unsigned* bb = key + kLocal;         // 6 x (kLocal / 16) bounding boxes
  unsigned short* perm = reinterpret_cast<unsigned short*>(bb + 6 * (kLocal / 16));  // slot per position
  unsigned char* axis = reinterpret_cast<unsigned char*>(perm + kLocal);

  const int b = blockIdx.y, s = blockIdx.x;
  const KdCloud c = clouds[b];
  const int cnt = seg_count(c.n, s, m0);
  if (cnt == 0) return;
  uint32_t* v = vals + (size_t)b * span + (size_t)s * m0;
  const int tid = threadIdx.x;
  for (int i = tid; i < m0; i += kLocalThreads) {
    if (i < cnt) {
      const uint32_t id = v[i];
      float4 p = c.pts[id];
      sx[i] = f2ord(p.x); sy[i] = f2ord(p.y); sz[i] = f2ord(p.z);
      sid[i] = id;
    } else {
      sx[i] = sy[i] = sz[i] = 0xffffffffu;  // padding sorts last on every axis
      sid[i] = 0xffffffffu;
    }
    perm[i] = (unsigned short)i;
  }
  __syncthreads();
  for (int m = m0; m > kLeaf; m >>= 1) {
    const int nseg = m0 / m;
    for (int q = tid; q < nseg * 6; q += kLocalThreads) bb[q] = (q % 6 < 3) ? 0xffffffffu : 0u;
    __syncthreads();
    // bounding box of every sub-segment (real points only): lanes of a warp (or
    // of a half-warp when m == 16) share a sub-segment, so reduce before the atomics
    for (int i = tid; i < m0; i += kLocalThreads) {
      const int slot = perm[i];
      const bool real = sid[slot] != 0xffffffffu;
      unsigned lo[3] = {real ? sx[slot] : 0xffffffffu, real ? sy[slot] : 0xffffffffu, real ? sz[slot] : 0xffffffffu};
      unsigned hi[3] = {real ? sx[slot] : 0u, real ? sy[slot] : 0u, real ? sz[slot] : 0u};
      const unsigned mask = m >= 32 ? 0xffffffffu : ((tid & 16) ? 0xffff0000u : 0x0000ffffu);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        lo[d] = __reduce_min_sync(mask, lo[d]);
        hi[d] = __reduce_max_sync(mask, hi[d]);
      }
      const bool leader = m >= 32 ? (tid & 31) == 0 : (tid & 15) == 0;
      if (leader && lo[0] <= hi[0]) {
        unsigned* q = bb + 6 * (i / m);
#pragma unroll
        for (int d = 0; d < 3; ++d) { atomicMin(q + d, lo[d]); atomicMax(q + 3 + d, hi[d]); }
      }
    }
    __syncthreads();
    for (int q = tid; q < nseg; q += kLocalThreads) axis[q] = (unsigned char)widest_axis(bb + 6 * q);
    __syncthreads();
    for (int i = tid; i < m0; i += kLocalThreads) {
      const int slot = perm[i];
      const int a = axis[i / m];
      key[i] = a == 0 ? sx[slot] : (a == 1 ? sy[slot] : sz[slot]);
    }
    __syncthreads();
    // bitonic sort of (key, perm) inside every sub-segment of size m, ascending.
    // Pair t of a stage with stride jj is (i, i|jj), i = t with a zero bit inserted
    // at log2(jj): for jj <= 32 the 32 pairs of a warp stay inside the warp's own
    // 64-element window.  A stage with jj >= 32 follows one that wrote across
    // warps (stride 2*jj >= 64) or opens a new merge, so it starts with a block
    // barrier; stages with jj <= 16 follow warp-local stages and only need a
    // warp barrier.
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
  for (int i = tid; i < cnt; i += kLocalThreads) v[i] = sid[perm[i]];
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
  uint32_t* vals_a = d_vals_out;
  kd_iota_kernel<<<dim3(ceil_div(max_n, 256), B), 256, 0, st>>>(clouds.p, vals_a, span);
  ctx_count_launches(ctx, 1);
  int level = 0;
  for (; (span >> level) > kLocal; ++level) {
    const int segs = 1 << level, S = span >> level, jobs = B * segs;
    DBuf<int> job_n(ctx, jobs);
    DBuf<unsigned> bbox(ctx, (size_t)6 * jobs);
    kd_jobs_kernel<<<ceil_div(jobs, 128), 128, 0, st>>>(clouds.p, B, segs, S, job_n.p, bbox.p);
    const int seg_max = std::min(S, max_n);
    kd_bbox_kernel<<<dim3(std::max(1, std::min(ceil_div(seg_max, 2048), 64)), jobs), 256, 0, st>>>(clouds.p, vals_a, span,
                                                                                                 segs, S, bbox.p);
    kd_key_kernel<<<dim3(ceil_div(seg_max, 256), jobs), 256, 0, st>>>(clouds.p, vals_a, span, segs, S, bbox.p, keys_a.p);
    ctx_count_launches(ctx, 3);
    // jobs are laid out back to back with stride S: cloud b, segment s starts at (b*segs + s) * S
    bool in_b = radix_sort_pairs<uint32_t>(ctx, keys_a.p, keys_b.p, vals_a, vals_b.p, job_n.p, jobs, S, seg_max, 16);
    if (in_b) throw Error(PGS_CUDA_ERROR, "kd_order: unexpected sort buffer parity");
  }
  const int m0 = std::min(span, kLocal);
  const int segs = span / m0;
  const size_t smem = (size_t)kLocal * 4 * 5 + 6 * (kLocal / 16) * 4 + (size_t)kLocal * 2 + kLocal / 16;
  PGS_CUDA(cudaFuncSetAttribute(kd_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kd_local_kernel<<<dim3(segs, B), kLocalThreads, smem, st>>>(clouds.p, vals_a, span, segs, m0);
  ctx_count_launches(ctx, 1);
  PGS_LAUNCH_CHECK();
}

}  // namespace pgs
