// index.cu — the spatial index that replaces libnabo's kd-tree
// (KDTreeMatcher::init -> NNS::create).
//
// B200 shape instead of a pointer-built unbalanced kd-tree: points are put in
// balanced kd-tree order on the device (kdorder.cu), cut into leaves of kLeaf
// consecutive points (one 128-byte line of float4 each), and covered by an
// implicit complete binary tree of axis-aligned boxes.  No child pointers: node i has
// children 2i and 2i+1, stored adjacently (48 bytes = three float4 loads), and
// traversal needs no stack (see knn.cu).  Exactness of the search does not
// depend on the tree shape, only on the boxes being conservative, so parity
// with the reference's kd-tree is parity of RESULTS (exact (dist, index)
// minima), not of visit order.
#include "core.cuh"

namespace pgs {

namespace {

struct BuildJob {
  const float4* src;
  float4* dst;    // sorted points (n_leaves * kLeaf entries)
  float* nodes;   // 2*P*6 floats
  float* cells;   // 2*P*6 floats
  int n, n_leaves, P, depth;
  const uint32_t* order;  // order[j] = original index of the j-th point in tree order
};

__global__ void bbox_init_kernel(unsigned* bbox, int n_jobs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_jobs * 6) bbox[i] = (i % 6 < 3) ? 0xffffffffu : 0u;
}

__global__ void __launch_bounds__(256)
bbox_kernel(const BuildJob* __restrict__ jobs, const float* __restrict__ shift, unsigned* __restrict__ bbox) {
  const BuildJob job = jobs[blockIdx.y];
  float sx = 0.f, sy = 0.f, sz = 0.f;
  if (shift) { sx = shift[4 * blockIdx.y]; sy = shift[4 * blockIdx.y + 1]; sz = shift[4 * blockIdx.y + 2]; }
  unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < job.n; i += gridDim.x * blockDim.x) {
    float4 p = job.src[i];
    unsigned ux = f2ord(__fsub_rn(p.x, sx)), uy = f2ord(__fsub_rn(p.y, sy)), uz = f2ord(__fsub_rn(p.z, sz));
    lo[0] = min(lo[0], ux); hi[0] = max(hi[0], ux);
    lo[1] = min(lo[1], uy); hi[1] = max(hi[1], uy);
    lo[2] = min(lo[2], uz); hi[2] = max(hi[2], uz);
  }
  // one set of atomics per block: the boxes of a batch share a few L2 lines, and atomics on one line are
  // served one after the other (per warp they were most of this kernel's time)
  __shared__ unsigned slot[8][6];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    lo[d] = __reduce_min_sync(0xffffffffu, lo[d]);
    hi[d] = __reduce_max_sync(0xffffffffu, hi[d]);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { slot[threadIdx.x >> 5][d] = lo[d]; slot[threadIdx.x >> 5][3 + d] = hi[d]; }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    const bool is_lo = threadIdx.x < 3;
    unsigned v = slot[0][threadIdx.x];
    for (int q = 1; q < 8; ++q) v = is_lo ? min(v, slot[q][threadIdx.x]) : max(v, slot[q][threadIdx.x]);
    unsigned* bb = bbox + 6 * blockIdx.y;
    if (is_lo) atomicMin(bb + threadIdx.x, v);
    else atomicMax(bb + threadIdx.x, v);
  }
}

__device__ __forceinline__ unsigned spread10(unsigned v) {
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

__global__ void __launch_bounds__(256)
morton_kernel(const BuildJob* __restrict__ jobs, const float* __restrict__ shift,
              const unsigned* __restrict__ bbox, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
              int stride) {
  const BuildJob job = jobs[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= job.n) return;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  if (shift) { sx = shift[4 * blockIdx.y]; sy = shift[4 * blockIdx.y + 1]; sz = shift[4 * blockIdx.y + 2]; }
  const unsigned* bb = bbox + 6 * blockIdx.y;
  float lo[3] = {ord2f(bb[0]), ord2f(bb[1]), ord2f(bb[2])};
  float hi[3] = {ord2f(bb[3]), ord2f(bb[4]), ord2f(bb[5])};
  float4 p = job.src[i];
  float c[3] = {__fsub_rn(p.x, sx), __fsub_rn(p.y, sy), __fsub_rn(p.z, sz)};
  // one cubic grid over the longest extent keeps cells isotropic
  float ext = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
  // 8 bits per axis = 24-bit keys = 3 radix passes: the Morton order is only
  // used where spatial COHERENCE is wanted (the reading), never for searching
  float scale = ext > 0.f ? 255.0f / ext : 0.f;
  unsigned q[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float f = (c[d] - lo[d]) * scale;
    int v = (int)f;
    q[d] = (unsigned)max(0, min(255, v));
  }
  keys[(size_t)blockIdx.y * stride + i] = spread10(q[0]) | (spread10(q[1]) << 1) | (spread10(q[2]) << 2);
  vals[(size_t)blockIdx.y * stride + i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256)
gather_sorted_kernel(const BuildJob* __restrict__ jobs, const float* __restrict__ shift) {
  const BuildJob job = jobs[blockIdx.y];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= job.n_leaves * kLeaf) return;
  float4 o;
  if (j < job.n) {
    float sx = 0.f, sy = 0.f, sz = 0.f;
    if (shift) { sx = shift[4 * blockIdx.y]; sy = shift[4 * blockIdx.y + 1]; sz = shift[4 * blockIdx.y + 2]; }
    uint32_t src = job.order[j];
    float4 p = job.src[src];
    o = make_float4(__fsub_rn(p.x, sx), __fsub_rn(p.y, sy), __fsub_rn(p.z, sz), __int_as_float((int)src));
  } else {
    // padding: +inf coordinates give dist = +inf, index INT_MAX never wins a tie
    const float inf = __int_as_float(0x7f800000);
    o = make_float4(inf, inf, inf, __int_as_float(0x7fffffff));
  }
  job.dst[j] = o;
}

// Leaf boxes, and the kLeafLevels tree levels above them: a block owns an aligned run of 256
// leaves, so the subtree over them is reduced in shared memory without leaving the block.
constexpr int kLeafLevels = 8;  // log2 of leaf_box_kernel's block size
__global__ void __launch_bounds__(256) leaf_box_kernel(const BuildJob* __restrict__ jobs) {
  __shared__ float sb[2][256][6];
  const BuildJob job = jobs[blockIdx.y];
  const int first = blockIdx.x * 256;
  if (first >= job.P) return;
  const int leaf = first + threadIdx.x;
  const float inf = __int_as_float(0x7f800000);
  float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
  if (leaf < job.n_leaves) {
    const float4* lp = job.dst + (size_t)leaf * kLeaf;
#pragma unroll
    for (int j = 0; j < kLeaf; ++j) {
      if (leaf * kLeaf + j < job.n) {
        float4 p = lp[j];
        lo[0] = fminf(lo[0], p.x); hi[0] = fmaxf(hi[0], p.x);
        lo[1] = fminf(lo[1], p.y); hi[1] = fmaxf(hi[1], p.y);
        lo[2] = fminf(lo[2], p.z); hi[2] = fmaxf(hi[2], p.z);
      }
    }
  }
  // node layout {lo.x, lo.y, hi.x, hi.y, lo.z, hi.z}: three 8-byte halves that feed the
  // packed fp32x2 box test (box_lb_packed) straight from the load
  float* me = sb[0][threadIdx.x];
  me[0] = lo[0]; me[1] = lo[1]; me[2] = hi[0];
  me[3] = hi[1]; me[4] = lo[2]; me[5] = hi[2];
  if (leaf == 0) {  // node 0 does not exist (1-based heap): an empty box, so that copies of the array read defined values
    float* n0 = job.nodes;
    n0[0] = inf; n0[1] = inf; n0[2] = -inf; n0[3] = -inf; n0[4] = inf; n0[5] = -inf;
  }
  if (leaf < job.P) {
    float* nd = job.nodes + (size_t)(job.P + leaf) * 6;
#pragma unroll
    for (int e = 0; e < 6; ++e) nd[e] = me[e];
  }
  const int span = min(256, job.P);  // leaves of this block
  int cur = 0;
  for (int k = 1; (span >> k) >= 1; ++k) {
    __syncthreads();
    const int t = threadIdx.x;
    if (t < (span >> k)) {
      const float* a = sb[cur][2 * t];
      const float* b = sb[cur][2 * t + 1];
      float* o = sb[cur ^ 1][t];
      o[0] = fminf(a[0], b[0]); o[1] = fminf(a[1], b[1]); o[4] = fminf(a[4], b[4]);
      o[2] = fmaxf(a[2], b[2]); o[3] = fmaxf(a[3], b[3]); o[5] = fmaxf(a[5], b[5]);
      float* nd = job.nodes + (size_t)(((job.P + first) >> k) + t) * 6;
#pragma unroll
      for (int e = 0; e < 6; ++e) nd[e] = o[e];
    }
    cur ^= 1;
  }
}

// the levels above leaf_box_kernel's: one block per job walks them bottom-up
__global__ void __launch_bounds__(64) upper_levels_kernel(const BuildJob* __restrict__ jobs) {
  const BuildJob job = jobs[blockIdx.x];
  for (int width = job.P >> (kLeafLevels + 1); width >= 1; width >>= 1) {
    for (int i = threadIdx.x; i < width; i += blockDim.x) {
      const int node = width + i;
      const float* a = job.nodes + (size_t)(2 * node) * 6;
      float* o = job.nodes + (size_t)node * 6;
      o[0] = fminf(a[0], a[6]); o[1] = fminf(a[1], a[7]); o[4] = fminf(a[4], a[10]);
      o[2] = fmaxf(a[2], a[8]); o[3] = fmaxf(a[3], a[9]); o[5] = fmaxf(a[5], a[11]);
    }
    __syncthreads();
  }
}

// Exclusive cells, top-down.  cell(i) is an axis-aligned region that contains NO point of any
// leaf outside node i's subtree: the root's cell is all of space, and a pair of siblings L, R
// cuts their parent's cell along the axis a on which their point sets are best separated -
// cell(L).hi[a] = min(.., box(R).lo[a]), cell(R).lo[a] = max(.., box(L).hi[a]) (or mirrored).
// Every point of R has x_a >= box(R).lo[a], so none lies strictly inside cell(L), whether or
// not the two boxes overlap (an overlap only makes the cell smaller than the box).  A search
// whose candidate ball lies strictly inside cell(i) can therefore start at node i and never
// has to look above it (match_persistent_kernel in icp.cu): the stored bounds are compared with
// the same fp32 subtraction the distance uses, so the proof survives rounding (monotonicity).
// Same {lo.x, lo.y, hi.x, hi.y, lo.z, hi.z} layout as the boxes.
__global__ void __launch_bounds__(1024) cell_levels_kernel(const BuildJob* __restrict__ jobs) {
  const BuildJob job = jobs[blockIdx.x];
  const float inf = __int_as_float(0x7f800000);
  // position of lo[d] / hi[d] inside a 6-float record
  const int LO[3] = {0, 1, 4}, HI[3] = {2, 3, 5};
  if (threadIdx.x < 6) {
    const bool is_lo = threadIdx.x == 0 || threadIdx.x == 1 || threadIdx.x == 4;
    job.cells[6 + threadIdx.x] = is_lo ? -inf : inf;  // node 1 = root
    job.cells[threadIdx.x] = is_lo ? -inf : inf;      // node 0 unused
  }
  __syncthreads();
  for (int width = 1; width < job.P; width <<= 1) {
    for (int i = threadIdx.x; i < width; i += blockDim.x) {
      const int node = width + i;
      const float* bl = job.nodes + (size_t)(2 * node) * 6;
      const float* br = bl + 6;
      float cl[6], cr[6];
#pragma unroll
      for (int e = 0; e < 6; ++e) cl[e] = cr[e] = job.cells[(size_t)node * 6 + e];
      const bool both = bl[LO[0]] <= bl[HI[0]] && br[LO[0]] <= br[HI[0]];  // an empty box is (+inf, -inf)
      if (both) {
        float best = -inf;
        int axis = 0;
        bool l_below = true;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const float sa = br[LO[d]] - bl[HI[d]];  // L below R along d
          const float sb = bl[LO[d]] - br[HI[d]];  // R below L along d
          if (sa > best) { best = sa; axis = d; l_below = true; }
          if (sb > best) { best = sb; axis = d; l_below = false; }
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          if (d != axis) continue;
          if (l_below) {
            cl[HI[d]] = fminf(cl[HI[d]], br[LO[d]]);
            cr[LO[d]] = fmaxf(cr[LO[d]], bl[HI[d]]);
          } else {
            cl[LO[d]] = fmaxf(cl[LO[d]], br[HI[d]]);
            cr[HI[d]] = fminf(cr[HI[d]], bl[LO[d]]);
          }
        }
      }
      float* ol = job.cells + (size_t)(2 * node) * 6;
#pragma unroll
      for (int e = 0; e < 6; ++e) { ol[e] = cl[e]; ol[6 + e] = cr[e]; }
    }
    __syncthreads();
  }
}

}  // namespace

void kd_order_batched(Ctx* ctx, const std::vector<const float4*>& pts, const std::vector<int>& n, int span,
                      uint32_t* d_vals_out);

// Orders: kd (balanced kd-tree order, kdorder.cu) for anything that is searched;
// Morton (one radix sort) where the order only has to be spatially coherent
// (the reading cloud: adjacent lanes should walk adjacent nodes).
void build_indices(Ctx* ctx, const std::vector<const float4*>& d_pts, const std::vector<int>& n,
                   const float* d_shift, std::vector<std::unique_ptr<Index>>& out, IndexOrder order_kind,
                   const std::vector<const uint32_t*>* given_order, std::vector<DBuf<uint32_t>>* keep_order) {
  const bool want_boxes = order_kind == IndexOrder::Kd;  // a Morton-ordered cloud is never searched
  const int B = (int)d_pts.size();
  out.clear();
  if (B == 0) return;
  int max_n = 0, max_P = 0, max_leaves = 0;
  std::vector<BuildJob> jobs(B);
  for (int b = 0; b < B; ++b) {
    auto idx = std::make_unique<Index>();
    idx->ctx = ctx;
    idx->n = n[b];
    idx->n_leaves = ceil_div(n[b] > 0 ? n[b] : 1, kLeaf);
    idx->P = next_pow2(idx->n_leaves);
    idx->depth = 0;
    while ((1 << idx->depth) < idx->P) ++idx->depth;
    idx->pts.reset(ctx, (size_t)idx->n_leaves * kLeaf);
    idx->nodes.reset(ctx, (size_t)2 * idx->P * 6);
    // exclusive cells only for the matcher variants that use them (measured: no faster)
    const bool want_cells = want_boxes && (ctx->tune.match_mode == 1 || ctx->tune.match_mode == 3);
    if (want_cells) idx->cells.reset(ctx, (size_t)2 * idx->P * 6);
    jobs[b] = BuildJob{d_pts[b], idx->pts.p, idx->nodes.p, idx->cells.p, idx->n, idx->n_leaves, idx->P, idx->depth, nullptr};
    max_n = std::max(max_n, n[b]);
    max_P = std::max(max_P, idx->P);
    max_leaves = std::max(max_leaves, idx->n_leaves);
    out.push_back(std::move(idx));
  }
  cudaStream_t s = ctx->stream;
  // ---- the order ------------------------------------------------------------
  DBuf<uint32_t> ka, kb, va, vb;  // kept alive until the gather below has been enqueued (stream-ordered frees)
  DBuf<unsigned> bbox;
  DBuf<int> d_n;
  if (given_order) {
    for (int b = 0; b < B; ++b) jobs[b].order = (*given_order)[b];
  } else if (order_kind == IndexOrder::Kd) {
    const int span = max_P * kLeaf;
    va.reset(ctx, (size_t)B * span);
    kd_order_batched(ctx, d_pts, n, span, va.p);
    for (int b = 0; b < B; ++b) jobs[b].order = va.p + (size_t)b * span;
  } else {
    const int stride = ceil_div(max_n > 0 ? max_n : 1, kSortChunk) * kSortChunk;
    DBuf<BuildJob> d_jobs0(ctx, B);
    ctx->upload_small(d_jobs0.p, jobs.data(), sizeof(BuildJob) * B);
    d_n.reset(ctx, B);
    ctx->upload_small(d_n.p, n.data(), sizeof(int) * B);
    bbox.reset(ctx, (size_t)6 * B);
    ka.reset(ctx, (size_t)B * stride); kb.reset(ctx, (size_t)B * stride);
    va.reset(ctx, (size_t)B * stride); vb.reset(ctx, (size_t)B * stride);
    bbox_init_kernel<<<ceil_div(6 * B, 256), 256, 0, s>>>(bbox.p, B);
    if (max_n > 0) {
      int nb = std::min(ceil_div(max_n, 256 * 4), 64);
      bbox_kernel<<<dim3(nb, B), 256, 0, s>>>(d_jobs0.p, d_shift, bbox.p);
      morton_kernel<<<dim3(ceil_div(max_n, 256), B), 256, 0, s>>>(d_jobs0.p, d_shift, bbox.p, ka.p, va.p, stride);
      ctx_count_launches(ctx, 3);
    }
    bool in_b = radix_sort_pairs<uint32_t>(ctx, ka.p, kb.p, va.p, vb.p, d_n.p, B, stride, max_n, 24);
    for (int b = 0; b < B; ++b) jobs[b].order = (in_b ? vb.p : va.p) + (size_t)b * stride;
  }
  if (keep_order) {
    keep_order->clear();
    for (int b = 0; b < B; ++b) {
      keep_order->emplace_back(ctx, (size_t)std::max(n[b], 1));
      if (n[b] > 0)
        PGS_CUDA(cudaMemcpyAsync(keep_order->back().p, jobs[b].order, (size_t)n[b] * sizeof(uint32_t),
                                 cudaMemcpyDeviceToDevice, s));
    }
  }
  // ---- sorted points, leaf boxes, upper levels ------------------------------------
  DBuf<BuildJob> d_jobs(ctx, B);
  ctx->upload_small(d_jobs.p, jobs.data(), sizeof(BuildJob) * B);
  gather_sorted_kernel<<<dim3(ceil_div(max_leaves * kLeaf, 256), B), 256, 0, s>>>(d_jobs.p, d_shift);
  ctx_count_launches(ctx, 1);
  if (want_boxes) {
    leaf_box_kernel<<<dim3(ceil_div(max_P, 256), B), 256, 0, s>>>(d_jobs.p);
    ctx_count_launches(ctx, 1);
    if (max_P > (1 << kLeafLevels)) {
      upper_levels_kernel<<<B, 64, 0, s>>>(d_jobs.p);
      ctx_count_launches(ctx, 1);
    }
    if (out[0]->cells.p) {
      cell_levels_kernel<<<B, 1024, 0, s>>>(d_jobs.p);
      ctx_count_launches(ctx, 1);
    }
  }
  PGS_LAUNCH_CHECK();
}

namespace {
struct ShiftJob {
  const float4* src_pts;
  const float* src_nodes;
  const float* src_cells;
  float4* dst_pts;
  float* dst_nodes;
  float* dst_cells;
  int n_pts;    // n_leaves * kLeaf
  int n_nodes;  // 2 * P
};

// fl(x - m) is monotone in x, so the box of the shifted points is exactly the
// shifted box and the kd order is unchanged: a mean-centred copy of an index is
// one streaming pass, not a rebuild.
__global__ void __launch_bounds__(256)
shift_index_kernel(const ShiftJob* __restrict__ jobs, const float* __restrict__ shift) {
  const ShiftJob job = jobs[blockIdx.y];
  const float sx = shift[4 * blockIdx.y], sy = shift[4 * blockIdx.y + 1], sz = shift[4 * blockIdx.y + 2];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < job.n_pts) {
    float4 p = job.src_pts[i];
    // padding (+inf) stays +inf
    job.dst_pts[i] = make_float4(__fsub_rn(p.x, sx), __fsub_rn(p.y, sy), __fsub_rn(p.z, sz), p.w);
  }
  if (i < job.n_nodes) {
    const float* a = job.src_nodes + (size_t)i * 6;
    float* o = job.dst_nodes + (size_t)i * 6;
    // empty boxes (+inf / -inf) stay empty
    o[0] = __fsub_rn(a[0], sx); o[1] = __fsub_rn(a[1], sy); o[2] = __fsub_rn(a[2], sx);
    o[3] = __fsub_rn(a[3], sy); o[4] = __fsub_rn(a[4], sz); o[5] = __fsub_rn(a[5], sz);
    // the cells' faces are point coordinates (or +-inf): the same monotone shift keeps every
    // outside point outside
    if (!job.src_cells) return;
    const float* ca = job.src_cells + (size_t)i * 6;
    float* co = job.dst_cells + (size_t)i * 6;
    co[0] = __fsub_rn(ca[0], sx); co[1] = __fsub_rn(ca[1], sy); co[2] = __fsub_rn(ca[2], sx);
    co[3] = __fsub_rn(ca[3], sy); co[4] = __fsub_rn(ca[4], sz); co[5] = __fsub_rn(ca[5], sz);
  }
}
}  // namespace

void derive_shifted_indices(Ctx* ctx, const std::vector<const Index*>& src, const float* d_shift,
                            std::vector<std::unique_ptr<Index>>& out) {
  const int B = (int)src.size();
  out.clear();
  std::vector<ShiftJob> jobs(B);
  int max_items = 1;
  for (int b = 0; b < B; ++b) {
    auto idx = std::make_unique<Index>();
    idx->ctx = ctx;
    idx->n = src[b]->n; idx->n_leaves = src[b]->n_leaves; idx->P = src[b]->P; idx->depth = src[b]->depth;
    idx->pts.reset(ctx, (size_t)idx->n_leaves * kLeaf);
    idx->nodes.reset(ctx, (size_t)2 * idx->P * 6);
    if (src[b]->cells.p) idx->cells.reset(ctx, (size_t)2 * idx->P * 6);
    jobs[b] = ShiftJob{src[b]->pts.p, src[b]->nodes.p, src[b]->cells.p, idx->pts.p, idx->nodes.p, idx->cells.p,
                       idx->n_leaves * kLeaf, 2 * idx->P};
    max_items = std::max(max_items, std::max(jobs[b].n_pts, jobs[b].n_nodes));
    out.push_back(std::move(idx));
  }
  if (B == 0) return;
  DBuf<ShiftJob> d_jobs(ctx, B);
  ctx->upload_small(d_jobs.p, jobs.data(), sizeof(ShiftJob) * B);
  shift_index_kernel<<<dim3(ceil_div(max_items, 256), B), 256, 0, ctx->stream>>>(d_jobs.p, d_shift);
  ctx_count_launches(ctx, 1);
  PGS_LAUNCH_CHECK();
}

// Index build for clouds, sharing the kd order between builds of the same
// (unchanged) cloud: the SurfaceNormal filter and the matcher both need one.
void build_indices_for_clouds(Ctx* ctx, const std::vector<Cloud*>& clouds, const float* d_shift,
                              std::vector<std::unique_ptr<Index>>& out) {
  const int B = (int)clouds.size();
  std::vector<const float4*> pts(B);
  std::vector<int> n(B);
  bool all_cached = B > 0;
  for (int b = 0; b < B; ++b) {
    pts[b] = clouds[b]->feat.p;
    n[b] = (int)clouds[b]->n;
    all_cached = all_cached && clouds[b]->kd_order && clouds[b]->kd_order_n == clouds[b]->n;
  }
  if (all_cached) {
    std::vector<const uint32_t*> given(B);
    for (int b = 0; b < B; ++b) given[b] = clouds[b]->kd_order->p;
    build_indices(ctx, pts, n, d_shift, out, IndexOrder::Kd, &given, nullptr);
    return;
  }
  std::vector<DBuf<uint32_t>> keep;
  build_indices(ctx, pts, n, d_shift, out, IndexOrder::Kd, nullptr, &keep);
  for (int b = 0; b < B; ++b) {
    clouds[b]->kd_order = std::make_shared<DBuf<uint32_t>>(std::move(keep[b]));
    clouds[b]->kd_order_n = clouds[b]->n;
  }
}

// Unshifted index of every cloud, cached in the cloud (shared with its clones)
// until its points change.
void cached_indices_for_clouds(Ctx* ctx, const std::vector<Cloud*>& clouds, std::vector<std::shared_ptr<Index>>& out) {
  const int B = (int)clouds.size();
  out.assign(B, nullptr);
  std::vector<Cloud*> todo;
  for (int b = 0; b < B; ++b)
    if (!(clouds[b]->index_cache && clouds[b]->index_cache->n == clouds[b]->n && clouds[b]->kd_order)) todo.push_back(clouds[b]);
  if (!todo.empty()) {
    std::vector<std::unique_ptr<Index>> built;
    build_indices_for_clouds(ctx, todo, nullptr, built);
    for (size_t t = 0; t < todo.size(); ++t) todo[t]->index_cache = std::shared_ptr<Index>(std::move(built[t]));
  }
  for (int b = 0; b < B; ++b) out[b] = clouds[b]->index_cache;
}

}  // namespace pgs
