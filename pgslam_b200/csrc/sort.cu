// sort.cu — batched LSD radix sort and scan/compaction primitives.
//
// The sort orders points by Morton code for the spatial index (and by voxel id
// for VoxelGridDataPointsFilter).  It is stable, so equal keys keep input
// order — VoxelGrid's "sum in input order" semantics
// depend on that.
//
// Shape: 8 bits per pass, three kernels per pass (count, scan, scatter).  A
// block owns a tile of 8 warp-chunks (32*kSortRounds consecutive keys each); a
// warp ranks its chunk with __match_any_sync (round-major, lane-minor == input
// order, so ranks are stable by construction) and the block only meets twice,
// to turn per-warp digit counts into offsets.  The scanned table is
// 256 x tiles ints per job (a few thousand), not one row per warp.
#include "core.cuh"

namespace pgs {

namespace {

constexpr int kWarpsPerBlock = 8;

// digit histogram of one warp-chunk into cnt[256] (warp-private shared memory)
template <typename K>
__device__ __forceinline__ void warp_chunk_count(const K* __restrict__ kp, int base, int n, int shift, int lane,
                                                 int* __restrict__ cnt, K (&kreg)[kSortRounds]) {
  for (int d = lane; d < 256; d += 32) cnt[d] = 0;
  __syncwarp();
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    int i = base + r * 32 + lane;
    kreg[r] = (i < n) ? kp[r * 32 + lane] : K(0);
  }
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    int i = base + r * 32 + lane;
    bool valid = i < n;
    unsigned act = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      int d = (int)((kreg[r] >> shift) & 255);
      unsigned m = __match_any_sync(act, d);
      if (lane == __ffs(m) - 1) cnt[d] += __popc(m);
    }
    __syncwarp();
  }
}

// one block = one tile of kWarpsPerBlock warp-chunks; counts[job][digit][tile]
template <typename K>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
sort_count_kernel(const K* __restrict__ keys, const int* __restrict__ ns, int stride, int ntiles,
                  int shift, int* __restrict__ counts) {
  __shared__ int cnt[kWarpsPerBlock][256];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, tile = blockIdx.x;
  const int n = ns[b];
  const int base = (tile * kWarpsPerBlock + w) * kSortChunk;
  K kreg[kSortRounds];
  warp_chunk_count<K>(keys + (size_t)b * stride + base, base, n, shift, lane, cnt[w], kreg);
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int j = 0; j < kWarpsPerBlock; ++j) t += cnt[j][threadIdx.x];
  counts[((size_t)b * 256 + threadIdx.x) * ntiles + tile] = t;
}

// in-place exclusive scan of `total` ints per job; one block per job.
__global__ void __launch_bounds__(1024) sort_scan_kernel(int* __restrict__ counts, int total) {
  __shared__ int warp_tot[32];
  int* a = counts + (size_t)blockIdx.x * total;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int per = (total + 1023) / 1024;
  const int start = min(tid * per, total), end = min(start + per, total);
  int sum = 0;
  for (int i = start; i < end; ++i) sum += a[i];
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_tot[w] = incl;
  __syncthreads();
  if (w == 0) {
    int v = warp_tot[lane];
    int iv = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, iv, o);
      if (lane >= o) iv += u;
    }
    warp_tot[lane] = iv - v;
  }
  __syncthreads();
  int run = warp_tot[w] + incl - sum;
  for (int i = start; i < end; ++i) {
    int t = a[i];
    a[i] = run;
    run += t;
  }
}

template <typename K>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
sort_scatter_kernel(const K* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                    K* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                    const int* __restrict__ ns, int stride, int ntiles, int shift,
                    const int* __restrict__ offsets) {
  __shared__ int off[kWarpsPerBlock][256];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, tile = blockIdx.x;
  const int n = ns[b];
  if (tile * kWarpsPerBlock * kSortChunk >= n) return;
  const int base = (tile * kWarpsPerBlock + w) * kSortChunk;
  const K* kp = keys_in + (size_t)b * stride + base;
  const uint32_t* vp = vals_in + (size_t)b * stride + base;
  K kreg[kSortRounds];
  warp_chunk_count<K>(kp, base, n, shift, lane, off[w], kreg);
  __syncthreads();
  {
    // digit d: tile base from the scan, then an exclusive prefix over the warps
    int run = offsets[((size_t)b * 256 + threadIdx.x) * ntiles + tile];
#pragma unroll
    for (int j = 0; j < kWarpsPerBlock; ++j) {
      int t = off[j][threadIdx.x];
      off[j][threadIdx.x] = run;
      run += t;
    }
  }
  __syncthreads();
  K* ko = keys_out + (size_t)b * stride;
  uint32_t* vo = vals_out + (size_t)b * stride;
  uint32_t vreg[kSortRounds];
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    int i = base + r * 32 + lane;
    vreg[r] = (i < n) ? vp[r * 32 + lane] : 0u;
  }
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    int i = base + r * 32 + lane;
    bool valid = i < n;
    unsigned act = __ballot_sync(0xffffffffu, valid);
    int d = 0, pos = 0;
    unsigned m = 0;
    if (valid) {
      d = (int)((kreg[r] >> shift) & 255);
      m = __match_any_sync(act, d);
      pos = off[w][d] + __popc(m & ((1u << lane) - 1u));
    }
    __syncwarp();
    if (valid && lane == __ffs(m) - 1) off[w][d] += __popc(m);
    __syncwarp();
    if (valid) {
      ko[pos] = kreg[r];
      vo[pos] = vreg[r];
    }
  }
}

// ---- generic int exclusive scan (compaction) ------------------------------
constexpr int kScanTile = 2048;  // 256 threads x 8

__global__ void __launch_bounds__(256) scan_tile_kernel(const int* __restrict__ in, int* __restrict__ out,
                                                        int n, int* __restrict__ tile_tot) {
  __shared__ int warp_tot[8];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int base = blockIdx.x * kScanTile + tid * 8;
  int v[8];
  int sum = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[j] = (base + j < n) ? in[base + j] : 0;
    sum += v[j];
  }
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) warp_tot[w] = incl;
  __syncthreads();
  int woff = 0;
  for (int j = 0; j < w; ++j) woff += warp_tot[j];
  int run = woff + incl - sum;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (base + j < n) out[base + j] = run;
    run += v[j];
  }
  if (tid == 255) tile_tot[blockIdx.x] = run;
}

__global__ void __launch_bounds__(1024) scan_totals_kernel(int* __restrict__ tile_tot, int ntiles,
                                                            int* __restrict__ total) {
  // single block, sequential chunks of 1024 with a running carry
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < ntiles; base += 1024) {
    int i = base + tid;
    int v = (i < ntiles) ? tile_tot[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    if (w == 0) {
      int t = warp_tot[lane];
      int it = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, it, o);
        if (lane >= o) it += u;
      }
      warp_tot[lane] = it - t;
    }
    __syncthreads();
    int carry = carry_s;
    int excl = carry + warp_tot[w] + incl - v;
    if (i < ntiles) tile_tot[i] = excl;
    __syncthreads();
    if (tid == 1023) carry_s = excl + v;
    __syncthreads();
  }
  if (tid == 0) *total = carry_s;
}

__global__ void __launch_bounds__(256) scan_add_kernel(int* __restrict__ out, int n,
                                                       const int* __restrict__ tile_off) {
  const int base = blockIdx.x * kScanTile + threadIdx.x * 8;
  const int o = tile_off[blockIdx.x];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (base + j < n) out[base + j] += o;
}

__global__ void compact_indices_kernel(const int* __restrict__ keep, const int* __restrict__ pos, int n,
                                       int* __restrict__ src) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && keep[i]) src[pos[i]] = i;
}

__global__ void gather_rows_kernel(const float* __restrict__ in, float* __restrict__ out,
                                   const int* __restrict__ src, int64_t m, int span) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * span) return;
  int64_t j = t / span;
  int d = (int)(t - j * span);
  out[t] = in[(int64_t)src[j] * span + d];
}

}  // namespace

template <typename K>
bool radix_sort_pairs(Ctx* ctx, K* keys_a, K* keys_b, uint32_t* vals_a, uint32_t* vals_b,
                      const int* d_n, int n_jobs, int stride, int max_n, int key_bits) {
  if (n_jobs == 0 || max_n == 0) return false;
  const int ntiles = ceil_div(max_n, kSortChunk * kWarpsPerBlock);
  const int passes = (key_bits + 7) / 8;
  DBuf<int> counts(ctx, (size_t)n_jobs * 256 * ntiles);
  dim3 grid(ntiles, n_jobs);
  K* ki = keys_a;
  K* ko = keys_b;
  uint32_t* vi = vals_a;
  uint32_t* vo = vals_b;
  for (int p = 0; p < passes; ++p) {
    sort_count_kernel<K><<<grid, kWarpsPerBlock * 32, 0, ctx->stream>>>(ki, d_n, stride, ntiles, 8 * p, counts.p);
    sort_scan_kernel<<<n_jobs, 1024, 0, ctx->stream>>>(counts.p, 256 * ntiles);
    sort_scatter_kernel<K><<<grid, kWarpsPerBlock * 32, 0, ctx->stream>>>(ki, vi, ko, vo, d_n, stride, ntiles,
                                                                         8 * p, counts.p);
    ctx_count_launches(ctx, 3);
    std::swap(ki, ko);
    std::swap(vi, vo);
  }
  PGS_LAUNCH_CHECK();
  return (passes & 1) != 0;
}

template bool radix_sort_pairs<uint32_t>(Ctx*, uint32_t*, uint32_t*, uint32_t*, uint32_t*, const int*, int, int, int, int);
template bool radix_sort_pairs<uint64_t>(Ctx*, uint64_t*, uint64_t*, uint32_t*, uint32_t*, const int*, int, int, int, int);

void exclusive_scan_int(Ctx* ctx, const int* d_in, int* d_out, int n, int* d_total) {
  const int ntiles = ceil_div(n > 0 ? n : 1, kScanTile);
  DBuf<int> tile_tot(ctx, ntiles);
  scan_tile_kernel<<<ntiles, 256, 0, ctx->stream>>>(d_in, d_out, n, tile_tot.p);
  scan_totals_kernel<<<1, 1024, 0, ctx->stream>>>(tile_tot.p, ntiles, d_total);
  scan_add_kernel<<<ntiles, 256, 0, ctx->stream>>>(d_out, n, tile_tot.p);
  ctx_count_launches(ctx, 3);
  PGS_LAUNCH_CHECK();
}

void gather_cloud(Cloud& c, const int* d_src, int64_t m) {
  Ctx* ctx = c.ctx;
  c.touch();
  {
    DBuf<float4> nf(ctx, (size_t)m);
    if (m > 0) {
      gather_rows_kernel<<<ceil_div(m * 4, 256), 256, 0, ctx->stream>>>(
          reinterpret_cast<const float*>(c.feat.p), reinterpret_cast<float*>(nf.p), d_src, m, 4);
      ctx_count_launches(ctx, 1);
    }
    c.feat = std::move(nf);
  }
  for (auto& d : c.descs) {
    DBuf<float> nd(ctx, (size_t)m * d.span);
    if (m > 0) {
      gather_rows_kernel<<<ceil_div(m * d.span, 256), 256, 0, ctx->stream>>>(d.data.p, nd.p, d_src, m, d.span);
      ctx_count_launches(ctx, 1);
    }
    d.data = std::move(nd);
  }
  c.n = m;
  PGS_LAUNCH_CHECK();
}

int64_t compact_cloud(Cloud& c, const int* d_keep) {
  Ctx* ctx = c.ctx;
  const int n = (int)c.n;
  if (n == 0) return 0;
  DBuf<int> pos(ctx, n), total(ctx, 1), src(ctx, n);
  exclusive_scan_int(ctx, d_keep, pos.p, n, total.p);
  compact_indices_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(d_keep, pos.p, n, src.p);
  ctx_count_launches(ctx, 1);
  int m = 0;
  total.download(&m, 1);
  ctx->sync();
  gather_cloud(c, src.p, m);
  return m;
}

}  // namespace pgs
