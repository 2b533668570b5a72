// filters.cu — DataPointsFilters and RigidTransformation on the device
// (reached from Localizer.hpp:103,106,
// 314-326, LocalMap.hpp:97,222 and the reference/reading filter lists of every
// ICP call).  Every filter takes a LIST of clouds so a batch of independent
// registrations shares launches.
#include "filters.cuh"

#include <cmath>

#include "knn.cuh"
#include "solve.cuh"

namespace pgs {

namespace {

// ---------------------------------------------------------------------------
// RigidTransformation::compute
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rigid_kernel(float4* __restrict__ feat, float* __restrict__ normals, float* __restrict__ obs, int n, Xf T) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = feat[i];
  float3 o = xform_rn(T, p.x, p.y, p.z);
  feat[i] = make_float4(o.x, o.y, o.z, p.w);
  if (normals) {
    float3 v = rot_rn(T, normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
    normals[3 * i] = v.x; normals[3 * i + 1] = v.y; normals[3 * i + 2] = v.z;
  }
  if (obs) {
    float3 v = rot_rn(T, obs[3 * i], obs[3 * i + 1], obs[3 * i + 2]);
    obs[3 * i] = v.x; obs[3 * i + 1] = v.y; obs[3 * i + 2] = v.z;
  }
}

// ---------------------------------------------------------------------------
// element-wise filters
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
obsdir_kernel(const float4* __restrict__ feat, float* __restrict__ obs, int n, float sx, float sy, float sz) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = feat[i];
  obs[3 * i] = __fsub_rn(sx, p.x);
  obs[3 * i + 1] = __fsub_rn(sy, p.y);
  obs[3 * i + 2] = __fsub_rn(sz, p.z);
}

__global__ void __launch_bounds__(256)
orient_kernel(float* __restrict__ normals, const float* __restrict__ obs, int n, int toward) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float nx = normals[3 * i], ny = normals[3 * i + 1], nz = normals[3 * i + 2];
  float s = __fmul_rn(obs[3 * i], nx);
  s = __fadd_rn(s, __fmul_rn(obs[3 * i + 1], ny));
  s = __fadd_rn(s, __fmul_rn(obs[3 * i + 2], nz));
  bool flip = toward ? (s < 0.f) : (s > 0.f);
  if (flip) { normals[3 * i] = -nx; normals[3 * i + 1] = -ny; normals[3 * i + 2] = -nz; }
}

__global__ void __launch_bounds__(256)
noise_kernel(const float4* __restrict__ feat, float* __restrict__ noise, int n, int kinect, float min_r,
             float beam_angle, float beam_const, float gain) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = feat[i];
  float r2 = __fmul_rn(p.x, p.x);
  r2 = __fadd_rn(r2, __fmul_rn(p.y, p.y));
  r2 = __fadd_rn(r2, __fmul_rn(p.z, p.z));
  float v;
  if (kinect) {
    v = __fmul_rn(0.5f, 0.00285f);
    v = __fmul_rn(v, r2);
  } else {
    float r = __fsqrt_rn(r2);
    v = __fadd_rn(__fmul_rn(beam_angle, r), beam_const);
    if (v < min_r) v = min_r;
  }
  noise[i] = __fmul_rn(gain, v);
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256)
random_keep_kernel(int* __restrict__ keep, int n, uint64_t seed, float prob) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t h = splitmix64(splitmix64(seed) ^ ((uint64_t)i * 0xD1B54A32D192ED03ull));
  float r = __fmul_rn((float)(h >> 40), 1.0f / 16777216.0f);
  keep[i] = r < prob;
}

// MaxDensityDataPointsFilter: maximum density and how many points sit on it, then the
// keep decision (same counter-based uniform as RandomSampling)
__global__ void __launch_bounds__(256) density_max_kernel(const float* __restrict__ dens, int n, unsigned* __restrict__ out) {
  unsigned m = 0u;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, f2ord(dens[i]));
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}
__global__ void __launch_bounds__(256) density_count_kernel(const float* __restrict__ dens, int n, unsigned* __restrict__ io) {
  const float last = ord2f(io[0]);
  unsigned c = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) c += dens[i] == last;
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(io + 1, c);
}
__global__ void __launch_bounds__(256)
density_keep_kernel(const float* __restrict__ dens, int* __restrict__ keep, int n, float max_density, uint64_t seed,
                    const unsigned* __restrict__ stats) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float last = ord2f(stats[0]);
  const float sat_factor = (float)(1 - (int)(stats[1] / (unsigned)n));
  const float density = dens[i];
  bool k = true;
  if (density > max_density) {
    uint64_t h = splitmix64(splitmix64(seed) ^ ((uint64_t)i * 0xD1B54A32D192ED03ull));
    float r = __fmul_rn((float)(h >> 40), 1.0f / 16777216.0f);
    float accept = __fdiv_rn(max_density, density);
    if (density == last) accept = __fmul_rn(accept, sat_factor);
    k = r < accept;
  }
  keep[i] = k;
}

// RemoveNaN / FixStepSampling / Shadow keep flags (libpointmatcher DataPointsFilters/RemoveNaN.cpp,
// FixStepSampling.cpp, Shadow.cpp; fp32 in the contract's x,y,z order, no FMA)
__global__ void __launch_bounds__(256) nan_keep_kernel(const float4* __restrict__ feat, int* __restrict__ keep, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = feat[i];
  keep[i] = !(isnan(p.x) || isnan(p.y) || isnan(p.z));
}
__global__ void __launch_bounds__(256) step_keep_kernel(int* __restrict__ keep, int n, int step, int phase) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keep[i] = i >= phase && (i - phase) % step == 0;
}
__device__ __forceinline__ float3 normalized_rn(float x, float y, float z) {
  float s = __fmul_rn(x, x);
  s = __fadd_rn(s, __fmul_rn(y, y));
  s = __fadd_rn(s, __fmul_rn(z, z));
  if (s > 0.f) {
    const float len = __fsqrt_rn(s);
    return make_float3(__fdiv_rn(x, len), __fdiv_rn(y, len), __fdiv_rn(z, len));
  }
  return make_float3(x, y, z);
}
__global__ void __launch_bounds__(256)
shadow_keep_kernel(const float4* __restrict__ feat, const float* __restrict__ normals, int* __restrict__ keep, int n,
                   float eps) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = feat[i];
  const float3 a = normalized_rn(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
  const float3 b = normalized_rn(p.x, p.y, p.z);
  float d = __fmul_rn(a.x, b.x);
  d = __fadd_rn(d, __fmul_rn(a.y, b.y));
  d = __fadd_rn(d, __fmul_rn(a.z, b.z));
  keep[i] = fabsf(d) > eps;
}

__global__ void __launch_bounds__(256)
dist_keep_kernel(const float4* __restrict__ feat, int* __restrict__ keep, int n, int dim, float lim, int is_max) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = feat[i];
  float v, l;
  if (dim < 0) {
    v = __fmul_rn(p.x, p.x);
    v = __fadd_rn(v, __fmul_rn(p.y, p.y));
    v = __fadd_rn(v, __fmul_rn(p.z, p.z));
    l = __fmul_rn(lim, lim);
  } else {
    v = dim == 0 ? p.x : (dim == 1 ? p.y : p.z);
    l = lim;
  }
  keep[i] = is_max ? (v < l) : (v > l);
}

__global__ void __launch_bounds__(256)
box_keep_kernel(const float4* __restrict__ feat, int* __restrict__ keep, int n, float x0, float x1, float y0,
                float y1, float z0, float z1, int remove_inside) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = feat[i];
  bool in = p.x > x0 && p.x < x1 && p.y > y0 && p.y < y1 && p.z > z0 && p.z < z1;
  keep[i] = remove_inside ? !in : in;
}

// ---------------------------------------------------------------------------
// SurfaceNormalDataPointsFilter: per-point covariance of the kNN set, 3x3
// eigen-decomposition (fp64 cyclic Jacobi), smallest-eigenvalue eigenvector.
// ---------------------------------------------------------------------------
struct NormalJob {
  const float4* feat;
  const int32_t* ids;  // k x n
  int n;
  float* normals;  // 3n or null
  float* dens;     // n or null
  float* eigval;   // 3n or null
  float* eigvec;   // 9n or null
  float* matched;  // k*n or null ("matchedIds": the neighbour ids as floats)
  float* meandist; // n or null   ("meanDists": |point - mean of its neighbours|)
  int sort_eigen;  // eigenvalues ascending, eigenvectors permuted with them
};

__global__ void __launch_bounds__(128)
normals_kernel(const NormalJob* __restrict__ jobs, int k) {
  const NormalJob job = jobs[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= job.n) return;
  const int32_t* nb = job.ids + (size_t)i * k;
  if (job.matched)
    for (int j = 0; j < k; ++j) job.matched[(size_t)i * k + j] = (float)nb[j];
  double mean[3] = {0.0, 0.0, 0.0};
  int real = 0;
  for (int j = 0; j < k; ++j) {
    int id = nb[j];
    if (id < 0) continue;
    float4 p = job.feat[id];
    mean[0] += (double)p.x; mean[1] += (double)p.y; mean[2] += (double)p.z;
    ++real;
  }
  if (real == 0) return;
  mean[0] = mean[0] / (double)real; mean[1] = mean[1] / (double)real; mean[2] = mean[2] / (double)real;
  double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double maxr2 = 0.0;
  for (int j = 0; j < k; ++j) {
    int id = nb[j];
    if (id < 0) continue;
    float4 p = job.feat[id];
    double dx = (double)p.x - mean[0], dy = (double)p.y - mean[1], dz = (double)p.z - mean[2];
    C[0] += dx * dx; C[1] += dx * dy; C[2] += dx * dz;
    C[4] += dy * dy; C[5] += dy * dz; C[8] += dz * dz;
    double r2 = dx * dx + dy * dy + dz * dz;
    if (r2 > maxr2) maxr2 = r2;
  }
  C[0] = C[0] / (double)real; C[1] = C[1] / (double)real; C[2] = C[2] / (double)real;
  C[4] = C[4] / (double)real; C[5] = C[5] / (double)real; C[8] = C[8] / (double)real;
  C[3] = C[1]; C[6] = C[2]; C[7] = C[5];
  if (job.meandist) {
    const float4 me = job.feat[i];
    const double ex = (double)me.x - mean[0], ey = (double)me.y - mean[1], ez = (double)me.z - mean[2];
    job.meandist[i] = (float)sqrt(ex * ex + ey * ey + ez * ez);
  }
  double w[3], V[9];
  jacobi_sym<3>(C, w, V);
  if (job.sort_eigen) {
    int ord[3] = {0, 1, 2};
    for (int a = 1; a < 3; ++a)
      for (int b = a; b > 0 && w[ord[b]] < w[ord[b - 1]]; --b) {
        const int t = ord[b]; ord[b] = ord[b - 1]; ord[b - 1] = t;
      }
    double w2[3], V2[9];
    for (int e = 0; e < 3; ++e) {
      w2[e] = w[ord[e]];
      for (int d = 0; d < 3; ++d) V2[e * 3 + d] = V[ord[e] * 3 + d];
    }
    for (int e = 0; e < 3; ++e) w[e] = w2[e];
    for (int e = 0; e < 9; ++e) V[e] = V2[e];
  }
  double wmax = w[0] > w[1] ? w[0] : w[1];
  if (w[2] > wmax) wmax = w[2];
  int rank = 0;
  for (int e = 0; e < 3; ++e)
    if (w[e] > 3.0 * 1.1920928955078125e-07 * wmax) ++rank;
  if (job.dens) {
    double r = sqrt(maxr2);
    double vol = (4.0 / 3.0) * 3.14159265358979323846 * (r * r * r);
    job.dens[i] = (float)((double)real / vol);
  }
  if (rank >= 2) {
    int e0 = 0;
    for (int e = 1; e < 3; ++e)
      if (w[e] < w[e0]) e0 = e;
    if (job.normals)
      for (int d = 0; d < 3; ++d) {
        float v = (float)V[e0 * 3 + d];
        job.normals[3 * i + d] = v < -1.f ? -1.f : (v > 1.f ? 1.f : v);
      }
    if (job.eigval)
      for (int e = 0; e < 3; ++e) job.eigval[3 * i + e] = (float)w[e];
    if (job.eigvec)
      for (int e = 0; e < 9; ++e) job.eigvec[9 * i + e] = (float)V[e];
  }
}

// ---------------------------------------------------------------------------
// VoxelGridDataPointsFilter
// ---------------------------------------------------------------------------
__global__ void minmax_init_kernel(unsigned* bb) {
  if (threadIdx.x < 6) bb[threadIdx.x] = threadIdx.x < 3 ? 0xffffffffu : 0u;
}

__global__ void __launch_bounds__(256) minmax_kernel(const float4* __restrict__ feat, int n, unsigned* __restrict__ bb) {
  unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = feat[i];
    unsigned u[3] = {f2ord(p.x), f2ord(p.y), f2ord(p.z)};
#pragma unroll
    for (int d = 0; d < 3; ++d) { lo[d] = min(lo[d], u[d]); hi[d] = max(hi[d], u[d]); }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    lo[d] = __reduce_min_sync(0xffffffffu, lo[d]);
    hi[d] = __reduce_max_sync(0xffffffffu, hi[d]);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { atomicMin(bb + d, lo[d]); atomicMax(bb + 3 + d, hi[d]); }
  }
}

__global__ void minmax_decode_kernel(const unsigned* bb, float* out) {
  if (threadIdx.x < 6) out[threadIdx.x] = ord2f(bb[threadIdx.x]);
}

struct VoxelGeom {
  float vs[3], minb[3];
  unsigned long long nd[3];
};

__global__ void __launch_bounds__(256)
voxel_key_kernel(const float4* __restrict__ feat, int n, VoxelGeom g, uint64_t* __restrict__ keys,
                 uint32_t* __restrict__ vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = feat[i];
  float c[3] = {p.x, p.y, p.z};
  unsigned long long ijk[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float q = __fdiv_rn(c[d], g.vs[d]);
    q = __fsub_rn(q, g.minb[d]);
    ijk[d] = (unsigned long long)floorf(q);
  }
  keys[i] = ijk[0] + ijk[1] * g.nd[0] + ijk[2] * g.nd[0] * g.nd[1];
  vals[i] = (uint32_t)i;
}

struct DescPtr {
  float* data;
  int span;
};
constexpr int kMaxDesc = 8;
struct DescList {
  DescPtr d[kMaxDesc];
  int count;
};

// one thread per sorted position; the head of a voxel segment folds the rest
// of the segment into the FIRST point's column in input order (A4)
__global__ void __launch_bounds__(128)
voxel_reduce_kernel(float4* __restrict__ feat, const uint64_t* __restrict__ keys,
                    const uint32_t* __restrict__ vals, int n, VoxelGeom g, int use_centroid, DescList dl,
                    int* __restrict__ keep) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  uint64_t key = keys[j];
  if (j > 0 && keys[j - 1] == key) return;
  int e = j + 1;
  while (e < n && keys[e] == key) ++e;
  const int first = (int)vals[j];
  const float cnt = (float)(e - j);
  keep[first] = 1;
  if (use_centroid) {
    float4 acc = feat[first];
    for (int t = j + 1; t < e; ++t) {
      float4 p = feat[vals[t]];
      acc.x = __fadd_rn(acc.x, p.x); acc.y = __fadd_rn(acc.y, p.y); acc.z = __fadd_rn(acc.z, p.z);
    }
    acc.x = __fdiv_rn(acc.x, cnt); acc.y = __fdiv_rn(acc.y, cnt); acc.z = __fdiv_rn(acc.z, cnt);
    feat[first] = acc;
  } else {
    unsigned long long ijk[3] = {key % g.nd[0], (key / g.nd[0]) % g.nd[1], key / (g.nd[0] * g.nd[1])};
    float o[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float ctr = __fadd_rn((float)ijk[d], 0.5f);
      ctr = __fadd_rn(ctr, g.minb[d]);
      o[d] = __fmul_rn(g.vs[d], ctr);
    }
    float4 p = feat[first];
    feat[first] = make_float4(o[0], o[1], o[2], p.w);
  }
  for (int q = 0; q < dl.count; ++q) {
    float* D = dl.d[q].data;
    const int span = dl.d[q].span;
    for (int c = 0; c < span; ++c) {
      float acc = D[(size_t)first * span + c];
      for (int t = j + 1; t < e; ++t) acc = __fadd_rn(acc, D[(size_t)vals[t] * span + c]);
      D[(size_t)first * span + c] = __fdiv_rn(acc, cnt);
    }
  }
}

void voxel_grid(Ctx* ctx, const Module& m, Cloud& c) {
  const int n = (int)c.n;
  if (n == 0) return;
  c.touch();
  cudaStream_t s = ctx->stream;
  DBuf<unsigned> bb(ctx, 6);
  DBuf<float> bbf(ctx, 6);
  minmax_init_kernel<<<1, 32, 0, s>>>(bb.p);
  minmax_kernel<<<std::min(ceil_div(n, 1024), 128), 256, 0, s>>>(c.feat.p, n, bb.p);
  minmax_decode_kernel<<<1, 32, 0, s>>>(bb.p, bbf.p);
  ctx_count_launches(ctx, 3);
  float h[6];
  bbf.download(h, 6);
  ctx->sync();
  VoxelGeom g;
  g.vs[0] = (float)m.real("vSizeX"); g.vs[1] = (float)m.real("vSizeY"); g.vs[2] = (float)m.real("vSizeZ");
  double total_bits = 0;
  for (int d = 0; d < 3; ++d) {
    // A4: minBound = min/vSize, numDiv = 1 + maxBound - minBound (fp32, truncated)
    volatile float minb = h[d] / g.vs[d];
    volatile float maxb = h[3 + d] / g.vs[d];
    volatile float nf = 1.0f + maxb;
    nf = nf - minb;
    g.minb[d] = minb;
    g.nd[d] = (unsigned long long)nf;
    if (g.nd[d] == 0) g.nd[d] = 1;
    total_bits += std::log2((double)g.nd[d] + 1.0);
  }
  if (total_bits > 62.0) throw Error(PGS_INVALID_PARAMETER, "VoxelGridDataPointsFilter: voxel grid too fine for this cloud");
  const int key_bits = std::max(8, (int)std::ceil(total_bits) + 1);
  const int stride = ceil_div(n, kSortChunk) * kSortChunk;
  DBuf<uint64_t> ka(ctx, stride), kb(ctx, stride);
  DBuf<uint32_t> va(ctx, stride), vb(ctx, stride);
  DBuf<int> d_n(ctx, 1);
  ctx->upload_small(d_n.p, &n, sizeof(int));
  voxel_key_kernel<<<ceil_div(n, 256), 256, 0, s>>>(c.feat.p, n, g, ka.p, va.p);
  ctx_count_launches(ctx, 1);
  bool in_b = radix_sort_pairs<uint64_t>(ctx, ka.p, kb.p, va.p, vb.p, d_n.p, 1, stride, n, key_bits);
  const uint64_t* keys = in_b ? kb.p : ka.p;
  const uint32_t* vals = in_b ? vb.p : va.p;
  DescList dl;
  dl.count = 0;
  if (m.flag("averageExistingDescriptors")) {
    for (auto& d : c.descs) {
      if (dl.count == kMaxDesc) throw Error(PGS_INVALID_PARAMETER, "VoxelGridDataPointsFilter: too many descriptors");
      dl.d[dl.count++] = DescPtr{d.data.p, d.span};
    }
  }
  DBuf<int> keep(ctx, n);
  keep.zero();
  voxel_reduce_kernel<<<ceil_div(n, 128), 128, 0, s>>>(c.feat.p, keys, vals, n, g, m.flag("useCentroid") ? 1 : 0, dl, keep.p);
  ctx_count_launches(ctx, 1);
  PGS_LAUNCH_CHECK();
  compact_cloud(c, keep.p);
}

void surface_normals(Ctx* ctx, const Module& m, std::vector<Cloud*>& clouds) {
  const int k = (int)m.integer("knn");
  if (k > kMaxDynK) throw Error(PGS_INVALID_PARAMETER, "SurfaceNormalDataPointsFilter: knn > 256 is not supported");
  const float max_dist = (float)m.real("maxDist");
  // upstream smooths in place, point after point, so every normal depends on the already
  // smoothed normals of lower-indexed neighbours: an order-dependent recurrence with no
  // parallel statement that reproduces it
  if (m.flag("smoothNormals"))
    throw Error(PGS_INVALID_PARAMETER, "SurfaceNormalDataPointsFilter: smoothNormals is not supported");
  const int B = (int)clouds.size();
  std::vector<int> ns(B);
  for (int b = 0; b < B; ++b) ns[b] = (int)clouds[b]->n;
  std::vector<std::shared_ptr<Index>> idx;
  cached_indices_for_clouds(ctx, clouds, idx);
  std::vector<DBuf<int32_t>> ids(B);
  std::vector<const Index*> ip(B);
  std::vector<int32_t*> idp(B);
  std::vector<float*> d2p(B, nullptr);  // the covariance only needs WHICH points: no distances written
  for (int b = 0; b < B; ++b) {
    ids[b].reset(ctx, (size_t)std::max(ns[b], 1) * k);
    ip[b] = idx[b].get(); idp[b] = ids[b].p;
  }
  knn_self_batched(ctx, ip, k, max_dist, idp, d2p);
  std::vector<NormalJob> jobs(B);
  int max_n = 0;
  for (int b = 0; b < B; ++b) {
    Cloud& c = *clouds[b];
    NormalJob j{c.feat.p, ids[b].p, ns[b], nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                m.flag("sortEigen") ? 1 : 0};
    if (m.flag("keepNormals")) j.normals = c.add("normals", 3).data.p;
    if (m.flag("keepDensities")) j.dens = c.add("densities", 1).data.p;
    if (m.flag("keepEigenValues")) j.eigval = c.add("eigValues", 3).data.p;
    if (m.flag("keepEigenVectors")) j.eigvec = c.add("eigVectors", 9).data.p;
    if (m.flag("keepMatchedIds")) j.matched = c.add("matchedIds", k).data.p;
    if (m.flag("keepMeanDist")) j.meandist = c.add("meanDists", 1).data.p;
    // Desc vectors may reallocate on add(): re-read the pointers afterwards
    if (j.normals) j.normals = c.find("normals")->data.p;
    if (j.dens) j.dens = c.find("densities")->data.p;
    if (j.eigval) j.eigval = c.find("eigValues")->data.p;
    if (j.eigvec) j.eigvec = c.find("eigVectors")->data.p;
    if (j.matched) j.matched = c.find("matchedIds")->data.p;
    if (j.meandist) j.meandist = c.find("meanDists")->data.p;
    jobs[b] = j;
    max_n = std::max(max_n, ns[b]);
  }
  if (max_n == 0) return;
  DBuf<NormalJob> d_jobs(ctx, B);
  ctx->upload_small(d_jobs.p, jobs.data(), sizeof(NormalJob) * B);
  normals_kernel<<<dim3(ceil_div(max_n, 128), B), 128, 0, ctx->stream>>>(d_jobs.p, k);
  ctx_count_launches(ctx, 1);
  PGS_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------
// SamplingSurfaceNormalDataPointsFilter: recursive median cells of at most knn
// points, one normal per cell, subsampled.  The recursion is level-synchronous:
// the cell COUNTS depend only on n (left = count - count/2), so the host lays
// out every level's cells; the device orders the points of all splitting cells
// of a level with one stable radix sort keyed (cell, coordinate along the cell's
// cut dimension) and clips the inherited boxes at the cut values.
// ---------------------------------------------------------------------------
struct SsnCell {
  int first, count;
  int parent;  // cell of the previous level this one came from
  int side;    // 0: unsplit copy of the parent, 1: left half, 2: right half
};

__global__ void __launch_bounds__(256)
ssn_cut_kernel(const float* __restrict__ box, const SsnCell* __restrict__ cells, int n_cells, int knn, int* __restrict__ cut) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_cells) return;
  if (cells[s].count <= knn) { cut[s] = -1; return; }
  const float* b = box + 6 * (size_t)s;
  int c = 0;
  for (int d = 1; d < 3; ++d)
    if (__fsub_rn(b[3 + d], b[d]) > __fsub_rn(b[3 + c], b[c])) c = d;
  cut[s] = c;
}

__global__ void __launch_bounds__(256)
ssn_key_kernel(const float4* __restrict__ feat, const uint32_t* __restrict__ order, int n, const SsnCell* __restrict__ cells,
               int n_cells, const int* __restrict__ cut, uint64_t* __restrict__ keys) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int lo = 0, hi = n_cells;  // last cell with first <= j
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (cells[mid].first <= j) lo = mid; else hi = mid;
  }
  const int c = cut[lo];
  uint32_t low;
  if (c < 0) {
    low = (uint32_t)(j - cells[lo].first);  // a finished cell keeps its order
  } else {
    const float4 p = feat[order[j]];
    low = f2ord(c == 0 ? p.x : (c == 1 ? p.y : p.z));
  }
  keys[j] = ((uint64_t)lo << 32) | low;
}

// boxes of the next level's cells: the parent's inherited box, clipped at the cut value
// (the coordinate of the first point of the right half) along the parent's cut dimension
__global__ void __launch_bounds__(256)
ssn_child_box_kernel(const float4* __restrict__ feat, const uint32_t* __restrict__ order, const SsnCell* __restrict__ parents,
                     const int* __restrict__ cut, const float* __restrict__ box_in, const SsnCell* __restrict__ cells,
                     int n_cells, float* __restrict__ box_out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_cells) return;
  const SsnCell me = cells[s];
  float b[6];
  for (int d = 0; d < 6; ++d) b[d] = box_in[6 * (size_t)me.parent + d];
  if (me.side != 0) {
    const SsnCell par = parents[me.parent];
    const int c = cut[me.parent];
    const int left = par.count - par.count / 2;
    const float4 p = feat[order[par.first + left]];
    const float v = c == 0 ? p.x : (c == 1 ? p.y : p.z);
    if (me.side == 1) b[3 + c] = v; else b[c] = v;
  }
  for (int d = 0; d < 6; ++d) box_out[6 * (size_t)s + d] = b[d];
}

struct SsnOut {
  float* normals;
  float* dens;
  float* eigval;
  float* eigvec;
};

__global__ void __launch_bounds__(128)
ssn_fuse_kernel(float4* __restrict__ feat, const uint32_t* __restrict__ order, const SsnCell* __restrict__ cells, int n_cells,
                SsnOut out, DescList avg, int bin_method, float ratio, float max_box_dim, uint64_t seed, int need_eigen,
                int* __restrict__ keep) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_cells) return;
  const int first = cells[s].first, cnt = cells[s].count;
  if (cnt <= 0) return;
  float lo[3], hi[3];
  double mean[3] = {0.0, 0.0, 0.0};
  for (int j = 0; j < cnt; ++j) {
    const float4 p4 = feat[order[first + j]];
    const float p[3] = {p4.x, p4.y, p4.z};
    for (int d = 0; d < 3; ++d) {
      if (j == 0 || p[d] < lo[d]) lo[d] = p[d];
      if (j == 0 || p[d] > hi[d]) hi[d] = p[d];
      mean[d] += (double)p[d];
    }
  }
  float box_dim = __fsub_rn(hi[0], lo[0]);
  if (__fsub_rn(hi[1], lo[1]) > box_dim) box_dim = __fsub_rn(hi[1], lo[1]);
  if (__fsub_rn(hi[2], lo[2]) > box_dim) box_dim = __fsub_rn(hi[2], lo[2]);
  if (box_dim > max_box_dim) return;
  for (int d = 0; d < 3; ++d) mean[d] = mean[d] / (double)cnt;
  double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, maxr2 = 0.0;
  for (int j = 0; j < cnt; ++j) {
    const float4 p = feat[order[first + j]];
    const double dx = (double)p.x - mean[0], dy = (double)p.y - mean[1], dz = (double)p.z - mean[2];
    C[0] += dx * dx; C[1] += dx * dy; C[2] += dx * dz;
    C[4] += dy * dy; C[5] += dy * dz; C[8] += dz * dz;
    const double r2 = dx * dx + dy * dy + dz * dz;
    if (r2 > maxr2) maxr2 = r2;
  }
  C[3] = C[1]; C[6] = C[2]; C[7] = C[5];
  double w[3] = {1.0, 0.0, 0.0}, V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (need_eigen) {
    jacobi_sym<3>(C, w, V);
    double wmax = w[0] > w[1] ? w[0] : w[1];
    if (w[2] > wmax) wmax = w[2];
    int rank = 0;
    for (int e = 0; e < 3; ++e)
      if (w[e] > 3.0 * 1.1920928955078125e-07 * wmax) ++rank;
    if (rank < 2) return;
  }
  float normal[3] = {0.f, 0.f, 0.f};
  if (out.normals) {
    int e0 = 0;
    for (int e = 1; e < 3; ++e)
      if (w[e] < w[e0]) e0 = e;
    for (int d = 0; d < 3; ++d) {
      const float v = (float)V[e0 * 3 + d];
      normal[d] = v < -1.f ? -1.f : (v > 1.f ? 1.f : v);
    }
  }
  float density = 0.f;
  if (out.dens) {
    const double r = sqrt(maxr2);
    density = (float)((double)cnt / ((4.0 / 3.0) * 3.14159265358979323846 * (r * r * r)));
  }
  auto write = [&](int k) {
    if (out.normals) for (int d = 0; d < 3; ++d) out.normals[3 * (size_t)k + d] = normal[d];
    if (out.dens) out.dens[k] = density;
    if (out.eigval) for (int e = 0; e < 3; ++e) out.eigval[3 * (size_t)k + e] = (float)w[e];
    if (out.eigvec) for (int e = 0; e < 9; ++e) out.eigvec[9 * (size_t)k + e] = (float)V[e];
  };
  if (!bin_method) {
    for (int j = 0; j < cnt; ++j) {
      const int k = (int)order[first + j];
      const uint64_t h = splitmix64(splitmix64(seed) ^ ((uint64_t)k * 0xD1B54A32D192ED03ull));
      const float r = __fmul_rn((float)(h >> 40), 1.0f / 16777216.0f);
      if (r < ratio) { keep[k] = 1; write(k); }
    }
  } else {
    const int k = (int)order[first];
    keep[k] = 1;
    for (int q = 0; q < avg.count; ++q) {
      float* D = avg.d[q].data;
      const int span = avg.d[q].span;
      for (int d = 0; d < span; ++d) {
        float acc = 0.f;
        for (int j = 0; j < cnt; ++j) acc = __fadd_rn(acc, D[(size_t)order[first + j] * span + d]);
        D[(size_t)k * span + d] = __fdiv_rn(acc, (float)cnt);
      }
    }
    feat[k] = make_float4((float)mean[0], (float)mean[1], (float)mean[2], 1.f);
    write(k);
  }
}

__global__ void ssn_iota_kernel(uint32_t* __restrict__ v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (uint32_t)i;
}

void sampling_surface_normals(Ctx* ctx, const Module& m, Cloud& c) {
  const int n = (int)c.n;
  if (n == 0) return;
  cudaStream_t s = ctx->stream;
  const int knn = (int)m.integer("knn");
  const bool bin = m.integer("samplingMethod") == 1;
  // ---- the cell layout of every level (depends on n and knn only) ------------------
  std::vector<std::vector<SsnCell>> levels;
  levels.push_back({SsnCell{0, n, 0, 0}});
  while (true) {
    const auto& cur = levels.back();
    bool any = false;
    std::vector<SsnCell> next;
    next.reserve(cur.size() * 2);
    for (int i = 0; i < (int)cur.size(); ++i) {
      if (cur[i].count > knn) {
        const int right = cur[i].count / 2, left = cur[i].count - right;
        next.push_back(SsnCell{cur[i].first, left, i, 1});
        next.push_back(SsnCell{cur[i].first + left, right, i, 2});
        any = true;
      } else {
        next.push_back(SsnCell{cur[i].first, cur[i].count, i, 0});
      }
    }
    if (!any) break;
    levels.push_back(std::move(next));
  }
  size_t max_cells = 1;
  for (auto& l : levels) max_cells = std::max(max_cells, l.size());
  // ---- device state -------------------------------------------------------------------
  const int stride = ceil_div(n, kSortChunk) * kSortChunk;
  DBuf<uint64_t> ka(ctx, stride), kb(ctx, stride);
  DBuf<uint32_t> va(ctx, stride), vb(ctx, stride);
  DBuf<int> d_n(ctx, 1), cut(ctx, max_cells);
  DBuf<SsnCell> cells_a(ctx, max_cells), cells_b(ctx, max_cells);
  DBuf<float> box_a(ctx, 6 * max_cells), box_b(ctx, 6 * max_cells);
  ctx->upload_small(d_n.p, &n, sizeof(int));
  ssn_iota_kernel<<<ceil_div(n, 256), 256, 0, s>>>(va.p, n);
  {
    // the root's inherited box is the cloud's bounding box
    DBuf<unsigned> bb(ctx, 6);
    minmax_init_kernel<<<1, 32, 0, s>>>(bb.p);
    minmax_kernel<<<std::min(ceil_div(n, 1024), 128), 256, 0, s>>>(c.feat.p, n, bb.p);
    minmax_decode_kernel<<<1, 32, 0, s>>>(bb.p, box_a.p);
    ctx_count_launches(ctx, 4);
  }
  uint32_t* order = va.p;
  uint32_t* order_alt = vb.p;
  SsnCell* cells = cells_a.p;
  SsnCell* cells_next = cells_b.p;
  float* box = box_a.p;
  float* box_next = box_b.p;
  auto upload_cells = [&](SsnCell* dst, const std::vector<SsnCell>& v) {
    // level tables can exceed the small-upload ring: plain async copy from a staging vector that
    // outlives the copy (synchronised below, once per level)
    PGS_CUDA(cudaMemcpyAsync(dst, v.data(), v.size() * sizeof(SsnCell), cudaMemcpyHostToDevice, s));
  };
  upload_cells(cells, levels[0]);
  for (size_t l = 0; l + 1 < levels.size(); ++l) {
    const int nc = (int)levels[l].size(), nn = (int)levels[l + 1].size();
    ssn_cut_kernel<<<ceil_div(nc, 256), 256, 0, s>>>(box, cells, nc, knn, cut.p);
    ssn_key_kernel<<<ceil_div(n, 256), 256, 0, s>>>(c.feat.p, order, n, cells, nc, cut.p, ka.p);
    ctx_count_launches(ctx, 2);
    int bits = 32;
    while ((1 << (bits - 32)) < nc) ++bits;
    // vals travel with the keys: (ka, order) -> sorted
    uint32_t* v_in = order;
    uint32_t* v_out = order_alt;
    const bool in_b = radix_sort_pairs<uint64_t>(ctx, ka.p, kb.p, v_in, v_out, d_n.p, 1, stride, n, bits);
    if (in_b) std::swap(order, order_alt);
    upload_cells(cells_next, levels[l + 1]);
    ssn_child_box_kernel<<<ceil_div(nn, 256), 256, 0, s>>>(c.feat.p, order, cells, cut.p, box, cells_next, nn, box_next);
    ctx_count_launches(ctx, 1);
    std::swap(cells, cells_next);
    std::swap(box, box_next);
    ctx->sync();  // the pageable level table has been consumed
  }
  ctx->sync();
  // ---- one normal per cell, subsampling ----------------------------------------------
  c.touch();
  SsnOut out{nullptr, nullptr, nullptr, nullptr};
  if (m.flag("keepNormals")) c.add("normals", 3);
  if (m.flag("keepDensities")) c.add("densities", 1);
  if (m.flag("keepEigenValues")) c.add("eigValues", 3);
  if (m.flag("keepEigenVectors")) c.add("eigVectors", 9);
  if (m.flag("keepNormals")) out.normals = c.find("normals")->data.p;
  if (m.flag("keepDensities")) out.dens = c.find("densities")->data.p;
  if (m.flag("keepEigenValues")) out.eigval = c.find("eigValues")->data.p;
  if (m.flag("keepEigenVectors")) out.eigvec = c.find("eigVectors")->data.p;
  DescList avg;
  avg.count = 0;
  if (bin && m.flag("averageExistingDescriptors")) {
    for (auto& d : c.descs) {
      float* p = d.data.p;
      if (p == out.normals || p == out.dens || p == out.eigval || p == out.eigvec) continue;  // overwritten below
      if (avg.count == kMaxDesc) throw Error(PGS_INVALID_PARAMETER, "SamplingSurfaceNormalDataPointsFilter: too many descriptors");
      avg.d[avg.count++] = DescPtr{p, d.span};
    }
  }
  DBuf<int> keep(ctx, n);
  keep.zero();
  const int nc = (int)levels.back().size();
  const int need_eigen = (out.normals || out.eigval || out.eigvec) ? 1 : 0;
  ssn_fuse_kernel<<<ceil_div(nc, 128), 128, 0, s>>>(c.feat.p, order, cells, nc, out, avg, bin ? 1 : 0, (float)m.real("ratio"),
                                                    (float)m.real("maxBoxDim"), (uint64_t)m.integer("seed"), need_eigen, keep.p);
  ctx_count_launches(ctx, 1);
  PGS_LAUNCH_CHECK();
  compact_cloud(c, keep.p);
}

}  // namespace

bool is_rigid(const double* T) {
  double det = T[0] * (T[5] * T[10] - T[9] * T[6]) - T[4] * (T[1] * T[10] - T[9] * T[2]) +
               T[8] * (T[1] * T[6] - T[5] * T[2]);
  return std::fabs(1.0 - det) <= 0.001;
}

Xf xf_from_T(const double* T) {
  Xf x;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) x.m[r * 4 + c] = (float)T[c * 4 + r];
  return x;
}

void rigid_transform_cloud(Cloud& c, const double* T) {
  if (!is_rigid(T)) throw Error(PGS_TRANSFORMATION_ERROR, "RigidTransformation: Error, rotation matrix is not orthogonal.");
  if (c.n == 0) return;
  c.touch();
  Desc* nrm = c.find("normals");
  Desc* obs = c.find("observationDirections");
  rigid_kernel<<<ceil_div(c.n, 256), 256, 0, c.ctx->stream>>>(c.feat.p, nrm ? nrm->data.p : nullptr,
                                                               obs ? obs->data.p : nullptr, (int)c.n, xf_from_T(T));
  ctx_count_launches(c.ctx, 1);
  PGS_LAUNCH_CHECK();
}

void apply_filter(Ctx* ctx, const Module& m, std::vector<Cloud*>& clouds) {
  cudaStream_t s = ctx->stream;
  const std::string& name = m.name;
  if (name == "IdentityDataPointsFilter") return;
  if (name == "SurfaceNormalDataPointsFilter") { surface_normals(ctx, m, clouds); return; }
  for (Cloud* cp : clouds) {
    Cloud& c = *cp;
    const int n = (int)c.n;
    if (name == "VoxelGridDataPointsFilter") { voxel_grid(ctx, m, c); continue; }
    if (name == "SamplingSurfaceNormalDataPointsFilter") { sampling_surface_normals(ctx, m, c); continue; }
    if (name == "ObservationDirectionDataPointsFilter") {
      float* o = c.add("observationDirections", 3).data.p;
      if (n) obsdir_kernel<<<ceil_div(n, 256), 256, 0, s>>>(c.feat.p, o, n, (float)m.real("x"), (float)m.real("y"), (float)m.real("z"));
    } else if (name == "OrientNormalsDataPointsFilter") {
      Desc* nrm = c.find("normals");
      Desc* obs = c.find("observationDirections");
      if (!nrm) throw Error(PGS_INVALID_FIELD, "OrientNormalsDataPointsFilter: Error, cannot find normals in descriptors.");
      if (!obs) throw Error(PGS_INVALID_FIELD, "OrientNormalsDataPointsFilter: Error, cannot find observation directions in descriptors.");
      if (n) orient_kernel<<<ceil_div(n, 256), 256, 0, s>>>(nrm->data.p, obs->data.p, n, m.flag("towardCenter") ? 1 : 0);
    } else if (name == "SimpleSensorNoiseDataPointsFilter") {
      static const float tab[5][3] = {{0.012f, 0.0068f, 0.0008f}, {0.028f, 0.0013f, 0.0001f},
                                      {0.018f, 0.0006f, 0.0015f}, {0.f, 0.f, 0.f}, {0.004f, 0.0053f, -0.0092f}};
      int st = (int)m.integer("sensorType");
      if (st < 0 || st > 4) throw Error(PGS_INVALID_PARAMETER, "SimpleSensorNoiseDataPointsFilter: unknown sensorType");
      float* o = c.add("simpleSensorNoise", 1).data.p;
      if (n) noise_kernel<<<ceil_div(n, 256), 256, 0, s>>>(c.feat.p, o, n, st == 3, tab[st][0], tab[st][1], tab[st][2], (float)m.real("gain"));
    } else if (name == "RandomSamplingDataPointsFilter") {
      if (!n) continue;
      DBuf<int> keep(ctx, n);
      random_keep_kernel<<<ceil_div(n, 256), 256, 0, s>>>(keep.p, n, (uint64_t)m.integer("seed"), (float)m.real("prob"));
      ctx_count_launches(ctx, 1);
      compact_cloud(c, keep.p);
      continue;
    } else if (name == "MaxDensityDataPointsFilter") {
      Desc* dn = c.find("densities");
      if (!dn) throw Error(PGS_INVALID_FIELD, "MaxDensityDataPointsFilter: Error, no densities found in descriptors.");
      if (!n) continue;
      DBuf<int> keep(ctx, n);
      DBuf<unsigned> stats(ctx, 2);
      stats.zero();
      const int nb = std::min(ceil_div(n, 1024), 256);
      density_max_kernel<<<nb, 256, 0, s>>>(dn->data.p, n, stats.p);
      density_count_kernel<<<nb, 256, 0, s>>>(dn->data.p, n, stats.p);
      density_keep_kernel<<<ceil_div(n, 256), 256, 0, s>>>(dn->data.p, keep.p, n, (float)m.real("maxDensity"),
                                                           (uint64_t)m.integer("seed"), stats.p);
      ctx_count_launches(ctx, 3);
      compact_cloud(c, keep.p);
      continue;
    } else if (name == "BoundingBoxDataPointsFilter") {
      if (!n) continue;
      DBuf<int> keep(ctx, n);
      box_keep_kernel<<<ceil_div(n, 256), 256, 0, s>>>(c.feat.p, keep.p, n, (float)m.real("xMin"), (float)m.real("xMax"),
                                                       (float)m.real("yMin"), (float)m.real("yMax"), (float)m.real("zMin"),
                                                       (float)m.real("zMax"), m.flag("removeInside") ? 1 : 0);
      ctx_count_launches(ctx, 1);
      compact_cloud(c, keep.p);
      continue;
    } else if (name == "RemoveNaNDataPointsFilter") {
      if (!n) continue;
      DBuf<int> keep(ctx, n);
      nan_keep_kernel<<<ceil_div(n, 256), 256, 0, s>>>(c.feat.p, keep.p, n);
      ctx_count_launches(ctx, 1);
      compact_cloud(c, keep.p);
      continue;
    } else if (name == "FixStepSamplingDataPointsFilter") {
      const int64_t step = m.integer("startStep");
      if (m.integer("endStep") != step || m.real("stepMult") != 1.0)
        throw Error(PGS_INVALID_PARAMETER,
                    "FixStepSamplingDataPointsFilter: step schedules across calls (endStep != startStep or stepMult != 1) "
                    "are not supported");
      if (!n) continue;
      uint64_t x = (uint64_t)m.integer("seed");
      for (int r = 0; r < 2; ++r) {  // splitmix64 twice, as random_keep_kernel hashes its seed
        x += 0x9E3779B97F4A7C15ull;
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
        x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
        x = x ^ (x >> 31);
      }
      DBuf<int> keep(ctx, n);
      step_keep_kernel<<<ceil_div(n, 256), 256, 0, s>>>(keep.p, n, (int)step, (int)(x % (uint64_t)step));
      ctx_count_launches(ctx, 1);
      compact_cloud(c, keep.p);
      continue;
    } else if (name == "ShadowDataPointsFilter") {
      Desc* nrm = c.find("normals");
      if (!nrm) throw Error(PGS_INVALID_FIELD, "ShadowDataPointsFilter, Error: cannot find normals in descriptors");
      if (!n) continue;
      DBuf<int> keep(ctx, n);
      shadow_keep_kernel<<<ceil_div(n, 256), 256, 0, s>>>(c.feat.p, nrm->data.p, keep.p, n, (float)m.real("eps"));
      ctx_count_launches(ctx, 1);
      compact_cloud(c, keep.p);
      continue;
    } else if (name == "MaxDistDataPointsFilter" || name == "MinDistDataPointsFilter") {
      if (!n) continue;
      const bool is_max = name[1] == 'a';
      DBuf<int> keep(ctx, n);
      dist_keep_kernel<<<ceil_div(n, 256), 256, 0, s>>>(c.feat.p, keep.p, n, (int)m.integer("dim"),
                                                        (float)m.real(is_max ? "maxDist" : "minDist"), is_max ? 1 : 0);
      ctx_count_launches(ctx, 1);
      compact_cloud(c, keep.p);
      continue;
    } else {
      throw Error(PGS_INVALID_ELEMENT, "DataPointsFilter " + name + " is registered but has no device implementation");
    }
    if (n) ctx_count_launches(ctx, 1);
  }
  PGS_LAUNCH_CHECK();
}

void apply_filters(Ctx* ctx, const std::vector<Module>& ms, std::vector<Cloud*>& clouds) {
  for (auto& m : ms) apply_filter(ctx, m, clouds);
}

}  // namespace pgs
