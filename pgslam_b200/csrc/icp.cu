// icp.cu — the ICP iteration loop on the device (SURVEY.md §3.3 "THE HOT LOOP";
// ICP::compute / computeWithTransformedReference reached from LoopCloser.hpp:98
// and Localizer.hpp:126).
//
// Per iteration, three kernels over all active pairs (blockIdx.y = pair):
//   match      stepReading = T_iter * reading (fp32, fused, never stored) ->
//              exact nearest neighbour in the reference index          (A7+A9)
//   select     exact order statistic of the match distances by MSB radix
//              select -> TrimmedDist / MedianDist limits                (A10)
//   accumulate weights + ErrorElements + normal equations in one pass: the
//              reference point and normal are gathered by match position, the
//              6x6 / 3x3 sums are reduced in fp64 (warp shuffle -> block ->
//              fixed-order sum over blocks), and the LAST block to finish
//              solves, composes T_iter and runs the transformation checkers
//              on one thread                                       (A11-A14)
// so an iteration needs no host round trip; the host only polls a mapped flag
// to stop launching once every pair has converged.
#include "icp.cuh"

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <thread>

#include "knn.cuh"
#include "solve.cuh"

// The matcher kernel waits on scattered node / leaf loads (top stall: long scoreboard, L1
// throughput 82 %): 16 blocks of 128 per SM (32 registers, 40 bytes spilled) hide more of
// that latency than 10 blocks at 48 registers - 5.5 % less match time, measured.
#ifndef PGS_MATCH_MIN_BLOCKS
#define PGS_MATCH_MIN_BLOCKS 16
#endif
#define PGS_MATCH_BOUNDS __launch_bounds__(128, PGS_MATCH_MIN_BLOCKS)

namespace pgs {

namespace {

constexpr float kInfF = __builtin_inff();

__host__ __device__ inline int tri(int c, int r) { return c * (c + 1) / 2 + r; }  // r <= c

// number of blocks that reduce a cloud of n points: a function of n alone, so
// that fp64 sums are bit-identical whatever batch the cloud is processed in
__host__ __device__ inline int reduce_slices(int n, int per_block, int max_blocks) {
  int s = (n + per_block - 1) / per_block;
  return s < 1 ? 1 : (s > max_blocks ? max_blocks : s);
}
constexpr int kAccPerBlock = 4096;
constexpr int kAccMaxBlocks = 128;
#ifndef PGS_ACC_UNROLL
#define PGS_ACC_UNROLL 4
#endif
constexpr int kAccUnroll = PGS_ACC_UNROLL;  // matches whose loads are in flight together in accumulate_kernel

// ---------------------------------------------------------------------------
// per-pair accumulation of one weighted match (A.5 / A.7), all in fp64
// ---------------------------------------------------------------------------
__device__ __forceinline__ void accumulate_match(double* acc, int minimizer, const float3 pf, const float4 qf,
                                                 float3 nf, double w, int force_mode = FORCE_NONE) {
  // force2D: features and normals are cut to x,y (PointToPlane.cpp compute_in_place); with
  // n.z = 0 the (p x n).z, n.x, n.y rows and the residual are exactly the 2-D ones
  if (force_mode == FORCE_2D) nf.z = 0.f;
  const double p[3] = {(double)pf.x, (double)pf.y, (double)pf.z};
  if (minimizer == MIN_P2POINT) {
    const double q[3] = {(double)qf.x, (double)qf.y, (double)qf.z};
    acc[0] += w;
#pragma unroll
    for (int d = 0; d < 3; ++d) { acc[1 + d] += w * p[d]; acc[4 + d] += w * q[d]; }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r) acc[7 + c * 3 + r] += w * q[r] * p[c];
    double dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
    acc[27] += sqrt(dx * dx + dy * dy + dz * dz);
  } else {
    const double n[3] = {(double)nf.x, (double)nf.y, (double)nf.z};
    double F[6];
    F[0] = p[1] * n[2] - p[2] * n[1];
    F[1] = p[2] * n[0] - p[0] * n[2];
    F[2] = p[0] * n[1] - p[1] * n[0];
    F[3] = n[0]; F[4] = n[1]; F[5] = n[2];
    double e = (p[0] - (double)qf.x) * n[0] + (p[1] - (double)qf.y) * n[1] + (p[2] - (double)qf.z) * n[2];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      double wf = w * F[c];
#pragma unroll
      for (int r = 0; r <= c; ++r) acc[c * (c + 1) / 2 + r] += wf * F[r];
      acc[21 + c] -= wf * e;
    }
    acc[27] += w * (e * e);
  }
  acc[28] += 1.0;
  acc[29] += w;
}

// A.6 per-pair terms of the Censi covariance; acc2[0..20] = H, [21..41] = D D^T
__device__ __forceinline__ void accumulate_cov(double* acc2, const float3 pf, const float4 qf, const float3 nf,
                                               double alpha, double beta, double gamma, const double* t) {
  double p[3] = {(double)pf.x, (double)pf.y, (double)pf.z}, q[3] = {(double)qf.x, (double)qf.y, (double)qf.z};
  double n[3] = {(double)nf.x, (double)nf.y, (double)nf.z};
  double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  double rp = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  double rq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  if (!(nn > 0.0) || !(rp > 0.0) || !(rq > 0.0)) return;
  for (int d = 0; d < 3; ++d) n[d] = n[d] / nn;
  double dp[3] = {p[0] / rp, p[1] / rp, p[2] / rp};
  double dq[3] = {q[0] / rq, q[1] / rq, q[2] / rq};
  double na = n[2] * dp[1] - n[1] * dp[2];
  double nb = n[0] * dp[2] - n[2] * dp[0];
  double ng = n[1] * dp[0] - n[0] * dp[1];
  double E = n[0] * (p[0] - gamma * p[1] + beta * p[2] + t[0] - q[0]);
  E += n[1] * (gamma * p[0] + p[1] - alpha * p[2] + t[1] - q[1]);
  E += n[2] * (-beta * p[0] + alpha * p[1] + p[2] + t[2] - q[2]);
  double Np = n[0] * (dp[0] - gamma * dp[1] + beta * dp[2]);
  Np += n[1] * (gamma * dp[0] + dp[1] - alpha * dp[2]);
  Np += n[2] * (-beta * dp[0] + alpha * dp[1] + dp[2]);
  double Nq = -(n[0] * dq[0] + n[1] * dq[1] + n[2] * dq[2]);
  double g[6] = {n[0], n[1], n[2], rp * na, rp * nb, rp * ng};
  double en = E + rp * Np;
  double u[6] = {n[0] * Np, n[1] * Np, n[2] * Np, na * en, nb * en, ng * en};
  double v[6] = {n[0] * Nq, n[1] * Nq, n[2] * Nq, rq * na * Nq, rq * nb * Nq, rq * ng * Nq};
#pragma unroll
  for (int c = 0; c < 6; ++c)
#pragma unroll
    for (int r = 0; r <= c; ++r) {
      acc2[c * (c + 1) / 2 + r] += g[c] * g[r];
      acc2[21 + c * (c + 1) / 2 + r] += u[c] * u[r] + v[c] * v[r];
    }
}

// SurfaceNormalOutlierFilter: weight 0 iff the unit normals' dot product is below
// cos(maxAngle); fp32, fixed order (NaN from a zero normal compares false -> kept)
__device__ __forceinline__ bool sn_reject(float ax, float ay, float az, float bx, float by, float bz, float eps) {
  float na = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), __fmul_rn(az, az)));
  float nb = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(bx, bx), __fmul_rn(by, by)), __fmul_rn(bz, bz)));
  float ux = __fdiv_rn(ax, na), uy = __fdiv_rn(ay, na), uz = __fdiv_rn(az, na);
  float vx = __fdiv_rn(bx, nb), vy = __fdiv_rn(by, nb), vz = __fdiv_rn(bz, nb);
  float dot = __fadd_rn(__fadd_rn(__fmul_rn(ux, vx), __fmul_rn(uy, vy)), __fmul_rn(uz, vz));
  return dot < eps;
}

__device__ void xf_from_T_dev(const double* T, Xf& x) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) x.m[r * 4 + c] = (float)T[c * 4 + r];
}

__device__ void chk_push(PairState& st, const double* T) {
  if (st.nhist == kChkHist) {
    for (int i = 0; i + 1 < kChkHist; ++i) {
      for (int d = 0; d < 4; ++d) st.q[i][d] = st.q[i + 1][d];
      for (int d = 0; d < 3; ++d) st.t[i][d] = st.t[i + 1][d];
    }
    st.nhist--;
  }
  quat_from_T(T, st.q[st.nhist]);
  st.t[st.nhist][0] = T[12]; st.t[st.nhist][1] = T[13]; st.t[st.nhist][2] = T[14];
  st.nhist++;
}

__device__ void pair_finished(int* n_active, volatile int* h_done) {
  int left = atomicSub(n_active, 1) - 1;
  if (left == 0) {
    *h_done = 1;
    __threadfence_system();
  }
}

// ErrorMinimizer::compute + T_iter update + TransformationCheckers (one thread)
__device__ void finish_iteration(PairState& st, const IcpParams& P, const double* acc, int* n_active,
                                 volatile int* h_done) {
  st.kept = acc[28];
  st.wsum = acc[29];
  st.resid = acc[27];
  if (!(acc[28] > 0.0)) {  // "no point to minimize"
    st.status = PGS_CONVERGENCE_ERROR;
    st.active = 0;
    pair_finished(n_active, h_done);
    return;
  }
  double Tinc[16];
  if (P.minimizer == MIN_P2POINT) {
    const double W = acc[0];
    double mp[3], mq[3], M[9];
    for (int d = 0; d < 3; ++d) { mp[d] = acc[1 + d] / W; mq[d] = acc[4 + d] / W; }
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r) M[c * 3 + r] = acc[7 + c * 3 + r] - W * mq[r] * mp[c];
    double U[9], S[3], V[9], R[9];
    svd3(M, U, S, V);
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r) {
        double s = 0.0;
        for (int k = 0; k < 3; ++k) s += U[k * 3 + r] * V[k * 3 + c];
        R[c * 3 + r] = s;
      }
    double det = R[0] * (R[4] * R[8] - R[7] * R[5]) - R[3] * (R[1] * R[8] - R[7] * R[2]) +
                 R[6] * (R[1] * R[5] - R[4] * R[2]);
    if (det < 0.0) {
      for (int r = 0; r < 3; ++r) V[6 + r] = -V[6 + r];
      for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) {
          double s = 0.0;
          for (int k = 0; k < 3; ++k) s += U[k * 3 + r] * V[k * 3 + c];
          R[c * 3 + r] = s;
        }
    }
    m4_identity(Tinc);
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r) Tinc[c * 4 + r] = R[c * 3 + r];
    for (int r = 0; r < 3; ++r) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += R[k * 3 + r] * mp[k];
      Tinc[12 + r] = mq[r] - s;
    }
  } else {
    double A[36], b[6], x[6];
    if (P.force_mode == FORCE_NONE) {
      for (int c = 0; c < 6; ++c) {
        for (int r = 0; r <= c; ++r) {
          A[c * 6 + r] = acc[tri(c, r)];
          A[r * 6 + c] = acc[tri(c, r)];
        }
        b[c] = acc[21 + c];
      }
      solve6(A, b, x);
      angle_axis_to_T(x, Tinc);
    } else {
      // unknowns [yaw, tx, ty(, tz)] = rows/columns 2,3,4(,5) of the 6-DOF system
      const int nd = P.force_mode == FORCE_2D ? 3 : 4;
      for (int c = 0; c < nd; ++c) {
        for (int r = 0; r <= c; ++r) {
          A[c * nd + r] = acc[tri(c + 2, r + 2)];
          A[r * nd + c] = acc[tri(c + 2, r + 2)];
        }
        b[c] = acc[21 + c + 2];
      }
      x[3] = 0.0;
      if (nd == 3) solve_sym<3>(A, b, x);
      else solve_sym<4>(A, b, x);
      m4_identity(Tinc);
      const double cs = cos(x[0]), sn = sin(x[0]);
      Tinc[0] = cs; Tinc[4] = -sn;
      Tinc[1] = sn; Tinc[5] = cs;
      Tinc[12] = x[1]; Tinc[13] = x[2]; Tinc[14] = nd == 4 ? x[3] : 0.0;
    }
  }
  for (int i = 0; i < 16; ++i) { st.T_prev[i] = st.T_iter[i]; st.T_inc[i] = Tinc[i]; }
  st.xf_prev = st.xf;
  m4_mul(Tinc, st.T_iter, st.T_iter);
  st.iterations++;

  bool iterate = true;
  int status = PGS_OK;
  if (P.max_iterations > 0) {
    st.counter++;
    if (st.counter >= P.max_iterations) { iterate = false; st.max_reached = 1; }
  }
  chk_push(st, st.T_iter);
  if (P.has_diff) {
    const int sl = P.smooth_length;
    if (st.nhist > sl) {
      double c0 = 0.0, c1 = 0.0;
      for (int i = st.nhist - 1; i >= st.nhist - sl; --i) {
        c0 += fabs(quat_angular_distance(st.q[i], st.q[i - 1]));
        double dx = st.t[i][0] - st.t[i - 1][0], dy = st.t[i][1] - st.t[i - 1][1], dz = st.t[i][2] - st.t[i - 1][2];
        c1 += sqrt(dx * dx + dy * dy + dz * dz);
      }
      c0 = c0 / (double)sl;
      c1 = c1 / (double)sl;
      if (isnan(c0) || isnan(c1)) status = PGS_CONVERGENCE_ERROR;
      else if (c0 < P.min_diff_rot && c1 < P.min_diff_trans) iterate = false;
    }
  }
  if (P.has_bound) {
    double qn[4];
    quat_from_T(st.T_iter, qn);
    double dx = st.T_iter[12] - st.t0[0], dy = st.T_iter[13] - st.t0[1], dz = st.T_iter[14] - st.t0[2];
    if (quat_angular_distance(qn, st.q0) > P.max_rot || sqrt(dx * dx + dy * dy + dz * dz) > P.max_trans)
      status = PGS_CONVERGENCE_ERROR;
  }
  if (st.iterations >= P.hard_iteration_cap) { iterate = false; st.max_reached = 1; }
  if (status == PGS_OK && iterate) {
    const double* T = st.T_iter;
    double det = T[0] * (T[5] * T[10] - T[9] * T[6]) - T[4] * (T[1] * T[10] - T[9] * T[2]) +
                 T[8] * (T[1] * T[6] - T[5] * T[2]);
    if (!(fabs(1.0 - det) <= 0.001)) status = PGS_TRANSFORMATION_ERROR;
  }
  xf_from_T_dev(st.T_iter, st.xf);
  if (status != PGS_OK) st.status = status;
  if (status != PGS_OK || !iterate) {
    st.active = 0;
    pair_finished(n_active, h_done);
  }
}

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mean_kernel(const float4* const* __restrict__ pts, const int* __restrict__ ns, double* __restrict__ partials,
            unsigned* __restrict__ tickets, float* __restrict__ shift /*4 per job*/) {
  __shared__ double sh[8][3];
  __shared__ bool last;
  const int b = blockIdx.y;
  const int n = ns[b];
  // the summation order must depend on this cloud only, never on what else is
  // in the batch: the number of slices is a function of n alone
  const unsigned slices = (unsigned)reduce_slices(n, 2048, 64);
  if (blockIdx.x >= slices) return;
  const float4* p = pts[b];
  double s[3] = {0.0, 0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += slices * blockDim.x) {
    float4 v = p[i];
    s[0] += (double)v.x; s[1] += (double)v.y; s[2] += (double)v.z;
  }
  for (int d = 0; d < 3; ++d) s[d] = warp_sum(s[d]);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[w][0] = s[0]; sh[w][1] = s[1]; sh[w][2] = s[2]; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int j = 0; j < 8; ++j) t += sh[j][threadIdx.x];
    partials[((size_t)b * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(&tickets[b], 1u) == slices - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (unsigned j = 0; j < slices; ++j) t += __ldcg(&partials[((size_t)b * gridDim.x + j) * 3 + threadIdx.x]);
    shift[4 * b + threadIdx.x] = n > 0 ? (float)(t / (double)n) : 0.f;
  }
  if (threadIdx.x == 3) shift[4 * b + 3] = 0.f;
  if (threadIdx.x == 0) tickets[b] = 0;
}

__global__ void shift_points_kernel(float4* __restrict__ pts, int n, const float* __restrict__ shift) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pts[i];
  pts[i] = make_float4(__fsub_rn(p.x, shift[0]), __fsub_rn(p.y, shift[1]), __fsub_rn(p.z, shift[2]), p.w);
}

// per-pair state initialisation: pose algebra in fp64 on the device so that the
// reference mean never has to visit the host
__global__ void init_state_kernel(PairState* __restrict__ states, const double* const* __restrict__ T_refIn_refMean,
                                  int n_pairs) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  PairState& st = states[p];
  for (int i = 0; i < 16; ++i) st.T_refIn_refMean[i] = T_refIn_refMean[p][i];
  double inv[16];
  m4_rigid_inv(st.T_refIn_refMean, inv);
  m4_mul(inv, st.T_init, st.T_refMean_dataIn);
  xf_from_T_dev(st.T_refMean_dataIn, st.xf0);
  m4_identity(st.T_iter);
  m4_identity(st.T_prev);
  m4_identity(st.T_inc);
  xf_from_T_dev(st.T_iter, st.xf);
  st.xf_prev = st.xf;
  st.iterations = 0;
  st.max_reached = 0;
  st.counter = 0;
  st.nhist = 0;
  chk_push(st, st.T_iter);
  for (int d = 0; d < 4; ++d) st.q0[d] = st.q[0][d];
  for (int d = 0; d < 3; ++d) st.t0[d] = st.t[0][d];
  st.ticket = 0;
  st.ticket2 = 0;
  st.ticket3 = 0;
  st.sel_prefix = st.sel_mask = 0;
  st.sel_rank = 0;
  st.kept = st.wsum = st.resid = st.overlap = 0.0;
  for (int i = 0; i < 36; ++i) st.cov[i] = 0.0;
  st.lim_lo = -kInfF;
  st.lim_hi = kInfF;
}

// T_refIn_refMean from the device-side mean (identity + translation)
__global__ void mean_pose_kernel(const float* __restrict__ shift, double* __restrict__ T, int n_jobs) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_jobs) return;
  double* M = T + 16 * b;
  m4_identity(M);
  M[12] = (double)shift[4 * b]; M[13] = (double)shift[4 * b + 1]; M[14] = (double)shift[4 * b + 2];
}

struct PreJob {
  float4* feat;
  float* normals;
  float* obs;
  int n;
};
// transformations.apply(reading, T_refMean_dataIn) for every pair
__global__ void __launch_bounds__(256)
pretransform_kernel(const PreJob* __restrict__ jobs, const PairState* __restrict__ states) {
  const PreJob job = jobs[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= job.n) return;
  const Xf T = states[blockIdx.y].xf0;
  float4 p = job.feat[i];
  float3 o = xform_rn(T, p.x, p.y, p.z);
  job.feat[i] = make_float4(o.x, o.y, o.z, p.w);
  if (job.normals) {
    float3 v = rot_rn(T, job.normals[3 * i], job.normals[3 * i + 1], job.normals[3 * i + 2]);
    job.normals[3 * i] = v.x; job.normals[3 * i + 1] = v.y; job.normals[3 * i + 2] = v.z;
  }
  if (job.obs) {
    float3 v = rot_rn(T, job.obs[3 * i], job.obs[3 * i + 1], job.obs[3 * i + 2]);
    job.obs[3 * i] = v.x; job.obs[3 * i + 1] = v.y; job.obs[3 * i + 2] = v.z;
  }
}

// descriptor rows re-ordered to the index' sorted order (float4 per point)
__global__ void __launch_bounds__(256)
gather_vec3_sorted_kernel(const float4* __restrict__ sorted_pts, int n, const float* __restrict__ src, int span,
                          float4* __restrict__ out4, float* __restrict__ out1) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int o = __float_as_int(sorted_pts[j].w);
  if (span == 3) out4[j] = make_float4(src[3 * o], src[3 * o + 1], src[3 * o + 2], 0.f);
  else out1[j] = src[o];
}

// the same for the span-3 descriptor of every reference of a batch in one launch
struct GatherJob {
  const float4* sorted_pts;
  const float* src;
  float4* out4;  // 2 float4 per sorted position: {point, normal} = one 32-byte sector per match
  int n;
};
__global__ void __launch_bounds__(256) gather_vec3_sorted_batched_kernel(const GatherJob* __restrict__ jobs) {
  const GatherJob job = jobs[blockIdx.y];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= job.n) return;
  const float4 p = job.sorted_pts[j];
  const int o = __float_as_int(p.w);
  job.out4[2 * j] = p;
  job.out4[2 * j + 1] = make_float4(job.src[3 * o], job.src[3 * o + 1], job.src[3 * o + 2], 0.f);
}

constexpr int kSeedLevels = 10;  // levels above the seed leaf searched top-down (a 2^10-leaf neighbourhood)
__global__ void PGS_MATCH_BOUNDS
match_kernel(const PairView* __restrict__ views, PairState* __restrict__ states, float maxr2) {
  PairState& st = states[blockIdx.y];
  if (!st.active) return;
  const PairView v = views[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.n_r) return;
  const Xf T = st.xf;
  float4 r = v.reading[i];
  float3 q = xform_rn(T, r.x, r.y, r.z);
  Best1 acc;
  acc.init(maxr2);
  // Temporal coherence: after the first iteration the previous match is almost always still
  // the answer.  It is offered first (a tight bound before any box is tested); the search
  // then walks down from the match's ancestor kSeedLevels up - the neighbourhood a query
  // drifts within between iterations - and climbs the rest of the way to the root testing
  // sibling boxes, which almost never qualify.  Measured against a pure climb from the seed
  // leaf and a pure descent from the root: 8 % less match time over a
  // registration, most of it in the first iterations, where the old leaf is a poor seed.
  const int pp = st.iterations > 0 ? v.match_pos[i] : -1;
  if (pp >= 0) {
    const float4 c = __ldg(v.tree.pts + pp);
    acc.offer(dist2_rn(q.x, q.y, q.z, c.x, c.y, c.z), __float_as_int(c.w), pp);
    const int h = min(kSeedLevels, v.tree.depth);
    knn_traverse_from(v.tree, (unsigned)(v.tree.P + pp / kLeaf) >> h, v.tree.depth - h, q.x, q.y, q.z, acc, pp, pp);
    knn_climb(v.tree, pp / kLeaf, q.x, q.y, q.z, acc, pp, pp, h, false);
  } else {
    knn_traverse(v.tree, q.x, q.y, q.z, acc);
  }
  v.match_pos[i] = acc.pos;
  v.match_d2[i] = acc.dist();
}

// ---------------------------------------------------------------------------
// Cell test (index.cu cell_levels_kernel): true iff every reference point OUTSIDE the cell of
// `node` is strictly farther from q than `bound`, i.e. the search may start at `node`.
// m = the smallest of the six face gaps, each computed with the subtraction the distance
// itself uses; a point beyond a face has |q_a - p_a| >= that gap after rounding (monotone),
// so its distance is >= fl(m*m).  Strict '>' keeps equal-distance lower-index points reachable.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool cell_contains(const unsigned long long* __restrict__ cells8, unsigned node,
                                              const QueryPk& q, float bound) {
  const unsigned long long* __restrict__ cb = cells8 + (size_t)node * 3;
  float ax, ay, bx, by, cl, ch;
  unpack_f32x2(sub_f32x2(q.xy, __ldg(cb)), ax, ay);      // q - lo
  unpack_f32x2(sub_f32x2(__ldg(cb + 1), q.xy), bx, by);  // hi - q
  unpack_f32x2(sub_f32x2(__ldg(cb + 2), q.zz), cl, ch);  // lo.z - q.z, hi.z - q.z
  const float m = fminf(fminf(fminf(ax, ay), fminf(bx, by)), fminf(-cl, ch));
  return m > 0.f && __fmul_rn(m, m) > bound;
}

// match_kernel with the cell test instead of the fixed 10-level descent + climb: the search
// starts at the lowest ancestor of the previous match's leaf whose cell holds the candidate
// ball, and never climbs.
__global__ void PGS_MATCH_BOUNDS
match_cells_kernel(const PairView* __restrict__ views, PairState* __restrict__ states, float maxr2) {
  PairState& st = states[blockIdx.y];
  if (!st.active) return;
  const PairView v = views[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.n_r) return;
  const Xf T = st.xf;
  float4 r = v.reading[i];
  float3 q = xform_rn(T, r.x, r.y, r.z);
  Best1 acc;
  acc.init(maxr2);
  const int pp = st.iterations > 0 ? v.match_pos[i] : -1;
  unsigned node = 1u;
  int depth = 0;
  if (pp >= 0) {
    const float4 c = __ldg(v.tree.pts + pp);
    acc.offer(dist2_rn(q.x, q.y, q.z, c.x, c.y, c.z), __float_as_int(c.w), pp);
    const unsigned long long* __restrict__ cells8 = reinterpret_cast<const unsigned long long*>(v.tree.cells);
    const QueryPk qp = pack_query(q.x, q.y, q.z);
    node = (unsigned)(v.tree.P + pp / kLeaf);
    depth = v.tree.depth;
    while (depth > 0 && !cell_contains(cells8, node, qp, acc.bound())) { node >>= 1; --depth; }
  }
  knn_traverse_from(v.tree, node, depth, q.x, q.y, q.z, acc, 0, -1);
  v.match_pos[i] = acc.pos;
  v.match_d2[i] = acc.dist();
}

// ---------------------------------------------------------------------------
// Persistent matcher: warps pull ranges of queries from a ticket counter and every LANE pulls
// its next query as soon as its search ends (dynamic refill), so a lane whose walk was short
// does not idle until the longest walk of its warp finishes.  A search is a small state
// machine - CELL (climb one ancestor, one cell test), PAIR (test the two children of a node),
// LEAF (scan 8 points) - and each trip of the warp's loop runs the phase most lanes wait in.
// A pending sibling is revisited through its parent's PAIR step (restricted to that child), so
// there is no separate "pop" phase to diverge into.  Results are the same exact (distance,
// index) minima as match_kernel's: the traversal order cannot change an exact minimum.
// ---------------------------------------------------------------------------
#ifndef PGS_PM_MIN_BLOCKS
#define PGS_PM_MIN_BLOCKS 10
#endif
constexpr int kPmRange = 2048;  // queries per ticket

enum { PH_IDLE = 0, PH_CELL = 1, PH_PAIR = 2, PH_LEAF = 3 };

template <bool kUseCells>
__global__ void __launch_bounds__(128, PGS_PM_MIN_BLOCKS)
match_persistent_kernel(const PairView* __restrict__ views, PairState* __restrict__ states, float maxr2,
                        int ranges_per_pair, unsigned n_tickets, unsigned* __restrict__ ticket, int refill_min,
                        int pair_w, int leaf_w) {
  const unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (;;) {
    unsigned t = 0;
    if (lane == 0) t = atomicAdd(ticket, 1u);
    t = __shfl_sync(FULL, t, 0);
    if (t >= n_tickets) return;
    const int pair = (int)(t / (unsigned)ranges_per_pair);
    PairState& st = states[pair];
    if (!st.active) continue;
    const PairView v = views[pair];
    int next = (int)(t % (unsigned)ranges_per_pair) * kPmRange;
    if (next >= v.n_r) continue;
    const int q1 = min(next + kPmRange, v.n_r);
    const bool seeded = st.iterations > 0;
    const int D = v.tree.depth, P = v.tree.P;
    const ulonglong2* __restrict__ nodes16 = reinterpret_cast<const ulonglong2*>(v.tree.nodes);
    const unsigned long long* __restrict__ cells8 = reinterpret_cast<const unsigned long long*>(v.tree.cells);

    int qi = -1, pos = -1, depth = 0, phase = PH_IDLE;
    unsigned node = 1u, trail = 0u, mode = 0u;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    unsigned long long key = 0ull;
    for (;;) {
      // ---- refill -------------------------------------------------------------------
      const unsigned bi = __ballot_sync(FULL, phase == PH_IDLE);
      if (bi) {
        const int nidle = __popc(bi);
        if (next < q1) {
          if (nidle >= refill_min || bi == FULL) {
            const int my = next + __popc(bi & lt_mask);
            if (phase == PH_IDLE && my < q1) {
              qi = my;
              const Xf T = st.xf;
              const float4 r = v.reading[qi];
              const float3 q = xform_rn(T, r.x, r.y, r.z);
              qx = q.x; qy = q.y; qz = q.z;
              key = make_key(maxr2, 0x7fffffff);
              pos = -1;
              node = 1u; depth = 0; trail = 0u; mode = 0u;
              phase = D == 0 ? PH_LEAF : PH_PAIR;
              const int pp = seeded ? v.match_pos[qi] : -1;
              if (pp >= 0) {
                const float4 c = __ldg(v.tree.pts + pp);
                const unsigned long long nk = make_key(dist2_rn(qx, qy, qz, c.x, c.y, c.z), __float_as_int(c.w));
                if (nk < key) { key = nk; pos = pp; }
                if (kUseCells && D > 0) { node = (unsigned)(P + pp / kLeaf); depth = D; phase = PH_CELL; }
              }
            }
            next += nidle;
          }
        } else if (bi == FULL) {
          break;
        }
      }
      const QueryPk qp = pack_query(qx, qy, qz);
      // ---- CELL: cheap, runs whenever a lane is looking for its start node ------------------
      if (kUseCells && phase == PH_CELL) {
        if (depth == 0 || cell_contains(cells8, node, qp, key_dist(key))) {
          trail = 0u; mode = 0u;
          phase = depth == D ? PH_LEAF : PH_PAIR;
        } else {
          node >>= 1; --depth;
        }
      }
      // ---- PAIR or LEAF: the phase that serves more lanes per issue slot --------------------
      const unsigned bp = __ballot_sync(FULL, phase == PH_PAIR);
      const unsigned bl = __ballot_sync(FULL, phase == PH_LEAF);
      bool go_pop = false;
      if (bp && (!bl || __popc(bp) * pair_w >= __popc(bl) * leaf_w)) {
        if (phase == PH_PAIR) {
          const ulonglong2* __restrict__ c = nodes16 + (size_t)node * 3;  // both children: 48 bytes
          const ulonglong2 a = __ldg(c), b = __ldg(c + 1), e = __ldg(c + 2);
          const float lb0 = box_lb_packed(qp, a.x, a.y, b.x);
          const float lb1 = box_lb_packed(qp, b.y, e.x, e.y);
          const float bound = key_dist(key);
          bool ok;
          unsigned child, pend;
          if (mode & 2u) {  // revisit of a pending sibling: only that child is of interest
            child = mode & 1u;
            ok = (child ? lb1 : lb0) <= bound;
            pend = 0u;
          } else {
            const bool near1 = lb1 < lb0;
            const float lbn = near1 ? lb1 : lb0, lbf = near1 ? lb0 : lb1;
            child = near1 ? 1u : 0u;
            ok = lbn <= bound;
            pend = (lbf <= bound) ? 1u : 0u;
          }
          mode = 0u;
          if (ok) {
            trail = (trail << 1) | pend;
            node = node * 2u + child;
            ++depth;
            if (depth == D) phase = PH_LEAF;
          } else {
            go_pop = true;
          }
        }
      } else if (bl) {
        if (phase == PH_LEAF) {
          const int leaf = (int)node - P;
          if (leaf < v.tree.n_leaves) {
            const float4* __restrict__ lp = v.tree.pts + (size_t)leaf * kLeaf;
#pragma unroll
            for (int j = 0; j < kLeaf; ++j) {
              const float4 p = __ldg(lp + j);
              const unsigned long long nk = make_key(dist2_rn_packed(qp.xy, qz, p.x, p.y, p.z), __float_as_int(p.w));
              if (nk < key) { key = nk; pos = leaf * kLeaf + j; }
            }
          }
          go_pop = true;
        }
      }
      if (go_pop) {
        if (trail == 0u) {
          v.match_pos[qi] = pos;
          v.match_d2[qi] = pos < 0 ? kInfF : key_dist(key);
          phase = PH_IDLE;
        } else {
          const int up = __ffs(trail) - 1;
          node >>= up; depth -= up; trail >>= up;  // node: the child whose sibling is pending
          mode = 2u | ((node & 1u) ^ 1u);
          node >>= 1; --depth; trail >>= 1;        // its parent; the sibling is entered from there
          phase = PH_PAIR;
        }
      }
    }
  }
}

// match_kernel with the seed-path search (knn_seed_path): one sibling box per level instead of the
// two child boxes of a descent, no separate climb.
__global__ void PGS_MATCH_BOUNDS
match_path_kernel(const PairView* __restrict__ views, PairState* __restrict__ states, float maxr2) {
  PairState& st = states[blockIdx.y];
  if (!st.active) return;
  const PairView v = views[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.n_r) return;
  const Xf T = st.xf;
  float4 r = v.reading[i];
  float3 q = xform_rn(T, r.x, r.y, r.z);
  Best1 acc;
  acc.init(maxr2);
  const int pp = st.iterations > 0 ? v.match_pos[i] : -1;
  if (pp >= 0) {
    const float4 c = __ldg(v.tree.pts + pp);
    acc.offer(dist2_rn(q.x, q.y, q.z, c.x, c.y, c.z), __float_as_int(c.w), pp);
    knn_seed_path(v.tree, pp / kLeaf, q.x, q.y, q.z, acc);
  } else {
    knn_traverse(v.tree, q.x, q.y, q.z, acc);
  }
  v.match_pos[i] = acc.pos;
  v.match_d2[i] = acc.dist();
}

// ---------------------------------------------------------------------------
// Queue matcher.  ncu on match_kernel (profiles/r2_match_blocks.md): the first descent and the
// first leaf scan run with all 32 lanes, everything after them - revisits of pending siblings,
// whose number is heavy-tailed - with ~8.  Here a warp owns kMqBatches x 32 consecutive queries
// and alternates between
//   phase 1  32 new queries in lock step: transform, previous match as the bound, descent
//            from the root to the first leaf, leaf scan.  A query with no pending sibling is
//            finished; the others are pushed onto the warp's shared-memory queue;
//   phase 2  every lane pops a queued search and runs one revisit (walk back to the next
//            pending sibling whose box still qualifies, descend, scan) per trip, popping the
//            next search as soon as its own ends - the lanes stay busy whatever the length
//            of the individual walks.  When fewer than kMqLow searches are left the in-flight
//            ones are parked in the queue and the warp produces 32 more.
// Same exact (distance, index) minimum as match_kernel: only the schedule differs.
// ---------------------------------------------------------------------------
#ifndef PGS_MQ_MIN_BLOCKS
#define PGS_MQ_MIN_BLOCKS 12
#endif
constexpr int kMqCap = 64;  // queue slots per warp (parked <= kMqLow + 32 new survivors)
constexpr int kMqLow = 16;

struct MqQueue {
  float qx[kMqCap], qy[kMqCap], qz[kMqCap];
  unsigned long long key[kMqCap];
  int pos[kMqCap], qi[kMqCap], depth[kMqCap];
  unsigned node[kMqCap], trail[kMqCap];
};

struct MqState {
  float qx, qy, qz;
  unsigned long long key;
  int pos, qi, depth;
  unsigned node, trail;
};

__device__ __forceinline__ void mq_push(MqQueue& Q, int& qlen, bool want, const MqState& s, unsigned lt_mask) {
  const unsigned b = __ballot_sync(0xffffffffu, want);
  if (want) {
    const int slot = qlen + __popc(b & lt_mask);
    Q.qx[slot] = s.qx; Q.qy[slot] = s.qy; Q.qz[slot] = s.qz;
    Q.key[slot] = s.key; Q.pos[slot] = s.pos; Q.qi[slot] = s.qi; Q.depth[slot] = s.depth;
    Q.node[slot] = s.node; Q.trail[slot] = s.trail;
  }
  qlen += __popc(b);
  __syncwarp();
}

__device__ __forceinline__ void mq_pop(const MqQueue& Q, int& qlen, bool& has, MqState& s, unsigned lt_mask) {
  const unsigned b = __ballot_sync(0xffffffffu, !has);
  const int n = min(__popc(b), qlen);
  const int rank = __popc(b & lt_mask);
  if (!has && rank < n) {
    const int slot = qlen - 1 - rank;
    s.qx = Q.qx[slot]; s.qy = Q.qy[slot]; s.qz = Q.qz[slot];
    s.key = Q.key[slot]; s.pos = Q.pos[slot]; s.qi = Q.qi[slot]; s.depth = Q.depth[slot];
    s.node = Q.node[slot]; s.trail = Q.trail[slot];
    has = true;
  }
  qlen -= n;
  __syncwarp();
}

// descend from (node, depth) while the nearer child qualifies, then scan the leaf reached
__device__ __forceinline__ void mq_descend_scan(const TreeView& t, MqState& s) {
  const ulonglong2* __restrict__ nodes16 = reinterpret_cast<const ulonglong2*>(t.nodes);
  const QueryPk q = pack_query(s.qx, s.qy, s.qz);
  bool at_leaf = true;
  while (s.depth < t.depth) {
    const ulonglong2* __restrict__ c = nodes16 + (size_t)s.node * 3;
    const ulonglong2 a = __ldg(c), b = __ldg(c + 1), e = __ldg(c + 2);
    const float lb0 = box_lb_packed(q, a.x, a.y, b.x);
    const float lb1 = box_lb_packed(q, b.y, e.x, e.y);
    const float bound = key_dist(s.key);
    const bool near1 = lb1 < lb0;
    const float lbn = near1 ? lb1 : lb0, lbf = near1 ? lb0 : lb1;
    if (!(lbn <= bound)) { at_leaf = false; break; }
    s.trail = (s.trail << 1) | ((lbf <= bound) ? 1u : 0u);
    s.node = s.node * 2 + (near1 ? 1u : 0u);
    ++s.depth;
  }
  if (at_leaf) {
    const int leaf = (int)s.node - t.P;
    if (leaf < t.n_leaves) {
      const float4* __restrict__ lp = t.pts + (size_t)leaf * kLeaf;
#pragma unroll
      for (int j = 0; j < kLeaf; ++j) {
        const float4 p = __ldg(lp + j);
        const unsigned long long nk = make_key(dist2_rn_packed(q.xy, s.qz, p.x, p.y, p.z), __float_as_int(p.w));
        if (nk < s.key) { s.key = nk; s.pos = leaf * kLeaf + j; }
      }
    }
  }
}

// walk back to the deepest pending sibling whose box still qualifies; false = the search is over
__device__ __forceinline__ bool mq_walk_back(const TreeView& t, MqState& s) {
  const unsigned long long* __restrict__ nodes8 = reinterpret_cast<const unsigned long long*>(t.nodes);
  const QueryPk q = pack_query(s.qx, s.qy, s.qz);
  while (s.trail != 0u) {
    const int up = __ffs(s.trail) - 1;
    s.node >>= up; s.depth -= up; s.trail >>= up;
    s.node ^= 1u; s.trail ^= 1u;
    const unsigned long long* __restrict__ nb = nodes8 + (size_t)s.node * 3;
    if (box_lb_packed(q, __ldg(nb), __ldg(nb + 1), __ldg(nb + 2)) <= key_dist(s.key)) return true;
  }
  return false;
}

template <int kMinBlocks>
__global__ void __launch_bounds__(128, kMinBlocks)
match_queue_kernel(const PairView* __restrict__ views, PairState* __restrict__ states, float maxr2, int batches) {
  __shared__ MqQueue queues[4];
  PairState& st = states[blockIdx.y];
  if (!st.active) return;
  const PairView v = views[blockIdx.y];
  const unsigned FULL = 0xffffffffu;
  const int warp = threadIdx.x >> 5;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int first = (blockIdx.x * 4 + warp) * batches * 32;  // this warp's queries
  if (first >= v.n_r) return;
  const int nb = min(batches, (v.n_r - first + 31) / 32);
  const bool seeded = st.iterations > 0;
  MqQueue& Q = queues[warp];
  int qlen = 0;
  bool has = false;
  MqState s;
  s.qx = s.qy = s.qz = 0.f; s.key = 0ull; s.pos = -1; s.qi = -1; s.depth = 0; s.node = 1u; s.trail = 0u;
  for (int b = 0; b < nb; ++b) {
    // ---- park what is in flight, then phase 1 on 32 new queries -----------------------------
    mq_push(Q, qlen, has, s, lt_mask);
    has = false;
    const int i = first + b * 32 + (int)lane;
    bool survivor = false;
    if (i < v.n_r) {
      const Xf T = st.xf;
      const float4 r = v.reading[i];
      const float3 q = xform_rn(T, r.x, r.y, r.z);
      s.qx = q.x; s.qy = q.y; s.qz = q.z;
      s.key = make_key(maxr2, 0x7fffffff);
      s.pos = -1; s.qi = i; s.node = 1u; s.depth = 0; s.trail = 0u;
      const int pp = seeded ? v.match_pos[i] : -1;
      if (pp >= 0) {
        const float4 c = __ldg(v.tree.pts + pp);
        const unsigned long long nk = make_key(dist2_rn(q.x, q.y, q.z, c.x, c.y, c.z), __float_as_int(c.w));
        if (nk < s.key) { s.key = nk; s.pos = pp; }
      }
      mq_descend_scan(v.tree, s);
      if (s.trail == 0u) {
        v.match_pos[i] = s.pos;
        v.match_d2[i] = s.pos < 0 ? kInfF : key_dist(s.key);
      } else {
        survivor = true;
      }
    }
    mq_push(Q, qlen, survivor, s, lt_mask);
    // ---- phase 2 ----------------------------------------------------------------------------
    const bool last = b == nb - 1;
    for (;;) {
      mq_pop(Q, qlen, has, s, lt_mask);
      const int busy = __popc(__ballot_sync(FULL, has));
      if (busy == 0) break;
      if (!last && busy + qlen < kMqLow) break;
      if (has) {
        if (mq_walk_back(v.tree, s)) {
          mq_descend_scan(v.tree, s);
        } else {
          v.match_pos[s.qi] = s.pos;
          v.match_d2[s.qi] = s.pos < 0 ? kInfF : key_dist(s.key);
          has = false;
        }
      }
    }
  }
}

// KDTreeMatcher knn > 1: every reading point keeps its K nearest (ascending, ties -> lower
// original index); matches are stored [point][neighbour] as sorted positions.  The climb is
// seeded from the leaf of the previous nearest match.
template <int K>
__global__ void __launch_bounds__(128)
match_k_kernel(const PairView* __restrict__ views, PairState* __restrict__ states, float maxr2, int k) {
  PairState& st = states[blockIdx.y];
  if (!st.active) return;
  const PairView v = views[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.n_r) return;
  const Xf T = st.xf;
  float4 r = v.reading[i];
  float3 q = xform_rn(T, r.x, r.y, r.z);
  BestK<K> acc;
  acc.init(maxr2);
  // same seeded search as match_kernel, from the previous NEAREST match
  const int pp = st.iterations > 0 ? v.match_pos[(size_t)i * k] : -1;
  if (pp >= 0) {
    const float4 c = __ldg(v.tree.pts + pp);
    acc.offer(dist2_rn(q.x, q.y, q.z, c.x, c.y, c.z), __float_as_int(c.w), pp);
    const int h = min(kSeedLevels, v.tree.depth);
    knn_traverse_from(v.tree, (unsigned)(v.tree.P + pp / kLeaf) >> h, v.tree.depth - h, q.x, q.y, q.z, acc, pp, pp);
    knn_climb(v.tree, pp / kLeaf, q.x, q.y, q.z, acc, pp, pp, h, false);
  } else {
    knn_traverse(v.tree, q.x, q.y, q.z, acc);
  }
  int* op = v.match_pos + (size_t)i * k;
  float* od = v.match_d2 + (size_t)i * k;
#pragma unroll
  for (int e = 0; e < K; ++e) {
    if (e < k) {
      const int id = key_id(acc.key[e]);
      const bool found = id != 0x7fffffff;
      op[e] = found ? v.ref_inv[id] : -1;
      od[e] = found ? key_dist(acc.key[e]) : kInfF;
    }
  }
}

// ---- re-ordering of the reading by matched leaf ---------------------------------------------
// The matcher is bound by L1 wavefronts (profiles/r2_match_l1.md): the lanes of a warp walk
// different nodes.  After the first iteration every query has a match, and matches move little
// afterwards, so the queries are re-ordered ONCE by the leaf of their first match (stable, so
// Morton order survives inside a leaf): the lanes of a warp then share their seed leaf, the path
// down to it and most of its neighbourhood.  The order depends only on the pair's own data.
__global__ void __launch_bounds__(256)
resort_keys_kernel(const PairView* __restrict__ views, const PairState* __restrict__ states, uint32_t* __restrict__ keys,
                   uint32_t* __restrict__ vals, int stride) {
  const PairView v = views[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.n_r) return;
  const int pp = states[blockIdx.y].active ? v.match_pos[i] : 0;
  keys[(size_t)blockIdx.y * stride + i] = pp >= 0 ? (uint32_t)(pp / kLeaf) : (uint32_t)v.tree.P;
  vals[(size_t)blockIdx.y * stride + i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256)
resort_gather_kernel(const PairView* __restrict__ views, const uint32_t* __restrict__ order, int stride,
                     float4* __restrict__ t4, int* __restrict__ ti, float* __restrict__ tf) {
  const PairView v = views[blockIdx.y];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= v.n_r) return;
  const size_t o = (size_t)blockIdx.y * stride + j;
  const uint32_t src = order[o];
  t4[o] = v.reading[src];
  ti[o] = v.match_pos[src];
  tf[o] = v.match_d2[src];
}

__global__ void __launch_bounds__(256)
resort_store_kernel(const PairView* __restrict__ views, int stride, const float4* __restrict__ t4,
                    const int* __restrict__ ti, const float* __restrict__ tf) {
  const PairView v = views[blockIdx.y];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= v.n_r) return;
  const size_t o = (size_t)blockIdx.y * stride + j;
  const_cast<float4*>(v.reading)[j] = t4[o];
  v.match_pos[j] = ti[o];
  v.match_d2[j] = tf[o];
}

__global__ void deactivate_kernel(PairState* st, int status) {
  st->active = 0;
  st->status = status;
}

__global__ void inverse_positions_kernel(const float4* __restrict__ sorted_pts, int n, int* __restrict__ inv) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) inv[__float_as_int(sorted_pts[j].w)] = j;
}

// Exact quantile of the valid match distances (Matches::getDistsQuantile, A.3)
// as a 3-pass MSB radix select over the fp32 bit patterns (11 + 11 + 10 bits;
// non-negative floats order like unsigned ints).  Each pass is one multi-block
// kernel: blocks histogram their slice in shared memory, merge into the pair's
// global 2048-bin histogram, and the last block to finish picks the digit that
// holds the wanted rank.  An exact order statistic is algorithm-independent, so
// the limit is bit-identical to the reference's nth_element.
constexpr int kSelBins = 2048;
constexpr int kSelBlocks = 16;

__global__ void __launch_bounds__(256)
select_pass_kernel(const PairView* __restrict__ views, PairState* __restrict__ states, IcpParams P, int jq, int pass,
                   int* n_active, volatile int* h_done) {
  __shared__ unsigned hist[kSelBins];
  __shared__ bool last;
  PairState& st = states[blockIdx.y];
  if (!st.active) return;
  const PairView v = views[blockIdx.y];
  const int tid = threadIdx.x, lane = tid & 31;
  const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
  const unsigned dmask = pass == 2 ? 1023u : 2047u;
  const unsigned prefix = pass == 0 ? 0u : st.sel_prefix;
  const unsigned mask = pass == 0 ? 0u : st.sel_mask;
  for (int d = tid; d < kSelBins; d += 256) hist[d] = 0;
  __syncthreads();
  // four tiles per trip: their loads are issued together (the kernel is one wave of blocks
  // waiting on DRAM, not on arithmetic)
  for (int base = blockIdx.x * 256; base < v.n_m; base += gridDim.x * 256 * 4) {
    unsigned uu[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = base + r * (int)gridDim.x * 256 + tid;
      uu[r] = i < v.n_m ? __float_as_uint(v.match_d2[i]) : 0u;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const unsigned u = uu[r];
      // getDistsQuantile: dist != inf and dist > 0 (an out-of-range slot reads as 0)
      const bool valid = (u != 0u) && (u < 0x7f800000u) && ((u & mask) == prefix);
      unsigned act = __ballot_sync(0xffffffffu, valid);
      if (valid) {
        unsigned d = (u >> shift) & dmask;
        unsigned m = __match_any_sync(act, d);
        if (lane == __ffs(m) - 1) atomicAdd(&hist[d], (unsigned)__popc(m));
      }
    }
  }
  __syncthreads();
  for (int d = tid; d < kSelBins; d += 256)
    if (hist[d]) atomicAdd(&v.sel_hist[d], hist[d]);
  __threadfence();
  __syncthreads();
  if (tid == 0) last = (atomicAdd(&st.ticket3, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  // the last block owns the merged histogram: read it, then clear it for the next pass
  for (int d = tid; d < kSelBins; d += 256) {
    hist[d] = __ldcg(&v.sel_hist[d]);
    v.sel_hist[d] = 0;
  }
  __syncthreads();
  // the digit that holds the wanted rank, found by the first warp: lane l owns bins 64l .. 64l+63
  if (tid >= 32) return;
  unsigned t = 0;
  for (int d = lane * 64; d < lane * 64 + 64; ++d) t += hist[d];
  unsigned incl = t;  // inclusive prefix over the lanes
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  const unsigned M = __shfl_sync(0xffffffffu, incl, 31);
  unsigned long long rank = st.sel_rank;
  if (pass == 0) {
    const double q = P.q_ratio[jq];
    // upstream multiplies in T = float: values.size() * quantile (Matches.cpp getDistsQuantile)
    rank = (q == 1.0) ? (M ? M - 1 : 0) : (unsigned long long)__fmul_rn((float)M, (float)q);
    if (M && rank >= M) rank = M - 1;
  }
  if (pass == 0 && M == 0) {  // "no outlier to filter"
    if (lane == 0) {
      st.ticket3 = 0;
      st.status = PGS_CONVERGENCE_ERROR;
      st.active = 0;
      pair_finished(n_active, h_done);
    }
    return;
  }
  // first lane whose inclusive count exceeds the rank (the last lane if none does)
  const unsigned over = __ballot_sync(0xffffffffu, (unsigned long long)incl > rank);
  const int l = over ? __ffs(over) - 1 : 31;
  const unsigned before = __shfl_sync(0xffffffffu, incl - t, l);
  // inside that lane's 64 bins: two bins per lane
  const unsigned h0 = hist[l * 64 + 2 * lane], h1 = hist[l * 64 + 2 * lane + 1];
  unsigned inc2 = h0 + h1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned up = __shfl_up_sync(0xffffffffu, inc2, o);
    if (lane >= o) inc2 += up;
  }
  const unsigned long long rel = rank - before;  // rank inside lane l's bins
  const unsigned over2 = __ballot_sync(0xffffffffu, (unsigned long long)inc2 > rel);
  const int l2 = over2 ? __ffs(over2) - 1 : 31;
  if (lane == l2) {
    const unsigned ex2 = inc2 - (h0 + h1);
    // the last bin of the range is taken when nothing exceeds the rank (cannot happen for a
    // consistent histogram; kept as the bound the serial scan had)
    const bool second = over2 ? ((unsigned long long)ex2 + h0 <= rel) : true;
    const int d = l * 64 + 2 * lane + (second ? 1 : 0);
    const unsigned cum = before + ex2 + (second ? h0 : 0u);
    st.ticket3 = 0;
    st.sel_rank = rank - cum;
    const unsigned np = prefix | ((unsigned)d << shift);
    st.sel_prefix = np;
    st.sel_mask = mask | (dmask << shift);
    if (pass == 2) {
      const float limit = __fmul_rn(P.q_factor[jq], __uint_as_float(np));
      const float hi = jq == 0 ? P.fixed_hi : st.lim_hi;
      st.lim_hi = fminf(hi, limit);
      st.lim_lo = P.fixed_lo;
    }
  }
}

void launch_match_k(int k, dim3 grid, cudaStream_t s, const PairView* views, PairState* states, float maxr2) {
  if (k == 2) match_k_kernel<2><<<grid, 128, 0, s>>>(views, states, maxr2, k);
  else if (k == 3) match_k_kernel<3><<<grid, 128, 0, s>>>(views, states, maxr2, k);
  else if (k == 4) match_k_kernel<4><<<grid, 128, 0, s>>>(views, states, maxr2, k);
  else if (k <= 6) match_k_kernel<6><<<grid, 128, 0, s>>>(views, states, maxr2, k);
  else if (k <= 8) match_k_kernel<8><<<grid, 128, 0, s>>>(views, states, maxr2, k);
  else if (k <= 12) match_k_kernel<12><<<grid, 128, 0, s>>>(views, states, maxr2, k);
  else if (k <= 16) match_k_kernel<16><<<grid, 128, 0, s>>>(views, states, maxr2, k);
  else if (k <= 24) match_k_kernel<24><<<grid, 128, 0, s>>>(views, states, maxr2, k);
  else match_k_kernel<32><<<grid, 128, 0, s>>>(views, states, maxr2, k);
}

// ---------------------------------------------------------------------------
// VarTrimmedDistOutlierFilter (optimizeInlierRatio): sort the valid squared
// distances (batched radix sort, invalid ones keyed to the end), then one block
// per pair runs the fp64 running sum and the FRMS criterion
//     FRMS(j) = (1 / (j/N)^lambda)^2 * (1/j) * (s_1 + ... + s_j)
// over j in (floor(minRatio N), floor(maxRatio N)], takes the first minimum and
// reads the quantile at the optimised ratio straight from the sorted array.
// ---------------------------------------------------------------------------
struct VarIn {
  const float* d2;
  int n;              // k x N_r
  const int* active;  // null = always
};

__global__ void __launch_bounds__(256)
var_keys_kernel(const VarIn* __restrict__ in, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, int stride) {
  const VarIn job = in[blockIdx.y];
  if (job.active && !*job.active) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= job.n) return;
  const unsigned u = __float_as_uint(job.d2[i]);
  const bool valid = (u != 0u) && (u < 0x7f800000u);  // dist > 0 and dist != inf
  keys[(size_t)blockIdx.y * stride + i] = valid ? u : 0xffffffffu;
  vals[(size_t)blockIdx.y * stride + i] = (uint32_t)i;
}

__global__ void __launch_bounds__(1024)
var_limit_kernel(const VarIn* __restrict__ in, const uint32_t* __restrict__ sorted, int stride, float min_ratio,
                 float max_ratio, double lambda, float* __restrict__ limit_out, int* __restrict__ fail_out) {
  __shared__ double warp_tot[32];
  __shared__ double best_v[32];
  __shared__ int best_i[32];
  __shared__ int s_M;
  const VarIn job = in[blockIdx.x];
  if (job.active && !*job.active) return;
  const uint32_t* __restrict__ s = sorted + (size_t)blockIdx.x * stride;
  const int n = job.n, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) {  // number of valid distances: first position holding the invalid key
    int lo = 0, hi = n;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s[mid] == 0xffffffffu) hi = mid; else lo = mid + 1;
    }
    s_M = lo;
  }
  __syncthreads();
  const int M = s_M;
  if (M == 0) {
    if (tid == 0) { fail_out[blockIdx.x] = 1; limit_out[blockIdx.x] = 0.f; }
    return;
  }
  int min_el = (int)floorf(__fmul_rn(min_ratio, (float)n));
  int max_el = (int)floorf(__fmul_rn(max_ratio, (float)n));
  if (max_el > M) max_el = M;
  if (min_el > max_el - 1) min_el = max_el - 1;
  if (min_el < 0) min_el = 0;
  const int chunk = (max_el + 1023) / 1024;
  const int j0 = min(tid * chunk, max_el), j1 = min(j0 + chunk, max_el);
  double sum = 0.0;
  for (int j = j0; j < j1; ++j) sum += (double)__uint_as_float(s[j]);
  // exclusive scan of the 1024 chunk sums
  double incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    double w = warp_tot[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    warp_tot[lane] = wi - w;  // exclusive
  }
  __syncthreads();
  double cum = warp_tot[wid] + (incl - sum);
  double bv = __longlong_as_double(0x7ff0000000000000ll);
  int bi = 0x7fffffff;
  for (int j = j0; j < j1; ++j) {
    cum += (double)__uint_as_float(s[j]);
    if (j < min_el) continue;
    const double id = (double)(j + 1);
    const double deno = pow(id / (double)n, lambda);
    const double inv = 1.0 / deno;
    const double frms = inv * inv * (1.0 / id) * cum;
    if (frms < bv) { bv = frms; bi = j; }
  }
  // first minimum over the block: smaller value, then lower index
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) { best_v[wid] = bv; best_i[wid] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 32; ++w)
      if (best_v[w] < bv || (best_v[w] == bv && best_i[w] < bi)) { bv = best_v[w]; bi = best_i[w]; }
    if (bi == 0x7fffffff) bi = min_el;
    const float ratio = __fdiv_rn((float)bi, (float)n);
    const double q = (double)ratio;
    unsigned long long rank = (q == 1.0) ? (unsigned long long)(M - 1) : (unsigned long long)__fmul_rn((float)M, ratio);
    if (rank >= (unsigned long long)M) rank = M - 1;
    limit_out[blockIdx.x] = __uint_as_float(s[rank]);
    fail_out[blockIdx.x] = 0;
  }
}

// fold the VarTrimmed limit of quantile slot jq into the pair's limits (what the third
// select pass does for a fixed ratio)
__global__ void var_apply_kernel(PairState* __restrict__ states, IcpParams P, int jq, const float* __restrict__ limit,
                                 const int* __restrict__ fail, int n_pairs, int* n_active, volatile int* h_done) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  PairState& st = states[p];
  if (!st.active) return;
  if (fail[p]) {  // "no outlier to filter"
    st.status = PGS_CONVERGENCE_ERROR;
    st.active = 0;
    pair_finished(n_active, h_done);
    return;
  }
  const float hi = jq == 0 ? P.fixed_hi : st.lim_hi;
  st.lim_hi = fminf(hi, __fmul_rn(P.q_factor[jq], limit[p]));
  st.lim_lo = P.fixed_lo;
}

// no quantile-based filter in the chain: the limits are the fixed ones
__global__ void fixed_limits_kernel(PairState* __restrict__ states, IcpParams P, int n_pairs) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  states[p].lim_hi = P.fixed_hi;
  states[p].lim_lo = P.fixed_lo;
}

#ifndef PGS_ACC_MIN_BLOCKS
#define PGS_ACC_MIN_BLOCKS 2
#endif
__global__ void __launch_bounds__(256, PGS_ACC_MIN_BLOCKS)
accumulate_kernel(const PairView* __restrict__ views, PairState* __restrict__ states, IcpParams P, int* n_active,
                  volatile int* h_done, unsigned* __restrict__ pm_ticket) {
  __shared__ double sh[8][kAcc];
  __shared__ double tot[kAcc];
  __shared__ bool last;
  // the persistent matcher's ticket counter is cleared here for the next iteration (this kernel
  // runs between two match launches of the same stream)
  if (pm_ticket && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *pm_ticket = 0u;
  PairState& st = states[blockIdx.y];
  if (!st.active) return;
  const PairView v = views[blockIdx.y];
  const unsigned slices = (unsigned)reduce_slices(v.n_m, kAccPerBlock, kAccMaxBlocks);
  if (blockIdx.x >= slices) return;
  const Xf T = st.xf;
  const float lo = st.lim_lo, hi = st.lim_hi;
  // The kernel holds 30 fp64 accumulators (128 registers, 2 blocks / SM), so there are few warps
  // to hide the three dependent DRAM latencies of a match (position + distance -> reading point
  // -> matched record): with one match at a time they were the whole run time (ncu: long
  // scoreboard on the first use of every load, issue slots 24 % busy).  kAccUnroll matches are
  // therefore loaded together - their streaming loads first, then an L2 prefetch of each matched
  // record - before any arithmetic; the matches are still accumulated in the same order.
  const float4* __restrict__ rec = v.ref_normals ? v.ref_normals : v.tree.pts;
  const int rstride = v.ref_normals ? 2 : 1;
  const int stride = (int)slices * blockDim.x;
  double acc[kAcc];
#pragma unroll
  for (int j = 0; j < kAcc; ++j) acc[j] = 0.0;
  for (int m0 = blockIdx.x * blockDim.x + threadIdx.x; m0 < v.n_m; m0 += kAccUnroll * stride) {
    int pos[kAccUnroll];
    float4 r[kAccUnroll];
#pragma unroll
    for (int u = 0; u < kAccUnroll; ++u) {
      const int m = m0 + u * stride;
      pos[u] = -1;
      r[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < v.n_m) {
        const int pp = v.match_pos[m];
        const float d = v.match_d2[m];
        r[u] = v.reading[P.knn == 1 ? m : m / P.knn];
        // ErrorElements skips dist == inf; OutlierFilters weights are 0/1 here
        bool use = pp >= 0 && d < kInfF;
        if (P.has_outliers) use = use && d <= hi && d >= lo;
        pos[u] = use ? pp : -1;
      }
    }
#pragma unroll
    for (int u = 0; u < kAccUnroll; ++u)
      if (pos[u] >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + (size_t)rstride * pos[u]));
#pragma unroll
    for (int u = 0; u < kAccUnroll; ++u) {
      if (pos[u] < 0) continue;
      if (P.has_sn && v.rd_normals && v.ref_normals) {
        const int m = m0 + u * stride;
        float4 a = v.rd_normals[P.knn == 1 ? m : m / P.knn];
        float3 ar = rot_rn(T, a.x, a.y, a.z);
        float4 b = v.ref_normals[2 * pos[u] + 1];
        if (sn_reject(ar.x, ar.y, ar.z, b.x, b.y, b.z, P.sn_eps)) continue;
      }
      float3 p = xform_rn(T, r[u].x, r[u].y, r[u].z);
      // matched point and its normal sit in one 32-byte record when the reference has normals
      float4 q = rec[(size_t)rstride * pos[u]];
      float3 n = make_float3(0.f, 0.f, 0.f);
      if (P.minimizer != MIN_P2POINT) {
        float4 n4 = v.ref_normals[2 * pos[u] + 1];
        n = make_float3(n4.x, n4.y, n4.z);
      }
      accumulate_match(acc, P.minimizer, p, q, n, 1.0, P.force_mode);
    }
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 30; ++j) {
    double s = warp_sum(acc[j]);
    if (lane == 0) sh[w][j] = s;
  }
  __syncthreads();
  if (threadIdx.x < 30) {
    double t = 0.0;
    for (int j = 0; j < 8; ++j) t += sh[j][threadIdx.x];
    v.partials[(size_t)blockIdx.x * kAcc2 + threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(&st.ticket, 1u) == slices - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x < 30) {
    double t = 0.0;
    for (unsigned j = 0; j < slices; ++j) t += __ldcg(&v.partials[(size_t)j * kAcc2 + threadIdx.x]);
    tot[threadIdx.x] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    st.ticket = 0;
    for (int j = 0; j < 30; ++j) st.acc[j] = tot[j];
    finish_iteration(st, P, tot, n_active, h_done);
  }
}

// after the loop: covariance (WithCov) and the sensor-noise overlap, from the
// LAST iteration's matches, limits and transform (what lastErrorElements holds)
__global__ void __launch_bounds__(256)
final_accumulate_kernel(const PairView* __restrict__ views, PairState* __restrict__ states, IcpParams P) {
  __shared__ double sh[8][kAcc2];
  __shared__ bool last;
  PairState& st = states[blockIdx.y];
  if (st.status != PGS_OK || st.iterations == 0) return;
  const PairView v = views[blockIdx.y];
  const unsigned slices = (unsigned)reduce_slices(v.n_m, kAccPerBlock, kAccMaxBlocks);
  if (blockIdx.x >= slices) return;
  const Xf T = st.xf_prev;
  const float lo = st.lim_lo, hi = st.lim_hi;
  const bool want_cov = P.minimizer == MIN_P2PLANE_COV;
  const bool want_overlap = v.rd_normals != nullptr && v.rd_noise != nullptr && P.minimizer != MIN_P2POINT;
  double alpha = 0, beta = 0, gamma = 0, t[3] = {0, 0, 0};
  if (want_cov) {
    const double* Ti = st.T_inc;
    beta = -asin(Ti[2]);
    alpha = atan2(Ti[6], Ti[10]);
    double cb = cos(beta);
    gamma = atan2(Ti[1] / cb, Ti[0] / cb);
    t[0] = Ti[12]; t[1] = Ti[13]; t[2] = Ti[14];
  }
  double acc[kAcc2];
#pragma unroll
  for (int j = 0; j < kAcc2; ++j) acc[j] = 0.0;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < v.n_m; m += slices * blockDim.x) {
    const int i = P.knn == 1 ? m : m / P.knn;
    const int pos = v.match_pos[m];
    const float d = v.match_d2[m];
    bool use = pos >= 0 && d < kInfF;
    if (P.has_outliers) use = use && d <= hi && d >= lo;
    if (!use) continue;
    if (P.has_sn && v.rd_normals && v.ref_normals) {
      float4 a = v.rd_normals[i];
      float3 ar = rot_rn(T, a.x, a.y, a.z);
      float4 b = v.ref_normals[2 * pos + 1];
      if (sn_reject(ar.x, ar.y, ar.z, b.x, b.y, b.z, P.sn_eps)) continue;
    }
    float4 r = v.reading[i];
    float3 p = xform_rn(T, r.x, r.y, r.z);
    float4 q = v.tree.pts[pos];
    if (want_cov) {
      float4 n4 = v.ref_normals[2 * pos + 1];
      accumulate_cov(acc, p, q, make_float3(n4.x, n4.y, n4.z), alpha, beta, gamma, t);
    }
    if (want_overlap) {
      float4 rn = v.rd_normals[i];
      float3 nr = rot_rn(T, rn.x, rn.y, rn.z);
      double n[3] = {(double)nr.x, (double)nr.y, (double)nr.z};
      double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      double e = ((double)p.x - (double)q.x) * (n[0] / nn) + ((double)p.y - (double)q.y) * (n[1] / nn) +
                 ((double)p.z - (double)q.z) * (n[2] / nn);
      if (fabs(e) < (double)v.rd_noise[i]) acc[42] += 1.0;
    }
    acc[43] += 1.0;
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < kAcc2; ++j) {
    double s = warp_sum(acc[j]);
    if (lane == 0) sh[w][j] = s;
  }
  __syncthreads();
  if (threadIdx.x < kAcc2) {
    double s = 0.0;
    for (int j = 0; j < 8; ++j) s += sh[j][threadIdx.x];
    v.partials[(size_t)blockIdx.x * kAcc2 + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(&st.ticket2, 1u) == slices - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x < kAcc2) {
    double s = 0.0;
    for (unsigned j = 0; j < slices; ++j) s += __ldcg(&v.partials[(size_t)j * kAcc2 + threadIdx.x]);
    st.acc2[threadIdx.x] = s;
  }
  if (threadIdx.x == 0) st.ticket2 = 0;
}

__global__ void finish_kernel(const PairView* __restrict__ views, PairState* __restrict__ states, IcpParams P,
                              int n_pairs) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  PairState& st = states[p];
  const PairView v = views[p];
  double tmp[16];
  m4_mul(st.T_iter, st.T_refMean_dataIn, tmp);
  m4_mul(st.T_refIn_refMean, tmp, st.T_out);
  if (st.status != PGS_OK || st.iterations == 0) return;
  const double denom = (double)v.n_m;  // knn * N_r
  const bool want_overlap = v.rd_normals != nullptr && v.rd_noise != nullptr && P.minimizer != MIN_P2POINT;
  st.overlap = want_overlap ? (st.acc2[43] > 0.0 ? st.acc2[42] / st.acc2[43] : 0.0) : st.wsum / denom;
  if (P.minimizer == MIN_P2PLANE_COV) {
    double H[36], DD[36], Hi[36], t1[36];
    for (int c = 0; c < 6; ++c)
      for (int r = 0; r <= c; ++r) {
        H[c * 6 + r] = H[r * 6 + c] = st.acc2[tri(c, r)];
        DD[c * 6 + r] = DD[r * 6 + c] = st.acc2[21 + tri(c, r)];
      }
    inv6_sym(H, Hi);
    for (int c = 0; c < 6; ++c)
      for (int r = 0; r < 6; ++r) {
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s += Hi[k * 6 + r] * DD[c * 6 + k];
        t1[c * 6 + r] = s;
      }
    const double s2 = P.sensor_std_dev * P.sensor_std_dev;
    for (int c = 0; c < 6; ++c)
      for (int r = 0; r < 6; ++r) {
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s += t1[k * 6 + r] * Hi[c * 6 + k];
        st.cov[c * 6 + r] = s2 * s;
      }
  }
}

// ---- fine-grained module kernels (explicit Matches / OutlierWeights) -------
__global__ void __launch_bounds__(256)
weights_kernel(const float* __restrict__ d2, int64_t nk, int has_outliers, float lo, float hi, float* __restrict__ w) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nk) return;
  float d = d2[i];
  // TrimmedDist/MaxDist/MedianDist: dist <= limit; MinDist: dist >= limit;
  // empty filter list: 1 unless dist == inf (A.3)
  w[i] = has_outliers ? ((d <= hi && d >= lo) ? 1.f : 0.f) : ((d == kInfF) ? 0.f : 1.f);
}

struct ExplicitJob {
  const float4* reading;
  const float4* reference;
  const float* ref_normals;  // 3 per point or null
  const float* rd_normals;
  const float* rd_noise;
  const int32_t* ids;
  const float* d2;
  const float* w;
  int64_t n_r;
  int k;
  double* partials;
};

__global__ void __launch_bounds__(256)
explicit_accumulate_kernel(ExplicitJob job, int minimizer, int pass, double alpha, double beta, double gamma,
                           double t0, double t1, double t2, int force_mode) {
  __shared__ double sh[8][kAcc2];
  double acc[kAcc2];
#pragma unroll
  for (int j = 0; j < kAcc2; ++j) acc[j] = 0.0;
  const double t[3] = {t0, t1, t2};
  const int64_t total = job.n_r * job.k;
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < total; m += (int64_t)gridDim.x * blockDim.x) {
    const float d = job.d2[m];
    const float w = job.w[m];
    if (d == kInfF || w == 0.f) continue;
    const int64_t i = m / job.k;
    const int id = job.ids[m];
    float4 r = job.reading[i];
    float4 q = job.reference[id];
    float3 p = make_float3(r.x, r.y, r.z);
    float3 n = make_float3(0.f, 0.f, 0.f);
    if (job.ref_normals) n = make_float3(job.ref_normals[3 * id], job.ref_normals[3 * id + 1], job.ref_normals[3 * id + 2]);
    if (pass == 0) {
      accumulate_match(acc, minimizer, p, q, n, (double)w, force_mode);
    } else {
      if (minimizer == MIN_P2PLANE_COV) accumulate_cov(acc, p, q, n, alpha, beta, gamma, t);
      if (job.rd_normals && job.rd_noise && minimizer != MIN_P2POINT) {
        double nr[3] = {(double)job.rd_normals[3 * i], (double)job.rd_normals[3 * i + 1], (double)job.rd_normals[3 * i + 2]};
        double nn = sqrt(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]);
        double e = ((double)p.x - (double)q.x) * (nr[0] / nn) + ((double)p.y - (double)q.y) * (nr[1] / nn) +
                   ((double)p.z - (double)q.z) * (nr[2] / nn);
        if (fabs(e) < (double)job.rd_noise[i]) acc[42] += 1.0;
      }
      acc[43] += 1.0;
    }
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < kAcc2; ++j) {
    double s = warp_sum(acc[j]);
    if (lane == 0) sh[w][j] = s;
  }
  __syncthreads();
  if (threadIdx.x < kAcc2) {
    double s = 0.0;
    for (int j = 0; j < 8; ++j) s += sh[j][threadIdx.x];
    job.partials[(size_t)blockIdx.x * kAcc2 + threadIdx.x] = s;
  }
}

// single-block quantile for the fine-grained OutlierFilters path
__global__ void __launch_bounds__(1024)
quantile_kernel(const float* __restrict__ d2, int64_t nk, double q, float* __restrict__ out, int* __restrict__ fail) {
  __shared__ unsigned hist[256];
  __shared__ unsigned s_prefix, s_mask, s_fail;
  __shared__ unsigned long long s_rank;
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) { s_prefix = 0; s_mask = 0; s_fail = 0; s_rank = 0; }
  __syncthreads();
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    const unsigned prefix = s_prefix, mask = s_mask;
    for (int64_t base = 0; base < nk; base += 1024) {
      const int64_t i = base + tid;
      unsigned u = 0;
      bool valid = false;
      if (i < nk) {
        u = __float_as_uint(d2[i]);
        valid = (u != 0u) && (u < 0x7f800000u) && ((u & mask) == prefix);
      }
      unsigned act = __ballot_sync(0xffffffffu, valid);
      if (valid) {
        unsigned d = (u >> shift) & 255u;
        unsigned m = __match_any_sync(act, d);
        if (lane == __ffs(m) - 1) atomicAdd(&hist[d], (unsigned)__popc(m));
      }
    }
    __syncthreads();
    if (tid == 0) {
      unsigned long long rank = s_rank;
      if (pass == 0) {
        unsigned long long M = 0;
        for (int d = 0; d < 256; ++d) M += hist[d];
        if (M == 0) s_fail = 1;
        rank = (q == 1.0) ? (M ? M - 1 : 0) : (unsigned long long)__fmul_rn((float)M, (float)q);
        if (M && rank >= M) rank = M - 1;
      }
      unsigned long long cum = 0;
      int d = 0;
      for (; d < 255; ++d) {
        if (cum + hist[d] > rank) break;
        cum += hist[d];
      }
      s_rank = rank - cum;
      s_prefix = prefix | ((unsigned)d << shift);
      s_mask = mask | (255u << shift);
    }
    __syncthreads();
    if (s_fail) break;
  }
  if (tid == 0) {
    *fail = (int)s_fail;
    *out = __uint_as_float(s_prefix);
  }
}

__global__ void solve_only_kernel(PairState* st, const double* acc, IcpParams P, int* n_active,
                                  volatile int* h_done) {
  if (threadIdx.x == 0 && blockIdx.x == 0) finish_iteration(*st, P, acc, n_active, h_done);
}

void solve_only_kernel_launch(Ctx* ctx, PairState* st, const double* acc, const IcpParams& prm, int* n_active) {
  solve_only_kernel<<<1, 32, 0, ctx->stream>>>(st, acc, prm, n_active, ctx->d_progress);
  ctx_count_launches(ctx, 2);
}

__global__ void __launch_bounds__(256)
ratio_kernel(const float* __restrict__ d2, const float* __restrict__ w, int64_t nk, double* __restrict__ out) {
  double kept = 0.0, wsum = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nk; i += (int64_t)gridDim.x * blockDim.x) {
    if (d2[i] != kInfF && w[i] != 0.f) { kept += 1.0; wsum += (double)w[i]; }
  }
  kept = warp_sum(kept);
  wsum = warp_sum(wsum);
  if ((threadIdx.x & 31) == 0) {  // counts of 0/1 weights: exact in fp64 in any order
    atomicAdd(out, kept);
    atomicAdd(out + 1, wsum);
  }
}

}  // namespace

void weights_ratio_device(Ctx* ctx, const float* d_d2, const float* d_w, int64_t nk, double* kept, double* wsum) {
  DBuf<double> out(ctx, 2);
  out.zero();
  if (nk > 0) {
    ratio_kernel<<<std::min(ceil_div(nk, 1024), 296), 256, 0, ctx->stream>>>(d_d2, d_w, nk, out.p);
    ctx_count_launches(ctx, 1);
  }
  double h[2];
  out.download(h, 2);
  ctx->sync();
  *kept = h[0];
  *wsum = h[1];
}

// ===========================================================================
// host side
// ===========================================================================
void outlier_limits_params(const std::vector<Module>& filters, IcpParams* p) {
  p->has_outliers = 0;
  p->n_quant = 0;
  p->fixed_hi = kInfF;
  p->fixed_lo = -kInfF;
  p->has_sn = 0;
  p->sn_eps = -2.f;
  for (auto& f : filters) {
    if (f.name == "NullOutlierFilter") {
      // weight 1 for everything; combined with others it is a no-op.  Alone it
      // keeps even dist == inf, which ErrorElements then skips anyway.
      p->has_outliers = 1;
      continue;
    }
    p->has_outliers = 1;
    if (f.name == "TrimmedDistOutlierFilter" || f.name == "MedianDistOutlierFilter" ||
        f.name == "VarTrimmedDistOutlierFilter") {
      if (p->n_quant == kMaxQuant) throw Error(PGS_INVALID_PARAMETER, "too many quantile-based outlier filters");
      const bool trimmed = f.name[0] == 'T', var = f.name[0] == 'V';
      p->q_ratio[p->n_quant] = trimmed ? f.real("ratio") : 0.5;
      p->q_factor[p->n_quant] = (trimmed || var) ? 1.0f : (float)f.real("factor");
      p->q_var[p->n_quant] = var ? 1 : 0;
      if (var) {
        p->q_min[p->n_quant] = (float)f.real("minRatio");
        p->q_max[p->n_quant] = (float)f.real("maxRatio");
        p->q_lambda[p->n_quant] = f.real("lambda");
      }
      p->n_quant++;
    } else if (f.name == "MaxDistOutlierFilter") {
      float m = (float)f.real("maxDist");
      p->fixed_hi = std::min(p->fixed_hi, m * m);
    } else if (f.name == "MinDistOutlierFilter") {
      float m = (float)f.real("minDist");
      p->fixed_lo = std::max(p->fixed_lo, m * m);
    } else if (f.name == "SurfaceNormalOutlierFilter") {
      // several of them: the largest cosine is the binding one
      p->has_sn = 1;
      p->sn_eps = std::max(p->sn_eps, (float)std::cos(f.real("maxAngle")));
    } else {
      throw Error(PGS_INVALID_ELEMENT, "OutlierFilter " + f.name + " has no device implementation");
    }
  }
}

// PointToPlaneErrorMinimizer{force2D, force4DOF}; force2D wins when both are set, as in
// upstream's compute_in_place (the 2-D branch is taken first)
int force_mode_of(const Module& minimizer) {
  if (minimizer.name == "PointToPointErrorMinimizer") return FORCE_NONE;
  const bool f2 = minimizer.flag("force2D"), f4 = minimizer.flag("force4DOF");
  if (!f2 && !f4) return FORCE_NONE;
  // the covariance estimate is the 6-DOF one whatever the solve was restricted to; force2D cuts the
  // error elements to 2-D upstream, which that formula cannot take
  if (minimizer.name == "PointToPlaneWithCovErrorMinimizer" && f2)
    throw Error(PGS_INVALID_PARAMETER, "PointToPlaneWithCovErrorMinimizer: force2D is not supported with the covariance "
                                       "estimate (force4DOF is)");
  return f2 ? FORCE_2D : FORCE_4DOF;
}

IcpParams params_from_chain(const ChainConfig& cfg) {
  IcpParams p;
  std::memset(&p, 0, sizeof(p));
  const std::string& mn = cfg.minimizer.name;
  if (mn == "PointToPlaneErrorMinimizer") p.minimizer = MIN_P2PLANE;
  else if (mn == "PointToPlaneWithCovErrorMinimizer") p.minimizer = MIN_P2PLANE_COV;
  else if (mn == "PointToPointErrorMinimizer") p.minimizer = MIN_P2POINT;
  else throw Error(PGS_INVALID_ELEMENT, "ErrorMinimizer " + mn + " has no device implementation");
  p.force_mode = force_mode_of(cfg.minimizer);
  p.sensor_std_dev = p.minimizer == MIN_P2PLANE_COV ? cfg.minimizer.real("sensorStdDev") : 0.01;
  p.max_iterations = 0;
  p.has_diff = 0;
  p.has_bound = 0;
  p.smooth_length = 3;
  for (auto& c : cfg.checkers) {
    if (c.name == "CounterTransformationChecker") {
      p.max_iterations = (int)c.integer("maxIterationCount");
      if (p.max_iterations <= 0) p.max_iterations = 1;
    } else if (c.name == "DifferentialTransformationChecker") {
      p.has_diff = 1;
      p.min_diff_rot = c.real("minDiffRotErr");
      p.min_diff_trans = c.real("minDiffTransErr");
      p.smooth_length = (int)c.integer("smoothLength");
    } else if (c.name == "BoundTransformationChecker") {
      p.has_bound = 1;
      p.max_rot = c.real("maxRotationNorm");
      p.max_trans = c.real("maxTranslationNorm");
    }
  }
  p.hard_iteration_cap = p.max_iterations > 0 ? p.max_iterations : 1000;
  outlier_limits_params(cfg.outlier_filters, &p);
  if (cfg.matcher.name != "KDTreeMatcher") throw Error(PGS_INVALID_ELEMENT, "Matcher " + cfg.matcher.name + " has no device implementation");
  p.knn = (int)cfg.matcher.integer("knn");
  if (p.knn < 1 || p.knn > 32) throw Error(PGS_INVALID_PARAMETER, "KDTreeMatcher: knn must be in [1, 32]");
  float md = (float)cfg.matcher.real("maxDist");
  p.max_r2 = std::isinf(md) ? kInfF : md * md;
  return p;
}

IcpEngine::IcpEngine(Ctx* ctx, const ChainConfig& cfg) : ctx_(ctx), cfg_(cfg) { params_ = params_from_chain(cfg); }

void IcpEngine::prepare_references(std::vector<std::unique_ptr<Cloud>>& refs, bool centre_first,
                                   std::vector<std::unique_ptr<PreparedRef>>& out) {
  Ctx* ctx = ctx_;
  cudaStream_t s = ctx->stream;
  const int B = (int)refs.size();
  std::vector<Cloud*> rp(B);
  for (int b = 0; b < B; ++b) rp[b] = refs[b].get();
  DBuf<float> shift(ctx, (size_t)4 * B);
  auto Tmean_sp = std::make_shared<DBuf<double>>(ctx, (size_t)16 * B);
  DBuf<double>& Tmean = *Tmean_sp;
  auto compute_mean = [&]() {
    std::vector<const float4*> pts(B);
    std::vector<int> ns(B);
    int max_n = 0;
    for (int b = 0; b < B; ++b) { pts[b] = rp[b]->feat.p; ns[b] = (int)rp[b]->n; max_n = std::max(max_n, ns[b]); }
    DBuf<const float4*> d_pts(ctx, B);
    DBuf<int> d_n(ctx, B);
    ctx->upload_small(d_pts.p, pts.data(), sizeof(float4*) * B);
    ctx->upload_small(d_n.p, ns.data(), sizeof(int) * B);
    const int nb = reduce_slices(max_n, 2048, 64);
    DBuf<double> partials(ctx, (size_t)B * nb * 3);
    DBuf<unsigned> tickets(ctx, B);
    tickets.zero();
    mean_kernel<<<dim3(nb, B), 256, 0, s>>>(d_pts.p, d_n.p, partials.p, tickets.p, shift.p);
    mean_pose_kernel<<<ceil_div(B, 64), 64, 0, s>>>(shift.p, Tmean.p, B);
    ctx_count_launches(ctx, 2);
  };
  if (centre_first) {
    // ICPSequence::setMap: mean-centre, THEN reference filters (A16)
    compute_mean();
    for (int b = 0; b < B; ++b)
      if (rp[b]->n) {
        rp[b]->touch();
        shift_points_kernel<<<ceil_div(rp[b]->n, 256), 256, 0, s>>>(rp[b]->feat.p, (int)rp[b]->n, shift.p + 4 * b);
        ctx_count_launches(ctx, 1);
      }
    apply_filters(ctx, cfg_.reference_filters, rp);
  } else {
    // ICP::compute: reference filters, THEN mean-centre (§3.3)
    apply_filters(ctx, cfg_.reference_filters, rp);
    compute_mean();
  }
  std::vector<int> ns(B);
  for (int b = 0; b < B; ++b) ns[b] = (int)rp[b]->n;
  std::vector<std::unique_ptr<Index>> idx;
  {
    // the un-shifted index is shared with the SurfaceNormal filter (cached in
    // the cloud); the matcher's mean-centred index is a shifted copy of it
    std::vector<std::shared_ptr<Index>> base;
    cached_indices_for_clouds(ctx, rp, base);
    if (centre_first) {
      // setMap centred the points themselves before the filters: nothing to shift
      idx.resize(B);
      std::vector<const Index*> src(B);
      for (int b = 0; b < B; ++b) src[b] = base[b].get();
      DBuf<float> zero(ctx, (size_t)4 * B);
      zero.zero();
      derive_shifted_indices(ctx, src, zero.p, idx);
    } else {
      std::vector<const Index*> src(B);
      for (int b = 0; b < B; ++b) src[b] = base[b].get();
      derive_shifted_indices(ctx, src, shift.p, idx);
    }
  }
  out.clear();
  std::vector<GatherJob> gjobs;
  int gmax = 0;
  for (int b = 0; b < B; ++b) {
    auto pr = std::make_unique<PreparedRef>();
    pr->index = std::move(idx[b]);
    const Desc* nrm = rp[b]->find("normals");
    pr->has_normals = nrm != nullptr;
    if (nrm && ns[b] > 0) {
      pr->normals_sorted.reset(ctx, (size_t)2 * ns[b]);
      gjobs.push_back(GatherJob{pr->index->pts.p, nrm->data.p, pr->normals_sorted.p, ns[b]});
      gmax = std::max(gmax, ns[b]);
    }
    if (params_.knn > 1 && ns[b] > 0) {
      pr->inv_pos.reset(ctx, (size_t)ns[b]);
      inverse_positions_kernel<<<ceil_div(ns[b], 256), 256, 0, s>>>(pr->index->pts.p, ns[b], pr->inv_pos.p);
      ctx_count_launches(ctx, 1);
    }
    pr->cloud = std::move(refs[b]);
    pr->T_mean = Tmean_sp;
    pr->T_mean_off = (size_t)16 * b;
    out.push_back(std::move(pr));
  }
  DBuf<GatherJob> d_gjobs(ctx, std::max<size_t>(gjobs.size(), 1));
  if (!gjobs.empty()) {
    ctx->upload_small(d_gjobs.p, gjobs.data(), sizeof(GatherJob) * gjobs.size());
    gather_vec3_sorted_batched_kernel<<<dim3(ceil_div(gmax, 256), (unsigned)gjobs.size()), 256, 0, s>>>(d_gjobs.p);
    ctx_count_launches(ctx, 1);
  }
  PGS_LAUNCH_CHECK();  // no host sync: the reference mean never leaves the device
}

namespace {
// clouds that already live on the device
struct ResidentSource : PairSource {
  const std::vector<const Cloud*>& rd;
  const std::vector<const Cloud*>& rf;
  ResidentSource(const std::vector<const Cloud*>& a, const std::vector<const Cloud*>& b) : rd(a), rf(b) {}
  void fetch(Ctx*, int lo, int hi, FetchedPairs& out) override {
    out.readings.assign(rd.begin() + lo, rd.begin() + hi);
    out.references.assign(rf.begin() + lo, rf.begin() + hi);
  }
};
}  // namespace

void IcpEngine::run_batch(const std::vector<const Cloud*>& readings, const std::vector<const Cloud*>& references,
                          const double* T_inits, pgs_icp_result* results) {
  ResidentSource src(readings, references);
  run_batch_source((int)readings.size(), src, T_inits, results);
}

// A batch is cut into contiguous chunks; `batch_streams` workers (own stream + host thread) pull
// the next chunk as soon as they finish one, so the latency-bound tail of one chunk's ICP loop
// overlaps the wide kernels of the others, and a 4096-pair batch (BASELINE C4) never holds more
// than workers x 2 chunks of temporaries.  Results do not depend on the cut.
void IcpEngine::run_batch_source(int P, PairSource& src, const double* T_inits, pgs_icp_result* results) {
  if (P <= 0) return;
  constexpr int kMinPairsPerStream = 4;
  int S = ctx_->profiling ? 1 : std::min(ctx_->batch_streams, P / kMinPairsPerStream);
  if (S <= 1) {
    FetchedPairs f;
    src.fetch(ctx_, 0, P, f);
    src.before_run(ctx_, f);
    run_direct(f.readings, f.references, T_inits, results);
    return;
  }
  int chunk = std::max(1, ctx_->tune.batch_chunk);
  if ((long long)S * chunk > P) chunk = std::max(kMinPairsPerStream, ceil_div(P, S));
  S = std::min(S, ceil_div(P, chunk));
  // everything queued so far on this context's stream (uploads, filters) happens first
  if (!ctx_->fork_ev) PGS_CUDA(cudaEventCreateWithFlags(&ctx_->fork_ev, cudaEventDisableTiming));
  PGS_CUDA(cudaEventRecord(ctx_->fork_ev, ctx_->stream));
  std::vector<std::thread> threads;
  std::vector<std::unique_ptr<Error>> errors(S);
  std::atomic<int> next_pair{0};
  // fixed-size chunks handed out first come, first served.  (Guided self-scheduling - smaller chunks
  // towards the end so that the workers finish together - was measured on 512- and 4096-pair batches:
  // 4 448 vs 4 572 and 4 683 vs 4 702 registrations/s; the small chunks cost more than the tail they trim.)
  auto grab = [&next_pair, P, chunk](int& lo, int& hi) {
    const int cur = next_pair.fetch_add(chunk);
    if (cur >= P) return false;
    lo = cur;
    hi = std::min(P, cur + chunk);
    return true;
  };
  for (int w = 0; w < S; ++w) {
    Ctx* wc = ctx_->worker(w);
    PGS_CUDA(cudaStreamWaitEvent(wc->stream, ctx_->fork_ev, 0));
    threads.emplace_back([this, wc, w, &grab, &src, T_inits, results, &errors]() {
      try {
        PGS_CUDA(cudaSetDevice(wc->device));
        IcpEngine sub(wc, cfg_);
        int lo = 0, hi = 0, nlo = 0, nhi = 0;
        bool have = grab(lo, hi);
        auto fcur = std::make_unique<FetchedPairs>();
        if (have) src.fetch(wc, lo, hi, *fcur);
        while (have) {
          const bool have_next = grab(nlo, nhi);
          auto fnext = std::make_unique<FetchedPairs>();
          if (have_next) src.fetch(wc, nlo, nhi, *fnext);
          src.before_run(wc, *fcur);
          sub.run_direct(fcur->readings, fcur->references, T_inits ? T_inits + (size_t)16 * lo : nullptr, results + lo);
          fcur = std::move(fnext);
          lo = nlo;
          hi = nhi;
          have = have_next;
        }
      } catch (const Error& e) {
        errors[w] = std::make_unique<Error>(e);
      } catch (const std::exception& e) {
        errors[w] = std::make_unique<Error>(PGS_CUDA_ERROR, e.what());
      }
    });
  }
  for (auto& t : threads) t.join();
  for (int w = 0; w < S; ++w) {
    ctx_->launches += ctx_->workers[w]->launches;
    ctx_->workers[w]->launches = 0;
  }
  for (auto& e : errors)
    if (e) throw *e;
  // every worker synchronised its stream before returning its results
}

void IcpEngine::run_direct(const std::vector<const Cloud*>& readings, const std::vector<const Cloud*>& references,
                           const double* T_inits, pgs_icp_result* results) {
  const int P = (int)readings.size();
  if (ctx_->profiling) {
    for (auto& e : idx_ev_)
      if (!e) PGS_CUDA(cudaEventCreate(&e));
    PGS_CUDA(cudaEventRecord(idx_ev_[0], ctx_->stream));
  }
  // PGS_TRACE_STALL=<ms>: report the host-side phases of any (sub-)batch slower than that
  static const double stall_ms = std::getenv("PGS_TRACE_STALL") ? std::atof(std::getenv("PGS_TRACE_STALL")) : 0.0;
  const auto now = [] { return std::chrono::steady_clock::now(); };
  const auto t0 = now();
  std::vector<std::unique_ptr<Cloud>> refs(P);
  for (int p = 0; p < P; ++p) refs[p] = references[p]->clone(ctx_);
  const auto t1 = now();
  std::vector<std::unique_ptr<PreparedRef>> prepared;
  prepare_references(refs, false, prepared);
  const auto t2 = now();
  if (ctx_->profiling) PGS_CUDA(cudaEventRecord(idx_ev_[1], ctx_->stream));
  have_idx_ev_ = ctx_->profiling;
  std::vector<const PreparedRef*> rp(P);
  for (int p = 0; p < P; ++p) rp[p] = prepared[p].get();
  run_prepared(readings, rp, T_inits, results);
  const auto t3 = now();
  if (stall_ms > 0.0) {
    const auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    if (ms(t0, t3) > stall_ms)
      std::fprintf(stderr, "[pgs] slow batch of %d on stream %p: clone %.1f ms, prepare_references %.1f ms, run_prepared %.1f ms\n",
                   P, (void*)ctx_->stream, ms(t0, t1), ms(t1, t2), ms(t2, t3));
  }
}

void IcpEngine::set_map(const Cloud& map) {
  if (map.n == 0) throw Error(PGS_CONVERGENCE_ERROR, "ICPSequence::setMap: empty map");
  std::vector<std::unique_ptr<Cloud>> refs(1);
  refs[0] = map.clone();
  std::vector<std::unique_ptr<PreparedRef>> prepared;
  prepare_references(refs, true, prepared);
  map_ = std::move(prepared[0]);
}

void IcpEngine::run_sequence(const Cloud& reading, const double* T_init, pgs_icp_result* out) {
  if (!map_) throw Error(PGS_INVALID_FIELD, "ICPSequence: no map set (call setMap first)");
  std::vector<const Cloud*> rd{&reading};
  std::vector<const PreparedRef*> rp{map_.get()};
  run_prepared(rd, rp, T_init, out);
}

void IcpEngine::run_prepared(const std::vector<const Cloud*>& readings, const std::vector<const PreparedRef*>& refs,
                             const double* T_inits, pgs_icp_result* results) {
  Ctx* ctx = ctx_;
  cudaStream_t s = ctx->stream;
  const int P = (int)readings.size();
  const IcpParams& prm = params_;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  if (ctx->profiling) {
    for (auto& e : ev) PGS_CUDA(cudaEventCreate(&e));
    PGS_CUDA(cudaEventRecord(ev[0], s));
  }

  // ---- reading side: copy, reading filters ---------------------------------
  std::vector<std::unique_ptr<Cloud>> rds(P);
  std::vector<Cloud*> rdp(P);
  for (int p = 0; p < P; ++p) { rds[p] = readings[p]->clone(ctx); rdp[p] = rds[p].get(); }
  apply_filters(ctx, cfg_.reading_filters, rdp);

  // ---- initial state ---------------------------------------------------------
  std::vector<PairState> hs(P);
  std::vector<const double*> hTm(P);
  int n_active = 0;
  for (int p = 0; p < P; ++p) {
    PairState& st = hs[p];
    std::memset(&st, 0, sizeof(st));
    if (T_inits) std::memcpy(st.T_init, T_inits + 16 * p, 16 * sizeof(double));
    else { st.T_init[0] = st.T_init[5] = st.T_init[10] = st.T_init[15] = 1.0; }
    hTm[p] = refs[p]->T_mean->p + refs[p]->T_mean_off;
    st.status = PGS_OK;
    if (!is_rigid(st.T_init)) st.status = PGS_TRANSFORMATION_ERROR;
    else if (prm.minimizer != MIN_P2POINT && !refs[p]->has_normals) st.status = PGS_INVALID_FIELD;
    else if (rdp[p]->n == 0 || refs[p]->index->n == 0) st.status = PGS_CONVERGENCE_ERROR;
    st.active = st.status == PGS_OK;
    if (st.status == PGS_TRANSFORMATION_ERROR) {  // keep the pose algebra finite
      std::memset(st.T_init, 0, sizeof(st.T_init));
      st.T_init[0] = st.T_init[5] = st.T_init[10] = st.T_init[15] = 1.0;
    }
    n_active += st.active;
  }
  DBuf<PairState> d_states(ctx, P);
  DBuf<const double*> d_Tm(ctx, (size_t)P);
  ctx->upload_small(d_states.p, hs.data(), sizeof(PairState) * P);
  ctx->upload_small(d_Tm.p, hTm.data(), sizeof(double*) * P);
  init_state_kernel<<<ceil_div(P, 64), 64, 0, s>>>(d_states.p, d_Tm.p, P);
  ctx_count_launches(ctx, 1);

  // ---- transformations.apply(reading, T_refMean_dataIn), Morton order --------
  int max_nr = 0;
  {
    std::vector<PreJob> pj(P);
    for (int p = 0; p < P; ++p) {
      Desc* nrm = rdp[p]->find("normals");
      Desc* obs = rdp[p]->find("observationDirections");
      pj[p] = PreJob{rdp[p]->feat.p, nrm ? nrm->data.p : nullptr, obs ? obs->data.p : nullptr, (int)rdp[p]->n};
      max_nr = std::max(max_nr, (int)rdp[p]->n);
    }
    if (max_nr > 0) {
      DBuf<PreJob> d_pj(ctx, P);
      ctx->upload_small(d_pj.p, pj.data(), sizeof(PreJob) * P);
      for (int p = 0; p < P; ++p) rdp[p]->touch();
      pretransform_kernel<<<dim3(ceil_div(max_nr, 256), P), 256, 0, s>>>(d_pj.p, d_states.p);
      ctx_count_launches(ctx, 1);
    }
  }
  // readingStepDataPointsFilters (§3.3): upstream re-applies them every iteration to a fresh
  // copy of the reading as it stands here (filtered, moved by T_refMean_dataIn), before the
  // step transform.  Every filter of this library is a deterministic function of its input,
  // so each iteration would produce this same cloud: apply once.
  if (!cfg_.reading_step_filters.empty()) {
    apply_filters(ctx, cfg_.reading_step_filters, rdp);
    max_nr = 0;
    for (int p = 0; p < P; ++p) {
      max_nr = std::max(max_nr, (int)rdp[p]->n);
      if (rdp[p]->n == 0 && hs[p].active) {  // nothing left to match: "no outlier to filter"
        hs[p].active = 0;
        hs[p].status = PGS_CONVERGENCE_ERROR;
        --n_active;
        deactivate_kernel<<<1, 1, 0, s>>>(d_states.p + p, PGS_CONVERGENCE_ERROR);
        ctx_count_launches(ctx, 1);
      }
    }
  }
  std::vector<std::unique_ptr<Index>> rd_sorted;
  {
    std::vector<const float4*> pts(P);
    std::vector<int> ns(P);
    for (int p = 0; p < P; ++p) { pts[p] = rdp[p]->feat.p; ns[p] = (int)rdp[p]->n; }
    build_indices(ctx, pts, ns, nullptr, rd_sorted, IndexOrder::Morton);
  }

  // ---- per-pair views ----------------------------------------------------------
  const int knn = prm.knn;
  if ((int64_t)max_nr * knn > 0x7fffffff) throw Error(PGS_INVALID_PARAMETER, "knn x reading points exceeds 2^31");
  const int max_nm = max_nr * knn;
  const int acc_blocks = reduce_slices(max_nm, kAccPerBlock, kAccMaxBlocks);
  std::vector<PairView> hv(P);
  std::vector<DBuf<int>> mpos(P);
  std::vector<DBuf<float>> md2(P);
  std::vector<DBuf<double>> partials(P);
  DBuf<unsigned> sel_hist(ctx, (size_t)P * kSelBins);
  sel_hist.zero();
  std::vector<DBuf<float4>> rdn(P);
  std::vector<DBuf<float>> rdnoise(P);
  for (int p = 0; p < P; ++p) {
    const int nr = (int)rdp[p]->n;
    mpos[p].reset(ctx, (size_t)std::max(nr, 1) * knn);
    md2[p].reset(ctx, (size_t)std::max(nr, 1) * knn);
    partials[p].reset(ctx, (size_t)acc_blocks * kAcc2);
    PairView v;
    std::memset(&v, 0, sizeof(v));
    v.reading = rd_sorted[p]->pts.p;
    v.n_r = nr;
    v.n_m = nr * knn;
    v.ref_inv = refs[p]->inv_pos.p;
    v.tree = refs[p]->index->view();
    v.ref_normals = refs[p]->has_normals ? refs[p]->normals_sorted.p : nullptr;
    const Desc* nrm = rdp[p]->find("normals");
    const Desc* noise = rdp[p]->find("simpleSensorNoise");
    const bool for_overlap = nrm && noise && prm.minimizer != MIN_P2POINT;
    if (nrm && nr > 0 && (for_overlap || prm.has_sn)) {
      rdn[p].reset(ctx, (size_t)nr);
      gather_vec3_sorted_kernel<<<ceil_div(nr, 256), 256, 0, s>>>(rd_sorted[p]->pts.p, nr, nrm->data.p, 3, rdn[p].p, nullptr);
      ctx_count_launches(ctx, 1);
      v.rd_normals = rdn[p].p;
    }
    if (for_overlap && nr > 0) {
      rdnoise[p].reset(ctx, (size_t)nr);
      gather_vec3_sorted_kernel<<<ceil_div(nr, 256), 256, 0, s>>>(rd_sorted[p]->pts.p, nr, noise->data.p, 1, nullptr, rdnoise[p].p);
      ctx_count_launches(ctx, 1);
      v.rd_noise = rdnoise[p].p;
    }
    v.match_pos = mpos[p].p;
    v.match_d2 = md2[p].p;
    v.partials = partials[p].p;
    v.sel_hist = sel_hist.p + (size_t)p * kSelBins;
    hv[p] = v;
  }
  DBuf<PairView> d_views(ctx, P);
  ctx->upload_small(d_views.p, hv.data(), sizeof(PairView) * P);

  // ---- VarTrimmedDist work space: sort buffers for the k x N distances of every pair --------
  bool any_var = false;
  for (int jq = 0; jq < prm.n_quant; ++jq) any_var = any_var || prm.q_var[jq];
  const int var_stride = ceil_div(std::max(max_nm, 1), kSortChunk) * kSortChunk;
  DBuf<uint32_t> var_ka, var_kb, var_va, var_vb;
  DBuf<VarIn> d_var_in;
  DBuf<int> d_var_n, var_fail;
  DBuf<float> var_limit;
  if (any_var) {
    var_ka.reset(ctx, (size_t)P * var_stride); var_kb.reset(ctx, (size_t)P * var_stride);
    var_va.reset(ctx, (size_t)P * var_stride); var_vb.reset(ctx, (size_t)P * var_stride);
    std::vector<VarIn> vin(P);
    std::vector<int> vn(P);
    for (int p = 0; p < P; ++p) {
      vin[p] = VarIn{hv[p].match_d2, hv[p].n_m, &d_states.p[p].active};
      vn[p] = hv[p].n_m;
    }
    d_var_in.reset(ctx, P); d_var_n.reset(ctx, P); var_fail.reset(ctx, P); var_limit.reset(ctx, P);
    ctx->upload_small(d_var_in.p, vin.data(), sizeof(VarIn) * P);
    ctx->upload_small(d_var_n.p, vn.data(), sizeof(int) * P);
  }

  // ---- the loop ------------------------------------------------------------------
  ctx->ensure_progress();
  DBuf<int> d_nactive(ctx, 1);
  ctx->upload_small(d_nactive.p, &n_active, sizeof(int));
  *ctx->h_progress = (n_active == 0) ? 1 : 0;
  if (ctx->profiling) PGS_CUDA(cudaEventRecord(ev[1], s));
  int launched = 0;
  std::vector<cudaEvent_t> kev;  // profiling only: 4 marks per iteration
  const int max_it = prm.hard_iteration_cap;
  const dim3 gm(ceil_div(std::max(max_nr, 1), 128), P), ga(acc_blocks, P);
  // the quantile select is an exact order statistic, so its grid may follow the batch: about four
  // blocks per SM over all pairs (a lone pair gets one load trip per thread instead of eight)
  const int sel_blocks = std::max(kSelBlocks, ceil_div(4 * ctx->num_sms, std::max(P, 1)));
  const dim3 gs(std::max(1, std::min(sel_blocks, ceil_div(max_nm, 1024))), P);
  if (prm.n_quant == 0) {
    fixed_limits_kernel<<<ceil_div(P, 64), 64, 0, s>>>(d_states.p, prm, P);
    ctx_count_launches(ctx, 1);
  }
  int match_mode = ctx->tune.match_mode;
  if (match_mode == 1 || match_mode == 3)
    for (int p = 0; p < P; ++p)
      if (!hv[p].tree.cells) match_mode = match_mode == 1 ? 0 : 2;  // an index built before the option was set
  const int pm_blocks = ctx->tune.pm_blocks, pm_refill = ctx->tune.pm_refill;
  const int pm_pair_w = ctx->tune.pm_pair_w, pm_leaf_w = ctx->tune.pm_leaf_w;
  const int resort_it = ctx->tune.resort_it;
  const int pm_ranges = ceil_div(std::max(max_nr, 1), kPmRange);
  const unsigned pm_tickets = (unsigned)pm_ranges * (unsigned)P;
  DBuf<unsigned> pm_ticket;
  if (knn == 1 && (match_mode == 2 || match_mode == 3)) {
    // cleared by accumulate_kernel between two match launches
    pm_ticket.reset(ctx, 1);
    pm_ticket.zero();
  }
  // small references: tensor-core distance tiles instead of the tree (option "dense_max_ref")
  DenseJobs dense_jobs;
  bool use_dense = knn == 1 && ctx->tune.dense_max_ref > 0 && max_nr > 0;
  for (int p = 0; p < P && use_dense; ++p)
    use_dense = refs[p]->index->n > 0 && refs[p]->index->n <= std::min(ctx->tune.dense_max_ref, kDenseMaxRef);
  if (use_dense) {
    std::vector<DenseQuery> dq(P);
    for (int p = 0; p < P; ++p) {
      if (!refs[p]->dense) {
        refs[p]->dense = std::make_unique<DenseRef>();
        dense_prepare(ctx, *refs[p]->index, *refs[p]->dense);
      }
      dq[p] = DenseQuery{refs[p]->dense.get(), hv[p].tree, hv[p].reading, hv[p].n_r, nullptr, hv[p].match_d2, hv[p].match_pos,
                         &d_states.p[p].xf, &d_states.p[p].active};
    }
    dense_upload_jobs(ctx, dq, dense_jobs);
  }
  for (int it = 0; it < max_it; ++it) {
    if (*ctx->h_progress) break;
    if (it >= 2) {
      // stay at most two iterations ahead of the device, then look at the flag
      PGS_CUDA(cudaEventSynchronize(ctx->loop_ev[it & 1]));
      if (*ctx->h_progress) break;
    }
    auto mark = [&]() {
      if (!ctx->profiling) return;
      cudaEvent_t e;
      PGS_CUDA(cudaEventCreate(&e));
      PGS_CUDA(cudaEventRecord(e, s));
      kev.push_back(e);
    };
    mark();
    if (use_dense) {
      dense_launch(ctx, dense_jobs, prm.max_r2);
    } else if (knn == 1) {
      if (match_mode == 0) {
        match_kernel<<<gm, 128, 0, s>>>(d_views.p, d_states.p, prm.max_r2);
      } else if (match_mode == 1) {
        match_cells_kernel<<<gm, 128, 0, s>>>(d_views.p, d_states.p, prm.max_r2);
      } else if (match_mode == 5) {
        match_path_kernel<<<gm, 128, 0, s>>>(d_views.p, d_states.p, prm.max_r2);
      } else if (match_mode == 4) {
        const int mq_batches = std::max(1, ctx->tune.mq_batches);
        const dim3 gq(ceil_div(std::max(max_nr, 1), 128 * mq_batches), P);
        if (ctx->tune.mq_blocks <= 10) match_queue_kernel<10><<<gq, 128, 0, s>>>(d_views.p, d_states.p, prm.max_r2, mq_batches);
        else if (ctx->tune.mq_blocks <= 12) match_queue_kernel<12><<<gq, 128, 0, s>>>(d_views.p, d_states.p, prm.max_r2, mq_batches);
        else match_queue_kernel<16><<<gq, 128, 0, s>>>(d_views.p, d_states.p, prm.max_r2, mq_batches);
      } else {
        const dim3 gp((unsigned)std::min<long long>((long long)ctx->num_sms * pm_blocks, (long long)pm_tickets * 4 + 1));
        if (match_mode == 2)
          match_persistent_kernel<false><<<gp, 128, 0, s>>>(d_views.p, d_states.p, prm.max_r2, pm_ranges, pm_tickets,
                                                           pm_ticket.p, pm_refill, pm_pair_w, pm_leaf_w);
        else
          match_persistent_kernel<true><<<gp, 128, 0, s>>>(d_views.p, d_states.p, prm.max_r2, pm_ranges, pm_tickets,
                                                          pm_ticket.p, pm_refill, pm_pair_w, pm_leaf_w);
      }
    } else {
      launch_match_k(knn, gm, s, d_views.p, d_states.p, prm.max_r2);
    }
    if (it == resort_it && knn == 1 && max_nr > 0) {
      // once, right after the first matches: group the queries by matched leaf (see resort_keys_kernel)
      bool plain = true;
      int max_depth = 0;
      for (int p = 0; p < P; ++p) {
        plain = plain && hv[p].rd_normals == nullptr && hv[p].rd_noise == nullptr;
        max_depth = std::max(max_depth, hv[p].tree.depth);
      }
      if (plain) {
        const int rs_stride = ceil_div(max_nr, kSortChunk) * kSortChunk;
        DBuf<uint32_t> ka(ctx, (size_t)P * rs_stride), kb(ctx, (size_t)P * rs_stride), va(ctx, (size_t)P * rs_stride),
            vb(ctx, (size_t)P * rs_stride);
        DBuf<float4> t4(ctx, (size_t)P * rs_stride);
        DBuf<int> ti(ctx, (size_t)P * rs_stride);
        DBuf<float> tf(ctx, (size_t)P * rs_stride);
        DBuf<int> d_nr(ctx, P);
        std::vector<int> nrs(P);
        for (int p = 0; p < P; ++p) nrs[p] = hv[p].n_r;
        ctx->upload_small(d_nr.p, nrs.data(), sizeof(int) * P);
        const dim3 gr(ceil_div(max_nr, 256), P);
        resort_keys_kernel<<<gr, 256, 0, s>>>(d_views.p, d_states.p, ka.p, va.p, rs_stride);
        const bool in_b = radix_sort_pairs<uint32_t>(ctx, ka.p, kb.p, va.p, vb.p, d_nr.p, P, rs_stride, max_nr, max_depth + 1);
        resort_gather_kernel<<<gr, 256, 0, s>>>(d_views.p, in_b ? vb.p : va.p, rs_stride, t4.p, ti.p, tf.p);
        resort_store_kernel<<<gr, 256, 0, s>>>(d_views.p, rs_stride, t4.p, ti.p, tf.p);
        ctx_count_launches(ctx, 3);
      }
    }
    mark();
    for (int jq = 0; jq < prm.n_quant; ++jq) {
      if (prm.q_var[jq]) {
        var_keys_kernel<<<dim3(ceil_div(std::max(max_nm, 1), 256), P), 256, 0, s>>>(d_var_in.p, var_ka.p, var_va.p, var_stride);
        const bool in_b = radix_sort_pairs<uint32_t>(ctx, var_ka.p, var_kb.p, var_va.p, var_vb.p, d_var_n.p, P,
                                                     var_stride, max_nm, 32);
        var_limit_kernel<<<P, 1024, 0, s>>>(d_var_in.p, in_b ? var_kb.p : var_ka.p, var_stride, prm.q_min[jq],
                                            prm.q_max[jq], prm.q_lambda[jq], var_limit.p, var_fail.p);
        var_apply_kernel<<<ceil_div(P, 64), 64, 0, s>>>(d_states.p, prm, jq, var_limit.p, var_fail.p, P, d_nactive.p,
                                                        ctx->d_progress);
        ctx_count_launches(ctx, 3);
        continue;
      }
      for (int pass = 0; pass < 3; ++pass)
        select_pass_kernel<<<gs, 256, 0, s>>>(d_views.p, d_states.p, prm, jq, pass, d_nactive.p, ctx->d_progress);
    }
    mark();
    accumulate_kernel<<<ga, 256, 0, s>>>(d_views.p, d_states.p, prm, d_nactive.p, ctx->d_progress, pm_ticket.p);
    mark();
    for (int jq = 0; jq < prm.n_quant; ++jq) ctx_count_launches(ctx, prm.q_var[jq] ? 0 : 3);
    ctx_count_launches(ctx, 2);
    PGS_CUDA(cudaEventRecord(ctx->loop_ev[it & 1], s));
    ++launched;
  }
  if (ctx->profiling) PGS_CUDA(cudaEventRecord(ev[2], s));

  // ---- covariance / overlap / final pose -------------------------------------------
  bool need_final = prm.minimizer == MIN_P2PLANE_COV;
  for (int p = 0; p < P; ++p) need_final = need_final || hv[p].rd_noise != nullptr;
  if (need_final) {
    final_accumulate_kernel<<<ga, 256, 0, s>>>(d_views.p, d_states.p, prm);
    ctx_count_launches(ctx, 1);
  }
  finish_kernel<<<ceil_div(P, 64), 64, 0, s>>>(d_views.p, d_states.p, prm, P);
  ctx_count_launches(ctx, 1);
  PGS_LAUNCH_CHECK();
  d_states.download(hs.data(), P);
  if (ctx->profiling) PGS_CUDA(cudaEventRecord(ev[3], s));
  ctx->sync();
  for (int p = 0; p < P; ++p) {
    const PairState& st = hs[p];
    pgs_icp_result& r = results[p];
    std::memset(&r, 0, sizeof(r));
    std::memcpy(r.T, st.T_out, sizeof(r.T));
    std::memcpy(r.covariance, st.cov, sizeof(r.covariance));
    r.iterations = st.iterations;
    r.max_iterations_reached = st.max_reached;
    r.status = st.status;
    r.n_reading = rdp[p]->n;
    r.n_reference = refs[p]->index->n;
    if (st.status == PGS_OK && st.iterations > 0 && rdp[p]->n > 0) {
      r.overlap = st.overlap;
      r.weighted_point_used_ratio = st.wsum / ((double)rdp[p]->n * knn);
      r.point_used_ratio = st.kept / ((double)rdp[p]->n * knn);
      r.residual = st.resid;
    }
  }
  if (ctx->profiling) {
    pgs_stage_times& t = ctx->times;
    PGS_CUDA(cudaEventElapsedTime(&t.filters_ms, ev[0], ev[1]));
    PGS_CUDA(cudaEventElapsedTime(&t.loop_ms, ev[1], ev[2]));
    PGS_CUDA(cudaEventElapsedTime(&t.total_ms, ev[0], ev[3]));
    t.iterations_launched = launched;
    t.index_ms = 0.f;
    if (have_idx_ev_) PGS_CUDA(cudaEventElapsedTime(&t.index_ms, idx_ev_[0], idx_ev_[1]));
    have_idx_ev_ = false;
    t.match_ms = t.select_ms = t.accumulate_ms = 0.f;
    for (size_t i = 0; i + 3 < kev.size(); i += 4) {
      float a = 0, b = 0, c = 0;
      PGS_CUDA(cudaEventElapsedTime(&a, kev[i], kev[i + 1]));
      PGS_CUDA(cudaEventElapsedTime(&b, kev[i + 1], kev[i + 2]));
      PGS_CUDA(cudaEventElapsedTime(&c, kev[i + 2], kev[i + 3]));
      t.match_ms += a;
      t.select_ms += b;
      t.accumulate_ms += c;
      if (std::getenv("PGS_TRACE_LOOP"))
        std::fprintf(stderr, "[pgs] iteration %zu: match %.3f ms, select %.3f ms, accumulate %.3f ms\n", i / 4, a, b, c);
    }
    for (auto& e : kev) cudaEventDestroy(e);
    for (auto& e : ev) cudaEventDestroy(e);
  }
}

// ---------------------------------------------------------------------------
// fine-grained modules
// ---------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
sn_weights_kernel(const float* __restrict__ rd_normals, const float* __restrict__ ref_normals,
                  const int32_t* __restrict__ ids, int64_t nk, int k, float eps, float* __restrict__ w) {
  int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= nk) return;
  const int id = ids[m];
  if (id < 0) { w[m] = 0.f; return; }
  const int64_t i = m / k;
  if (sn_reject(rd_normals[3 * i], rd_normals[3 * i + 1], rd_normals[3 * i + 2], ref_normals[3 * (int64_t)id],
                ref_normals[3 * (int64_t)id + 1], ref_normals[3 * (int64_t)id + 2], eps))
    w[m] = 0.f;
}
}  // namespace

void outlier_weights_device(Ctx* ctx, const std::vector<Module>& filters, const float* d_d2, int64_t nk, float* d_w,
                            const Cloud* reading, const Cloud* reference, const int32_t* d_ids, int k) {
  IcpParams p;
  std::memset(&p, 0, sizeof(p));
  outlier_limits_params(filters, &p);
  float hi = p.fixed_hi, lo = p.fixed_lo;
  if (p.n_quant > 0) {
    DBuf<float> q(ctx, kMaxQuant);
    DBuf<int> fail(ctx, kMaxQuant);
    for (int j = 0; j < p.n_quant; ++j) {
      if (p.q_var[j]) {
        if (nk > 0x7fffffff) throw Error(PGS_INVALID_PARAMETER, "VarTrimmedDistOutlierFilter: too many matches");
        const int n = (int)nk;
        const int stride = ceil_div(std::max(n, 1), kSortChunk) * kSortChunk;
        DBuf<uint32_t> ka(ctx, stride), kb(ctx, stride), va(ctx, stride), vb(ctx, stride);
        DBuf<VarIn> din(ctx, 1);
        DBuf<int> dn(ctx, 1);
        VarIn vin{d_d2, n, nullptr};
        ctx->upload_small(din.p, &vin, sizeof(vin));
        ctx->upload_small(dn.p, &n, sizeof(int));
        if (n > 0) {
          var_keys_kernel<<<dim3(ceil_div(n, 256), 1), 256, 0, ctx->stream>>>(din.p, ka.p, va.p, stride);
          const bool in_b = radix_sort_pairs<uint32_t>(ctx, ka.p, kb.p, va.p, vb.p, dn.p, 1, stride, n, 32);
          var_limit_kernel<<<1, 1024, 0, ctx->stream>>>(din.p, in_b ? kb.p : ka.p, stride, p.q_min[j], p.q_max[j],
                                                        p.q_lambda[j], q.p + j, fail.p + j);
          ctx_count_launches(ctx, 2);
        } else {
          int one = 1;
          ctx->upload_small(fail.p + j, &one, sizeof(int));
        }
        continue;
      }
      quantile_kernel<<<1, 1024, 0, ctx->stream>>>(d_d2, nk, p.q_ratio[j], q.p + j, fail.p + j);
      ctx_count_launches(ctx, 1);
    }
    float hq[kMaxQuant];
    int hf[kMaxQuant];
    q.download(hq, p.n_quant);
    fail.download(hf, p.n_quant);
    ctx->sync();
    for (int j = 0; j < p.n_quant; ++j) {
      if (hf[j]) throw Error(PGS_CONVERGENCE_ERROR, "no outlier to filter");
      volatile float lim = p.q_factor[j] * hq[j];
      hi = std::min(hi, (float)lim);
    }
  }
  if (nk > 0) {
    // a chain made only of SurfaceNormal filters starts from weights 1 (not from "dist != inf")
    bool only_sn = p.has_sn;
    for (auto& f : filters) only_sn = only_sn && f.name == "SurfaceNormalOutlierFilter";
    weights_kernel<<<ceil_div(nk, 256), 256, 0, ctx->stream>>>(d_d2, nk, p.has_outliers, only_sn ? -kInfF : lo,
                                                               only_sn ? kInfF : hi, d_w);
    ctx_count_launches(ctx, 1);
    if (p.has_sn && reading && reference && d_ids) {
      const Desc* a = reading->find("normals");
      const Desc* b = reference->find("normals");
      if (a && b) {  // "surface normals not available: skipping filtering" otherwise
        sn_weights_kernel<<<ceil_div(nk, 256), 256, 0, ctx->stream>>>(a->data.p, b->data.p, d_ids, nk, k, p.sn_eps, d_w);
        ctx_count_launches(ctx, 1);
      }
    }
  }
  PGS_LAUNCH_CHECK();
}

void minimize_device(Ctx* ctx, const Module& minimizer, const Cloud& reading, const Cloud& reference,
                     const int32_t* d_ids, const float* d_d2, const float* d_w, int k, pgs_min_result* out) {
  ChainConfig tmp;
  int mz;
  if (minimizer.name == "PointToPlaneErrorMinimizer") mz = MIN_P2PLANE;
  else if (minimizer.name == "PointToPlaneWithCovErrorMinimizer") mz = MIN_P2PLANE_COV;
  else if (minimizer.name == "PointToPointErrorMinimizer") mz = MIN_P2POINT;
  else throw Error(PGS_INVALID_ELEMENT, "ErrorMinimizer " + minimizer.name + " has no device implementation");
  const Desc* rn = reference.find("normals");
  if (mz != MIN_P2POINT && !rn) throw Error(PGS_INVALID_FIELD, "Cannot find descriptor normals in the reference cloud");
  const Desc* rdn = reading.find("normals");
  const Desc* rdz = reading.find("simpleSensorNoise");
  std::memset(out, 0, sizeof(*out));
  out->T[0] = out->T[5] = out->T[10] = out->T[15] = 1.0;
  const int64_t total = reading.n * k;
  const int nb = std::max(1, std::min(ceil_div(total, 1024), ctx->num_sms));
  DBuf<double> partials(ctx, (size_t)nb * kAcc2);
  ExplicitJob job{reading.feat.p, reference.feat.p, rn ? rn->data.p : nullptr, rdn ? rdn->data.p : nullptr,
                  rdz ? rdz->data.p : nullptr, d_ids, d_d2, d_w, reading.n, k, partials.p};
  std::vector<double> hp((size_t)nb * kAcc2);
  auto reduce = [&](double* acc) {
    partials.download(hp.data(), hp.size());
    ctx->sync();
    for (int j = 0; j < kAcc2; ++j) {
      double s = 0.0;
      for (int b = 0; b < nb; ++b) s += hp[(size_t)b * kAcc2 + j];
      acc[j] = s;
    }
  };
  const int force_mode = force_mode_of(minimizer);
  explicit_accumulate_kernel<<<nb, 256, 0, ctx->stream>>>(job, mz, 0, 0, 0, 0, 0, 0, 0, force_mode);
  ctx_count_launches(ctx, 1);
  PGS_LAUNCH_CHECK();
  double acc[kAcc2];
  reduce(acc);
  if (!(acc[28] > 0.0)) throw Error(PGS_CONVERGENCE_ERROR, "no point to minimize");
  out->kept = (int64_t)acc[28];
  out->point_used_ratio = acc[28] / (double)total;
  out->weighted_point_used_ratio = acc[29] / (double)total;
  out->residual = acc[27];
  // the 6x6 / 3x3 solve of this module-level call runs on the device too, in
  // the same single-thread routine the fused loop uses
  DBuf<PairState> st(ctx, 1);
  PairState hs;
  std::memset(&hs, 0, sizeof(hs));
  hs.T_init[0] = hs.T_init[5] = hs.T_init[10] = hs.T_init[15] = 1.0;
  hs.active = 1;
  ctx->upload_small(st.p, &hs, sizeof(hs));
  DBuf<double> Tm(ctx, 16);
  double I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  ctx->upload_small(Tm.p, I, sizeof(I));
  DBuf<const double*> Tmp(ctx, 1);
  const double* Tm_ptr = Tm.p;
  ctx->upload_small(Tmp.p, &Tm_ptr, sizeof(Tm_ptr));
  init_state_kernel<<<1, 32, 0, ctx->stream>>>(st.p, Tmp.p, 1);
  DBuf<double> d_acc(ctx, kAcc2);
  ctx->upload_small(d_acc.p, acc, sizeof(double) * kAcc2);
  IcpParams prm;
  std::memset(&prm, 0, sizeof(prm));
  prm.minimizer = mz;
  prm.force_mode = force_mode;
  prm.knn = 1;
  prm.hard_iteration_cap = 1 << 30;
  prm.sensor_std_dev = mz == MIN_P2PLANE_COV ? minimizer.real("sensorStdDev") : 0.01;
  ctx->ensure_progress();
  DBuf<int> na(ctx, 1);
  int one = 1 << 20;
  ctx->upload_small(na.p, &one, sizeof(int));
  solve_only_kernel_launch(ctx, st.p, d_acc.p, prm, na.p);
  st.download(&hs, 1);
  ctx->sync();
  if (hs.status != PGS_OK && hs.status != PGS_TRANSFORMATION_ERROR) throw Error(hs.status, "error minimizer failed");
  std::memcpy(out->T, hs.T_inc, sizeof(out->T));
  const bool want_overlap = rdn && rdz && mz != MIN_P2POINT;
  out->overlap = out->weighted_point_used_ratio;
  if (mz == MIN_P2PLANE_COV || want_overlap) {
    const double* Ti = hs.T_inc;
    double beta = -std::asin(Ti[2]);
    double alpha = std::atan2(Ti[6], Ti[10]);
    double cb = std::cos(beta);
    double gamma = std::atan2(Ti[1] / cb, Ti[0] / cb);
    explicit_accumulate_kernel<<<nb, 256, 0, ctx->stream>>>(job, mz, 1, alpha, beta, gamma, Ti[12], Ti[13], Ti[14], force_mode);
    ctx_count_launches(ctx, 1);
    PGS_LAUNCH_CHECK();
    double acc2[kAcc2];
    reduce(acc2);
    if (want_overlap) out->overlap = acc2[43] > 0.0 ? acc2[42] / acc2[43] : 0.0;
    if (mz == MIN_P2PLANE_COV) {
      std::memcpy(hs.acc2, acc2, sizeof(acc2));
      hs.wsum = acc[29];
      hs.iterations = 1;
      hs.status = PGS_OK;
      ctx->upload_small(st.p, &hs, sizeof(hs));
      PairView v;
      std::memset(&v, 0, sizeof(v));
      v.n_r = (int)reading.n;
      v.n_m = (int)total;
      DBuf<PairView> dv(ctx, 1);
      ctx->upload_small(dv.p, &v, sizeof(v));
      finish_kernel<<<1, 32, 0, ctx->stream>>>(dv.p, st.p, prm, 1);
      ctx_count_launches(ctx, 1);
      st.download(&hs, 1);
      ctx->sync();
      std::memcpy(out->covariance, hs.cov, sizeof(out->covariance));
    }
  }
}

}  // namespace pgs
