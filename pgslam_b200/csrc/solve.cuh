// solve.cuh — small dense fp64 algebra that runs on ONE device thread at the
// tail of the accumulation kernel, so that no ICP iteration needs a host
// round trip: 6x6 normal-equation solve, 3x3 SVD, angle-axis, quaternions and
// the transformation checkers.
//
// Only +,-,*,/,sqrt are used in the solvers (the library is built with
// -fmad=false), which makes them reproducible against the CPU statement of the
// same algorithms; sin/cos/atan2/asin are used only where the reference uses
// them (angle-axis, angular distance, Euler angles of the covariance).
#pragma once

#include "common.cuh"

namespace pgs {

#ifdef __CUDACC__

// cyclic Jacobi eigen-decomposition of a symmetric N x N matrix (col-major).
template <int N>
__device__ void jacobi_sym(double* A, double* w, double* V) {
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) V[j * N + i] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < N; ++p)
      for (int q = p + 1; q < N; ++q) off += fabs(A[q * N + p]);
    if (off == 0.0) break;
    for (int p = 0; p < N; ++p)
      for (int q = p + 1; q < N; ++q) {
        double apq = A[q * N + p];
        if (apq == 0.0) continue;
        double app = A[p * N + p], aqq = A[q * N + q];
        double g = 100.0 * fabs(apq);
        if (sweep > 3 && fabs(app) + g == fabs(app) && fabs(aqq) + g == fabs(aqq)) {
          A[q * N + p] = 0.0;
          A[p * N + q] = 0.0;
          continue;
        }
        double h = aqq - app, t;
        if (fabs(h) + g == fabs(h)) {
          t = apq / h;
        } else {
          double theta = 0.5 * h / apq;
          t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
          if (theta < 0.0) t = -t;
        }
        double c = 1.0 / sqrt(t * t + 1.0);
        double s = t * c;
        double tau = s / (1.0 + c);
        double hh = t * apq;
        A[p * N + p] = app - hh;
        A[q * N + q] = aqq + hh;
        A[q * N + p] = 0.0;
        A[p * N + q] = 0.0;
        for (int r = 0; r < N; ++r) {
          if (r != p && r != q) {
            double arp = A[p * N + r], arq = A[q * N + r];
            double nrp = arp - s * (arq + arp * tau);
            double nrq = arq + s * (arp - arq * tau);
            A[p * N + r] = nrp; A[r * N + p] = nrp;
            A[q * N + r] = nrq; A[r * N + q] = nrq;
          }
        }
        for (int r = 0; r < N; ++r) {
          double vrp = V[p * N + r], vrq = V[q * N + r];
          V[p * N + r] = vrp - s * (vrq + vrp * tau);
          V[q * N + r] = vrq + s * (vrp - vrq * tau);
        }
      }
  }
  for (int i = 0; i < N; ++i) w[i] = A[i * N + i];
}

// A x = b for symmetric PSD N x N: Cholesky, else minimum-norm via eigen-decomposition
template <int N>
__device__ inline int solve_sym(const double* A, const double* b, double* x) {
  double L[N * N];
  double maxd = 0.0;
  for (int i = 0; i < N; ++i)
    if (A[i * N + i] > maxd) maxd = A[i * N + i];
  bool ok = maxd > 0.0;
  for (int i = 0; i < N * N; ++i) L[i] = 0.0;
  for (int j = 0; j < N && ok; ++j) {
    double d = A[j * N + j];
    for (int k = 0; k < j; ++k) d -= L[k * N + j] * L[k * N + j];
    if (!(d > 1e-12 * maxd)) { ok = false; break; }
    double ljj = sqrt(d);
    L[j * N + j] = ljj;
    for (int i = j + 1; i < N; ++i) {
      double s = A[j * N + i];
      for (int k = 0; k < j; ++k) s -= L[k * N + i] * L[k * N + j];
      L[j * N + i] = s / ljj;
    }
  }
  if (ok) {
    double y[N];
    for (int i = 0; i < N; ++i) {
      double s = b[i];
      for (int k = 0; k < i; ++k) s -= L[k * N + i] * y[k];
      y[i] = s / L[i * N + i];
    }
    for (int i = N - 1; i >= 0; --i) {
      double s = y[i];
      for (int k = i + 1; k < N; ++k) s -= L[i * N + k] * x[k];
      x[i] = s / L[i * N + i];
    }
    return N;
  }
  double B[N * N], w[N], V[N * N];
  for (int i = 0; i < N * N; ++i) B[i] = A[i];
  jacobi_sym<N>(B, w, V);
  double wmax = 0.0;
  for (int i = 0; i < N; ++i)
    if (fabs(w[i]) > wmax) wmax = fabs(w[i]);
  for (int i = 0; i < N; ++i) x[i] = 0.0;
  int rank = 0;
  for (int e = 0; e < N; ++e) {
    if (!(w[e] > 1e-12 * wmax)) continue;
    ++rank;
    double vb = 0.0;
    for (int i = 0; i < N; ++i) vb += V[e * N + i] * b[i];
    vb = vb / w[e];
    for (int i = 0; i < N; ++i) x[i] += vb * V[e * N + i];
  }
  return rank;
}
__device__ inline int solve6(const double* A, const double* b, double* x) { return solve_sym<6>(A, b, x); }

__device__ inline void inv6_sym(const double* H, double* Hi) {
  for (int c = 0; c < 6; ++c) {
    double e[6] = {0, 0, 0, 0, 0, 0}, x[6];
    e[c] = 1.0;
    solve6(H, e, x);
    for (int r = 0; r < 6; ++r) Hi[c * 6 + r] = x[r];
  }
}

// 3x3 SVD through the eigen-decomposition of M^T M, U completed by Gram-Schmidt
__device__ inline void svd3(const double* M, double* U, double* S, double* V) {
  double MtM[9], w[3], Vv[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += M[i * 3 + k] * M[j * 3 + k];
      MtM[j * 3 + i] = s;
    }
  jacobi_sym<3>(MtM, w, Vv);
  int ord[3] = {0, 1, 2};
  for (int a = 0; a < 2; ++a)
    for (int b2 = a + 1; b2 < 3; ++b2)
      if (w[ord[b2]] > w[ord[a]]) { int t = ord[a]; ord[a] = ord[b2]; ord[b2] = t; }
  for (int c = 0; c < 3; ++c) {
    for (int r = 0; r < 3; ++r) V[c * 3 + r] = Vv[ord[c] * 3 + r];
    S[c] = w[ord[c]] > 0.0 ? sqrt(w[ord[c]]) : 0.0;
  }
  double smax = S[0];
  int good = 0;
  for (int c = 0; c < 3; ++c) {
    double u[3];
    for (int r = 0; r < 3; ++r) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += M[k * 3 + r] * V[c * 3 + k];
      u[r] = s;
    }
    for (int p = 0; p < good; ++p) {
      double dp = u[0] * U[p * 3] + u[1] * U[p * 3 + 1] + u[2] * U[p * 3 + 2];
      for (int r = 0; r < 3; ++r) u[r] -= dp * U[p * 3 + r];
    }
    double nrm = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    if (S[c] > 1e-13 * smax && nrm > 0.0) {
      for (int r = 0; r < 3; ++r) U[c * 3 + r] = u[r] / nrm;
    } else {
      double best[3] = {0, 0, 0};
      double bestn = -1.0;
      for (int e = 0; e < 3; ++e) {
        double v[3] = {0, 0, 0};
        v[e] = 1.0;
        for (int p = 0; p < good; ++p) {
          double dp = v[0] * U[p * 3] + v[1] * U[p * 3 + 1] + v[2] * U[p * 3 + 2];
          for (int r = 0; r < 3; ++r) v[r] -= dp * U[p * 3 + r];
        }
        double vn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        if (vn > bestn) { bestn = vn; for (int r = 0; r < 3; ++r) best[r] = v[r] / vn; }
      }
      for (int r = 0; r < 3; ++r) U[c * 3 + r] = best[r];
    }
    ++good;
  }
}

__device__ inline void m4_identity(double* M) {
  for (int i = 0; i < 16; ++i) M[i] = 0.0;
  M[0] = M[5] = M[10] = M[15] = 1.0;
}
__device__ inline void m4_mul(const double* A, const double* B, double* C) {
  double R[16];
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += A[k * 4 + r] * B[c * 4 + k];
      R[c * 4 + r] = s;
    }
  for (int i = 0; i < 16; ++i) C[i] = R[i];
}
__device__ inline void m4_rigid_inv(const double* M, double* O) {
  double R[16];
  m4_identity(R);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[c * 4 + r] = M[r * 4 + c];
  for (int r = 0; r < 3; ++r) {
    double s = 0.0;
    for (int k = 0; k < 3; ++k) s += R[k * 4 + r] * M[12 + k];
    R[12 + r] = -s;
  }
  for (int i = 0; i < 16; ++i) O[i] = R[i];
}

// A.5: AngleAxis(|x0..2|, x0..2/|.|) with translation x3..5; zero angle -> R = I
__device__ inline void angle_axis_to_T(const double* x, double* T) {
  m4_identity(T);
  double th = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  if (th > 0.0 && isfinite(th)) {
    double ax = x[0] / th, ay = x[1] / th, az = x[2] / th;
    double c = cos(th), s = sin(th), v = 1.0 - c;
    T[0] = c + ax * ax * v;       T[4] = ax * ay * v - az * s;  T[8] = ax * az * v + ay * s;
    T[1] = ay * ax * v + az * s;  T[5] = c + ay * ay * v;       T[9] = ay * az * v - ax * s;
    T[2] = az * ax * v - ay * s;  T[6] = az * ay * v + ax * s;  T[10] = c + az * az * v;
  }
  T[12] = x[3]; T[13] = x[4]; T[14] = x[5];
}

__device__ inline void quat_from_T(const double* T, double* q /* w,x,y,z */) {
  double m00 = T[0], m11 = T[5], m22 = T[10];
  double tr = m00 + m11 + m22;
  if (tr > 0.0) {
    double s = sqrt(tr + 1.0);
    q[0] = 0.5 * s;
    s = 0.5 / s;
    q[1] = (T[6] - T[9]) * s;
    q[2] = (T[8] - T[2]) * s;
    q[3] = (T[1] - T[4]) * s;
  } else {
    int i = 0;
    if (m11 > m00) i = 1;
    if (m22 > (i == 0 ? m00 : m11)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    double s = sqrt(T[i * 4 + i] - T[j * 4 + j] - T[k * 4 + k] + 1.0);
    q[1 + i] = 0.5 * s;
    s = 0.5 / s;
    q[0] = (T[j * 4 + k] - T[k * 4 + j]) * s;
    q[1 + j] = (T[i * 4 + j] + T[j * 4 + i]) * s;
    q[1 + k] = (T[i * 4 + k] + T[k * 4 + i]) * s;
  }
}

__device__ inline double quat_angular_distance(const double* a, const double* b) {
  double w = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
  double x = -a[0] * b[1] + a[1] * b[0] - a[2] * b[3] + a[3] * b[2];
  double y = -a[0] * b[2] + a[1] * b[3] + a[2] * b[0] - a[3] * b[1];
  double z = -a[0] * b[3] - a[1] * b[2] + a[2] * b[1] + a[3] * b[0];
  return 2.0 * atan2(sqrt(x * x + y * y + z * z), fabs(w));
}

#endif  // __CUDACC__

}  // namespace pgs
