// does ptxas contract mul.rn.f32x2 + add.rn.f32x2 into a fused FFMA2?  (experiment)
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__global__ void k(const float* a, const float* b, float* packed, float* scalar, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (2 * i + 1 >= n) return;
  unsigned long long A = pk(a[2 * i], a[2 * i + 1]), B = pk(b[2 * i], b[2 * i + 1]), M, S;
  asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(M) : "l"(A));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(S) : "l"(M), "l"(B));
  float x, y;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(S));
  packed[2 * i] = x; packed[2 * i + 1] = y;
  scalar[2 * i] = __fadd_rn(__fmul_rn(a[2 * i], a[2 * i]), b[2 * i]);
  scalar[2 * i + 1] = __fadd_rn(__fmul_rn(a[2 * i + 1], a[2 * i + 1]), b[2 * i + 1]);
}
int main() {
  const int n = 1 << 20;
  float *a, *b, *p, *s;
  cudaMallocManaged(&a, n * 4); cudaMallocManaged(&b, n * 4); cudaMallocManaged(&p, n * 4); cudaMallocManaged(&s, n * 4);
  unsigned st = 12345;
  for (int i = 0; i < n; ++i) { st = st * 1664525u + 1013904223u; a[i] = (st >> 8) * (1.0f / 16777216.0f) * 3.f; st = st * 1664525u + 1013904223u; b[i] = (st >> 8) * (1.0f / 16777216.0f); }
  k<<<n / 512, 256>>>(a, b, p, s, n);
  cudaDeviceSynchronize();
  int diff = 0, host_diff = 0;
  for (int i = 0; i < n; ++i) {
    if (p[i] != s[i]) ++diff;
    volatile float m = a[i] * a[i];
    volatile float h = m + b[i];
    if (h != s[i]) ++host_diff;
  }
  printf("packed != scalar in %d of %d; scalar != host unfused in %d\n", diff, n, host_diff);
  return 0;
}
