// common.cuh — shared host/device plumbing for libpgslam_b200 (sm_100a only).
//
// Everything in this library is compiled with -fmad=false: the numeric
// contract fixes the fp32 operation order of squared distances
// and rigid transforms, and the fp64 solvers rely on +,-,*,/,sqrt only, so a
// fused multiply-add anywhere on those paths would break bit-parity with the
// CPU statement of the algorithm.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pgslam_b200.h"

namespace pgs {

// ---------------------------------------------------------------------------
// errors: C++ exceptions inside the library, mapped to status codes at the ABI
// (one code per libpointmatcher exception type, SURVEY.md §8b "Errors").
// ---------------------------------------------------------------------------
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define PGS_CUDA(expr)                                                            \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess)                                                        \
      throw ::pgs::Error(PGS_CUDA_ERROR, std::string(#expr) + ": " +              \
                                             cudaGetErrorString(_e) + " at " +   \
                                             __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)

#define PGS_LAUNCH_CHECK() PGS_CUDA(cudaGetLastError())

// ---------------------------------------------------------------------------
// context: one device, one stream.  Handles created from different contexts
// are independent (pgslam-MT drives the localizer and the loop closer from two
// host threads, LocalizerMT.hpp:47 / LoopCloserMT.hpp:41).
// ---------------------------------------------------------------------------
// Tuning knobs of the hot kernels (pgs_ctx_set_option; defaults from PGS_* environment
// variables).  None of them changes a result - only how the work is scheduled.
struct Tuning {
  int match_mode;     // 0 descent from 10 levels up + climb, 1 cell-guided start, 2 persistent lanes from
                      // the root, 3 persistent lanes + cell-guided start, 4 queue matcher (uniform first phase,
                      // compacted revisits)
  int pm_blocks;      // persistent matcher: resident blocks per SM
  int pm_refill;      // ... idle lanes that trigger a refill
  int pm_pair_w, pm_leaf_w;  // ... PAIR runs when pairs * pair_w >= leaves * leaf_w
  int mq_batches;     // queue matcher (mode 4): batches of 32 queries per warp
  int mq_blocks;      // ... resident blocks per SM it is compiled for (10, 12 or 16)
  int resort_it;      // iteration after whose matches the reading is re-ordered by matched leaf (-1 never)
  int dense_max_ref;  // KDTreeMatcher k = 1 on a reference of at most this many points uses the tensor-core
                      // distance tiles (dense.cu) instead of the tree; 0 = never (default)
  int dense_count_fallbacks;  // read back how many queries needed the exact fallback (one host sync)
  int batch_chunk;    // pairs per chunk a batch worker pulls (pgs_icp_run_batch)
  Tuning();
};

struct Ctx {
  int device = 0;
  Tuning tune;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int num_sms = 148;
  std::string last_error;
  uint64_t launches = 0;  // kernels launched through this context (bench.py gpu_launches)
  bool profiling = false;
  pgs_stage_times times{};
  // pinned staging + mapped progress words for the ICP loop
  void* pinned = nullptr;
  void* pinned_dev = nullptr;  // device alias of the pinned ring (the ring is mapped)
  size_t pinned_bytes = 0, pinned_head = 0;
  volatile int* h_progress = nullptr;  // mapped host word: set to 1 by the device when no pair is active
  volatile int* d_progress = nullptr;  // device alias of h_progress
  cudaEvent_t loop_ev[2] = {nullptr, nullptr};
  cudaStream_t copy_stream = nullptr;  // uploads from pinned host memory overlap with compute
  cudaEvent_t copy_ev[2] = {nullptr, nullptr};
  void ensure_progress();
  // A large batch of independent pairs is split over `batch_streams` worker contexts, each
  // with its own stream and host thread, so that one sub-batch's latency-bound tails (the
  // small select / solve kernels, host bookkeeping between launches) overlap the others'
  // wide kernels.  Results do not depend on the split (per-pair reduction order).
  int batch_streams = 4;
  std::vector<Ctx*> workers;  // owned; created on first use
  cudaEvent_t fork_ev = nullptr;
  Ctx* worker(int i);
  void destroy_resources();

  void* alloc(size_t bytes);
  void free(void* p);
  void upload_small(void* dst, const void* src, size_t bytes);
  // Worker contexts of a batch sleep in their waits instead of spinning: a box that runs one
  // rank per GPU has 8 x (4 workers + main + uploader) host threads on far fewer cores, and a
  // spinning waiter that loses its core finds out late that the device went idle.  The
  // single-registration path keeps the spinning waits (lowest latency).
  bool blocking_waits = false;
  cudaEvent_t cross_ev = nullptr;  // joins with other contexts' streams (api.cu Borrow)
  cudaEvent_t sync_ev = nullptr;
  void sync() {
    if (!blocking_waits) { PGS_CUDA(cudaStreamSynchronize(stream)); return; }
    if (!sync_ev) PGS_CUDA(cudaEventCreateWithFlags(&sync_ev, cudaEventDisableTiming | cudaEventBlockingSync));
    PGS_CUDA(cudaEventRecord(sync_ev, stream));
    PGS_CUDA(cudaEventSynchronize(sync_ev));
  }
};

static inline void ctx_count_launches(Ctx* ctx, int n) { ctx->launches += (uint64_t)n; }

// RAII device buffer bound to a context's stream-ordered pool.
template <typename T>
struct DBuf {
  Ctx* ctx = nullptr;
  T* p = nullptr;
  size_t n = 0;
  DBuf() {}
  DBuf(Ctx* c, size_t count) { reset(c, count); }
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  DBuf(DBuf&& o) noexcept : ctx(o.ctx), p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DBuf& operator=(DBuf&& o) noexcept {
    if (this != &o) { release(); ctx = o.ctx; p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DBuf() { release(); }
  void reset(Ctx* c, size_t count) {
    release();
    ctx = c;
    n = count;
    p = static_cast<T*>(c->alloc((count ? count : 1) * sizeof(T)));
  }
  // take ownership of a block allocated elsewhere in the same pool (freed on ctx's stream)
  void adopt(Ctx* c, T* ptr, size_t count) {
    release();
    ctx = c;
    p = ptr;
    n = count;
  }
  void release() {
    if (p && ctx) ctx->free(p);
    p = nullptr;
    n = 0;
  }
  void zero() { PGS_CUDA(cudaMemsetAsync(p, 0, (n ? n : 1) * sizeof(T), ctx->stream)); }
  void upload(const T* h, size_t count) {
    PGS_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  }
  void download(T* h, size_t count) const {
    PGS_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
  }
};

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
#ifdef __CUDACC__

// squared distance in the contract's order: ((dx*dx)+dy*dy)+dz*dz, fp32, no FMA
__device__ __forceinline__ float dist2_rn(float qx, float qy, float qz, float px, float py, float pz) {
  float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
  float d = __fmul_rn(dx, dx);
  d = __fadd_rn(d, __fmul_rn(dy, dy));
  d = __fadd_rn(d, __fmul_rn(dz, dz));
  return d;
}

// ---- packed fp32x2 arithmetic (sm_100a: FADD2 / FMUL2 issue two IEEE-rn fp32 operations
// per slot; each half rounds exactly like the scalar instruction, so results are unchanged).
// Only SUB and MUL are used packed: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into a
// fused FFMA2 even under -fmad=false (measured: 19 % of a*a+b results differ), which would
// break the no-FMA contract, so sums of products stay scalar.
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long sub_f32x2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long mul_f32x2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// dist2_rn with the x,y halves packed: same operations, same order, 6 issue slots instead of 8
__device__ __forceinline__ float dist2_rn_packed(unsigned long long qxy, float qz, float px, float py, float pz) {
  const unsigned long long dxy = sub_f32x2(qxy, pack_f32x2(px, py));
  const unsigned long long sq = mul_f32x2(dxy, dxy);
  float sx, sy;
  unpack_f32x2(sq, sx, sy);
  const float dz = __fsub_rn(qz, pz);
  float d = __fadd_rn(sx, sy);
  d = __fadd_rn(d, __fmul_rn(dz, dz));
  return d;
}

// Lower bound of dist2_rn over an axis-aligned box, same operation order, so by monotonicity
// of IEEE rounding it never exceeds the computed distance to any point inside the box.
// The node is stored {lo.x, lo.y, hi.x, hi.y, lo.z, hi.z} (index.cu): the three
// 8-byte halves arrive packed from the load, the six subtractions issue as three FADD2
// (q.z - hi.z is taken as the exact negation of hi.z - q.z) and the x,y squares as one FMUL2.
struct QueryPk {
  unsigned long long xy, zz;
  float z;
};
__device__ __forceinline__ QueryPk pack_query(float qx, float qy, float qz) {
  return QueryPk{pack_f32x2(qx, qy), pack_f32x2(qz, qz), qz};
}
__device__ __forceinline__ float box_lb_packed(const QueryPk& q, unsigned long long lxy, unsigned long long hxy,
                                               unsigned long long lhz) {
  float ax, ay, bx, by, cl, ch;
  unpack_f32x2(sub_f32x2(lxy, q.xy), ax, ay);
  unpack_f32x2(sub_f32x2(q.xy, hxy), bx, by);
  unpack_f32x2(sub_f32x2(lhz, q.zz), cl, ch);
  const float dx = fmaxf(fmaxf(ax, bx), 0.f);
  const float dy = fmaxf(fmaxf(ay, by), 0.f);
  const float dz = fmaxf(fmaxf(cl, -ch), 0.f);
  const unsigned long long dxy = pack_f32x2(dx, dy);
  float sx, sy;
  unpack_f32x2(mul_f32x2(dxy, dxy), sx, sy);
  float d = __fadd_rn(sx, sy);
  d = __fadd_rn(d, __fmul_rn(dz, dz));
  return d;
}

// 3x4 fp32 rigid transform, row r: (((T[r,0]*x)+T[r,1]*y)+T[r,2]*z)+T[r,3]
struct Xf {
  float m[12];  // row-major 3x4
};
__device__ __forceinline__ float3 xform_rn(const Xf& T, float x, float y, float z) {
  float3 o;
  o.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T.m[0], x), __fmul_rn(T.m[1], y)), __fmul_rn(T.m[2], z)), T.m[3]);
  o.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T.m[4], x), __fmul_rn(T.m[5], y)), __fmul_rn(T.m[6], z)), T.m[7]);
  o.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T.m[8], x), __fmul_rn(T.m[9], y)), __fmul_rn(T.m[10], z)), T.m[11]);
  return o;
}
__device__ __forceinline__ float3 rot_rn(const Xf& T, float x, float y, float z) {
  float3 o;
  o.x = __fadd_rn(__fadd_rn(__fmul_rn(T.m[0], x), __fmul_rn(T.m[1], y)), __fmul_rn(T.m[2], z));
  o.y = __fadd_rn(__fadd_rn(__fmul_rn(T.m[4], x), __fmul_rn(T.m[5], y)), __fmul_rn(T.m[6], z));
  o.z = __fadd_rn(__fadd_rn(__fmul_rn(T.m[8], x), __fmul_rn(T.m[9], y)), __fmul_rn(T.m[10], z));
  return o;
}

// order-preserving float <-> uint mapping (for atomic min/max and radix keys)
__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 16-byte load that does not pollute L1 (read-once data)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

#endif  // __CUDACC__

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
static inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace pgs
