// dense.cu — tensor-core distance tiles for SMALL reference clouds (BASELINE.json north_star:
// "tensor cores are used only for the dense brute-force distance-tile path on small clouds";
// a GEMM cannot produce the contract's fp32 distances, so it may only
// pre-filter candidates that are then re-tested exactly).
//
// For a reference cloud of <= kDenseMaxRef points the matcher can skip the tree:
//   score(q, p) = |p'|^2 - 2 q'.p'      (q' = q - c_b, p' = p - c_b, c_b = centre of ref block b)
// is one row x column of a [128 queries] x [128 reference points] tile computed by tcgen05.mma
// (kind::tf32, M = 128, N = 128, accumulators in TMEM, two tiles in flight).  fp32 operands are fed as two tf32
// pieces (x = hi + lo), laid out over K = 16 so that hi.hi + hi.lo + lo.hi products and a
// three-piece |p'|^2 are summed by the tensor core:
//   A row (query):      [-2qh.x -2qh.y -2qh.z | -2qh.x -2qh.y -2qh.z | -2ql.x -2ql.y -2ql.z | 1 1 1 | 0 0 0 0]
//   B row (reference):  [  ph.x   ph.y   ph.z |   pl.x   pl.y   pl.z |   ph.x   ph.y   ph.z | n1 n2 n3 | 0 0 0 0]
// The reference tiles are prepared once per cloud (dense_prepare) in the canonical K-major,
// no-swizzle shared-memory layout and pulled in with one 8 KB bulk copy (TMA engine,
// cp.async.bulk + mbarrier) per block.  Every thread owns one query = one TMEM lane, reads its
// 128 scores back with tcgen05.ld and keeps the 4 smallest approximate distances; at the end
// those 4 candidates are re-tested with the contract's exact fp32 distance, and the answer is
// CERTIFIED: every other point's approximate distance is >= the 4th best, and approximate and
// exact distances differ by at most E_b = 2^-16 (|q'| + R_b)^2 in block b (far blocks are
// excluded geometrically), so if (4th best - max E_b) > exact best, no other point can win or
// tie.  A query that cannot be certified (more than 4 points inside the error window) falls back
// to the exact tree search.  Result: the same exact (distance, index) minimum as the tree search.
#include "dense.cuh"

#include "knn.cuh"

namespace pgs {

namespace {

constexpr int kRefBlock = 128;  // N of one MMA = reference points per tile (and the centring granularity)
constexpr int kQTile = 128;     // M = queries per CTA = TMEM lanes
constexpr int kTileFloats = 16 * kRefBlock;  // 4 K-chunks x 128 rows x 4 floats = 8 KB
constexpr float kInfF = __builtin_inff();

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- reference side: block centres, radii and MMA-ready tiles ---------------------------------
__global__ void __launch_bounds__(kRefBlock)
dense_prepare_kernel(const float4* __restrict__ pts /* kd order, padded */, int n, float* __restrict__ tiles,
                     float4* __restrict__ info) {
  __shared__ float red[kRefBlock / 32][7];
  const int b = blockIdx.x, t = threadIdx.x, i = b * kRefBlock + t;
  const bool valid = i < n;
  const float4 p = valid ? pts[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  // centre = middle of the block's bounding box (any point works; a tight one keeps E small)
  float lo[3] = {valid ? p.x : kInfF, valid ? p.y : kInfF, valid ? p.z : kInfF};
  float hi[3] = {valid ? p.x : -kInfF, valid ? p.y : -kInfF, valid ? p.z : -kInfF};
#pragma unroll
  for (int d = 0; d < 3; ++d)
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
  if ((t & 31) == 0)
    for (int d = 0; d < 3; ++d) { red[t >> 5][d] = lo[d]; red[t >> 5][3 + d] = hi[d]; }
  __syncthreads();
  float c[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float l = kInfF, h = -kInfF;
    for (int w = 0; w < kRefBlock / 32; ++w) { l = fminf(l, red[w][d]); h = fmaxf(h, red[w][3 + d]); }
    c[d] = 0.5f * (l + h);
  }
  const float px = p.x - c[0], py = p.y - c[1], pz = p.z - c[2];
  float r2 = valid ? px * px + py * py + pz * pz : 0.f;
  for (int o = 16; o > 0; o >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(0xffffffffu, r2, o));
  __syncthreads();
  if ((t & 31) == 0) red[t >> 5][6] = r2;
  __syncthreads();
  if (t == 0) {
    float m = 0.f;
    for (int w = 0; w < kRefBlock / 32; ++w) m = fmaxf(m, red[w][6]);
    info[b] = make_float4(c[0], c[1], c[2], sqrtf(m) * 1.0001f);
  }
  float row[16];
  if (valid) {
    const float hx = trunc_tf32(px), hy = trunc_tf32(py), hz = trunc_tf32(pz);
    const double nn = (double)px * px + (double)py * py + (double)pz * pz;
    const float n1 = trunc_tf32((float)nn);
    const float n2 = trunc_tf32((float)(nn - (double)n1));
    const float n3 = (float)(nn - (double)n1 - (double)n2);
    const float v[16] = {hx, hy, hz, px - hx, py - hy, pz - hz, hx, hy, hz, n1, n2, n3, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 16; ++k) row[k] = v[k];
  } else {
    // padding rows never win: score = +1e30
#pragma unroll
    for (int k = 0; k < 16; ++k) row[k] = 0.f;
    row[9] = 1e30f;
  }
  float* tile = tiles + (size_t)b * kTileFloats;
#pragma unroll
  for (int ch = 0; ch < 4; ++ch)
    reinterpret_cast<float4*>(tile)[ch * kRefBlock + t] = make_float4(row[4 * ch], row[4 * ch + 1], row[4 * ch + 2], row[4 * ch + 3]);
}

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned ok = 0;
  unsigned long long spins = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!ok && ++spins > (1ull << 28)) __trap();  // a lost completion must not hang the device
  }
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 B; SBO = bytes between 8-row groups, LBO =
// bytes between the two 16-byte K chunks of one MMA (both in units of 16 B); version 1 = sm_100
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kRefBlock >> 3) << 17) | ((uint32_t)(kQTile >> 4) << 24);

struct DenseJob {
  const float* tiles;      // n_blocks x 16 KB
  const float4* info;      // n_blocks x {cx, cy, cz, R}
  TreeView tree;           // the reference index: exact re-test of the candidates, fallback search
  int n, n_blocks;
  const float4* queries;
  int nq;
  int32_t* ids;     // original reference index (matcher module), or
  int* out_pos;     // sorted position (ICP loop); exactly one of the two is set
  float* d2;
  const Xf* xf;     // ICP loop: transform applied to every query first, or null
  const int* active;  // ICP loop: skip the job when *active == 0, or null
};

// min of 32 scores (3-input min on sm_100a: 16 instructions)
__device__ __forceinline__ float min32(const uint32_t (&v)[32]) {
  float m[11];
#pragma unroll
  for (int i = 0; i < 10; ++i)
    m[i] = fminf(fminf(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1])), __uint_as_float(v[3 * i + 2]));
  m[10] = fminf(__uint_as_float(v[30]), __uint_as_float(v[31]));
  const float a = fminf(fminf(m[0], m[1]), m[2]), b = fminf(fminf(m[3], m[4]), m[5]);
  const float c = fminf(fminf(m[6], m[7]), m[8]), d = fminf(m[9], m[10]);
  return fminf(fminf(a, b), fminf(c, d));
}

// Software pipeline per CTA (128 threads = 128 queries = 128 TMEM lanes), two stages in flight:
//   trip b:  [TMA]  bulk copy of reference tile b+1          -> sB[(b+1)&1]
//            [all]  A rows of block b+1 (q - c_{b+1}, tf32 hi/lo) -> sA[(b+1)&1]
//            [t0]   tcgen05.mma of block b+1 -> TMEM buffer (b+1)&1, commit -> bar_mma[(b+1)&1]
//            [all]  epilogue of block b from TMEM buffer b&1 (its MMA was issued one trip earlier)
__global__ void __launch_bounds__(kQTile)
dense_knn1_kernel(const DenseJob* __restrict__ jobs, float maxr2, unsigned* __restrict__ fallbacks) {
  __shared__ __align__(1024) float sB[2][kTileFloats];     // 2 x 8 KB: [chunk][ref row][4]
  __shared__ __align__(1024) float sA[2][16 * kQTile];     // 2 x 8 KB: [chunk][query row][4]
  __shared__ __align__(8) uint64_t bar_tile[2], bar_mma[2];
  __shared__ uint32_t tmem_base_s;
  const DenseJob job = jobs[blockIdx.y];
  const int t = threadIdx.x, warp = t >> 5;
  const int qi = blockIdx.x * kQTile + t;
  if (blockIdx.x * kQTile >= job.nq) return;
  if (job.active && !*job.active) return;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(2 * kRefBlock));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (t == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_tile[i], 1); mbar_init(&bar_mma[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  const bool live = qi < job.nq;
  float4 q4 = live ? job.queries[qi] : make_float4(0.f, 0.f, 0.f, 0.f);
  if (job.xf) {  // ICP loop: stepReading = T_iter * reading, fused (never stored)
    const float3 o = xform_rn(*job.xf, q4.x, q4.y, q4.z);
    q4 = make_float4(o.x, o.y, o.z, q4.w);
  }
  // the 4 smallest approximate distances seen so far (ascending) and their sorted positions
  float bd0 = kInfF, bd1 = kInfF, bd2 = kInfF, bd3 = kInfF;
  int bp0 = -1, bp1 = -1, bp2 = -1, bp3 = -1;
  float e_max = 0.f;
  float q2_cur = 0.f, q2_next = 0.f;
  const int nb = job.n_blocks;

  // stage block `blk`: tile copy, A rows, MMA issue
  auto stage = [&](int blk) {
    const int s = blk & 1;
    const float4 cb = __ldg(job.info + blk);
    if (t == 0) {
      mbar_expect_tx(&bar_tile[s], kTileFloats * 4);
      bulk_copy_g2s(sB[s], job.tiles + (size_t)blk * kTileFloats, kTileFloats * 4, &bar_tile[s]);
    }
    const float qx = q4.x - cb.x, qy = q4.y - cb.y, qz = q4.z - cb.z;
    const float hx = trunc_tf32(qx), hy = trunc_tf32(qy), hz = trunc_tf32(qz);
    const float lx = qx - hx, ly = qy - hy, lz = qz - hz;
    float4* a4 = reinterpret_cast<float4*>(sA[s]);
    a4[0 * kQTile + t] = make_float4(-2.f * hx, -2.f * hy, -2.f * hz, -2.f * hx);
    a4[1 * kQTile + t] = make_float4(-2.f * hy, -2.f * hz, -2.f * lx, -2.f * ly);
    a4[2 * kQTile + t] = make_float4(-2.f * lz, 1.f, 1.f, 1.f);
    a4[3 * kQTile + t] = make_float4(0.f, 0.f, 0.f, 0.f);
    q2_next = qx * qx + qy * qy + qz * qz;
    // certification bookkeeping for this block (see the header comment)
    const float qn = sqrtf(q2_next);
    const float gap = fmaxf(qn - cb.w, 0.f);
    const float lb = gap * gap * 0.999999f;
    const float eb = (qn + cb.w) * (qn + cb.w) * (1.0f / 65536.0f);
    if (!(lb > bd3)) e_max = fmaxf(e_max, eb);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
    // also orders every lane's TMEM reads of the previous use of this accumulator buffer before the MMA
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (t == 0) {
      mbar_wait(&bar_tile[s], (unsigned)(blk >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a0 = smem_u32(sA[s]), b0 = smem_u32(sB[s]);
      const uint32_t d = tmem_base + (uint32_t)(s * kRefBlock);
      // two K = 8 steps: chunks {0,1} then {2,3}
      umma_tf32(d, umma_desc(a0, kQTile * 16, 128), umma_desc(b0, kRefBlock * 16, 128), kIdesc, 0u);
      umma_tf32(d, umma_desc(a0 + 2 * kQTile * 16, kQTile * 16, 128), umma_desc(b0 + 2 * kRefBlock * 16, kRefBlock * 16, 128), kIdesc, 1u);
      umma_commit(&bar_mma[s]);
    }
  };

  stage(0);
  q2_cur = q2_next;
  for (int b = 0; b < nb; ++b) {
    if (b + 1 < nb) stage(b + 1);
    const int s = b & 1;
    mbar_wait(&bar_mma[s], (unsigned)(b >> 1) & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- epilogue: my row's 128 scores; a 32-column chunk is looked at only if its minimum can enter
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * kRefBlock);
#pragma unroll 1
    for (int c0 = 0; c0 < kRefBlock; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(lane_addr + (uint32_t)c0, v);
      if (min32(v) + q2_cur < bd3) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float da = __uint_as_float(v[j]) + q2_cur;
          if (da < bd3) {
            const int pos = b * kRefBlock + c0 + j;
            if (da < bd2) {
              bd3 = bd2; bp3 = bp2;
              if (da < bd1) {
                bd2 = bd1; bp2 = bp1;
                if (da < bd0) { bd1 = bd0; bp1 = bp0; bd0 = da; bp0 = pos; }
                else { bd1 = da; bp1 = pos; }
              } else { bd2 = da; bp2 = pos; }
            } else { bd3 = da; bp3 = pos; }
          }
        }
      }
    }
    q2_cur = q2_next;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * kRefBlock));
  if (!live) return;

  // ---- exact re-test of the candidates, certification, fallback -----------------------------------
  Best1 acc;
  acc.init(kInfF);
  const int cand[4] = {bp0, bp1, bp2, bp3};
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (cand[k] >= 0 && cand[k] < job.n) {
      const float4 p = __ldg(job.tree.pts + cand[k]);
      acc.offer(dist2_rn(q4.x, q4.y, q4.z, p.x, p.y, p.z), __float_as_int(p.w), cand[k]);
    }
  const bool certified = job.n <= 4 || (acc.pos >= 0 && bd3 - e_max > acc.bound());
  if (!certified) {
    if (fallbacks) atomicAdd(fallbacks, 1u);
    knn_traverse(job.tree, q4.x, q4.y, q4.z, acc);  // the candidates' best is already the bound
  }
  const bool found = acc.pos >= 0 && acc.bound() <= maxr2;
  if (job.out_pos) {  // ICP loop: sorted position + distance, as match_kernel writes them
    job.out_pos[qi] = found ? acc.pos : -1;
    job.d2[qi] = found ? key_dist(acc.key) : kInfF;
  } else {
    job.ids[qi] = found ? key_id(acc.key) : -1;
    if (job.d2) job.d2[qi] = found ? key_dist(acc.key) : kInfF;
  }
}

}  // namespace

void dense_prepare(Ctx* ctx, const Index& idx, DenseRef& out) {
  if (idx.n > kDenseMaxRef) throw Error(PGS_INVALID_ARGUMENT, "dense matcher: reference cloud too large");
  out.n = idx.n;
  out.n_blocks = ceil_div(std::max(idx.n, 1), kRefBlock);
  out.tiles.reset(ctx, (size_t)out.n_blocks * kTileFloats);
  out.info.reset(ctx, (size_t)out.n_blocks);
  dense_prepare_kernel<<<out.n_blocks, kRefBlock, 0, ctx->stream>>>(idx.pts.p, idx.n, out.tiles.p, out.info.p);
  ctx_count_launches(ctx, 1);
  PGS_LAUNCH_CHECK();
}

void dense_upload_jobs(Ctx* ctx, const std::vector<DenseQuery>& queries, DenseJobs& out) {
  std::vector<DenseJob> jobs;
  out.max_q = 0;
  for (auto& q : queries) {
    jobs.push_back(DenseJob{q.ref->tiles.p, q.ref->info.p, q.tree, q.ref->n, q.ref->n_blocks, q.queries, q.nq, q.ids, q.out_pos,
                            q.d2, q.xf, q.active});
    out.max_q = std::max(out.max_q, q.nq);
  }
  out.n_jobs = (int)jobs.size();
  out.table.reset(ctx, std::max<size_t>(jobs.size(), 1) * sizeof(DenseJob));
  if (!jobs.empty()) ctx->upload_small(out.table.p, jobs.data(), jobs.size() * sizeof(DenseJob));
}

void dense_launch(Ctx* ctx, const DenseJobs& jobs, float maxr2) {
  if (jobs.n_jobs == 0 || jobs.max_q == 0) return;
  dense_knn1_kernel<<<dim3(ceil_div(jobs.max_q, kQTile), (unsigned)jobs.n_jobs), kQTile, 0, ctx->stream>>>(
      reinterpret_cast<const DenseJob*>(jobs.table.p), maxr2, nullptr);
  ctx_count_launches(ctx, 1);
  PGS_LAUNCH_CHECK();
}

unsigned dense_knn1(Ctx* ctx, const std::vector<DenseQuery>& queries, float max_dist, bool count_fallbacks) {
  if (queries.empty()) return 0;
  std::vector<DenseJob> jobs;
  int max_q = 0;
  for (auto& q : queries) {
    jobs.push_back(DenseJob{q.ref->tiles.p, q.ref->info.p, q.tree, q.ref->n, q.ref->n_blocks, q.queries, q.nq, q.ids, q.out_pos,
                            q.d2, q.xf, q.active});
    max_q = std::max(max_q, q.nq);
  }
  if (max_q == 0) return 0;
  DBuf<DenseJob> d_jobs(ctx, jobs.size());
  ctx->upload_small(d_jobs.p, jobs.data(), jobs.size() * sizeof(DenseJob));
  DBuf<unsigned> fb(ctx, 1);
  fb.zero();
  const float inf = __builtin_inff();
  const float maxr2 = (max_dist == inf) ? inf : max_dist * max_dist;
  dense_knn1_kernel<<<dim3(ceil_div(max_q, kQTile), (unsigned)jobs.size()), kQTile, 0, ctx->stream>>>(
      d_jobs.p, maxr2, count_fallbacks ? fb.p : nullptr);
  ctx_count_launches(ctx, 1);
  PGS_LAUNCH_CHECK();
  unsigned h = 0;
  if (count_fallbacks) {
    fb.download(&h, 1);
    ctx->sync();
  }
  return h;
}

}  // namespace pgs
