// knn.cu — batched exact kNN kernels (generic k) on top of knn.cuh.
// One thread per query; a self-search takes its queries in tree order, so the lanes of a
// warp share leaves and walk nearly the same nodes (broadcast loads, less divergence).
//
// k is a template parameter (the candidate list lives in registers); a request
// for k neighbours runs the smallest instantiated K >= k and reports the first
// k entries, which is exact because the K-list is sorted.
#include "knn.cuh"

namespace pgs {

namespace {

constexpr int kSelfGroup = 1;  // self-search seed: the aligned pair of leaves (16 points) around the query

struct KnnJobDev {
  TreeView tree;
  const float4* queries;
  const int* qperm;
  int nq;
  int32_t* ids;
  float* d2;
  int self;  // queries ARE tree.pts (self-kNN): seed from the query's own leaf group
};

template <int K>
__global__ void __launch_bounds__(128)
knn_kernel(const KnnJobDev* __restrict__ jobs, int k, float maxr2) {
  const KnnJobDev job = jobs[blockIdx.y];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= job.nq) return;
  const float4 q = job.queries[j];
  BestK<K> acc;
  acc.init(maxr2);
  if (job.self && K <= kSeedSlots) {
    // the lanes that share a group of leaves share its candidates: sort them once per lane
    // without a branch, then climb from above the group on a path those lanes have in common
    if constexpr (K <= kSeedSlots) {
      knn_seed_group<K>(job.tree, j / kLeaf, kSelfGroup, q.x, q.y, q.z, maxr2, acc);
      knn_climb(job.tree, j / kLeaf, q.x, q.y, q.z, acc, 0, -1, kSelfGroup, false);
    }
  } else if (job.self) {
    // K > 16: neighbours in tree order are mostly neighbours in space: offering the +-K
    // window first gives a tight k-th distance before the tree is touched, so the climb
    // from the query's own leaf enters few sibling subtrees
    const int skip_lo = max(0, j - K), skip_hi = min(job.tree.n - 1, j + K);
    for (int p = skip_lo; p <= skip_hi; ++p) {
      float4 c = job.tree.pts[p];
      float dd = dist2_rn(q.x, q.y, q.z, c.x, c.y, c.z);
      acc.offer(dd, __float_as_int(c.w), p);
    }
    knn_climb(job.tree, j / kLeaf, q.x, q.y, q.z, acc, skip_lo, skip_hi);
  } else {
    knn_traverse(job.tree, q.x, q.y, q.z, acc);
  }
  // a self-search's queries are the index' own points: .w is the original index = the result column
  const int col = job.self ? __float_as_int(q.w) : (job.qperm ? job.qperm[j] : j);
  int32_t* oi = job.ids + (size_t)col * k;
  float* od = job.d2 + (size_t)col * k;
#pragma unroll
  for (int e = 0; e < K; ++e) {
    if (e < k) {
      const int id = key_id(acc.key[e]);
      oi[e] = (id == 0x7fffffff) ? -1 : id;
      if (job.d2) od[e] = (id == 0x7fffffff) ? __int_as_float(0x7f800000) : key_dist(acc.key[e]);
    }
  }
}

// knn > 32: run-time k, keys in local memory (BestDyn)
__global__ void __launch_bounds__(128)
knn_dyn_kernel(const KnnJobDev* __restrict__ jobs, int k, float maxr2) {
  const KnnJobDev job = jobs[blockIdx.y];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= job.nq) return;
  const float4 q = job.queries[j];
  unsigned long long keys[kMaxDynK];
  BestDyn acc{keys, k};
  acc.init(maxr2);
  if (job.self) {
    // neighbours in tree order are mostly neighbours in space: the +-k window gives a bound
    // before the tree is touched, then the climb from the query's own leaf
    const int skip_lo = max(0, j - k), skip_hi = min(job.tree.n - 1, j + k);
    for (int p = skip_lo; p <= skip_hi; ++p) {
      const float4 c = job.tree.pts[p];
      acc.offer(dist2_rn(q.x, q.y, q.z, c.x, c.y, c.z), __float_as_int(c.w), p);
    }
    knn_climb(job.tree, j / kLeaf, q.x, q.y, q.z, acc, skip_lo, skip_hi);
  } else {
    knn_traverse(job.tree, q.x, q.y, q.z, acc);
  }
  const int col = job.self ? __float_as_int(q.w) : (job.qperm ? job.qperm[j] : j);
  int32_t* oi = job.ids + (size_t)col * k;
  float* od = job.d2 ? job.d2 + (size_t)col * k : nullptr;
  for (int e = 0; e < k; ++e) {
    const int id = key_id(keys[e]);
    oi[e] = (id == 0x7fffffff) ? -1 : id;
    if (od) od[e] = (id == 0x7fffffff) ? __int_as_float(0x7f800000) : key_dist(keys[e]);
  }
}

template <int K>
void launch_k(Ctx* ctx, dim3 grid, const KnnJobDev* d_jobs, int k, float maxr2) {
  knn_kernel<K><<<grid, 128, 0, ctx->stream>>>(d_jobs, k, maxr2);
}

void launch(Ctx* ctx, const std::vector<KnnJobDev>& jobs, int k, float max_dist) {
  if (jobs.empty()) return;
  if (k < 1 || k > kMaxDynK) throw Error(PGS_INVALID_PARAMETER, "knn must be in [1, " + std::to_string(kMaxDynK) + "]");
  int max_q = 0;
  for (auto& j : jobs) max_q = std::max(max_q, j.nq);
  if (max_q == 0) return;
  DBuf<KnnJobDev> d_jobs(ctx, jobs.size());
  ctx->upload_small(d_jobs.p, jobs.data(), jobs.size() * sizeof(KnnJobDev));
  const float inf = __builtin_inff();
  float maxr2 = (max_dist == inf) ? inf : max_dist * max_dist;
  dim3 grid(ceil_div(max_q, 128), (unsigned)jobs.size());
  if (k == 1) launch_k<1>(ctx, grid, d_jobs.p, k, maxr2);
  else if (k == 2) launch_k<2>(ctx, grid, d_jobs.p, k, maxr2);
  else if (k == 3) launch_k<3>(ctx, grid, d_jobs.p, k, maxr2);
  else if (k == 4) launch_k<4>(ctx, grid, d_jobs.p, k, maxr2);
  else if (k == 5) launch_k<5>(ctx, grid, d_jobs.p, k, maxr2);
  else if (k == 6) launch_k<6>(ctx, grid, d_jobs.p, k, maxr2);
  else if (k <= 8) launch_k<8>(ctx, grid, d_jobs.p, k, maxr2);
  else if (k <= 10) launch_k<10>(ctx, grid, d_jobs.p, k, maxr2);
  else if (k <= 12) launch_k<12>(ctx, grid, d_jobs.p, k, maxr2);
  else if (k <= 16) launch_k<16>(ctx, grid, d_jobs.p, k, maxr2);
  else if (k <= 20) launch_k<20>(ctx, grid, d_jobs.p, k, maxr2);
  else if (k <= 24) launch_k<24>(ctx, grid, d_jobs.p, k, maxr2);
  else if (k <= 32) launch_k<32>(ctx, grid, d_jobs.p, k, maxr2);
  else knn_dyn_kernel<<<grid, 128, 0, ctx->stream>>>(d_jobs.p, k, maxr2);
  ctx_count_launches(ctx, 1);
  PGS_LAUNCH_CHECK();
}

}  // namespace

void knn_batched(Ctx* ctx, const std::vector<KnnJob>& jobs, int k, float max_dist) {
  std::vector<KnnJobDev> dj;
  for (auto& j : jobs) dj.push_back(KnnJobDev{j.tree, j.queries, j.qperm, j.nq, j.ids, j.d2, 0});
  launch(ctx, dj, k, max_dist);
}

void knn_self_batched(Ctx* ctx, const std::vector<const Index*>& idx, int k, float max_dist,
                      const std::vector<int32_t*>& ids, const std::vector<float*>& d2) {
  // queries = the index' own sorted points; column = original index (pts.w)
  std::vector<KnnJobDev> dj;
  for (size_t b = 0; b < idx.size(); ++b)
    dj.push_back(KnnJobDev{idx[b]->view(), idx[b]->pts.p, nullptr, idx[b]->n, ids[b], d2[b], 1});
  launch(ctx, dj, k, max_dist);
}

}  // namespace pgs
