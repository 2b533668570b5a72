// knn.cu — batched exact kNN kernels (generic k) on top of knn.cuh.
// One thread per query; queries arrive in Morton order so the lanes of a warp
// walk nearly the same nodes (broadcast loads, little divergence).
#include "knn.cuh"

namespace pgs {

namespace {

struct KnnJobDev {
  TreeView tree;
  const float4* queries;
  const int* qperm;
  int nq;
  int32_t* ids;
  float* d2;
};

template <int KCAP>
__global__ void __launch_bounds__(128)
knn_kernel(const KnnJobDev* __restrict__ jobs, int k, float maxr2) {
  const KnnJobDev job = jobs[blockIdx.y];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= job.nq) return;
  float4 q = job.queries[j];
  BestK<KCAP> acc;
  acc.init(k);
  knn_traverse(job.tree, q.x, q.y, q.z, maxr2, acc);
  const int col = job.qperm ? job.qperm[j] : j;
  int32_t* oi = job.ids + (size_t)col * k;
  float* od = job.d2 + (size_t)col * k;
#pragma unroll
  for (int e = 0; e < KCAP; ++e) {
    if (e < k) {
      oi[e] = (acc.id[e] == 0x7fffffff) ? -1 : acc.id[e];
      od[e] = acc.d[e];
    }
  }
}

template <>
__global__ void __launch_bounds__(128)
knn_kernel<1>(const KnnJobDev* __restrict__ jobs, int k, float maxr2) {
  const KnnJobDev job = jobs[blockIdx.y];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= job.nq) return;
  float4 q = job.queries[j];
  Best1 acc;
  acc.init();
  knn_traverse(job.tree, q.x, q.y, q.z, maxr2, acc);
  const int col = job.qperm ? job.qperm[j] : j;
  job.ids[col] = (acc.pos < 0) ? -1 : acc.id;
  job.d2[col] = acc.d;
}

__global__ void perm_from_sorted_kernel(const float4* __restrict__ pts, int n, int* __restrict__ perm) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) perm[j] = __float_as_int(pts[j].w);
}

void launch(Ctx* ctx, const std::vector<KnnJobDev>& jobs, int k, float max_dist) {
  if (jobs.empty()) return;
  if (k < 1 || k > 32) throw Error(PGS_INVALID_PARAMETER, "knn must be in [1, 32]");
  int max_q = 0;
  for (auto& j : jobs) max_q = std::max(max_q, j.nq);
  if (max_q == 0) return;
  DBuf<KnnJobDev> d_jobs(ctx, jobs.size());
  ctx->upload_small(d_jobs.p, jobs.data(), jobs.size() * sizeof(KnnJobDev));
  const float inf = __builtin_inff();
  float maxr2 = (max_dist == inf) ? inf : max_dist * max_dist;
  dim3 grid(ceil_div(max_q, 128), (unsigned)jobs.size());
  if (k == 1) knn_kernel<1><<<grid, 128, 0, ctx->stream>>>(d_jobs.p, k, maxr2);
  else if (k <= 8) knn_kernel<8><<<grid, 128, 0, ctx->stream>>>(d_jobs.p, k, maxr2);
  else if (k <= 16) knn_kernel<16><<<grid, 128, 0, ctx->stream>>>(d_jobs.p, k, maxr2);
  else knn_kernel<32><<<grid, 128, 0, ctx->stream>>>(d_jobs.p, k, maxr2);
  ctx_count_launches(ctx, 1);
  PGS_LAUNCH_CHECK();
}

}  // namespace

void knn_batched(Ctx* ctx, const std::vector<KnnJob>& jobs, int k, float max_dist) {
  std::vector<KnnJobDev> dj;
  for (auto& j : jobs) dj.push_back(KnnJobDev{j.tree, j.queries, j.qperm, j.nq, j.ids, j.d2});
  launch(ctx, dj, k, max_dist);
}

void knn_self_batched(Ctx* ctx, const std::vector<const Index*>& idx, int k, float max_dist,
                      const std::vector<int32_t*>& ids, const std::vector<float*>& d2) {
  // queries = the index' own sorted points; column = original index (pts.w)
  std::vector<KnnJobDev> dj;
  std::vector<DBuf<int>> perms;
  perms.reserve(idx.size());
  for (size_t b = 0; b < idx.size(); ++b) {
    perms.emplace_back(ctx, (size_t)std::max(idx[b]->n, 1));
    if (idx[b]->n > 0) {
      perm_from_sorted_kernel<<<ceil_div(idx[b]->n, 256), 256, 0, ctx->stream>>>(idx[b]->pts.p, idx[b]->n,
                                                                                 perms.back().p);
      ctx_count_launches(ctx, 1);
    }
    dj.push_back(KnnJobDev{idx[b]->view(), idx[b]->pts.p, perms.back().p, idx[b]->n, ids[b], d2[b]});
  }
  launch(ctx, dj, k, max_dist);
}

}  // namespace pgs
