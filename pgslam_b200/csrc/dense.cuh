// dense.cuh — tensor-core distance-tile matcher for small reference clouds (dense.cu).
#pragma once

#include <vector>

#include "core.cuh"

namespace pgs {

constexpr int kDenseMaxRef = 16384;  // reference points a dense search accepts (128 tiles of 128)

struct DenseRef {
  DBuf<float> tiles;   // n_blocks x 8 KB, MMA-ready (K-major, no swizzle)
  DBuf<float4> info;   // n_blocks x {centre.xyz, radius}
  int n = 0, n_blocks = 0;
};

// tiles of an index' kd-ordered points (once per reference cloud)
void dense_prepare(Ctx* ctx, const Index& idx, DenseRef& out);

struct DenseQuery {
  const DenseRef* ref;
  TreeView tree;          // the reference index (exact re-test of the candidates, fallback search)
  const float4* queries;
  int nq;
  int32_t* ids;           // original reference index or -1 (matcher module) ...
  float* d2;              // may be null
  int* out_pos = nullptr; // ... or sorted position (ICP loop), with d2 mandatory
  const Xf* xf = nullptr; // transform applied to every query first (device pointer) or null
  const int* active = nullptr;  // the job is skipped when *active == 0
};
// the same in two steps for a loop that repeats the search: the job table is uploaded once
struct DenseJobs {
  DBuf<unsigned char> table;
  int n_jobs = 0, max_q = 0;
};
void dense_upload_jobs(Ctx* ctx, const std::vector<DenseQuery>& queries, DenseJobs& out);
void dense_launch(Ctx* ctx, const DenseJobs& jobs, float maxr2);

// exact k = 1 nearest neighbour (ties -> lower index) of every query; returns the number of
// queries that needed the exact fallback scan when count_fallbacks (one host sync)
unsigned dense_knn1(Ctx* ctx, const std::vector<DenseQuery>& queries, float max_dist, bool count_fallbacks);

}  // namespace pgs
