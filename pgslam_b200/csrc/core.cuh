// core.cuh — device-resident DataPoints and the batched primitives every stage
// is built from (sort, spatial index, kNN).  All primitives are batched over a
// list of clouds with blockIdx.y = cloud, because the throughput configuration
// (4096 loop-closure candidate pairs) runs many independent
// registrations concurrently on one GPU.
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "common.cuh"

namespace pgs {

// ---------------------------------------------------------------------------
// DataPoints on the device (types.h:20).  features: N float4 {x,y,z,1};
// each descriptor block is stored point-major (span floats per point), which
// is exactly PM's column-major span x N layout.
// ---------------------------------------------------------------------------
struct Desc {
  std::string label;
  int span = 0;
  DBuf<float> data;
};

struct Index;

struct Cloud {
  Ctx* ctx = nullptr;
  int64_t n = 0;
  DBuf<float4> feat;
  std::vector<Desc> descs;
  // kd order of `feat` (index.cu), valid while the points are untouched; every
  // function that moves, adds or removes points calls touch()
  std::shared_ptr<DBuf<uint32_t>> kd_order;
  int64_t kd_order_n = -1;
  std::shared_ptr<Index> index_cache;  // index of the un-shifted points, same validity
  void touch() { kd_order.reset(); kd_order_n = -1; index_cache.reset(); }
  // set when the features are still being uploaded on the context's copy stream:
  // the compute stream joins that copy at the FIRST USE of the cloud, not at its
  // creation, so an upload queued for the next batch overlaps this batch's kernels
  mutable cudaEvent_t ready = nullptr;
  void wait_ready() const {
    if (!ready) return;
    cudaStreamWaitEvent(ctx->stream, ready, 0);
    cudaEventDestroy(ready);
    ready = nullptr;
  }
  ~Cloud() { wait_ready(); }

  explicit Cloud(Ctx* c) : ctx(c) {}
  Desc* find(const std::string& label) {
    for (auto& d : descs)
      if (d.label == label) return &d;
    return nullptr;
  }
  const Desc* find(const std::string& label) const { return const_cast<Cloud*>(this)->find(label); }
  Desc& add(const std::string& label, int span);  // (re)allocates span*n floats, zeroed
  void remove(const std::string& label);
  // deep copy; `target` = the context (stream) the copy is made on and belongs to
  std::unique_ptr<Cloud> clone(Ctx* target = nullptr) const;
};

void concatenate_cloud(Cloud& a, const Cloud& b);  // DP::concatenate

// keep points flagged in `keep` (n ints, 0/1), stable; returns the new count.
// One host sync (the new size is data dependent).
int64_t compact_cloud(Cloud& c, const int* d_keep);
// keep points listed in d_src[0..m) (ascending source indices)
void gather_cloud(Cloud& c, const int* d_src, int64_t m);

// ---------------------------------------------------------------------------
// batched LSD radix sort of (key, value) pairs, 8 bits per pass.
// Layout: job b owns [b*stride, b*stride + n[b]) of every array.
// Returns true if the sorted result is in the *_b arrays.
// ---------------------------------------------------------------------------
static constexpr int kSortRounds = 16;
static constexpr int kSortChunk = 32 * kSortRounds;  // keys per warp-chunk

template <typename K>
bool radix_sort_pairs(Ctx* ctx, K* keys_a, K* keys_b, uint32_t* vals_a, uint32_t* vals_b,
                      const int* d_n, int n_jobs, int stride, int max_n, int key_bits);

// device-wide exclusive scan of ints (single job), used by compaction
void exclusive_scan_int(Ctx* ctx, const int* d_in, int* d_out, int n, int* d_total);

// ---------------------------------------------------------------------------
// spatial index: points in balanced kd-tree order (kdorder.cu), cut into leaves of kLeaf
// consecutive points, under an implicit complete binary tree of AABBs
// (1-based heap numbering: children of i are 2i, 2i+1; leaf j is node P+j).
// ---------------------------------------------------------------------------
static constexpr int kLeaf = 8;  // 8 float4 = one 128-byte line per leaf

struct TreeView {
  const float4* pts;   // sorted; w = original index (int bits); padded to n_leaves*kLeaf
  const float* nodes;  // 2*P nodes x {lo.x, lo.y, hi.x, hi.y, lo.z, hi.z}; node 0 unused
  const float* cells;  // 2*P exclusive cells, same layout (index.cu cell_levels_kernel); may be null
  int n;
  int n_leaves;
  int P;      // leaves rounded up to a power of two
  int depth;  // log2(P)
};

struct Index {
  Ctx* ctx = nullptr;
  int n = 0, n_leaves = 0, P = 1, depth = 0;
  DBuf<float4> pts;
  DBuf<float> nodes;
  DBuf<float> cells;
  TreeView view() const { return TreeView{pts.p, nodes.p, cells.p, n, n_leaves, P, depth}; }
};

// Build one index per input cloud (features only).  If d_shift != nullptr,
// job b's points are translated by -shift[b] (fp32 subtraction, the
// mean-centring of ICP::compute) before being stored; boxes are built from the
// stored coordinates.
enum class IndexOrder { Kd, Morton };
void build_indices(Ctx* ctx, const std::vector<const float4*>& d_pts, const std::vector<int>& n,
                   const float* d_shift /* 4 floats per job or nullptr */,
                   std::vector<std::unique_ptr<Index>>& out, IndexOrder order_kind = IndexOrder::Kd,
                   const std::vector<const uint32_t*>* given_order = nullptr,
                   std::vector<DBuf<uint32_t>>* keep_order = nullptr);
// same for clouds; caches / reuses the kd order in the Cloud objects
void build_indices_for_clouds(Ctx* ctx, const std::vector<Cloud*>& clouds, const float* d_shift,
                              std::vector<std::unique_ptr<Index>>& out);
// un-shifted index of each cloud, built once and cached in the cloud
void cached_indices_for_clouds(Ctx* ctx, const std::vector<Cloud*>& clouds, std::vector<std::shared_ptr<Index>>& out);
// mean-centred copies of existing indices (one streaming kernel)
void derive_shifted_indices(Ctx* ctx, const std::vector<const Index*>& src, const float* d_shift,
                            std::vector<std::unique_ptr<Index>>& out);

// ---------------------------------------------------------------------------
// exact kNN (eps = 0, ties -> lower original index), k <= 32.
// Queries are float4 (w ignored).  Outputs k x nq column-major, ORIGINAL
// reference indices; unfound -> -1 / +inf.  If d_qperm != nullptr, query j's
// result is written to column d_qperm[j] (queries given in sorted order).
// ---------------------------------------------------------------------------
struct KnnJob {
  TreeView tree;
  const float4* queries;
  const int* qperm;
  int nq;
  int32_t* ids;
  float* d2;
};
void knn_batched(Ctx* ctx, const std::vector<KnnJob>& jobs, int k, float max_dist);

// self-kNN of an index' own (sorted) points, results in original order.
void knn_self_batched(Ctx* ctx, const std::vector<const Index*>& idx, int k, float max_dist,
                      const std::vector<int32_t*>& ids, const std::vector<float*>& d2);

}  // namespace pgs
