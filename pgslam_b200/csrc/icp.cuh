// icp.cuh — the fused ICP engine (PM::ICP / PM::ICPSequence, types.h:24-25;
// loop spec SURVEY.md §3.3).  One engine instance == one ICPChainBase: it owns
// the parsed chain and, for ICPSequence, the prepared map.
#pragma once

#include <memory>
#include <vector>

#include "core.cuh"
#include "dense.cuh"
#include "filters.cuh"
#include "modules.h"

namespace pgs {

constexpr int kChkHist = 16;  // DifferentialTransformationChecker history (smoothLength <= 15)
constexpr int kAcc = 32;      // accumulator slots of one iteration
constexpr int kAcc2 = 44;     // accumulator slots of the final (covariance / overlap) pass
constexpr int kMaxQuant = 4;

enum { MIN_P2PLANE = 1, MIN_P2PLANE_COV = 2, MIN_P2POINT = 3 };

// chain parameters, uniform over a batch; passed to kernels by value
struct IcpParams {
  int minimizer;
  double sensor_std_dev;
  int max_iterations;  // Counter; 0 = no counter
  int has_diff;
  double min_diff_rot, min_diff_trans;
  int smooth_length;
  int has_bound;
  double max_rot, max_trans;
  int has_outliers;    // 0: weights = (dist != inf)
  int n_quant;         // quantile-based limits: hi = min_j factor_j * quantile(ratio_j)
  double q_ratio[kMaxQuant];
  float q_factor[kMaxQuant];
  int q_var[kMaxQuant];      // 1: VarTrimmedDist (ratio optimised per iteration), else fixed ratio
  float q_min[kMaxQuant], q_max[kMaxQuant];  // VarTrimmedDist minRatio / maxRatio
  double q_lambda[kMaxQuant];
  float fixed_hi, fixed_lo;  // MaxDist^2 / MinDist^2 limits
  float max_r2;        // matcher maxDist^2
  int hard_iteration_cap;
  int has_sn;          // SurfaceNormalOutlierFilter present
  float sn_eps;        // cos(maxAngle)
  int knn;             // KDTreeMatcher knn: every reading point is paired with its k nearest
  int force_mode;      // PointToPlane: 0 = 6-DOF, 1 = force2D (theta,x,y), 2 = force4DOF (yaw,x,y,z)
};

enum { FORCE_NONE = 0, FORCE_2D = 1, FORCE_4DOF = 2 };

struct PairState {
  double T_init[16];
  double T_refIn_refMean[16], T_refMean_dataIn[16];
  double T_iter[16], T_prev[16], T_inc[16], T_out[16];
  Xf xf;       // fp32 T_iter, applied to the pre-transformed reading each iteration
  Xf xf_prev;  // fp32 T_prev: the transform the last matches were computed with
  Xf xf0;      // fp32 T_refMean_dataIn
  int active, iterations, max_reached, status;
  int counter, nhist;
  double q[kChkHist][4], t[kChkHist][3];
  double q0[4], t0[3];
  float lim_lo, lim_hi;
  unsigned ticket;   // last-block detection
  unsigned ticket2;
  unsigned ticket3;  // select passes
  unsigned sel_prefix, sel_mask;
  unsigned long long sel_rank;
  double acc[kAcc];
  double acc2[kAcc2];
  double kept, wsum, resid, overlap;
  double cov[36];
};

struct PairView {
  const float4* reading;  // pre-transformed reading, Morton order (w = original index)
  int n_r;
  int n_m;                     // matches = knn * n_r, stored [point][neighbour] like PM's k x N Matches
  const int* ref_inv;          // original reference index -> sorted position (knn > 1 only)
  TreeView tree;               // reference index
  const float4* ref_normals;   // {point, normal} per sorted reference position (2 float4), or null
  const float4* rd_normals;    // per sorted reading position (pre-transformed), or null
  const float* rd_noise;       // per sorted reading position, or null
  int* match_pos;
  float* match_d2;
  double* partials;  // [grid.x][kAcc2]
  unsigned* sel_hist;  // 2048 bins, zero between passes
};

// reference side of a registration, ready for matching
struct PreparedRef {
  std::unique_ptr<Cloud> cloud;   // filtered reference (original order, NOT centred unless setMap)
  std::unique_ptr<Index> index;   // centred, sorted
  DBuf<float4> normals_sorted;
  DBuf<int> inv_pos;              // original index -> sorted position, built on first use (knn > 1)
  bool has_normals = false;
  // T_refIn_refMean (identity + the reference mean) stays on the device: 16 doubles at
  // T_mean->p + T_mean_off, shared by the references prepared together
  std::shared_ptr<DBuf<double>> T_mean;
  size_t T_mean_off = 0;
  // tensor-core distance tiles of a small reference (dense.cu), built on first use
  mutable std::unique_ptr<DenseRef> dense;
};

// Where a batch's clouds come from.  fetch() makes the pairs [lo, hi) available on `ctx` and may
// start asynchronous uploads; before_run() is called right before those pairs are registered
// (the point where the compute stream joins the uploads).  A worker fetches its NEXT chunk before
// it registers the current one, so uploads overlap the registration of the chunk before.
struct FetchedPairs {
  std::vector<std::unique_ptr<Cloud>> owned;  // uploaded for this chunk; freed with it
  std::vector<const Cloud*> readings, references;
};
struct PairSource {
  virtual ~PairSource() {}
  virtual void fetch(Ctx* ctx, int lo, int hi, FetchedPairs& out) = 0;
  virtual void before_run(Ctx* ctx, FetchedPairs& f) {
    (void)ctx;
    (void)f;
  }
};

class IcpEngine {
 public:
  IcpEngine(Ctx* ctx, const ChainConfig& cfg);
  Ctx* ctx() const { return ctx_; }
  const ChainConfig& config() const { return cfg_; }

  // ICP::operator() on P independent pairs
  void run_batch(const std::vector<const Cloud*>& readings, const std::vector<const Cloud*>& references,
                 const double* T_inits, pgs_icp_result* results);
  // the same for P pairs that `src` delivers chunk by chunk (host-resident clouds: BASELINE C4)
  void run_batch_source(int P, PairSource& src, const double* T_inits, pgs_icp_result* results);
  // ICPSequence
  void set_map(const Cloud& map);
  bool has_map() const { return map_ != nullptr; }
  void run_sequence(const Cloud& reading, const double* T_init, pgs_icp_result* out);

 private:
  void prepare_references(std::vector<std::unique_ptr<Cloud>>& refs, bool centre_first,
                          std::vector<std::unique_ptr<PreparedRef>>& out);
  void run_direct(const std::vector<const Cloud*>& readings, const std::vector<const Cloud*>& references,
                  const double* T_inits, pgs_icp_result* results);
  void run_prepared(const std::vector<const Cloud*>& readings, const std::vector<const PreparedRef*>& refs,
                    const double* T_inits, pgs_icp_result* results);
  IcpParams params_;
  Ctx* ctx_;
  ChainConfig cfg_;
  std::unique_ptr<PreparedRef> map_;
  cudaEvent_t idx_ev_[2] = {nullptr, nullptr};  // profiling: reference-side preparation
  bool have_idx_ev_ = false;
};

// fine-grained modules (pgs_matcher / pgs_outliers / pgs_minimizer)
IcpParams params_from_chain(const ChainConfig& cfg);
void outlier_limits_params(const std::vector<Module>& filters, IcpParams* p);
// OutlierFilters::compute on device arrays (k x n dists) -> weights
void outlier_weights_device(Ctx* ctx, const std::vector<Module>& filters, const float* d_d2, int64_t nk, float* d_w,
                            const Cloud* reading = nullptr, const Cloud* reference = nullptr,
                            const int32_t* d_ids = nullptr, int k = 1);
// ErrorElements + ErrorMinimizer::compute on explicit matches
void minimize_device(Ctx* ctx, const Module& minimizer, const Cloud& reading, const Cloud& reference,
                     const int32_t* d_ids, const float* d_d2, const float* d_w, int k, pgs_min_result* out);

// ErrorElements ratios only: kept = #(w != 0 and dist != inf), wsum = sum of those weights
void weights_ratio_device(Ctx* ctx, const float* d_d2, const float* d_w, int64_t nk, double* kept, double* wsum);

}  // namespace pgs
