// callsites.cpp — pgslam's own call sequences, spelled against the adapter the
// way the reference spells them against libpointmatcher, to show the drop-in
// compiles and behaves:
//   LoopCloser::ProcessVertex / CheckIcpResult / ComputeResidualError
//       (LoopCloser.hpp:83-110, 308-340, 343-365)
//   Localizer::ComputeOverlapWith                      (Localizer.hpp:282-348)
//   LocalMap::BuildCloudFromData                       (LocalMap.hpp:209-224)
//   Localizer::ProcessFirstCloud / ProcessData         (Localizer.hpp:103-148)
//   + the plugin registrar (PM::get().REG(...).create), T = double (tests/instantiation.cpp:6,10)
//     and the batched candidate loop (LoopCloser.hpp:266-297 -> ICP::computeBatch)
// Usage: callsites <reading.bin> <reference.bin> <icp.yaml> <filters.yaml>
// (clouds: int32 n, then n x 4 float32).  Prints key=value lines.
//        callsites --bench <points>   times the ProcessVertex + CheckIcpResult + ComputeResidualError
//        sequence on a synthetic pair of that size and prints one JSON line ({"dropin_ms": ...}).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

#include "pgslam_b200/pm_adapter.hpp"

// the one block a pgslam maintainer changes (types.h:19-27)
template <typename T>
struct Types {
  using PM = pgslam_b200::PointMatcher<T>;
  using DP = typename PM::DataPoints;
  using Matrix = typename PM::Matrix;
  using ICP = typename PM::ICP;
  using ICPSequence = typename PM::ICPSequence;
  using TransformationPtr = std::shared_ptr<typename PM::Transformation>;
  using DataPointsFilters = typename PM::DataPointsFilters;
};
using TY = Types<float>;
using PM = TY::PM;
using DP = TY::DP;
using Matrix = TY::Matrix;

static DP load_cloud(const char* path) {
  std::ifstream f(path, std::ios::binary);
  int32_t n = 0;
  f.read(reinterpret_cast<char*>(&n), 4);
  std::vector<float> buf(static_cast<size_t>(n) * 4);
  f.read(reinterpret_cast<char*>(buf.data()), buf.size() * 4);
  DP dp;
  dp.features.resize(4, n);
  for (int i = 0; i < n; ++i)
    for (int r = 0; r < 4; ++r) dp.features(r, i) = buf[static_cast<size_t>(i) * 4 + r];
  return dp;
}
static std::string slurp(const char* path) {
  std::ifstream f(path);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}
static void print_T(const char* key, const Matrix& T) {
  std::printf("%s=", key);
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) std::printf("%.9g%s", static_cast<double>(T(r, c)), (c == 3 && r == 3) ? "\n" : ",");
}

// ---- T = double through the same call sites: poses must come back in double ------------------
template <typename S>
static void run_typed(const DP& rd_f, const DP& rf_f, const std::string& yaml, const char* tag) {
  using PMS = pgslam_b200::PointMatcher<S>;
  typename PMS::DataPoints rd, rf;
  auto convert = [](const DP& in, typename PMS::DataPoints& out) {
    const int n = static_cast<int>(in.features.cols());
    out.features.resize(4, n);
    out.descriptors.resize(static_cast<int>(in.descriptors.rows()), n);
    for (int i = 0; i < n; ++i) {
      for (int r = 0; r < 4; ++r) out.features(r, i) = static_cast<S>(in.features(r, i));
      for (int r = 0; r < static_cast<int>(in.descriptors.rows()); ++r) out.descriptors(r, i) = static_cast<S>(in.descriptors(r, i));
    }
    for (auto& l : in.descriptorLabels) out.descriptorLabels.emplace_back(l.text, l.span);
  };
  convert(rd_f, rd);
  convert(rf_f, rf);
  typename PMS::ICP icp;
  std::istringstream iss{yaml};
  icp.loadFromYaml(iss);
  typename PMS::Matrix T = icp(rd, rf);
  std::printf("%s=", tag);
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) std::printf("%.17g%s", static_cast<double>(T(r, c)), (c == 3 && r == 3) ? "\n" : ",");
}

// ---- drop-in cost: what pgslam pays per loop-closure vertex through the adapter ------------------
static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static DP synthetic_cloud(int n, float shift) {
  // a corridor-like shell (two walls + floor), deterministic; only its size matters here
  DP dp;
  dp.features.resize(4, n);
  unsigned s = 12345u;
  auto rnd = [&s]() { s = s * 1664525u + 1013904223u; return static_cast<float>(s >> 8) * (1.0f / 16777216.0f); };
  for (int i = 0; i < n; ++i) {
    const float u = rnd() * 40.f - 20.f, v = rnd() * 3.f;
    const int face = i % 3;
    float x = u, y = face == 0 ? -4.f : (face == 1 ? 4.f : rnd() * 8.f - 4.f), z = face == 2 ? 0.f : v;
    x += 0.3f * std::sin(0.7f * u);
    dp.features(0, i) = x + shift; dp.features(1, i) = y + 0.5f * shift; dp.features(2, i) = z; dp.features(3, i) = 1.f;
  }
  return dp;
}
static int bench_main(int n) {
  const std::string yaml =
      "referenceDataPointsFilters:\n  - SurfaceNormalDataPointsFilter:\n      knn: 10\n"
      "matcher:\n  KDTreeMatcher:\n    knn: 1\n"
      "outlierFilters:\n  - TrimmedDistOutlierFilter:\n      ratio: 0.85\n"
      "errorMinimizer:\n  PointToPlaneWithCovErrorMinimizer\n"
      "transformationCheckers:\n  - CounterTransformationChecker:\n      maxIterationCount: 40\n"
      "  - DifferentialTransformationChecker:\n      minDiffRotErr: 0.001\n      minDiffTransErr: 0.001\n      smoothLength: 3\n";
  DP input_cloud = synthetic_cloud(n, 0.05f), candidate_cloud = synthetic_cloud(n, 0.f);
  TY::ICP icp_;
  {
    std::istringstream iss{yaml};
    icp_.loadFromYaml(iss);
  }
  double best = 1e30, best_cached = 1e30, first = 0;
  double ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // phases of the best cold repetition
  int iterations = 0;
  constexpr int kCold = 7, kReps = 10;  // the first call, six cold calls (best of), three with unchanged clouds
  for (int rep = 0; rep < kReps; ++rep) {
    if (rep < kCold) {  // cold: new host data every time, as for a new keyframe
      input_cloud.features(0, 0) += 1e-3f;
      candidate_cloud.features(0, 0) += 1e-3f;
    }
    const double t0 = now_ms();
    double tp[9];
    tp[0] = t0;
    // ProcessVertex (LoopCloser.hpp:98)
    Matrix T = icp_(input_cloud, candidate_cloud, Matrix::Identity(4, 4));
    tp[1] = now_ms();
    // CheckIcpResult (LoopCloser.hpp:308-340)
    volatile bool reached = icp_.getMaxNumIterationsReached();
    volatile float overlap = icp_.errorMinimizer->getOverlap();
    (void)reached; (void)overlap;
    // ComputeResidualError (LoopCloser.hpp:343-365)
    TY::ICP temp_icp;
    std::istringstream iss{yaml};
    temp_icp.loadFromYaml(iss);
    tp[2] = now_ms();
    DP reading(input_cloud);
    temp_icp.transformations.apply(reading, T);
    tp[3] = now_ms();
    DP reference(candidate_cloud);
    temp_icp.referenceDataPointsFilters.apply(reference);
    tp[4] = now_ms();
    temp_icp.matcher->init(reference);
    tp[5] = now_ms();
    auto matches = temp_icp.matcher->findClosests(reading);
    tp[6] = now_ms();
    auto w = temp_icp.outlierFilters.compute(reading, reference, matches);
    tp[7] = now_ms();
    volatile float residual = temp_icp.errorMinimizer->getResidualError(reading, reference, w, matches);
    (void)residual;
    tp[8] = now_ms();
    const double dt = tp[8] - t0;
    if (rep == 0) first = dt;
    else if (rep < kCold) {
      if (dt < best)
        for (int i = 0; i < 8; ++i) ph[i] = tp[i + 1] - tp[i];
      best = std::min(best, dt);
    } else best_cached = std::min(best_cached, dt);
    iterations = icp_.lastResult().iterations;
  }
  std::printf("{\"dropin_ms\": %.3f, \"dropin_ms_unchanged_clouds\": %.3f, \"first_call_ms\": %.3f, \"points\": %d, "
              "\"iterations\": %d, \"phases_ms\": {\"icp\": %.3f, \"checks_and_yaml\": %.3f, \"copy_transform\": %.3f, "
              "\"reference_filters\": %.3f, \"matcher_init\": %.3f, \"find_closests\": %.3f, \"outlier_weights\": %.3f, "
              "\"residual\": %.3f}, \"sequence\": \"LoopCloser::ProcessVertex + CheckIcpResult + ComputeResidualError "
              "through pm_adapter.hpp, host DataPoints in, host results out\"}\n",
              best, best_cached, first, n, iterations, ph[0], ph[1], ph[2], ph[3], ph[4], ph[5], ph[6], ph[7]);
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 3 && std::string(argv[1]) == "--bench") return bench_main(std::atoi(argv[2]));
  if (argc < 5) return 2;
  DP input_cloud = load_cloud(argv[1]);
  DP candidate_cloud = load_cloud(argv[2]);
  const std::string icp_config_buffer_ = slurp(argv[3]);
  const std::string filters_config = slurp(argv[4]);

  // Localizer::SetInputFiltersConfig + ProcessData (Localizer.hpp:74-78, 103-106)
  std::istringstream fss{filters_config};
  TY::DataPointsFilters input_filters_(fss);
  input_filters_.apply(input_cloud);
  input_filters_.apply(candidate_cloud);
  std::printf("descriptors=%zu\n", input_cloud.descriptorLabels.size());
  TY::TransformationPtr rigid_transformation_ = PM::get().REG(Transformation).create("RigidTransformation");

  // LoopCloser::SetIcpConfig + ProcessVertex (LoopCloser.hpp:59-74, 83-110)
  TY::ICP icp_;
  {
    std::istringstream iss{icp_config_buffer_};
    icp_.loadFromYaml(iss);
  }
  Matrix input_T_refkf_kf = Matrix::Identity(4, 4);
  Matrix T_refkf_kf_ = icp_(input_cloud, candidate_cloud, input_T_refkf_kf);
  print_T("T", T_refkf_kf_);
  // CheckIcpResult (LoopCloser.hpp:308-340)
  std::printf("max_iter_reached=%d\n", icp_.getMaxNumIterationsReached() ? 1 : 0);
  std::printf("iterations=%d\n", icp_.lastResult().iterations);
  std::printf("overlap=%.9g\n", static_cast<double>(icp_.errorMinimizer->getOverlap()));
  Matrix cov = icp_.errorMinimizer->getCovariance();
  std::printf("cov00=%.9g\n", static_cast<double>(cov(0, 0)));
  {
    // ComputeResidualError (LoopCloser.hpp:343-365)
    TY::ICP temp_icp;
    std::istringstream iss{icp_config_buffer_};
    temp_icp.loadFromYaml(iss);
    DP reading(input_cloud);
    temp_icp.transformations.apply(reading, T_refkf_kf_);
    temp_icp.matcher->init(candidate_cloud);
    auto matches = temp_icp.matcher->findClosests(reading);
    auto outlier_weights = temp_icp.outlierFilters.compute(reading, candidate_cloud, matches);
    float residual = temp_icp.errorMinimizer->getResidualError(reading, candidate_cloud, outlier_weights, matches);
    std::printf("residual=%.9g\n", static_cast<double>(residual));
  }
  {
    // ComputeOverlapWith (Localizer.hpp:309-347)
    using Matches = typename PM::Matches;
    using OutlierWeights = typename PM::OutlierWeights;
    using ErrorElements = typename PM::ErrorMinimizer::ErrorElements;
    TY::ICP temp_icp;
    std::istringstream iss{icp_config_buffer_};
    temp_icp.loadFromYaml(iss);
    DP reference(candidate_cloud);
    temp_icp.referenceDataPointsFilters.init();
    temp_icp.referenceDataPointsFilters.apply(reference);
    temp_icp.matcher->init(reference);
    DP reading(input_cloud);
    temp_icp.readingDataPointsFilters.init();
    temp_icp.readingDataPointsFilters.apply(reading);
    reading = rigid_transformation_->compute(reading, T_refkf_kf_);
    temp_icp.readingStepDataPointsFilters.init();
    temp_icp.readingStepDataPointsFilters.apply(reading);
    const Matches matches(temp_icp.matcher->findClosests(reading));
    const OutlierWeights outlierWeights(temp_icp.outlierFilters.compute(reading, reference, matches));
    ErrorElements matchedPoints(reading, reference, outlierWeights, matches);
    std::printf("weighted_ratio=%.9g\n", static_cast<double>(matchedPoints.weightedPointUsedRatio));
  }
  {
    // LocalMap::BuildCloudFromData (LocalMap.hpp:209-224) then ICPSequence (Localizer.hpp:126,148)
    DP cloud_ = candidate_cloud;
    Matrix T_refkf_world = Matrix::Identity(4, 4).inverse();
    cloud_.concatenate(rigid_transformation_->compute(input_cloud, T_refkf_world * T_refkf_kf_));
    std::printf("local_map_points=%u\n", cloud_.getNbPoints());
    TY::ICPSequence icp_sequence_;
    std::istringstream iss{icp_config_buffer_};
    icp_sequence_.loadFromYaml(iss);
    std::printf("has_map_before=%d\n", icp_sequence_.hasMap() ? 1 : 0);
    icp_sequence_.setMap(cloud_);
    Matrix T = icp_sequence_(input_cloud, T_refkf_kf_);
    print_T("T_seq", T);
    std::printf("seq_iterations=%d\n", icp_sequence_.lastResult().iterations);
    // ICPSequence without a map: identity, no throw (upstream logs a warning)
    TY::ICPSequence empty_seq;
    Matrix Te = empty_seq(input_cloud, T_refkf_kf_);
    std::printf("no_map_identity=%d\n", (Te(0, 0) == 1.f && Te(0, 3) == 0.f && Te(1, 0) == 0.f) ? 1 : 0);
  }
  {
    // the plugin registrar: PM::get().REG(kind).create(name, params), YAML-free
    PM::Parameters p;
    p["prob"] = "0.5";
    auto rs = PM::get().REG(DataPointsFilter).create("RandomSamplingDataPointsFilter", p);
    DP sub = rs->filter(candidate_cloud);
    std::printf("reg_filter_points=%u\n", sub.getNbPoints());
    std::printf("reg_filter_class=%s\n", rs->className.c_str());
    std::printf("reg_filter_seed_default=%s\n", rs->getParamValueString("seed").c_str());
    std::printf("reg_filter_nparams=%zu\n", rs->availableParameters().size());
    auto m = PM::get().REG(Matcher).create("KDTreeMatcher", {{"knn", "2"}});
    m->init(candidate_cloud);
    auto mm = m->findClosests(input_cloud);
    std::printf("reg_matcher_k=%d\n", static_cast<int>(mm.ids.rows()));
    std::printf("reg_matcher_id0=%d\n", mm.ids(0, 0));
    auto of = PM::get().REG(OutlierFilter).create("TrimmedDistOutlierFilter", {{"ratio", "0.5"}});
    auto w = of->compute(input_cloud, candidate_cloud, mm);
    double ws = 0;
    for (int i = 0; i < static_cast<int>(w.cols()); ++i) ws += w(0, i) + w(1, i);
    std::printf("reg_outlier_ratio=%.6f\n", ws / (2.0 * w.cols()));
    PM::OutlierFilters chain;
    chain.push_back(of);
    auto w2 = chain.compute(input_cloud, candidate_cloud, mm);
    std::printf("reg_outlier_chain_equal=%d\n", (w2(0, 0) == w(0, 0) && w2(1, 5) == w(1, 5)) ? 1 : 0);
    auto em = PM::get().REG(ErrorMinimizer).create("PointToPlaneErrorMinimizer");
    Matrix Tm = em->compute(input_cloud, candidate_cloud, w, mm);
    std::printf("reg_minimizer_t03=%.9g\n", static_cast<double>(Tm(0, 3)));
    auto ck = PM::get().REG(TransformationChecker).create("CounterTransformationChecker", {{"maxIterationCount", "7"}});
    std::printf("reg_checker_max=%d\n", ck->get<int>("maxIterationCount"));
    int caught = 0;
    try { PM::get().REG(DataPointsFilter).create("NoSuchFilter"); } catch (const PM::InvalidElement&) { caught |= 1; }
    try { PM::get().REG(OutlierFilter).create("TrimmedDistOutlierFilter", {{"ratio", "7"}}); } catch (const PM::InvalidParameter&) { caught |= 2; }
    try { PM::get().REG(Matcher).create("KDTreeMatcher", {{"bogus", "1"}}); } catch (const PM::InvalidParameter&) { caught |= 4; }
    std::printf("reg_errors=%d\n", caught);
    std::printf("reg_names=%zu\n", PM::get().REG(DataPointsFilter).names().size());
  }
  {
    // times ride along through a device filter
    DP timed = candidate_cloud;
    timed.times.resize(1, static_cast<int>(timed.features.cols()));
    for (int i = 0; i < static_cast<int>(timed.features.cols()); ++i) timed.times(0, i) = 1000000007ll + i;
    timed.timeLabels.push_back(DP::Label("stamp", 1));
    auto fs = PM::get().REG(DataPointsFilter).create("FixStepSamplingDataPointsFilter", {{"startStep", "3"}, {"endStep", "3"}});
    DP kept = fs->filter(timed);
    long bad = 0;
    for (int i = 0; i < static_cast<int>(kept.features.cols()); ++i) {
      const long src = static_cast<long>(kept.times(0, i) - 1000000007ll);
      if (src < 0 || src >= static_cast<long>(timed.features.cols()) || kept.features(0, i) != timed.features(0, static_cast<int>(src))) ++bad;
    }
    std::printf("times_cols=%d\n", static_cast<int>(kept.times.cols()));
    std::printf("times_points=%u\n", kept.getNbPoints());
    std::printf("times_bad=%ld\n", bad);
  }
  {
    // the candidate loop as one batch (LoopCloser.hpp:266-297 -> :98 per candidate)
    std::vector<const DP*> readings{&input_cloud, &input_cloud, &candidate_cloud, &input_cloud, &input_cloud,
                                    &input_cloud, &input_cloud, &input_cloud, &input_cloud};
    std::vector<const DP*> references(readings.size(), &candidate_cloud);
    std::vector<Matrix> inits(readings.size(), Matrix::Identity(4, 4));
    auto res = icp_.computeBatch(readings, references, inits);
    print_T("T_batch0", res[0].transformation);
    print_T("T_batch8", res[8].transformation);
    std::printf("batch_iterations0=%d\n", res[0].iterations);
    std::printf("batch_converged=%d\n", static_cast<int>(res[0].converged && res[1].converged && res[8].converged));
    std::printf("batch_pair2_status=%d\n", static_cast<int>(res[2].status));  // identical clouds: reported, not thrown
    auto res2 = icp_.computeBatch(readings, references, inits, {0, 0});  // two contexts of device 0
    int same = 1;
    for (size_t i = 0; i < res.size(); ++i)
      for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) same &= res[i].transformation(r, c) == res2[i].transformation(r, c);
    std::printf("batch_two_contexts_equal=%d\n", same);
  }
  {
    // DataPoints::save / load (csv) next to the input file
    const std::string path = std::string(argv[1]) + ".roundtrip.csv";
    input_cloud.save(path);
    DP back = DP::load(path);
    int same = back.getNbPoints() == input_cloud.getNbPoints() && back.descriptorLabels.size() == input_cloud.descriptorLabels.size();
    for (int i = 0; same && i < static_cast<int>(back.features.cols()); i += 37)
      same = back.features(0, i) == input_cloud.features(0, i) && back.descriptors(1, i) == input_cloud.descriptors(1, i);
    std::printf("file_roundtrip=%d\n", same);
  }
  run_typed<double>(input_cloud, candidate_cloud, icp_config_buffer_, "T_double");
  return 0;
}
