// callsites.cpp — pgslam's own call sequences, spelled against the adapter the
// way the reference spells them against libpointmatcher, to show the drop-in
// compiles and behaves:
//   LoopCloser::ProcessVertex / CheckIcpResult / ComputeResidualError
//       (LoopCloser.hpp:83-110, 308-340, 343-365)
//   Localizer::ComputeOverlapWith                      (Localizer.hpp:282-348)
//   LocalMap::BuildCloudFromData                       (LocalMap.hpp:209-224)
//   Localizer::ProcessFirstCloud / ProcessData         (Localizer.hpp:103-148)
// Usage: callsites <reading.bin> <reference.bin> <icp.yaml> <filters.yaml>
// (clouds: int32 n, then n x 4 float32).  Prints key=value lines.
#include <cstdio>
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

int main(int argc, char** argv) {
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
  }
  return 0;
}
