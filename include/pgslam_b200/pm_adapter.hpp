// pm_adapter.hpp — header-only C++14 adapter that gives libpgslam_b200's C ABI
// the class shapes pgslam binds to at /root/reference/src/pgslam/types.h:19-27:
//
//     using PM  = PointMatcher<T>;          ->  pgslam_b200::PointMatcher<T>
//     using DP  = typename PM::DataPoints;
//     using ICP = typename PM::ICP;         using ICPSequence = typename PM::ICPSequence;
//     using TransformationPtr = std::shared_ptr<typename PM::Transformation>;
//     using DataPointsFilters = typename PM::DataPointsFilters;
//
// Only the members pgslam actually touches are provided; each
// one cites the call site that needs it.  With Eigen available the matrix
// types are Eigen's (bit-compatible with libpointmatcher); without it a minimal
// column-major stand-in keeps this header compilable (this container has no
// Eigen).  Failures of the C ABI are re-thrown as the libpointmatcher exception
// types of the same name, so pgslam (which catches none) behaves identically.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <istream>
#include <iterator>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../pgslam_b200.h"

#if defined(PGSLAM_B200_USE_EIGEN) || (defined(__has_include) && __has_include(<Eigen/Core>))
#include <Eigen/Core>
#include <Eigen/LU>
#define PGSLAM_B200_HAVE_EIGEN 1
#endif

namespace pgslam_b200 {

#ifndef PGSLAM_B200_HAVE_EIGEN
// Minimal column-major dynamic matrix with the handful of Eigen members pgslam
// uses on poses and clouds: (r,c), rows(), cols(), Identity, inverse(), operator*.
template <typename S>
class MiniMatrix {
 public:
  MiniMatrix() : r_(0), c_(0) {}
  MiniMatrix(int r, int c) : r_(r), c_(c), d_(static_cast<size_t>(r) * c, S(0)) {}
  static MiniMatrix Identity(int r, int c) {
    MiniMatrix m(r, c);
    for (int i = 0; i < (r < c ? r : c); ++i) m(i, i) = S(1);
    return m;
  }
  static MiniMatrix Zero(int r, int c) { return MiniMatrix(r, c); }
  int rows() const { return r_; }
  int cols() const { return c_; }
  S& operator()(int r, int c) { return d_[static_cast<size_t>(c) * r_ + r]; }
  const S& operator()(int r, int c) const { return d_[static_cast<size_t>(c) * r_ + r]; }
  S* data() { return d_.data(); }
  const S* data() const { return d_.data(); }
  void resize(int r, int c) { r_ = r; c_ = c; d_.assign(static_cast<size_t>(r) * c, S(0)); }
  MiniMatrix operator*(const MiniMatrix& o) const {
    MiniMatrix m(r_, o.c_);
    for (int c = 0; c < o.c_; ++c)
      for (int k = 0; k < c_; ++k)
        for (int r = 0; r < r_; ++r) m(r, c) += (*this)(r, k) * o(k, c);
    return m;
  }
  // inverse of a homogeneous rigid transform (the only inverse pgslam takes:
  // LoopCloser.hpp:95, LocalMap.hpp:216, Localizer.hpp:119)
  MiniMatrix inverse() const {
    if (r_ != 4 || c_ != 4) throw std::logic_error("MiniMatrix::inverse: 4x4 rigid transforms only");
    MiniMatrix m = Identity(4, 4);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) m(r, c) = (*this)(c, r);
    for (int r = 0; r < 3; ++r) {
      S s = 0;
      for (int k = 0; k < 3; ++k) s += m(r, k) * (*this)(k, 3);
      m(r, 3) = -s;
    }
    return m;
  }

 private:
  int r_, c_;
  std::vector<S> d_;
};
#endif

template <typename T>
struct PointMatcher {
  static_assert(sizeof(T) == 4 || sizeof(T) == 8, "T must be float or double");
#ifdef PGSLAM_B200_HAVE_EIGEN
  using Matrix = Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic>;
  using IntMatrix = Eigen::Matrix<int, Eigen::Dynamic, Eigen::Dynamic>;
#else
  using Matrix = MiniMatrix<T>;
  using IntMatrix = MiniMatrix<int>;
#endif
  using TransformationParameters = Matrix;
  using OutlierWeights = Matrix;

  // ---- exceptions, as libpointmatcher names them -------------------------------
  struct ConvergenceError : std::runtime_error { using std::runtime_error::runtime_error; };
  struct TransformationError : std::runtime_error { using std::runtime_error::runtime_error; };
  struct InvalidParameter : std::runtime_error { using std::runtime_error::runtime_error; };
  struct InvalidField : std::runtime_error { using std::runtime_error::runtime_error; };
  struct InvalidModuleType : std::runtime_error { using std::runtime_error::runtime_error; };
  struct InvalidElement : std::runtime_error { using std::runtime_error::runtime_error; };

  // one context (device 0, own stream) per host thread: pgslam-MT drives the
  // localizer and the loop closer from two threads (LocalizerMT.hpp:47, LoopCloserMT.hpp:41)
  static pgs_ctx* context() {
    struct Holder {
      pgs_ctx* c = nullptr;
      Holder() {
        if (pgs_ctx_create(0, nullptr, &c) != PGS_OK) throw std::runtime_error(pgs_last_error(nullptr));
      }
      ~Holder() { pgs_ctx_destroy(c); }
    };
    static thread_local Holder h;
    return h.c;
  }
  static void check(pgs_status st) {
    if (st == PGS_OK) return;
    const std::string msg = pgs_last_error(context());
    switch (st) {
      case PGS_CONVERGENCE_ERROR: throw ConvergenceError(msg);
      case PGS_TRANSFORMATION_ERROR: throw TransformationError(msg);
      case PGS_INVALID_PARAMETER: throw InvalidParameter(msg);
      case PGS_INVALID_FIELD: throw InvalidField(msg);
      case PGS_INVALID_MODULE_TYPE: throw InvalidModuleType(msg);
      case PGS_INVALID_ELEMENT: throw InvalidElement(msg);
      default: throw std::runtime_error(msg);
    }
  }
  static void to_double16(const Matrix& M, double* out) {
    for (int c = 0; c < 4; ++c)
      for (int r = 0; r < 4; ++r) out[c * 4 + r] = static_cast<double>(M(r, c));
  }
  static Matrix from_double16(const double* in) {
    Matrix M = Matrix::Identity(4, 4);
    for (int c = 0; c < 4; ++c)
      for (int r = 0; r < 4; ++r) M(r, c) = static_cast<T>(in[c * 4 + r]);
    return M;
  }

  // ---- DataPoints (types.h:20) ----------------------------------------------------
  struct DataPoints {
    struct Label {
      std::string text;
      size_t span;
      Label(const std::string& t = "", size_t s = 0) : text(t), span(s) {}
    };
    using Labels = std::vector<Label>;
    Matrix features;      // 4 x N, last row 1
    Matrix descriptors;   // D x N
    Labels featureLabels, descriptorLabels;

    DataPoints() {}
    unsigned getNbPoints() const { return static_cast<unsigned>(features.cols()); }
    bool descriptorExists(const std::string& name) const {
      for (auto& l : descriptorLabels)
        if (l.text == name) return true;
      return false;
    }

    // host -> device (one H2D per call; the handle owns its device memory)
    struct Device {
      pgs_cloud* h = nullptr;
      ~Device() { pgs_cloud_destroy(h); }
    };
    std::unique_ptr<Device> upload() const {
      const int64_t n = features.cols();
      std::vector<float> f(static_cast<size_t>(n) * 4);
      for (int64_t i = 0; i < n; ++i)
        for (int r = 0; r < 4; ++r) f[i * 4 + r] = static_cast<float>(features(r, static_cast<int>(i)));
      std::unique_ptr<Device> d(new Device);
      check(pgs_cloud_create(context(), f.data(), n, 0, &d->h));
      int row = 0;
      for (auto& l : descriptorLabels) {
        std::vector<float> v(static_cast<size_t>(n) * l.span);
        for (int64_t i = 0; i < n; ++i)
          for (size_t r = 0; r < l.span; ++r)
            v[i * l.span + r] = static_cast<float>(descriptors(row + static_cast<int>(r), static_cast<int>(i)));
        check(pgs_cloud_set_descriptor(d->h, l.text.c_str(), static_cast<int>(l.span), v.data(), 0));
        row += static_cast<int>(l.span);
      }
      return d;
    }
    // device -> host
    void download(const pgs_cloud* h) {
      const int64_t n = pgs_cloud_num_points(h);
      std::vector<float> f(static_cast<size_t>(n) * 4);
      check(pgs_cloud_get_features(h, f.data(), 0));
      features.resize(4, static_cast<int>(n));
      for (int64_t i = 0; i < n; ++i)
        for (int r = 0; r < 4; ++r) features(r, static_cast<int>(i)) = static_cast<T>(f[i * 4 + r]);
      descriptorLabels.clear();
      int rows = 0;
      const int nd = pgs_cloud_num_descriptors(h);
      std::vector<std::pair<std::string, int>> info;
      for (int k = 0; k < nd; ++k) {
        char name[128];
        int span = 0;
        check(pgs_cloud_descriptor_info(h, k, name, sizeof(name), &span));
        info.emplace_back(name, span);
        rows += span;
      }
      descriptors.resize(rows, static_cast<int>(n));
      int row = 0;
      for (auto& it : info) {
        std::vector<float> v(static_cast<size_t>(n) * it.second);
        check(pgs_cloud_get_descriptor(h, it.first.c_str(), v.data(), 0));
        for (int64_t i = 0; i < n; ++i)
          for (int r = 0; r < it.second; ++r)
            descriptors(row + r, static_cast<int>(i)) = static_cast<T>(v[i * it.second + r]);
        descriptorLabels.emplace_back(it.first, static_cast<size_t>(it.second));
        row += it.second;
      }
    }
    // DP::concatenate (LocalMap.hpp:222)
    void concatenate(const DataPoints& other) {
      auto a = upload();
      auto b = other.upload();
      check(pgs_cloud_concatenate(a->h, b->h));
      download(a->h);
    }
  };

  struct Matches {
    Matrix dists;   // k x N, SQUARED distances
    IntMatrix ids;  // k x N
  };

  // ---- Transformation (Localizer.hpp:20,106; LocalMap.hpp:37,97,222) -----------------
  struct Transformation {
    virtual ~Transformation() {}
    virtual DataPoints compute(const DataPoints& input, const TransformationParameters& Tr) const = 0;
  };
  struct RigidTransformation : Transformation {
    DataPoints compute(const DataPoints& input, const TransformationParameters& Tr) const override {
      auto d = input.upload();
      double t[16];
      to_double16(Tr, t);
      check(pgs_rigid_transform(d->h, t));
      DataPoints out;
      out.download(d->h);
      return out;
    }
  };
  struct Transformations {
    void apply(DataPoints& cloud, const TransformationParameters& Tr) const {  // LoopCloser.hpp:352
      cloud = RigidTransformation().compute(cloud, Tr);
    }
  };

  // PM::get().REG(Transformation).create("RigidTransformation")
  struct TransformationRegistrar {
    std::shared_ptr<Transformation> create(const std::string& name) const {
      if (name != "RigidTransformation") throw InvalidElement("Trying to instanciate unknown element " + name);
      return std::make_shared<RigidTransformation>();
    }
  };
  struct Registry {
    TransformationRegistrar TransformationRegistrar_;
  };
  static const Registry& get() {
    static Registry r;
    return r;
  }
#ifndef REG
#define REG(name) name##Registrar_
#endif

  // ---- DataPointsFilters (types.h:27; Localizer.hpp:77,103) -----------------------------
  class DataPointsFilters {
   public:
    DataPointsFilters() : h_(nullptr), own_(false) {}
    explicit DataPointsFilters(std::istream& in) : h_(nullptr), own_(true) {
      const std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
      check(pgs_filters_create_from_yaml(context(), text.data(), text.size(), &h_));
    }
    DataPointsFilters(DataPointsFilters&& o) noexcept : h_(o.h_), own_(o.own_) { o.h_ = nullptr; }
    DataPointsFilters& operator=(DataPointsFilters&& o) noexcept {
      if (this != &o) { release(); h_ = o.h_; own_ = o.own_; o.h_ = nullptr; }
      return *this;
    }
    ~DataPointsFilters() { release(); }
    void init() {}
    void apply(DataPoints& cloud) {
      if (!h_) return;  // empty list
      auto d = cloud.upload();
      check(pgs_filters_apply(h_, d->h));
      cloud.download(d->h);
    }
    void borrow(pgs_filters* h) { release(); h_ = h; own_ = false; }

   private:
    void release() {
      if (own_ && h_) pgs_filters_destroy(h_);
      h_ = nullptr;
    }
    pgs_filters* h_;
    bool own_;
  };

  // ---- Matcher (Localizer.hpp:317,328; LoopCloser.hpp:356,358) -----------------------------
  struct Matcher {
    pgs_matcher* h = nullptr;  // borrowed from the owning ICP
    void init(const DataPoints& filteredReference) {
      auto d = filteredReference.upload();
      check(pgs_matcher_init(h, d->h));
    }
    Matches findClosests(const DataPoints& filteredReading) {
      auto d = filteredReading.upload();
      const int k = pgs_matcher_knn(h);
      const int n = static_cast<int>(filteredReading.features.cols());
      std::vector<int32_t> ids(static_cast<size_t>(k) * n);
      std::vector<float> d2(static_cast<size_t>(k) * n);
      check(pgs_matcher_find(h, d->h, ids.data(), d2.data(), 0));
      Matches m;
      m.ids.resize(k, n);
      m.dists.resize(k, n);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < k; ++j) {
          m.ids(j, i) = ids[static_cast<size_t>(i) * k + j];
          m.dists(j, i) = static_cast<T>(d2[static_cast<size_t>(i) * k + j]);
        }
      return m;
    }
  };

  static void flatten(const Matches& m, std::vector<int32_t>& ids, std::vector<float>& d2) {
    const int k = static_cast<int>(m.dists.rows()), n = static_cast<int>(m.dists.cols());
    ids.resize(static_cast<size_t>(k) * n);
    d2.resize(static_cast<size_t>(k) * n);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < k; ++j) {
        ids[static_cast<size_t>(i) * k + j] = m.ids(j, i);
        d2[static_cast<size_t>(i) * k + j] = static_cast<float>(m.dists(j, i));
      }
  }

  // ---- OutlierFilters (Localizer.hpp:330; LoopCloser.hpp:360) --------------------------------
  struct OutlierFilters {
    pgs_outliers* h = nullptr;
    OutlierWeights compute(const DataPoints& reading, const DataPoints& reference, const Matches& input) {
      auto dr = reading.upload();
      auto df = reference.upload();
      std::vector<int32_t> ids;
      std::vector<float> d2;
      flatten(input, ids, d2);
      const int k = static_cast<int>(input.dists.rows()), n = static_cast<int>(input.dists.cols());
      std::vector<float> w(ids.size());
      check(pgs_outliers_compute(h, dr->h, df->h, ids.data(), d2.data(), k, w.data(), 0));
      OutlierWeights out;
      out.resize(k, n);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < k; ++j) out(j, i) = static_cast<T>(w[static_cast<size_t>(i) * k + j]);
      return out;
    }
  };

  // ---- ErrorMinimizer (Localizer.hpp:238,278,332,347; LoopCloser.hpp:108,331,362) --------------
  struct ErrorMinimizer {
    pgs_minimizer* h = nullptr;
    pgs_min_result last;
    bool have_last = false;

    static pgs_min_result run(pgs_minimizer* h, const DataPoints& reading, const DataPoints& reference,
                              const OutlierWeights& weights, const Matches& matches) {
      auto dr = reading.upload();
      auto df = reference.upload();
      std::vector<int32_t> ids;
      std::vector<float> d2;
      flatten(matches, ids, d2);
      const int k = static_cast<int>(matches.dists.rows()), n = static_cast<int>(matches.dists.cols());
      std::vector<float> w(ids.size());
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < k; ++j) w[static_cast<size_t>(i) * k + j] = static_cast<float>(weights(j, i));
      pgs_min_result r;
      check(pgs_minimizer_compute(h, dr->h, df->h, ids.data(), d2.data(), w.data(), k, 0, &r));
      return r;
    }

    // ErrorElements(reading, reference, weights, matches) (Localizer.hpp:332,347): pgslam
    // reads only weightedPointUsedRatio, which is a count over weights and matches
    //  and needs no device work.
    struct ErrorElements {
      T pointUsedRatio;
      T weightedPointUsedRatio;
      ErrorElements(const DataPoints& reading, const DataPoints&, const OutlierWeights& weights,
                    const Matches& matches) {
        const int k = static_cast<int>(matches.dists.rows()), n = static_cast<int>(matches.dists.cols());
        long kept = 0;
        double wsum = 0.0;
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < k; ++j) {
            if (std::isinf(static_cast<double>(matches.dists(j, i)))) continue;
            if (weights(j, i) != T(0)) { ++kept; wsum += static_cast<double>(weights(j, i)); }
          }
        if (kept == 0) throw ConvergenceError("no point to minimize");
        (void)reading;
        pointUsedRatio = static_cast<T>(static_cast<double>(kept) / (static_cast<double>(k) * n));
        weightedPointUsedRatio = static_cast<T>(wsum / (static_cast<double>(k) * n));
      }
    };

    T getOverlap() const {  // Localizer.hpp:278, LoopCloser.hpp:331
      if (!have_last) throw ConvergenceError("getOverlap() before any ICP");
      return static_cast<T>(last.overlap);
    }
    Matrix getCovariance() const {  // Localizer.hpp:238, LoopCloser.hpp:108; order x,y,z,rx,ry,rz
      Matrix c = Matrix::Zero(6, 6);
      if (have_last)
        for (int cc = 0; cc < 6; ++cc)
          for (int r = 0; r < 6; ++r) c(r, cc) = static_cast<T>(last.covariance[cc * 6 + r]);
      return c;
    }
    T getResidualError(const DataPoints& reading, const DataPoints& reference, const OutlierWeights& weights,
                       const Matches& matches) const {  // LoopCloser.hpp:362
      return static_cast<T>(run(h, reading, reference, weights, matches).residual);
    }
  };

  // ---- ICP (types.h:24) -------------------------------------------------------------------------
  class ICP {
   public:
    // public members of ICPChainBase that pgslam reaches into
    DataPointsFilters readingDataPointsFilters, readingStepDataPointsFilters, referenceDataPointsFilters;
    Transformations transformations;
    std::shared_ptr<Matcher> matcher;
    OutlierFilters outlierFilters;
    std::shared_ptr<ErrorMinimizer> errorMinimizer;

    ICP() : matcher(std::make_shared<Matcher>()), errorMinimizer(std::make_shared<ErrorMinimizer>()), h_(nullptr) {
      setDefault();
    }
    ICP(const ICP&) = delete;
    ICP& operator=(const ICP&) = delete;
    virtual ~ICP() { pgs_icp_destroy(h_); }

    void setDefault() {
      pgs_icp* h = nullptr;
      check(pgs_icp_create_default(context(), &h));
      replace(h);
    }
    void loadFromYaml(std::istream& in) {  // Localizer.hpp:70,311; LoopCloser.hpp:73,348
      const std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
      pgs_icp* h = nullptr;
      check(pgs_icp_create_from_yaml(context(), text.data(), text.size(), &h));
      replace(h);
    }
    // ICP::operator()(reading, reference, T_init)  (LoopCloser.hpp:98)
    TransformationParameters operator()(const DataPoints& reading, const DataPoints& reference,
                                        const TransformationParameters& initial) {
      auto dr = reading.upload();
      auto df = reference.upload();
      double t[16];
      to_double16(initial, t);
      pgs_icp_result r;
      pgs_status st = pgs_icp_run(h_, dr->h, df->h, t, &r);
      return finish(st, r);
    }
    TransformationParameters operator()(const DataPoints& reading, const DataPoints& reference) {
      return (*this)(reading, reference, Matrix::Identity(4, 4));
    }
    // the accessor pgslam's author patched into libpointmatcher (LoopCloser.hpp:310-318)
    bool getMaxNumIterationsReached() const { return last_.max_iterations_reached != 0; }
    const pgs_icp_result& lastResult() const { return last_; }
    pgs_icp* handle() const { return h_; }

   protected:
    TransformationParameters finish(pgs_status st, const pgs_icp_result& r) {
      last_ = r;
      errorMinimizer->last.overlap = r.overlap;
      errorMinimizer->last.residual = r.residual;
      errorMinimizer->last.weighted_point_used_ratio = r.weighted_point_used_ratio;
      errorMinimizer->last.point_used_ratio = r.point_used_ratio;
      std::memcpy(errorMinimizer->last.covariance, r.covariance, sizeof(r.covariance));
      errorMinimizer->have_last = true;
      check(st);
      return from_double16(r.T);
    }
    void replace(pgs_icp* h) {
      pgs_icp_destroy(h_);
      h_ = h;
      std::memset(&last_, 0, sizeof(last_));
      readingDataPointsFilters.borrow(pgs_icp_reading_filters(h_));
      readingStepDataPointsFilters.borrow(pgs_icp_reading_step_filters(h_));
      referenceDataPointsFilters.borrow(pgs_icp_reference_filters(h_));
      matcher->h = pgs_icp_matcher(h_);
      outlierFilters.h = pgs_icp_outliers(h_);
      errorMinimizer->h = pgs_icp_minimizer(h_);
      errorMinimizer->have_last = false;
    }
    pgs_icp* h_;
    pgs_icp_result last_;
  };

  // ---- ICPSequence (types.h:25; Localizer.hpp:126,148,168,254) ---------------------------------
  class ICPSequence : public ICP {
   public:
    bool setMap(const DataPoints& map) {
      auto d = map.upload();
      check(pgs_icp_set_map(this->h_, d->h));
      return true;
    }
    bool hasMap() const { return pgs_icp_has_map(this->h_) != 0; }
    TransformationParameters operator()(const DataPoints& cloud, const TransformationParameters& initial) {
      auto d = cloud.upload();
      double t[16];
      PointMatcher::to_double16(initial, t);
      pgs_icp_result r;
      pgs_status st = pgs_icp_run_sequence(this->h_, d->h, t, &r);
      return this->finish(st, r);
    }
  };
};

}  // namespace pgslam_b200
