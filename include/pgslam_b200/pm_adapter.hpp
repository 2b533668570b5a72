// pm_adapter.hpp — header-only C++14 adapter that gives libpgslam_b200's C ABI
// the class shapes pgslam binds to at /root/reference/src/pgslam/types.h:19-27:
//
//     using PM  = PointMatcher<T>;          ->  pgslam_b200::PointMatcher<T>
//     using DP  = typename PM::DataPoints;
//     using ICP = typename PM::ICP;         using ICPSequence = typename PM::ICPSequence;
//     using TransformationPtr = std::shared_ptr<typename PM::Transformation>;
//     using DataPointsFilters = typename PM::DataPointsFilters;
//
// What is provided: the members pgslam touches (each citing its call site), the
// plugin surface north_star names - PM::get().REG(DataPointsFilter | Matcher | OutlierFilter |
// ErrorMinimizer | TransformationChecker | Inspector | Logger | Transformation).create(name,
// params) over polymorphic module classes with className / parameters / availableParameters() -
// and one extension: ICP::computeBatch for the loop-closure candidate loop
// (LoopCloser.hpp:266-297), on one or several GPUs of the process.
//
// T = float and T = double are both first class (tests/instantiation.cpp:6,10): poses,
// covariances and ratios cross the ABI as double and are returned in T without passing
// through float.  Clouds are stored on the device as fp32 whatever T is (the numeric contract
// of the path); a PointMatcher<double> cloud is rounded once on upload.
//
// Host <-> device traffic: a DataPoints keeps a handle to its device copy together with a
// 64-bit fingerprint of the host data it was made from.  An unchanged cloud is not uploaded
// again (and keeps the spatial index cached with the device copy), whatever object consumes it.
//
// Contexts: one per (host thread, device), created on first use and kept alive by every object
// created from it (shared_ptr), so an object may outlive or leave the thread that built it.
// PointMatcher<T>::setDevice(d) selects the device new objects of the calling thread use.
// With Eigen available the matrix types are Eigen's (bit-compatible with libpointmatcher);
// without it a minimal column-major stand-in keeps this header compilable.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <istream>
#include <iterator>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
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

namespace detail {

// ---- contexts --------------------------------------------------------------------------------
struct Ctx {
  pgs_ctx* h = nullptr;
  int device = 0;
  Ctx(int dev) : device(dev) {
    if (pgs_ctx_create(dev, nullptr, &h) != PGS_OK) throw std::runtime_error(pgs_last_error(nullptr));
  }
  ~Ctx() { pgs_ctx_destroy(h); }
  Ctx(const Ctx&) = delete;
  Ctx& operator=(const Ctx&) = delete;
};
using CtxPtr = std::shared_ptr<Ctx>;

inline int& thread_device() {
  static thread_local int d = 0;
  return d;
}
// the calling thread's context on `device` (-1: the thread's selected device).  pgslam-MT drives
// the localizer and the loop closer from two threads (LocalizerMT.hpp:47, LoopCloserMT.hpp:41):
// each gets its own stream.  Objects hold the shared_ptr, so the context outlives the thread.
inline CtxPtr context(int device = -1) {
  static thread_local std::map<int, CtxPtr> per_device;
  if (device < 0) device = thread_device();
  CtxPtr& c = per_device[device];
  if (!c) c = std::make_shared<Ctx>(device);
  return c;
}

// 64-bit fingerprint of host data (multiply-xorshift over 8-byte words)
inline uint64_t mix64(uint64_t h, uint64_t v) {
  h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
  h *= 0xff51afd7ed558ccdull;
  return h ^ (h >> 32);
}
inline uint64_t hash_bytes(const void* p, size_t n, uint64_t h) {
  const unsigned char* b = static_cast<const unsigned char*>(p);
  size_t i = 0;
  uint64_t a = h, c = ~h;  // two lanes keep the multiply latency off the critical path
  for (; i + 16 <= n; i += 16) {
    uint64_t w0, w1;
    std::memcpy(&w0, b + i, 8);
    std::memcpy(&w1, b + i + 8, 8);
    a = (a ^ w0) * 0x9fb21c651e98df25ull;
    a ^= a >> 29;
    c = (c ^ w1) * 0xc2b2ae3d27d4eb4full;
    c ^= c >> 31;
  }
  uint64_t tail = 0;
  if (i < n) std::memcpy(&tail, b + i, n - i < 8 ? n - i : 8);
  return mix64(mix64(a, c), tail ^ static_cast<uint64_t>(n));
}

struct DeviceCloud {
  CtxPtr ctx;
  pgs_cloud* h = nullptr;
  uint64_t fingerprint = 0;
  ~DeviceCloud() { pgs_cloud_destroy(h); }
};

using Parameters = std::map<std::string, std::string>;
struct KV {
  std::vector<const char*> ptrs;
  explicit KV(const Parameters& p) {
    for (auto& kv : p) { ptrs.push_back(kv.first.c_str()); ptrs.push_back(kv.second.c_str()); }
  }
  const char* const* data() const { return ptrs.empty() ? nullptr : ptrs.data(); }
  int count() const { return static_cast<int>(ptrs.size() / 2); }
};
enum Kind { K_DataPointsFilter = 0, K_Matcher, K_OutlierFilter, K_ErrorMinimizer, K_TransformationChecker, K_Inspector,
            K_Logger, K_Transformation };

}  // namespace detail

template <typename T>
struct PointMatcher {
  static_assert(sizeof(T) == 4 || sizeof(T) == 8, "T must be float or double");
#ifdef PGSLAM_B200_HAVE_EIGEN
  using Matrix = Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic>;
  using IntMatrix = Eigen::Matrix<int, Eigen::Dynamic, Eigen::Dynamic>;
  using Int64Matrix = Eigen::Matrix<std::int64_t, Eigen::Dynamic, Eigen::Dynamic>;
#else
  using Matrix = MiniMatrix<T>;
  using IntMatrix = MiniMatrix<int>;
  using Int64Matrix = MiniMatrix<std::int64_t>;
#endif
  using ScalarType = T;
  using TransformationParameters = Matrix;
  using OutlierWeights = Matrix;
  using Parameters = detail::Parameters;
  using CtxPtr = detail::CtxPtr;

  // ---- exceptions, as libpointmatcher names them -------------------------------
  struct ConvergenceError : std::runtime_error { using std::runtime_error::runtime_error; };
  struct TransformationError : std::runtime_error { using std::runtime_error::runtime_error; };
  struct InvalidParameter : std::runtime_error { using std::runtime_error::runtime_error; };
  struct InvalidField : std::runtime_error { using std::runtime_error::runtime_error; };
  struct InvalidModuleType : std::runtime_error { using std::runtime_error::runtime_error; };
  struct InvalidElement : std::runtime_error { using std::runtime_error::runtime_error; };

  // device used by objects the calling thread creates from now on (default 0)
  static void setDevice(int device) { detail::thread_device() = device; }
  static int getDevice() { return detail::thread_device(); }
  static CtxPtr context(int device = -1) { return detail::context(device); }

  static void raise(pgs_status st, const std::string& msg) {
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
  // the error text lives in the context the failing handle belongs to
  static void check(pgs_status st, const CtxPtr& ctx) {
    if (st != PGS_OK) raise(st, pgs_last_error(ctx ? ctx->h : nullptr));
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
      bool operator==(const Label& o) const { return text == o.text && span == o.span; }
    };
    using Labels = std::vector<Label>;
    Matrix features;      // 4 x N, last row 1
    Matrix descriptors;   // D x N
    Int64Matrix times;    // Tm x N (host side only: carried through device filters by point index)
    Labels featureLabels, descriptorLabels, timeLabels;

    DataPoints() {}
    unsigned getNbPoints() const { return static_cast<unsigned>(features.cols()); }
    bool descriptorExists(const std::string& name) const { return descriptorRow(name) >= 0; }
    int descriptorRow(const std::string& name, int* span = nullptr) const {
      int row = 0;
      for (auto& l : descriptorLabels) {
        if (l.text == name) {
          if (span) *span = static_cast<int>(l.span);
          return row;
        }
        row += static_cast<int>(l.span);
      }
      return -1;
    }
    unsigned getDescriptorDimension(const std::string& name) const {
      int span = 0;
      return descriptorRow(name, &span) < 0 ? 0u : static_cast<unsigned>(span);
    }
    Matrix getDescriptorCopyByName(const std::string& name) const {
      int span = 0;
      const int row = descriptorRow(name, &span);
      if (row < 0) throw InvalidField("Cannot find descriptor " + name);
      Matrix m(span, static_cast<int>(descriptors.cols()));
      for (int c = 0; c < static_cast<int>(descriptors.cols()); ++c)
        for (int r = 0; r < span; ++r) m(r, c) = descriptors(row + r, c);
      return m;
    }
    void addDescriptor(const std::string& name, const Matrix& d) {
      if (d.cols() != features.cols()) throw InvalidField("addDescriptor: wrong number of columns for " + name);
      if (descriptorExists(name)) removeDescriptor(name);
      const int n = static_cast<int>(features.cols()), old = static_cast<int>(descriptors.rows() * (descriptors.cols() ? 1 : 0));
      Matrix nd(old + static_cast<int>(d.rows()), n);
      for (int c = 0; c < n; ++c) {
        for (int r = 0; r < old; ++r) nd(r, c) = descriptors(r, c);
        for (int r = 0; r < static_cast<int>(d.rows()); ++r) nd(old + r, c) = d(r, c);
      }
      descriptors = nd;
      descriptorLabels.emplace_back(name, static_cast<size_t>(d.rows()));
    }
    void removeDescriptor(const std::string& name) {
      int span = 0;
      const int row = descriptorRow(name, &span);
      if (row < 0) return;
      const int n = static_cast<int>(descriptors.cols()), rows = static_cast<int>(descriptors.rows());
      Matrix nd(rows - span, n);
      for (int c = 0; c < n; ++c)
        for (int r = 0, o = 0; r < rows; ++r)
          if (r < row || r >= row + span) nd(o++, c) = descriptors(r, c);
      descriptors = nd;
      for (size_t i = 0; i < descriptorLabels.size(); ++i)
        if (descriptorLabels[i].text == name) { descriptorLabels.erase(descriptorLabels.begin() + static_cast<long>(i)); break; }
    }

    // DP::concatenate (LocalMap.hpp:222): columns appended on the HOST; descriptors (and times)
    // present in both clouds with equal span are kept, the others dropped, as upstream does
    void concatenate(const DataPoints& o) {
      const int na = static_cast<int>(features.cols()), nb = static_cast<int>(o.features.cols());
      const int fr = na ? static_cast<int>(features.rows()) : static_cast<int>(o.features.rows());
      if (na && nb && features.rows() != o.features.rows()) throw InvalidField("concatenate: feature dimensions differ");
      Matrix nf(fr, na + nb);
      for (int c = 0; c < na; ++c)
        for (int r = 0; r < fr; ++r) nf(r, c) = features(r, c);
      for (int c = 0; c < nb; ++c)
        for (int r = 0; r < fr; ++r) nf(r, na + c) = o.features(r, c);
      Labels kept;
      int rows = 0;
      for (auto& l : descriptorLabels) {
        int span = 0;
        if (o.descriptorRow(l.text, &span) >= 0 && span == static_cast<int>(l.span)) { kept.push_back(l); rows += span; }
      }
      Matrix nd(rows, na + nb);
      int out = 0;
      for (auto& l : kept) {
        const int ra = descriptorRow(l.text), rb = o.descriptorRow(l.text);
        for (int r = 0; r < static_cast<int>(l.span); ++r, ++out) {
          for (int c = 0; c < na; ++c) nd(out, c) = descriptors(ra + r, c);
          for (int c = 0; c < nb; ++c) nd(out, na + c) = o.descriptors(rb + r, c);
        }
      }
      if (times.rows() > 0 && times.rows() == o.times.rows() && timeLabels == o.timeLabels) {
        Int64Matrix nt(static_cast<int>(times.rows()), na + nb);
        for (int r = 0; r < static_cast<int>(times.rows()); ++r) {
          for (int c = 0; c < na; ++c) nt(r, c) = times(r, c);
          for (int c = 0; c < nb; ++c) nt(r, na + c) = o.times(r, c);
        }
        times = nt;
      } else {
        times = Int64Matrix();
        timeLabels.clear();
      }
      if (!na) featureLabels = o.featureLabels;
      features = nf;
      descriptors = nd;
      descriptorLabels = kept;
    }

    // ---- device copy ------------------------------------------------------------
    using DevicePtr = std::shared_ptr<detail::DeviceCloud>;
    uint64_t fingerprint() const {
      uint64_t h = detail::mix64(0x5bd1e995u, (static_cast<uint64_t>(features.rows()) << 32) | static_cast<uint64_t>(features.cols()));
      h = detail::hash_bytes(features.data(), sizeof(T) * static_cast<size_t>(features.rows()) * features.cols(), h);
      h = detail::mix64(h, static_cast<uint64_t>(descriptors.rows()));
      h = detail::hash_bytes(descriptors.data(), sizeof(T) * static_cast<size_t>(descriptors.rows()) * descriptors.cols(), h);
      for (auto& l : descriptorLabels) h = detail::mix64(detail::hash_bytes(l.text.data(), l.text.size(), h), l.span);
      return detail::mix64(h, times.rows() > 0 ? 1u : 0u);
    }
    static const char* indexLabel() { return "__pgs_point_index"; }

    // the device copy of this cloud on ctx's device: re-used while the host data is unchanged
    DevicePtr device(const CtxPtr& ctx) const {
      const uint64_t fp = fingerprint();
      if (dev_ && dev_->fingerprint == fp && dev_->ctx->device == ctx->device) return dev_;
      const int64_t n = features.cols();
      auto d = std::make_shared<detail::DeviceCloud>();
      d->ctx = ctx;
      d->fingerprint = fp;
      if (n > 0 && features.rows() != 4) throw InvalidField("features must be 4 x N (3-D homogeneous points)");
      if (sizeof(T) == 4) {  // PM's column-major 4 x N float block IS the ABI layout
        check(pgs_cloud_create(ctx->h, reinterpret_cast<const float*>(features.data()), n, 0, &d->h), ctx);
      } else {
        std::vector<float> f(static_cast<size_t>(n) * 4);
        const T* src = features.data();
        for (size_t i = 0; i < f.size(); ++i) f[i] = static_cast<float>(src[i]);
        check(pgs_cloud_create(ctx->h, f.data(), n, 0, &d->h), ctx);
      }
      int row = 0;
      std::vector<float> v;
      for (auto& l : descriptorLabels) {
        v.resize(static_cast<size_t>(n) * l.span);
        for (int64_t i = 0; i < n; ++i)
          for (size_t r = 0; r < l.span; ++r)
            v[static_cast<size_t>(i) * l.span + r] = static_cast<float>(descriptors(row + static_cast<int>(r), static_cast<int>(i)));
        check(pgs_cloud_set_descriptor(d->h, l.text.c_str(), static_cast<int>(l.span), v.data(), 0), ctx);
        row += static_cast<int>(l.span);
      }
      if (times.rows() > 0) {  // times follow the points through device filters by index
        if (n > (1 << 24)) throw InvalidField("times: more than 2^24 points");
        v.resize(static_cast<size_t>(n));
        for (int64_t i = 0; i < n; ++i) v[static_cast<size_t>(i)] = static_cast<float>(i);
        check(pgs_cloud_set_descriptor(d->h, indexLabel(), 1, v.data(), 0), ctx);
      }
      dev_ = d;
      return d;
    }
    // a device copy this object may MODIFY: never one shared with another DataPoints
    DevicePtr deviceForWrite(const CtxPtr& ctx) const {
      DevicePtr d = device(ctx);
      if (d.use_count() <= 2) return d;  // dev_ + d
      auto c = std::make_shared<detail::DeviceCloud>();
      c->ctx = d->ctx;
      c->fingerprint = d->fingerprint;
      check(pgs_cloud_copy(d->h, &c->h), d->ctx);
      dev_ = c;
      return c;
    }
    // device -> host: this object becomes the cloud held by `d` (and keeps it as its device copy)
    void adopt(const DevicePtr& d) {
      const CtxPtr& ctx = d->ctx;
      const int64_t n = pgs_cloud_num_points(d->h);
      std::vector<float> f(static_cast<size_t>(n) * 4);
      check(pgs_cloud_get_features(d->h, f.data(), 0), ctx);
      features.resize(4, static_cast<int>(n));
      {
        T* dst = features.data();
        for (size_t i = 0; i < f.size(); ++i) dst[i] = static_cast<T>(f[i]);
      }
      Labels labels;
      int rows = 0;
      bool has_index = false;
      const int nd = pgs_cloud_num_descriptors(d->h);
      for (int k = 0; k < nd; ++k) {
        char name[128];
        int span = 0;
        check(pgs_cloud_descriptor_info(d->h, k, name, sizeof(name), &span), ctx);
        if (std::string(name) == indexLabel()) { has_index = true; continue; }
        labels.emplace_back(name, static_cast<size_t>(span));
        rows += span;
      }
      descriptors.resize(rows, static_cast<int>(n));
      int row = 0;
      for (auto& l : labels) {
        f.resize(static_cast<size_t>(n) * l.span);
        check(pgs_cloud_get_descriptor(d->h, l.text.c_str(), f.data(), 0), ctx);
        for (int64_t i = 0; i < n; ++i)
          for (size_t r = 0; r < l.span; ++r)
            descriptors(row + static_cast<int>(r), static_cast<int>(i)) = static_cast<T>(f[static_cast<size_t>(i) * l.span + r]);
        row += static_cast<int>(l.span);
      }
      descriptorLabels = labels;
      if (times.rows() > 0) {
        if (has_index) {
          f.resize(static_cast<size_t>(n));
          check(pgs_cloud_get_descriptor(d->h, indexLabel(), f.data(), 0), ctx);
          Int64Matrix nt(static_cast<int>(times.rows()), static_cast<int>(n));
          for (int64_t i = 0; i < n; ++i) {
            int src = static_cast<int>(f[static_cast<size_t>(i)]);  // averaged indices (voxel centroids) round down
            if (src < 0) src = 0;
            if (src >= static_cast<int>(times.cols())) src = static_cast<int>(times.cols()) - 1;
            for (int r = 0; r < static_cast<int>(times.rows()); ++r) nt(r, static_cast<int>(i)) = times(r, src);
          }
          times = nt;
        } else if (times.cols() != n) {
          times = Int64Matrix();
          timeLabels.clear();
        }
      }
      d->fingerprint = fingerprint();
      dev_ = d;
    }
    void dropDeviceCopy() const { dev_.reset(); }

    // DataPoints::load / save: libpointmatcher's .csv, legacy ASCII .vtk and .ply files (pgs_cloud_load / _save)
    static DataPoints load(const std::string& fileName) {
      CtxPtr ctx = context();
      auto d = std::make_shared<detail::DeviceCloud>();
      d->ctx = ctx;
      check(pgs_cloud_load(ctx->h, fileName.c_str(), &d->h), ctx);
      DataPoints out;
      out.adopt(d);
      return out;
    }
    void save(const std::string& fileName) const {
      if (times.rows() > 0) {  // the file formats carry no times (and must not carry the index column)
        DataPoints plain = *this;
        plain.times = Int64Matrix();
        plain.timeLabels.clear();
        plain.dropDeviceCopy();
        plain.save(fileName);
        return;
      }
      CtxPtr ctx = context();
      auto d = device(ctx);
      check(pgs_cloud_save(d->h, fileName.c_str()), d->ctx);
    }

   private:
    mutable DevicePtr dev_;
  };

  struct Matches {
    Matrix dists;   // k x N, SQUARED distances
    IntMatrix ids;  // k x N
  };

  // ---- Parametrizable / Registrar (PointMatcherSupport) -----------------------------------------
  struct ParameterDoc {
    std::string name, doc, defaultValue, minValue, maxValue;
    char type;
  };
  using ParametersDoc = std::vector<ParameterDoc>;

  static ParametersDoc availableParametersOf(int kind, const std::string& name) {
    const int n = pgs_registrar_param_count(kind, name.c_str());
    if (n < 0) throw InvalidElement("Trying to instanciate unknown element " + name);
    ParametersDoc out;
    for (int i = 0; i < n; ++i) {
      const char *k = "", *doc = "", *def = "", *mn = "", *mx = "";
      char type = 's';
      pgs_registrar_param(kind, name.c_str(), i, &k, &doc, &def, &mn, &mx, &type);
      out.push_back(ParameterDoc{k, doc, def, mn, mx, type});
    }
    return out;
  }

  struct Parametrizable {
    std::string className;
    Parameters parameters;  // fully defaulted
    int kind = 0;
    virtual ~Parametrizable() {}
    ParametersDoc availableParameters() const { return availableParametersOf(kind, className); }
    std::string getParamValueString(const std::string& name) const {
      auto it = parameters.find(name);
      if (it == parameters.end()) throw InvalidParameter("Parameter " + name + " does not exist in class " + className);
      return it->second;
    }
    template <typename S>
    S get(const std::string& name) const {
      std::istringstream ss(getParamValueString(name));
      S v;
      ss >> v;
      return v;
    }

   protected:
    // validate (unknown module / unknown or out-of-range parameter) and fill in the defaults
    void configure(int k, const std::string& name, const Parameters& params) {
      kind = k;
      className = name;
      detail::KV kv(params);
      char err[512];
      const pgs_status st = pgs_module_validate(k, name.c_str(), kv.data(), kv.count(), err, sizeof(err));
      if (st != PGS_OK) raise(st, err);
      for (auto& d : availableParametersOf(k, name)) parameters[d.name] = d.defaultValue;
      for (auto& p : params) parameters[p.first] = p.second;
    }
  };

  // ---- Transformation (Localizer.hpp:20,106; LocalMap.hpp:37,97,222) -----------------
  struct Transformation : Parametrizable {
    virtual DataPoints compute(const DataPoints& input, const TransformationParameters& Tr) const = 0;
    virtual bool checkParameters(const TransformationParameters&) const { return true; }
  };
  struct RigidTransformation : Transformation {
    RigidTransformation() { this->configure(detail::K_Transformation, "RigidTransformation", Parameters()); }
    DataPoints compute(const DataPoints& input, const TransformationParameters& Tr) const override {
      CtxPtr ctx = context();
      auto src = input.device(ctx);
      auto d = std::make_shared<detail::DeviceCloud>();
      d->ctx = src->ctx;
      check(pgs_cloud_copy(src->h, &d->h), src->ctx);
      double t[16];
      to_double16(Tr, t);
      check(pgs_rigid_transform(d->h, t), d->ctx);
      DataPoints out = input;  // labels, times
      out.dropDeviceCopy();
      out.adopt(d);
      return out;
    }
    bool checkParameters(const TransformationParameters& Tr) const override {
      double det = 0;
      det = static_cast<double>(Tr(0, 0)) * (static_cast<double>(Tr(1, 1)) * Tr(2, 2) - static_cast<double>(Tr(1, 2)) * Tr(2, 1)) -
            static_cast<double>(Tr(0, 1)) * (static_cast<double>(Tr(1, 0)) * Tr(2, 2) - static_cast<double>(Tr(1, 2)) * Tr(2, 0)) +
            static_cast<double>(Tr(0, 2)) * (static_cast<double>(Tr(1, 0)) * Tr(2, 1) - static_cast<double>(Tr(1, 1)) * Tr(2, 0));
      return std::fabs(1.0 - det) <= 0.001;
    }
  };
  struct Transformations {
    void apply(DataPoints& cloud, const TransformationParameters& Tr) const {  // LoopCloser.hpp:352
      cloud = RigidTransformation().compute(cloud, Tr);
    }
  };

  // ---- DataPointsFilter / DataPointsFilters (types.h:27; Localizer.hpp:77,103) --------------------
  struct DataPointsFilter : Parametrizable {
    DataPointsFilter(const std::string& name, const Parameters& params) { this->configure(detail::K_DataPointsFilter, name, params); }
    virtual void init() {}
    virtual void inPlaceFilter(DataPoints& cloud) {
      CtxPtr ctx = context();
      pgs_filters* f = nullptr;
      check(pgs_filters_create(ctx->h, &f), ctx);
      detail::KV kv(this->parameters);
      pgs_status st = pgs_filters_append(f, this->className.c_str(), kv.data(), kv.count());
      if (st == PGS_OK) {
        auto d = cloud.deviceForWrite(ctx);
        st = pgs_filters_apply(f, d->h);
        if (st == PGS_OK) cloud.adopt(d);
        else cloud.dropDeviceCopy();
      }
      const std::string msg = st == PGS_OK ? "" : pgs_last_error(ctx->h);
      pgs_filters_destroy(f);
      if (st != PGS_OK) raise(st, msg);
    }
    virtual DataPoints filter(const DataPoints& input) {
      DataPoints out = input;
      inPlaceFilter(out);
      return out;
    }
  };

  class DataPointsFilters {
   public:
    DataPointsFilters() : h_(nullptr), own_(false) {}
    explicit DataPointsFilters(std::istream& in) : ctx_(context()), h_(nullptr), own_(true) {
      const std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
      check(pgs_filters_create_from_yaml(ctx_->h, text.data(), text.size(), &h_), ctx_);
    }
    DataPointsFilters(DataPointsFilters&& o) noexcept : ctx_(std::move(o.ctx_)), h_(o.h_), own_(o.own_) { o.h_ = nullptr; }
    DataPointsFilters& operator=(DataPointsFilters&& o) noexcept {
      if (this != &o) { release(); ctx_ = std::move(o.ctx_); h_ = o.h_; own_ = o.own_; o.h_ = nullptr; }
      return *this;
    }
    ~DataPointsFilters() { release(); }
    void init() {}
    size_t size() const { return h_ ? static_cast<size_t>(pgs_filters_count(h_)) : 0; }
    // DataPointsFilters is a vector of filters upstream: push_back appends to the chain
    void push_back(const std::shared_ptr<DataPointsFilter>& f) {
      if (!h_) {
        ctx_ = context();
        check(pgs_filters_create(ctx_->h, &h_), ctx_);
        own_ = true;
      }
      detail::KV kv(f->parameters);
      check(pgs_filters_append(h_, f->className.c_str(), kv.data(), kv.count()), ctx_);
    }
    void apply(DataPoints& cloud) {
      if (!h_ || pgs_filters_count(h_) == 0) return;  // empty list
      auto d = cloud.deviceForWrite(ctx_);
      const pgs_status st = pgs_filters_apply(h_, d->h);
      if (st != PGS_OK) cloud.dropDeviceCopy();
      check(st, ctx_);
      cloud.adopt(d);
    }
    void borrow(const CtxPtr& ctx, pgs_filters* h) { release(); ctx_ = ctx; h_ = h; own_ = false; }

   private:
    void release() {
      if (own_ && h_) pgs_filters_destroy(h_);
      h_ = nullptr;
    }
    CtxPtr ctx_;
    pgs_filters* h_;
    bool own_;
  };

  // ---- Matcher (Localizer.hpp:317,328; LoopCloser.hpp:356,358) -----------------------------
  struct Matcher : Parametrizable {
    Matcher() {}  // borrowed from an ICP (handle set by the owner)
    Matcher(const std::string& name, const Parameters& params) : ctx_(context()), own_(true) {
      this->configure(detail::K_Matcher, name, params);
      detail::KV kv(params);
      check(pgs_matcher_create(ctx_->h, name.c_str(), kv.data(), kv.count(), &h), ctx_);
    }
    Matcher(const Matcher&) = delete;
    Matcher& operator=(const Matcher&) = delete;
    ~Matcher() override {
      if (own_ && h) pgs_matcher_destroy(h);
    }
    virtual void init(const DataPoints& filteredReference) {
      ref_ = filteredReference.device(ctx_);  // the index reads this copy: keep it alive
      check(pgs_matcher_init(h, ref_->h), ctx_);
    }
    virtual Matches findClosests(const DataPoints& filteredReading) {
      auto d = filteredReading.device(ctx_);
      const int k = pgs_matcher_knn(h);
      const int n = static_cast<int>(filteredReading.features.cols());
      std::vector<int32_t> ids(static_cast<size_t>(k) * n);
      std::vector<float> d2(static_cast<size_t>(k) * n);
      check(pgs_matcher_find(h, d->h, ids.data(), d2.data(), 0), ctx_);
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
    void borrow(const CtxPtr& ctx, pgs_matcher* handle) { ctx_ = ctx; h = handle; own_ = false; }
    pgs_matcher* h = nullptr;

   private:
    CtxPtr ctx_;
    bool own_ = false;
    typename DataPoints::DevicePtr ref_;
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

  // ---- OutlierFilter / OutlierFilters (Localizer.hpp:330; LoopCloser.hpp:360) ----------------------
  static OutlierWeights outlier_compute(const CtxPtr& ctx, pgs_outliers* h, const DataPoints& reading,
                                        const DataPoints& reference, const Matches& input) {
    auto dr = reading.device(ctx);
    auto df = reference.device(ctx);
    std::vector<int32_t> ids;
    std::vector<float> d2;
    flatten(input, ids, d2);
    const int k = static_cast<int>(input.dists.rows()), n = static_cast<int>(input.dists.cols());
    std::vector<float> w(ids.size());
    check(pgs_outliers_compute(h, dr->h, df->h, ids.data(), d2.data(), k, w.data(), 0), ctx);
    OutlierWeights out;
    out.resize(k, n);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < k; ++j) out(j, i) = static_cast<T>(w[static_cast<size_t>(i) * k + j]);
    return out;
  }
  struct OutlierFilter : Parametrizable {
    OutlierFilter(const std::string& name, const Parameters& params) { this->configure(detail::K_OutlierFilter, name, params); }
    virtual OutlierWeights compute(const DataPoints& reading, const DataPoints& reference, const Matches& input) {
      CtxPtr ctx = context();
      pgs_outliers* o = nullptr;
      check(pgs_outliers_create(ctx->h, &o), ctx);
      detail::KV kv(this->parameters);
      pgs_status st = pgs_outliers_append(o, this->className.c_str(), kv.data(), kv.count());
      OutlierWeights w;
      std::string msg;
      if (st == PGS_OK) {
        try {
          w = outlier_compute(ctx, o, reading, reference, input);
        } catch (...) {
          pgs_outliers_destroy(o);
          throw;
        }
      } else {
        msg = pgs_last_error(ctx->h);
      }
      pgs_outliers_destroy(o);
      if (st != PGS_OK) raise(st, msg);
      return w;
    }
  };
  class OutlierFilters {
   public:
    OutlierFilters() {}
    OutlierFilters(OutlierFilters&& o) noexcept : ctx_(std::move(o.ctx_)), h(o.h), own_(o.own_) { o.h = nullptr; }
    OutlierFilters& operator=(OutlierFilters&& o) noexcept {
      if (this != &o) { release(); ctx_ = std::move(o.ctx_); h = o.h; own_ = o.own_; o.h = nullptr; }
      return *this;
    }
    ~OutlierFilters() { release(); }
    void push_back(const std::shared_ptr<OutlierFilter>& f) {
      if (!h) {
        ctx_ = context();
        check(pgs_outliers_create(ctx_->h, &h), ctx_);
        own_ = true;
      }
      detail::KV kv(f->parameters);
      check(pgs_outliers_append(h, f->className.c_str(), kv.data(), kv.count()), ctx_);
    }
    OutlierWeights compute(const DataPoints& reading, const DataPoints& reference, const Matches& input) {
      if (!h) {  // an empty list: weight 1 for every finite match (A.3)
        OutlierWeights w;
        w.resize(static_cast<int>(input.dists.rows()), static_cast<int>(input.dists.cols()));
        for (int c = 0; c < static_cast<int>(input.dists.cols()); ++c)
          for (int r = 0; r < static_cast<int>(input.dists.rows()); ++r)
            w(r, c) = std::isinf(static_cast<double>(input.dists(r, c))) ? T(0) : T(1);
        return w;
      }
      return outlier_compute(ctx_, h, reading, reference, input);
    }
    void borrow(const CtxPtr& ctx, pgs_outliers* handle) { release(); ctx_ = ctx; h = handle; own_ = false; }

   private:
    void release() {
      if (own_ && h) pgs_outliers_destroy(h);
      h = nullptr;
    }
    CtxPtr ctx_;
    pgs_outliers* h = nullptr;
    bool own_ = false;
  };

  // ---- ErrorMinimizer (Localizer.hpp:238,278,332,347; LoopCloser.hpp:108,331,362) --------------
  struct ErrorMinimizer : Parametrizable {
    ErrorMinimizer() { std::memset(&last, 0, sizeof(last)); }
    ErrorMinimizer(const std::string& name, const Parameters& params) : ctx_(context()), own_(true) {
      std::memset(&last, 0, sizeof(last));
      this->configure(detail::K_ErrorMinimizer, name, params);
      detail::KV kv(params);
      check(pgs_minimizer_create(ctx_->h, name.c_str(), kv.data(), kv.count(), &h), ctx_);
    }
    ErrorMinimizer(const ErrorMinimizer&) = delete;
    ErrorMinimizer& operator=(const ErrorMinimizer&) = delete;
    ~ErrorMinimizer() override {
      if (own_ && h) pgs_minimizer_destroy(h);
    }
    void borrow(const CtxPtr& ctx, pgs_minimizer* handle) { ctx_ = ctx; h = handle; own_ = false; }

    pgs_min_result run(const DataPoints& reading, const DataPoints& reference, const OutlierWeights& weights,
                       const Matches& matches) const {
      auto dr = reading.device(ctx_);
      auto df = reference.device(ctx_);
      std::vector<int32_t> ids;
      std::vector<float> d2;
      flatten(matches, ids, d2);
      const int k = static_cast<int>(matches.dists.rows()), n = static_cast<int>(matches.dists.cols());
      std::vector<float> w(ids.size());
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < k; ++j) w[static_cast<size_t>(i) * k + j] = static_cast<float>(weights(j, i));
      pgs_min_result r;
      check(pgs_minimizer_compute(h, dr->h, df->h, ids.data(), d2.data(), w.data(), k, 0, &r), ctx_);
      return r;
    }
    // ErrorMinimizer::compute(filteredReading, filteredReference, outlierWeights, matches)
    virtual TransformationParameters compute(const DataPoints& reading, const DataPoints& reference,
                                             const OutlierWeights& weights, const Matches& matches) {
      last = run(reading, reference, weights, matches);
      have_last = true;
      return from_double16(last.T);
    }

    // ErrorElements(reading, reference, weights, matches) (Localizer.hpp:332,347): pgslam
    // reads only weightedPointUsedRatio, which is a count over weights and matches
    //  and needs no device work.
    struct ErrorElements {
      T pointUsedRatio;
      T weightedPointUsedRatio;
      int nbRejectedMatches;
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
        nbRejectedMatches = static_cast<int>(static_cast<long>(k) * n - kept);
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
      return static_cast<T>(run(reading, reference, weights, matches).residual);
    }
    T getPointUsedRatio() const { return static_cast<T>(last.point_used_ratio); }
    T getWeightedPointUsedRatio() const { return static_cast<T>(last.weighted_point_used_ratio); }

    pgs_minimizer* h = nullptr;
    pgs_min_result last;
    bool have_last = false;

   private:
    CtxPtr ctx_;
    bool own_ = false;
  };

  // modules with no device object: validated names and parameters
  struct TransformationChecker : Parametrizable {
    TransformationChecker(const std::string& n, const Parameters& p) { this->configure(detail::K_TransformationChecker, n, p); }
  };
  struct Inspector : Parametrizable {
    Inspector(const std::string& n, const Parameters& p) { this->configure(detail::K_Inspector, n, p); }
  };
  struct Logger : Parametrizable {
    Logger(const std::string& n, const Parameters& p) { this->configure(detail::K_Logger, n, p); }
  };

  // ---- PM::get().REG(kind) ----------------------------------------------------------------------
  template <typename Interface, int kKind>
  struct Registrar {
    std::shared_ptr<Interface> create(const std::string& name, const Parameters& params = Parameters()) const {
      return std::make_shared<Interface>(name, params);
    }
    ParametersDoc availableParameters(const std::string& name) const { return availableParametersOf(kKind, name); }
    std::vector<std::string> names() const {
      std::vector<std::string> out;
      for (int i = 0; i < pgs_registrar_count(kKind); ++i) out.push_back(pgs_registrar_name(kKind, i));
      return out;
    }
    // "name: parameter (default, min..max) - doc" lines, one per parameter
    std::string description(const std::string& name) const {
      std::ostringstream ss;
      ss << name << "\n";
      for (auto& d : availableParametersOf(kKind, name))
        ss << "  " << d.name << " (default: " << d.defaultValue << ", min: " << d.minValue << ", max: " << d.maxValue << ") - "
           << d.doc << "\n";
      return ss.str();
    }
  };
  struct TransformationRegistrar {
    std::shared_ptr<Transformation> create(const std::string& name, const Parameters& params = Parameters()) const {
      if (name != "RigidTransformation") throw InvalidElement("Trying to instanciate unknown element " + name);
      if (!params.empty()) throw InvalidParameter("RigidTransformation takes no parameter");
      return std::make_shared<RigidTransformation>();
    }
    std::vector<std::string> names() const { return {"RigidTransformation"}; }
  };
  struct Registry {
    Registrar<DataPointsFilter, detail::K_DataPointsFilter> DataPointsFilterRegistrar_;
    Registrar<Matcher, detail::K_Matcher> MatcherRegistrar_;
    Registrar<OutlierFilter, detail::K_OutlierFilter> OutlierFilterRegistrar_;
    Registrar<ErrorMinimizer, detail::K_ErrorMinimizer> ErrorMinimizerRegistrar_;
    Registrar<TransformationChecker, detail::K_TransformationChecker> TransformationCheckerRegistrar_;
    Registrar<Inspector, detail::K_Inspector> InspectorRegistrar_;
    Registrar<Logger, detail::K_Logger> LoggerRegistrar_;
    TransformationRegistrar TransformationRegistrar_;
  };
  static const Registry& get() {
    static Registry r;
    return r;
  }
#ifndef REG
#define REG(name) name##Registrar_
#endif

  // ---- ICP (types.h:24) -------------------------------------------------------------------------
  struct BatchResult {
    TransformationParameters transformation;
    Matrix covariance;  // 6 x 6
    int iterations;
    bool maxNumIterationsReached;
    bool converged;     // false: ConvergenceError / TransformationError for this pair (status)
    pgs_status status;
    T overlap, residual, weightedPointUsedRatio;
  };

  class ICP {
   public:
    // public members of ICPChainBase that pgslam reaches into
    DataPointsFilters readingDataPointsFilters, readingStepDataPointsFilters, referenceDataPointsFilters;
    Transformations transformations;
    std::shared_ptr<Matcher> matcher;
    OutlierFilters outlierFilters;
    std::shared_ptr<ErrorMinimizer> errorMinimizer;

    ICP() : matcher(std::make_shared<Matcher>()), errorMinimizer(std::make_shared<ErrorMinimizer>()), ctx_(context()), h_(nullptr) {
      setDefault();
    }
    ICP(const ICP&) = delete;
    ICP& operator=(const ICP&) = delete;
    virtual ~ICP() {
      for (auto& p : peers_) pgs_icp_destroy(p.h);
      pgs_icp_destroy(h_);
    }

    void setDefault() {
      pgs_icp* h = nullptr;
      check(pgs_icp_create_default(ctx_->h, &h), ctx_);
      yaml_.clear();
      replace(h);
    }
    void loadFromYaml(std::istream& in) {  // Localizer.hpp:70,311; LoopCloser.hpp:73,348
      const std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
      pgs_icp* h = nullptr;
      check(pgs_icp_create_from_yaml(ctx_->h, text.data(), text.size(), &h), ctx_);
      yaml_ = text;
      replace(h);
    }
    // ICP::operator()(reading, reference, T_init)  (LoopCloser.hpp:98)
    TransformationParameters operator()(const DataPoints& reading, const DataPoints& reference,
                                        const TransformationParameters& initial) {
      auto dr = reading.device(ctx_);
      auto df = reference.device(ctx_);
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

    // Loop-closure candidate verification as ONE call (the candidate loop of
    // LoopCloser.hpp:266-297 + ProcessVertex :98 per candidate): P independent registrations,
    // cut into contiguous blocks over `devices` (default: this object's device).  A failing
    // pair does not throw: its BatchResult carries the status.  Results are bit-identical to
    // P calls of operator() whatever the devices.
    std::vector<BatchResult> computeBatch(const std::vector<const DataPoints*>& readings,
                                          const std::vector<const DataPoints*>& references,
                                          const std::vector<TransformationParameters>& initials = {},
                                          const std::vector<int>& devices = {}) {
      const int P = static_cast<int>(readings.size());
      if (references.size() != readings.size() || (!initials.empty() && initials.size() != readings.size()))
        throw InvalidParameter("computeBatch: readings, references and initial guesses differ in number");
      std::vector<pgs_icp*> handles;
      for (auto& p : peers_) p.used = false;
      bool own_used = false;
      if (devices.empty()) handles.push_back(h_);
      for (int d : devices) {
        if (d == ctx_->device && !own_used) {
          handles.push_back(h_);
          own_used = true;
        } else {
          handles.push_back(peer(d));
        }
      }
      // host staging: float, column-major 4 x N (features) / point-major descriptor blocks
      struct Staged {
        std::vector<float> feat;
        std::vector<std::vector<float>> desc;
        std::vector<const char*> labels;
        std::vector<int> spans;
        std::vector<const float*> data;
      };
      std::vector<Staged> stage(2 * static_cast<size_t>(P));
      std::vector<pgs_host_cloud> hr(P), hf(P);
      auto fill = [](const DataPoints& c, Staged& s, pgs_host_cloud& h) {
        const int64_t n = c.features.cols();
        if (n > 0 && c.features.rows() != 4) throw InvalidField("features must be 4 x N (3-D homogeneous points)");
        if (sizeof(T) == 4) {
          h.features4xN = reinterpret_cast<const float*>(c.features.data());
        } else {
          s.feat.resize(static_cast<size_t>(n) * 4);
          const T* src = c.features.data();
          for (size_t i = 0; i < s.feat.size(); ++i) s.feat[i] = static_cast<float>(src[i]);
          h.features4xN = s.feat.data();
        }
        h.n = n;
        int row = 0;
        for (auto& l : c.descriptorLabels) {
          s.desc.emplace_back(static_cast<size_t>(n) * l.span);
          std::vector<float>& v = s.desc.back();
          for (int64_t i = 0; i < n; ++i)
            for (size_t r = 0; r < l.span; ++r)
              v[static_cast<size_t>(i) * l.span + r] = static_cast<float>(c.descriptors(row + static_cast<int>(r), static_cast<int>(i)));
          s.labels.push_back(l.text.c_str());
          s.spans.push_back(static_cast<int>(l.span));
          row += static_cast<int>(l.span);
        }
        for (auto& v : s.desc) s.data.push_back(v.data());
        h.n_descriptors = static_cast<int>(s.labels.size());
        h.labels = s.labels.empty() ? nullptr : s.labels.data();
        h.spans = s.spans.empty() ? nullptr : s.spans.data();
        h.data = s.data.empty() ? nullptr : s.data.data();
      };
      for (int p = 0; p < P; ++p) {
        fill(*readings[p], stage[2 * static_cast<size_t>(p)], hr[p]);
        fill(*references[p], stage[2 * static_cast<size_t>(p) + 1], hf[p]);
      }
      std::vector<double> Ti;
      if (!initials.empty()) {
        Ti.resize(static_cast<size_t>(16) * P);
        for (int p = 0; p < P; ++p) to_double16(initials[p], Ti.data() + static_cast<size_t>(16) * p);
      }
      std::vector<pgs_icp_result> res(static_cast<size_t>(P > 0 ? P : 1));
      const pgs_status st = pgs_icp_run_batch_multi(handles.data(), static_cast<int>(handles.size()), P, hr.data(), hf.data(),
                                                    Ti.empty() ? nullptr : Ti.data(), 0, res.data());
      bool any_ok = P == 0;
      for (int p = 0; p < P; ++p) any_ok = any_ok || res[p].status == PGS_OK;
      if (st != PGS_OK && any_ok) check(st, ctx_);  // an infrastructure error, not "every pair failed"
      std::vector<BatchResult> out(static_cast<size_t>(P));
      for (int p = 0; p < P; ++p) {
        const pgs_icp_result& r = res[p];
        BatchResult& b = out[p];
        b.transformation = from_double16(r.T);
        b.covariance = Matrix::Zero(6, 6);
        for (int c = 0; c < 6; ++c)
          for (int rr = 0; rr < 6; ++rr) b.covariance(rr, c) = static_cast<T>(r.covariance[c * 6 + rr]);
        b.iterations = r.iterations;
        b.maxNumIterationsReached = r.max_iterations_reached != 0;
        b.status = static_cast<pgs_status>(r.status);
        b.converged = r.status == PGS_OK;
        b.overlap = static_cast<T>(r.overlap);
        b.residual = static_cast<T>(r.residual);
        b.weightedPointUsedRatio = static_cast<T>(r.weighted_point_used_ratio);
      }
      return out;
    }

   protected:
    TransformationParameters finish(pgs_status st, const pgs_icp_result& r) {
      last_ = r;
      errorMinimizer->last.overlap = r.overlap;
      errorMinimizer->last.residual = r.residual;
      errorMinimizer->last.weighted_point_used_ratio = r.weighted_point_used_ratio;
      errorMinimizer->last.point_used_ratio = r.point_used_ratio;
      std::memcpy(errorMinimizer->last.covariance, r.covariance, sizeof(r.covariance));
      errorMinimizer->have_last = true;
      check(st, ctx_);
      return from_double16(r.T);
    }
    void replace(pgs_icp* h) {
      pgs_icp_destroy(h_);
      for (auto& p : peers_) pgs_icp_destroy(p.h);
      peers_.clear();
      h_ = h;
      std::memset(&last_, 0, sizeof(last_));
      readingDataPointsFilters.borrow(ctx_, pgs_icp_reading_filters(h_));
      readingStepDataPointsFilters.borrow(ctx_, pgs_icp_reading_step_filters(h_));
      referenceDataPointsFilters.borrow(ctx_, pgs_icp_reference_filters(h_));
      matcher->borrow(ctx_, pgs_icp_matcher(h_));
      outlierFilters.borrow(ctx_, pgs_icp_outliers(h_));
      errorMinimizer->borrow(ctx_, pgs_icp_minimizer(h_));
      errorMinimizer->have_last = false;
    }
    // the same chain on another device (or a second context of this one), created on first use
    pgs_icp* peer(int device) {
      for (auto& p : peers_)
        if (p.device == device && !p.used) { p.used = true; return p.h; }
      Peer p;
      p.device = device;
      p.ctx = std::make_shared<detail::Ctx>(device);  // private context: the batch drives it from its own thread
      if (yaml_.empty()) check(pgs_icp_create_default(p.ctx->h, &p.h), p.ctx);
      else check(pgs_icp_create_from_yaml(p.ctx->h, yaml_.data(), yaml_.size(), &p.h), p.ctx);
      p.used = true;
      peers_.push_back(p);
      return peers_.back().h;
    }
    struct Peer {
      int device = 0;
      CtxPtr ctx;
      pgs_icp* h = nullptr;
      bool used = false;
    };
    CtxPtr ctx_;
    pgs_icp* h_;
    pgs_icp_result last_;
    std::string yaml_;
    std::vector<Peer> peers_;
  };

  // ---- ICPSequence (types.h:25; Localizer.hpp:126,148,168,254) ---------------------------------
  class ICPSequence : public ICP {
   public:
    bool setMap(const DataPoints& map) {
      map_ = map.device(this->ctx_);
      check(pgs_icp_set_map(this->h_, map_->h), this->ctx_);
      return true;
    }
    bool hasMap() const { return pgs_icp_has_map(this->h_) != 0; }
    TransformationParameters operator()(const DataPoints& cloud, const TransformationParameters& initial) {
      if (!hasMap()) return Matrix::Identity(4, 4);  // upstream: "Ignoring attempt to perform ICP with an empty map"
      auto d = cloud.device(this->ctx_);
      double t[16];
      PointMatcher::to_double16(initial, t);
      pgs_icp_result r;
      pgs_status st = pgs_icp_run_sequence(this->h_, d->h, t, &r);
      return this->finish(st, r);
    }
    TransformationParameters operator()(const DataPoints& cloud) { return (*this)(cloud, Matrix::Identity(4, 4)); }

   private:
    typename DataPoints::DevicePtr map_;
  };
};

}  // namespace pgslam_b200
