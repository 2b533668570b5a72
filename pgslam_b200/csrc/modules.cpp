// modules.cpp — Registrar / Parametrizable mirror, YAML subset parser and ICP
// chain loader (host only).  Behavioural spec: SURVEY.md §8a rows A17-A18 and
// Appendix A.10; the call sites that reach it are Localizer.hpp:55-78,
// LoopCloser.hpp:59-74 (YAML files slurped to strings, re-parsed per temp ICP).
#include "modules.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <sstream>

namespace pgs {

namespace {

struct ModuleDoc {
  Kind kind;
  const char* name;
  std::vector<ParamDoc> params;
};

const char* INF = "inf";
const char* NINF = "-inf";
const char* IMAX = "2147483647";

const std::vector<ModuleDoc>& registry() {
  static const std::vector<ModuleDoc> reg = {
      // ---- DataPointsFilters (A3-A6, A.9) --------------------------------
      {Kind::DataPointsFilter, "IdentityDataPointsFilter", {}},
      {Kind::DataPointsFilter, "RemoveNaNDataPointsFilter", {}},
      {Kind::DataPointsFilter, "FixStepSamplingDataPointsFilter",
       {{"startStep", "keep one point in `startStep`", "10", "1", IMAX, 'i'},
        {"endStep", "must equal startStep (step schedules across calls are not supported)", "10", "1", IMAX, 'i'},
        {"stepMult", "must be 1", "1", "0.0000001", INF, 'f'},
        {"seed", "seed of the phase (upstream: rand() % step, SURVEY H7)", "0", "0", "", 'i'}}},
      {Kind::DataPointsFilter, "ShadowDataPointsFilter",
       {{"eps", "minimum |cos| between the normal and the ray from the origin", "0.1", "0.0000001", "3.1416", 'f'}}},
      {Kind::DataPointsFilter, "RandomSamplingDataPointsFilter",
       {{"prob", "probability to keep a point", "0.75", "0", "1", 'f'},
        {"seed", "seed of the counter-based generator (SURVEY H7)", "0", "0", "", 'i'}}},
      {Kind::DataPointsFilter, "VoxelGridDataPointsFilter",
       {{"vSizeX", "voxel size, x", "1.0", NINF, INF, 'f'},
        {"vSizeY", "voxel size, y", "1.0", NINF, INF, 'f'},
        {"vSizeZ", "voxel size, z", "1.0", NINF, INF, 'f'},
        {"useCentroid", "centroid (1) or voxel centre (0)", "1", "0", "1", 'u'},
        {"averageExistingDescriptors", "average descriptors in a voxel", "1", "0", "1", 'u'}}},
      {Kind::DataPointsFilter, "SurfaceNormalDataPointsFilter",
       {{"knn", "neighbours used per point", "5", "3", IMAX, 'i'},
        {"maxDist", "maximum neighbour distance", INF, "0", INF, 'f'},
        {"epsilon", "approximation of the search: must be 0 (exact); see PGS_EPSILON_POLICY", "0", "0", INF, 'f'},
        {"keepNormals", "", "1", "0", "1", 'u'},
        {"keepDensities", "", "0", "0", "1", 'u'},
        {"keepEigenValues", "", "0", "0", "1", 'u'},
        {"keepEigenVectors", "", "0", "0", "1", 'u'},
        {"keepMatchedIds", "", "0", "0", "1", 'u'},
        {"keepMeanDist", "", "0", "0", "1", 'u'},
        {"sortEigen", "", "0", "0", "1", 'u'},
        {"smoothNormals", "", "0", "0", "1", 'u'}}},
      {Kind::DataPointsFilter, "ObservationDirectionDataPointsFilter",
       {{"x", "sensor x", "0", NINF, INF, 'f'},
        {"y", "sensor y", "0", NINF, INF, 'f'},
        {"z", "sensor z", "0", NINF, INF, 'f'}}},
      {Kind::DataPointsFilter, "OrientNormalsDataPointsFilter",
       {{"towardCenter", "orient normals toward the sensor", "1", "0", "1", 'u'}}},
      {Kind::DataPointsFilter, "SimpleSensorNoiseDataPointsFilter",
       {{"sensorType", "0 LMS-1xx, 1 URG-04LX, 2 UTM-30LX, 3 Kinect, 4 Tim3xx", "0", "0", IMAX, 'i'},
        {"gain", "uncertainty gain", "1", "1", INF, 'f'}}},
      {Kind::DataPointsFilter, "MaxDistDataPointsFilter",
       {{"dim", "axis, -1 = radial", "-1", "-1", "2", 'i'}, {"maxDist", "", "1", NINF, INF, 'f'}}},
      {Kind::DataPointsFilter, "MinDistDataPointsFilter",
       {{"dim", "axis, -1 = radial", "-1", "-1", "2", 'i'}, {"minDist", "", "1", NINF, INF, 'f'}}},
      {Kind::DataPointsFilter, "SamplingSurfaceNormalDataPointsFilter",
       {{"ratio", "ratio of points to keep with random subsampling", "0.5", "0.0000001", "1.0", 'f'},
        {"knn", "largest number of points in a cell", "7", "3", "64", 'i'},
        {"samplingMethod", "0: keep points of a cell at random, 1: one point at the cell mean", "0", "0", "1", 'i'},
        {"maxBoxDim", "cells wider than this are dropped", INF, "0", INF, 'f'},
        {"averageExistingDescriptors", "average the descriptors of a cell (samplingMethod 1)", "1", "0", "1", 'u'},
        {"keepNormals", "", "1", "0", "1", 'u'},
        {"keepDensities", "", "0", "0", "1", 'u'},
        {"keepEigenValues", "", "0", "0", "1", 'u'},
        {"keepEigenVectors", "", "0", "0", "1", 'u'},
        {"seed", "seed of the counter-based generator (SURVEY H7)", "0", "0", "", 'i'}}},
      {Kind::DataPointsFilter, "MaxDensityDataPointsFilter",
       {{"maxDensity", "points denser than this are subsampled towards it", "10", "0.0000001", INF, 'f'},
        {"seed", "seed of the counter-based generator (SURVEY H7)", "0", "0", "", 'i'}}},
      {Kind::DataPointsFilter, "BoundingBoxDataPointsFilter",
       {{"xMin", "", "-1", NINF, INF, 'f'}, {"xMax", "", "1", NINF, INF, 'f'},
        {"yMin", "", "-1", NINF, INF, 'f'}, {"yMax", "", "1", NINF, INF, 'f'},
        {"zMin", "", "-1", NINF, INF, 'f'}, {"zMax", "", "1", NINF, INF, 'f'},
        {"removeInside", "remove the points inside (1) or outside (0) the box", "1", "0", "1", 'u'}}},
      // ---- Matcher (A8, A9) ----------------------------------------------
      {Kind::Matcher, "KDTreeMatcher",
       {{"knn", "number of nearest neighbours", "1", "1", IMAX, 'i'},
        {"epsilon", "approximation of the search: must be 0 (exact); see PGS_EPSILON_POLICY", "0", "0", INF, 'f'},
        {"searchType", "libnabo search type (ignored: one exact GPU index)", "1", "0", "2", 'i'},
        {"maxDist", "maximum distance to consider", INF, "0", INF, 'f'}}},
      // ---- OutlierFilters (A10, A.3) -------------------------------------
      {Kind::OutlierFilter, "NullOutlierFilter", {}},
      {Kind::OutlierFilter, "TrimmedDistOutlierFilter",
       {{"ratio", "fraction of closest matches kept", "0.85", "0.0000001", "1.0", 'f'}}},
      {Kind::OutlierFilter, "MaxDistOutlierFilter", {{"maxDist", "", "1", "0.0000001", INF, 'f'}}},
      {Kind::OutlierFilter, "MinDistOutlierFilter", {{"minDist", "", "1", "0.0000001", INF, 'f'}}},
      {Kind::OutlierFilter, "MedianDistOutlierFilter", {{"factor", "", "3", "0.0000001", INF, 'f'}}},
      {Kind::OutlierFilter, "VarTrimmedDistOutlierFilter",
       {{"minRatio", "lower bound of the optimised inlier ratio", "0.05", "0.0000001", "1", 'f'},
        {"maxRatio", "upper bound of the optimised inlier ratio", "0.99", "0.0000001", "1", 'f'},
        {"lambda", "exponent of the FRMS criterion", "0.95", "", "", 'f'}}},
      {Kind::OutlierFilter, "SurfaceNormalOutlierFilter",
       {{"maxAngle", "max angle (rad) between the normals of a matched pair", "1.57", "0.0", "3.1416", 'f'}}},
      // ---- ErrorMinimizers (A12, A12d, A13) ------------------------------
      {Kind::ErrorMinimizer, "PointToPlaneErrorMinimizer",
       {{"force2D", "", "0", "0", "1", 'u'}, {"force4DOF", "", "0", "0", "1", 'u'}}},
      {Kind::ErrorMinimizer, "PointToPlaneWithCovErrorMinimizer",
       {{"force2D", "", "0", "0", "1", 'u'},
        {"force4DOF", "", "0", "0", "1", 'u'},
        {"sensorStdDev", "sensor standard deviation", "0.01", "0", INF, 'f'}}},
      {Kind::ErrorMinimizer, "PointToPointErrorMinimizer", {}},
      // ---- TransformationCheckers (A14, A.8) -----------------------------
      {Kind::TransformationChecker, "CounterTransformationChecker",
       {{"maxIterationCount", "", "40", "0", IMAX, 'i'}}},
      {Kind::TransformationChecker, "DifferentialTransformationChecker",
       {{"minDiffRotErr", "", "0.001", "0", "6.2831854", 'f'},
        {"minDiffTransErr", "", "0.001", "0", INF, 'f'},
        {"smoothLength", "", "3", "0", "15", 'i'}}},
      {Kind::TransformationChecker, "BoundTransformationChecker",
       {{"maxRotationNorm", "", "1", "0", INF, 'f'}, {"maxTranslationNorm", "", "1", "0", INF, 'f'}}},
      // ---- no-op plumbing modules ----------------------------------------
      {Kind::Inspector, "NullInspector", {}},
      {Kind::Logger, "NullLogger", {}},
      {Kind::Logger, "FileLogger",
       {{"infoFileName", "", "", "", "", 's'},
        {"warningFileName", "", "", "", "", 's'},
        {"displayLocation", "", "0", "0", "1", 'u'}}},
      {Kind::Transformation, "RigidTransformation", {}},
  };
  return reg;
}

bool parse_number(const std::string& s, double* out) {
  std::string t = s;
  while (!t.empty() && isspace((unsigned char)t.back())) t.pop_back();
  size_t b = 0;
  while (b < t.size() && isspace((unsigned char)t[b])) ++b;
  t = t.substr(b);
  if (t.empty()) return false;
  if (t == "inf" || t == "+inf" || t == ".inf" || t == "Inf") { *out = std::numeric_limits<double>::infinity(); return true; }
  if (t == "-inf" || t == "-.inf" || t == "-Inf") { *out = -std::numeric_limits<double>::infinity(); return true; }
  if (t == "true") { *out = 1; return true; }
  if (t == "false") { *out = 0; return true; }
  char* end = nullptr;
  double v = strtod(t.c_str(), &end);
  if (end == t.c_str() || *end != '\0') return false;
  *out = v;
  return true;
}

const char* kind_name(Kind k) {
  switch (k) {
    case Kind::DataPointsFilter: return "DataPointsFilter";
    case Kind::Matcher: return "Matcher";
    case Kind::OutlierFilter: return "OutlierFilter";
    case Kind::ErrorMinimizer: return "ErrorMinimizer";
    case Kind::TransformationChecker: return "TransformationChecker";
    case Kind::Inspector: return "Inspector";
    case Kind::Logger: return "Logger";
    case Kind::Transformation: return "Transformation";
  }
  return "?";
}

}  // namespace

const std::vector<ParamDoc>* module_params(Kind kind, const std::string& name) {
  for (auto& m : registry())
    if (m.kind == kind && name == m.name) return &m.params;
  return nullptr;
}

std::vector<std::string> registered_modules(Kind kind) {
  std::vector<std::string> out;
  for (auto& m : registry())
    if (m.kind == kind) out.push_back(m.name);
  return out;
}

Module create_module(Kind kind, const std::string& name, const Params& params) {
  const std::vector<ParamDoc>* docs = module_params(kind, name);
  if (!docs)
    throw Error(PGS_INVALID_ELEMENT, std::string("Trying to instanciate unknown element ") + name + " from " +
                                         kind_name(kind) + " registrar");
  Module m;
  m.kind = kind;
  m.name = name;
  for (auto& d : *docs) m.params[d.name] = d.def;
  for (auto& kv : params) {
    const ParamDoc* doc = nullptr;
    for (auto& d : *docs)
      if (kv.first == d.name) doc = &d;
    if (!doc)
      throw Error(PGS_INVALID_PARAMETER, "Parameter " + kv.first + " for module " + name + " was set but is not used");
    if (doc->type != 's') {
      double v;
      if (!parse_number(kv.second, &v))
        throw Error(PGS_INVALID_PARAMETER, "Parameter " + kv.first + " of " + name + ": cannot cast value '" + kv.second + "'");
      if ((doc->type == 'i' || doc->type == 'u') && v != std::floor(v))
        throw Error(PGS_INVALID_PARAMETER, "Parameter " + kv.first + " of " + name + " must be an integer, got " + kv.second);
      double lo, hi;
      if (doc->min[0] && parse_number(doc->min, &lo) && v < lo)
        throw Error(PGS_INVALID_PARAMETER, "Value " + kv.second + " of parameter " + kv.first + " in " + name +
                                               " is smaller than minimum admissible value " + doc->min);
      if (doc->max[0] && parse_number(doc->max, &hi) && v > hi)
        throw Error(PGS_INVALID_PARAMETER, "Value " + kv.second + " of parameter " + kv.first + " in " + name +
                                               " is larger than maximum admissible value " + doc->max);
    }
    m.params[kv.first] = kv.second;
  }
  // `epsilon` != 0 asks libnabo for an APPROXIMATE search whose answer depends on libnabo's own
  // tree shape and visit order (measured with the CPU checker: 8-49 % of the ids change, poses move by
  // centimetres).  This library searches exactly; silently doing so would return results the
  // configuration did not ask for, so the parameter is refused unless the caller opts in.
  auto eps = m.params.find("epsilon");
  if (eps != m.params.end() && (name == "KDTreeMatcher" || name == "SurfaceNormalDataPointsFilter")) {
    double v = 0;
    if (parse_number(eps->second, &v) && v != 0.0) {
      const char* pol = std::getenv("PGS_EPSILON_POLICY");
      if (!(pol && std::string(pol) == "exact"))
        throw Error(PGS_INVALID_PARAMETER, "Parameter epsilon = " + eps->second + " of " + name +
                                               ": approximate (eps > 0) nearest-neighbour search is not implemented - this "
                                               "library searches exactly.  Set epsilon: 0, or export PGS_EPSILON_POLICY=exact "
                                               "to accept the exact search in its place");
    }
  }
  return m;
}

double Module::real(const std::string& k) const {
  auto it = params.find(k);
  if (it == params.end()) throw Error(PGS_INVALID_PARAMETER, "Parameter " + k + " does not exist in " + name);
  double v = 0;
  if (!parse_number(it->second, &v)) throw Error(PGS_INVALID_PARAMETER, "Parameter " + k + ": bad value " + it->second);
  return v;
}
int64_t Module::integer(const std::string& k) const { return (int64_t)real(k); }

// ===========================================================================
// YAML subset
// ===========================================================================
namespace {

struct Line {
  int indent;
  std::string text;
};

std::string trim(const std::string& s) {
  size_t b = 0, e = s.size();
  while (b < e && isspace((unsigned char)s[b])) ++b;
  while (e > b && isspace((unsigned char)s[e - 1])) --e;
  return s.substr(b, e - b);
}

std::string unquote(const std::string& s) {
  std::string t = trim(s);
  if (t.size() >= 2 && ((t.front() == '"' && t.back() == '"') || (t.front() == '\'' && t.back() == '\'')))
    return t.substr(1, t.size() - 2);
  return t;
}

std::vector<Line> split_lines(const std::string& text) {
  std::vector<Line> out;
  std::istringstream in(text);
  std::string raw;
  while (std::getline(in, raw)) {
    // strip comments outside quotes
    std::string s;
    char quote = 0;
    for (size_t i = 0; i < raw.size(); ++i) {
      char c = raw[i];
      if (quote) {
        if (c == quote) quote = 0;
      } else if (c == '"' || c == '\'') {
        quote = c;
      } else if (c == '#' && (i == 0 || isspace((unsigned char)raw[i - 1]))) {
        break;
      }
      s.push_back(c);
    }
    while (!s.empty() && isspace((unsigned char)s.back())) s.pop_back();
    if (s.empty()) continue;
    int indent = 0;
    while (indent < (int)s.size() && (s[indent] == ' ' || s[indent] == '\t')) ++indent;
    std::string body = s.substr(indent);
    if (body == "---" || body == "..." || body.rfind("%YAML", 0) == 0) continue;
    out.push_back({indent, body});
  }
  return out;
}

bool is_dash(const std::string& t) { return !t.empty() && t[0] == '-' && (t.size() == 1 || t[1] == ' '); }

// position of the key/value ':' outside quotes and flow brackets, or npos
size_t find_colon(const std::string& t) {
  char quote = 0;
  int depth = 0;
  for (size_t i = 0; i < t.size(); ++i) {
    char c = t[i];
    if (quote) { if (c == quote) quote = 0; continue; }
    if (c == '"' || c == '\'') { quote = c; continue; }
    if (c == '{' || c == '[') ++depth;
    if (c == '}' || c == ']') --depth;
    if (c == ':' && depth == 0 && (i + 1 == t.size() || t[i + 1] == ' ')) return i;
  }
  return std::string::npos;
}

struct FlowParser {
  const std::string& s;
  size_t i = 0;
  explicit FlowParser(const std::string& str) : s(str) {}
  void ws() { while (i < s.size() && isspace((unsigned char)s[i])) ++i; }
  [[noreturn]] void fail(const char* what) { throw Error(PGS_INVALID_PARAMETER, std::string("YAML: ") + what + " in '" + s + "'"); }
  std::string scalar(const char* stops) {
    ws();
    std::string out;
    if (i < s.size() && (s[i] == '"' || s[i] == '\'')) {
      char q = s[i++];
      while (i < s.size() && s[i] != q) out.push_back(s[i++]);
      if (i >= s.size()) fail("unterminated quote");
      ++i;
      return out;
    }
    while (i < s.size() && !strchr(stops, s[i])) {
      if (s[i] == ':' && (i + 1 >= s.size() || s[i + 1] == ' ') && strchr(stops, ':')) break;
      out.push_back(s[i++]);
    }
    return trim(out);
  }
  YamlNode value() {
    ws();
    YamlNode n;
    if (i < s.size() && s[i] == '{') {
      ++i;
      n.type = YamlNode::Map;
      ws();
      if (i < s.size() && s[i] == '}') { ++i; return n; }
      while (true) {
        std::string key = scalar(",}:");
        ws();
        YamlNode v;
        if (i < s.size() && s[i] == ':') { ++i; v = value(); }
        n.map.emplace_back(key, v);
        ws();
        if (i < s.size() && s[i] == ',') { ++i; continue; }
        if (i < s.size() && s[i] == '}') { ++i; break; }
        fail("expected ',' or '}'");
      }
      return n;
    }
    if (i < s.size() && s[i] == '[') {
      ++i;
      n.type = YamlNode::Seq;
      ws();
      if (i < s.size() && s[i] == ']') { ++i; return n; }
      while (true) {
        n.seq.push_back(value());
        ws();
        if (i < s.size() && s[i] == ',') { ++i; continue; }
        if (i < s.size() && s[i] == ']') { ++i; break; }
        fail("expected ',' or ']'");
      }
      return n;
    }
    n.type = YamlNode::Scalar;
    n.scalar = scalar(",}]");
    if (n.scalar.empty() || n.scalar == "~" || n.scalar == "null") n.type = n.scalar.empty() ? YamlNode::Null : YamlNode::Scalar;
    return n;
  }
};

YamlNode parse_inline(const std::string& text) {
  std::string t = trim(text);
  YamlNode n;
  if (t.empty() || t == "~") return n;
  if (t[0] == '{' || t[0] == '[') {
    FlowParser fp(t);
    n = fp.value();
    fp.ws();
    if (fp.i != t.size()) fp.fail("trailing characters");
    return n;
  }
  n.type = YamlNode::Scalar;
  n.scalar = unquote(t);
  return n;
}

YamlNode parse_block(std::vector<Line>& L, size_t& i, int indent);

YamlNode parse_map(std::vector<Line>& L, size_t& i, int indent) {
  YamlNode n;
  n.type = YamlNode::Map;
  while (i < L.size() && L[i].indent == indent && !is_dash(L[i].text)) {
    size_t c = find_colon(L[i].text);
    if (c == std::string::npos) throw Error(PGS_INVALID_PARAMETER, "YAML: expected 'key: value' in '" + L[i].text + "'");
    std::string key = unquote(L[i].text.substr(0, c));
    std::string val = trim(L[i].text.substr(c + 1));
    ++i;
    YamlNode child;
    if (val.empty()) {
      if (i < L.size() && (L[i].indent > indent || (L[i].indent == indent && is_dash(L[i].text))))
        child = parse_block(L, i, L[i].indent);
    } else {
      child = parse_inline(val);
    }
    n.map.emplace_back(key, child);
  }
  return n;
}

YamlNode parse_seq(std::vector<Line>& L, size_t& i, int indent) {
  YamlNode n;
  n.type = YamlNode::Seq;
  while (i < L.size() && L[i].indent == indent && is_dash(L[i].text)) {
    std::string rest = L[i].text.substr(1);
    size_t skip = 0;
    while (skip < rest.size() && rest[skip] == ' ') ++skip;
    rest = rest.substr(skip);
    int rest_indent = indent + 1 + (int)skip;
    if (rest.empty()) {
      ++i;
      if (i < L.size() && L[i].indent > indent) n.seq.push_back(parse_block(L, i, L[i].indent));
      else n.seq.push_back(YamlNode());
    } else if (rest[0] != '{' && rest[0] != '[' && find_colon(rest) != std::string::npos) {
      L[i].indent = rest_indent;  // the item is a block map that starts on the dash line
      L[i].text = rest;
      n.seq.push_back(parse_map(L, i, rest_indent));
    } else {
      n.seq.push_back(parse_inline(rest));
      ++i;
    }
  }
  return n;
}

YamlNode parse_block(std::vector<Line>& L, size_t& i, int indent) {
  if (is_dash(L[i].text)) return parse_seq(L, i, indent);
  if (find_colon(L[i].text) == std::string::npos) {
    YamlNode n = parse_inline(L[i].text);
    ++i;
    return n;
  }
  return parse_map(L, i, indent);
}

}  // namespace

const YamlNode* YamlNode::get(const std::string& key) const {
  for (auto& kv : map)
    if (kv.first == key) return &kv.second;
  return nullptr;
}

YamlNode parse_yaml(const std::string& text) {
  std::vector<Line> L = split_lines(text);
  if (L.empty()) return YamlNode();
  size_t i = 0;
  YamlNode root = parse_block(L, i, L[0].indent);
  if (i != L.size()) throw Error(PGS_INVALID_PARAMETER, "YAML: unexpected indentation at '" + L[i].text + "'");
  return root;
}

Module module_from_yaml(Kind kind, const YamlNode& node) {
  if (node.type == YamlNode::Scalar) return create_module(kind, node.scalar, {});
  if (node.type == YamlNode::Map && node.map.size() == 1) {
    const std::string& name = node.map[0].first;
    const YamlNode& pv = node.map[0].second;
    Params p;
    if (pv.type == YamlNode::Map) {
      for (auto& kv : pv.map) {
        if (kv.second.type != YamlNode::Scalar)
          throw Error(PGS_INVALID_PARAMETER, "Parameter " + kv.first + " of " + name + " must be a scalar");
        p[kv.first] = kv.second.scalar;
      }
    } else if (pv.type != YamlNode::Null) {
      throw Error(PGS_INVALID_PARAMETER, "Parameters of " + name + " must be a map");
    }
    return create_module(kind, name, p);
  }
  throw Error(PGS_INVALID_MODULE_TYPE, std::string("Malformed module entry for ") + kind_name(kind));
}

std::vector<Module> module_list_from_yaml(Kind kind, const YamlNode& node) {
  std::vector<Module> out;
  if (node.type == YamlNode::Null) return out;
  if (node.type != YamlNode::Seq)
    throw Error(PGS_INVALID_MODULE_TYPE, std::string("Expected a list of ") + kind_name(kind) + " modules");
  for (auto& it : node.seq) out.push_back(module_from_yaml(kind, it));
  return out;
}

ChainConfig chain_default() {
  // ICPChainBase::setDefault (A17): RandomSampling(0.75) on the reading,
  // SamplingSurfaceNormal on the reference, TrimmedDist(0.85), KDTreeMatcher,
  // PointToPlane, Counter + Differential checkers
  ChainConfig c;
  c.reading_filters.push_back(create_module(Kind::DataPointsFilter, "RandomSamplingDataPointsFilter", {}));
  c.reference_filters.push_back(create_module(Kind::DataPointsFilter, "SamplingSurfaceNormalDataPointsFilter", {}));
  c.outlier_filters.push_back(create_module(Kind::OutlierFilter, "TrimmedDistOutlierFilter", {}));
  c.matcher = create_module(Kind::Matcher, "KDTreeMatcher", {});
  c.minimizer = create_module(Kind::ErrorMinimizer, "PointToPlaneErrorMinimizer", {});
  c.checkers.push_back(create_module(Kind::TransformationChecker, "CounterTransformationChecker", {}));
  c.checkers.push_back(create_module(Kind::TransformationChecker, "DifferentialTransformationChecker", {}));
  c.inspector = create_module(Kind::Inspector, "NullInspector", {});
  c.logger = create_module(Kind::Logger, "NullLogger", {});
  return c;
}

ChainConfig chain_from_yaml(const std::string& text) {
  YamlNode root = parse_yaml(text);
  if (root.type != YamlNode::Map) throw Error(PGS_INVALID_MODULE_TYPE, "ICP configuration must be a YAML map");
  ChainConfig c;
  c.matcher = create_module(Kind::Matcher, "KDTreeMatcher", {});
  c.minimizer = create_module(Kind::ErrorMinimizer, "PointToPlaneErrorMinimizer", {});
  c.inspector = create_module(Kind::Inspector, "NullInspector", {});
  c.logger = create_module(Kind::Logger, "NullLogger", {});
  for (auto& kv : root.map) {
    const std::string& k = kv.first;
    const YamlNode& v = kv.second;
    if (k == "readingDataPointsFilters") c.reading_filters = module_list_from_yaml(Kind::DataPointsFilter, v);
    else if (k == "readingStepDataPointsFilters") c.reading_step_filters = module_list_from_yaml(Kind::DataPointsFilter, v);
    else if (k == "referenceDataPointsFilters") c.reference_filters = module_list_from_yaml(Kind::DataPointsFilter, v);
    else if (k == "matcher") c.matcher = module_from_yaml(Kind::Matcher, v);
    else if (k == "outlierFilters") c.outlier_filters = module_list_from_yaml(Kind::OutlierFilter, v);
    else if (k == "errorMinimizer") c.minimizer = module_from_yaml(Kind::ErrorMinimizer, v);
    else if (k == "transformationCheckers") c.checkers = module_list_from_yaml(Kind::TransformationChecker, v);
    else if (k == "inspector") c.inspector = module_from_yaml(Kind::Inspector, v);
    else if (k == "logger") c.logger = module_from_yaml(Kind::Logger, v);
    else
      throw Error(PGS_INVALID_MODULE_TYPE, "Module type " + k + " does not exist");
  }
  return c;
}

}  // namespace pgs
