// cloud_io.cpp — DataPoints::load / save in libpointmatcher's three text formats (.csv, legacy
// ASCII .vtk POLYDATA, .ply ascii / binary) behind the C ABI.  Host side only:
// files are parsed into host arrays and handed to the same cloud constructors every other
// caller uses.  Column naming follows upstream's IO.cpp [UPSTREAM-RECALLED]: x y z -> features;
// nx ny nz -> `normals`; a scalar column `name` -> descriptor `name`; `name_x name_y name_z`
// (CSV / PLY) or VECTORS|NORMALS name (VTK) -> one span-3 descriptor; `name_0 .. name_k` -> span k+1.
#include "cloud_io.h"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>

namespace pgs {

namespace {

std::string lower_ext(const std::string& path) {
  const size_t dot = path.rfind('.');
  std::string e = dot == std::string::npos ? "" : path.substr(dot);
  for (auto& c : e) c = (char)std::tolower((unsigned char)c);
  return e;
}

[[noreturn]] void bad(const std::string& path, const std::string& why) {
  throw Error(PGS_INVALID_ARGUMENT, path + ": " + why);
}

std::vector<std::string> split_tokens(const std::string& line, const char* seps) {
  std::vector<std::string> out;
  size_t i = 0;
  while (i < line.size()) {
    while (i < line.size() && std::strchr(seps, line[i])) ++i;
    size_t j = i;
    while (j < line.size() && !std::strchr(seps, line[j])) ++j;
    if (j > i) out.push_back(line.substr(i, j - i));
    i = j;
  }
  return out;
}

// named scalar columns -> features + descriptors by upstream's naming conventions
void columns_to_cloud(const std::string& path, const std::vector<std::string>& names,
                      const std::vector<std::vector<float>>& cols, HostCloud& out) {
  std::map<std::string, int> at;
  for (size_t i = 0; i < names.size(); ++i) at[names[i]] = (int)i;
  for (const char* a : {"x", "y", "z"})
    if (!at.count(a)) bad(path, std::string("no '") + a + "' column (libpointmatcher needs x, y, z)");
  const int64_t n = names.empty() ? 0 : (int64_t)cols[0].size();
  out.n = n;
  out.features.assign((size_t)n * 4, 1.f);
  for (int64_t i = 0; i < n; ++i) {
    out.features[4 * i] = cols[at["x"]][i];
    out.features[4 * i + 1] = cols[at["y"]][i];
    out.features[4 * i + 2] = cols[at["z"]][i];
  }
  std::map<std::string, bool> used{{"x", true}, {"y", true}, {"z", true}};
  auto add = [&](const std::string& label, const std::vector<int>& src) {
    HostDesc d;
    d.label = label;
    d.span = (int)src.size();
    d.data.resize((size_t)n * d.span);
    for (int64_t i = 0; i < n; ++i)
      for (int r = 0; r < d.span; ++r) d.data[(size_t)i * d.span + r] = cols[src[r]][i];
    out.descs.push_back(std::move(d));
  };
  if (at.count("nx") && at.count("ny") && at.count("nz")) {
    add("normals", {at["nx"], at["ny"], at["nz"]});
    used["nx"] = used["ny"] = used["nz"] = true;
  }
  for (auto& nme : names) {
    if (used.count(nme)) continue;
    if (nme.size() > 2 && nme.compare(nme.size() - 2, 2, "_x") == 0) {
      const std::string base = nme.substr(0, nme.size() - 2);
      if (at.count(base + "_y") && at.count(base + "_z")) {
        add(base, {at[base + "_x"], at[base + "_y"], at[base + "_z"]});
        used[base + "_x"] = used[base + "_y"] = used[base + "_z"] = true;
        continue;
      }
    }
    if (nme.size() > 2 && nme.compare(nme.size() - 2, 2, "_0") == 0) {
      const std::string base = nme.substr(0, nme.size() - 2);
      std::vector<int> src;
      for (int k = 0; at.count(base + "_" + std::to_string(k)); ++k) {
        src.push_back(at[base + "_" + std::to_string(k)]);
        used[base + "_" + std::to_string(k)] = true;
      }
      add(base, src);
      continue;
    }
    add(nme, {at[nme]});
    used[nme] = true;
  }
}

// (name, values) scalar columns of a cloud, upstream's CSV naming
std::vector<std::pair<std::string, std::vector<float>>> cloud_columns(const HostCloud& c) {
  std::vector<std::pair<std::string, std::vector<float>>> cols;
  auto col = [&](const std::string& name, const float* base, int stride, int off) {
    std::vector<float> v((size_t)c.n);
    for (int64_t i = 0; i < c.n; ++i) v[i] = base[(size_t)i * stride + off];
    cols.emplace_back(name, std::move(v));
  };
  col("x", c.features.data(), 4, 0);
  col("y", c.features.data(), 4, 1);
  col("z", c.features.data(), 4, 2);
  for (auto& d : c.descs) {
    if (d.label == "normals" && d.span == 3) {
      const char* nn[3] = {"nx", "ny", "nz"};
      for (int r = 0; r < 3; ++r) col(nn[r], d.data.data(), 3, r);
    } else if (d.span == 1) {
      col(d.label, d.data.data(), 1, 0);
    } else if (d.span == 3) {
      const char* ax[3] = {"_x", "_y", "_z"};
      for (int r = 0; r < 3; ++r) col(d.label + ax[r], d.data.data(), 3, r);
    } else {
      for (int r = 0; r < d.span; ++r) col(d.label + "_" + std::to_string(r), d.data.data(), d.span, r);
    }
  }
  return cols;
}

void write_rows(std::FILE* f, const std::vector<std::pair<std::string, std::vector<float>>>& cols, int64_t n, char sep) {
  for (int64_t i = 0; i < n; ++i)
    for (size_t c = 0; c < cols.size(); ++c) std::fprintf(f, "%.9g%c", (double)cols[c].second[i], c + 1 == cols.size() ? '\n' : sep);
}

// ---- CSV ------------------------------------------------------------------------------------
void load_csv(const std::string& path, HostCloud& out) {
  std::ifstream f(path);
  if (!f) bad(path, "cannot open");
  std::vector<std::string> lines;
  std::string ln;
  while (std::getline(f, ln)) {
    size_t a = ln.find_first_not_of(" \t\r\n");
    if (a == std::string::npos || ln[a] == '#') continue;
    size_t b = ln.find_last_not_of(" \t\r\n");
    lines.push_back(ln.substr(a, b - a + 1));
  }
  std::vector<std::string> names{"x", "y", "z"};
  std::vector<std::vector<float>> cols(3);
  if (!lines.empty()) {
    const auto first = split_tokens(lines[0], ",; \t");
    bool header = false;
    for (auto& t : first)
      for (char ch : t)
        if ((std::isalpha((unsigned char)ch) && ch != 'e' && ch != 'E') || ch == '_') header = true;  // 'e' may be an exponent
    if (header) names = first;
    else {
      names = {"x", "y", "z"};
      for (size_t i = 3; i < first.size(); ++i) names.push_back("c" + std::to_string(i - 3));
    }
    cols.assign(names.size(), {});
    for (size_t r = header ? 1 : 0; r < lines.size(); ++r) {
      const auto t = split_tokens(lines[r], ",; \t");
      if (t.size() != names.size()) bad(path, "row " + std::to_string(r + 1) + " has " + std::to_string(t.size()) + " columns");
      for (size_t c = 0; c < t.size(); ++c) cols[c].push_back((float)std::strtod(t[c].c_str(), nullptr));
    }
  }
  columns_to_cloud(path, names, cols, out);
}

void save_csv(const std::string& path, const HostCloud& c) {
  std::FILE* f = std::fopen(path.c_str(), "w");
  if (!f) bad(path, "cannot create");
  const auto cols = cloud_columns(c);
  for (size_t i = 0; i < cols.size(); ++i) std::fprintf(f, "%s%c", cols[i].first.c_str(), i + 1 == cols.size() ? '\n' : ',');
  write_rows(f, cols, c.n, ',');
  std::fclose(f);
}

// ---- legacy ASCII VTK --------------------------------------------------------------------------
void load_vtk(const std::string& path, HostCloud& out) {
  std::ifstream f(path);
  if (!f) bad(path, "cannot open");
  std::string l1, l2;
  std::getline(f, l1);
  std::getline(f, l2);
  if (l1.compare(0, 14, "# vtk DataFile") != 0) bad(path, "not a legacy VTK file");
  std::vector<std::string> w;
  std::string t;
  while (f >> t) w.push_back(t);
  auto up = [](std::string s) {
    for (auto& ch : s) ch = (char)std::toupper((unsigned char)ch);
    return s;
  };
  if (w.empty() || up(w[0]) != "ASCII") bad(path, "only ASCII VTK files are supported");
  size_t i = 1;
  int64_t n = 0;
  bool have_points = false;
  auto floats = [&](size_t count, std::vector<float>& dst) {
    if (i + count > w.size()) bad(path, "truncated data block");
    dst.resize(count);
    for (size_t k = 0; k < count; ++k) dst[k] = (float)std::strtod(w[i + k].c_str(), nullptr);
    i += count;
  };
  auto block = [&](const std::string& name, int span) {
    HostDesc d;
    d.label = name;
    d.span = span;
    floats((size_t)span * n, d.data);
    out.descs.push_back(std::move(d));
  };
  while (i < w.size()) {
    const std::string k = up(w[i]);
    if (k == "DATASET") {
      const std::string kind = i + 1 < w.size() ? up(w[i + 1]) : "";
      if (kind != "POLYDATA" && kind != "UNSTRUCTURED_GRID") bad(path, "unsupported DATASET " + kind);
      i += 2;
    } else if (k == "POINTS") {
      n = std::strtoll(w[i + 1].c_str(), nullptr, 10);
      i += 3;
      std::vector<float> xyz;
      floats((size_t)3 * n, xyz);
      out.n = n;
      out.features.assign((size_t)n * 4, 1.f);
      for (int64_t p = 0; p < n; ++p)
        for (int d = 0; d < 3; ++d) out.features[4 * p + d] = xyz[3 * p + d];
      have_points = true;
    } else if (k == "VERTICES" || k == "LINES" || k == "POLYGONS" || k == "CELLS") {
      i += 3 + (size_t)std::strtoll(w[i + 2].c_str(), nullptr, 10);
    } else if (k == "CELL_TYPES") {
      i += 2 + (size_t)std::strtoll(w[i + 1].c_str(), nullptr, 10);
    } else if (k == "POINT_DATA") {
      i += 2;
    } else if (k == "NORMALS" || k == "VECTORS") {
      const std::string name = w[i + 1];
      i += 3;
      block(name, 3);
    } else if (k == "TENSORS") {
      const std::string name = w[i + 1];
      i += 3;
      block(name, 9);
    } else if (k == "SCALARS") {
      const std::string name = w[i + 1];
      int span = 1;
      i += 3;
      if (i < w.size() && !w[i].empty() && std::all_of(w[i].begin(), w[i].end(), [](char ch) { return std::isdigit((unsigned char)ch); })) {
        span = std::atoi(w[i].c_str());
        ++i;
      }
      if (i < w.size() && up(w[i]) == "LOOKUP_TABLE") i += 2;
      block(name, span);
    } else if (k == "COLOR_SCALARS") {
      const std::string name = w[i + 1];
      const int span = std::atoi(w[i + 2].c_str());
      i += 3;
      block(name, span);
    } else {
      bad(path, "unsupported VTK keyword " + w[i]);
    }
  }
  if (!have_points) bad(path, "no POINTS block");
}

void save_vtk(const std::string& path, const HostCloud& c) {
  std::FILE* f = std::fopen(path.c_str(), "w");
  if (!f) bad(path, "cannot create");
  std::fprintf(f, "# vtk DataFile Version 3.0\nFile created by pgslam_b200\nASCII\nDATASET POLYDATA\n");
  std::fprintf(f, "POINTS %lld float\n", (long long)c.n);
  for (int64_t i = 0; i < c.n; ++i)
    std::fprintf(f, "%.9g %.9g %.9g\n", (double)c.features[4 * i], (double)c.features[4 * i + 1], (double)c.features[4 * i + 2]);
  std::fprintf(f, "VERTICES %lld %lld\n", (long long)c.n, (long long)(2 * c.n));
  for (int64_t i = 0; i < c.n; ++i) std::fprintf(f, "1 %lld\n", (long long)i);
  if (!c.descs.empty()) std::fprintf(f, "POINT_DATA %lld\n", (long long)c.n);
  for (auto& d : c.descs) {
    int per_line = d.span;
    if (d.label == "normals" && d.span == 3) std::fprintf(f, "NORMALS %s float\n", d.label.c_str());
    else if (d.span == 1) std::fprintf(f, "SCALARS %s float 1\nLOOKUP_TABLE default\n", d.label.c_str());
    else if (d.span == 3) std::fprintf(f, "VECTORS %s float\n", d.label.c_str());
    else if (d.span == 9) { std::fprintf(f, "TENSORS %s float\n", d.label.c_str()); per_line = 3; }
    else std::fprintf(f, "SCALARS %s float %d\nLOOKUP_TABLE default\n", d.label.c_str(), d.span);
    const size_t total = (size_t)c.n * d.span;
    for (size_t k = 0; k < total; ++k) std::fprintf(f, "%.9g%c", (double)d.data[k], (k + 1) % per_line == 0 ? '\n' : ' ');
  }
  std::fclose(f);
}

// ---- PLY ------------------------------------------------------------------------------------------
void load_ply(const std::string& path, HostCloud& out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) bad(path, "cannot open");
  std::string raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  const size_t end = raw.find("end_header");
  if (raw.compare(0, 3, "ply") != 0 || end == std::string::npos) bad(path, "not a PLY file");
  std::istringstream hs(raw.substr(0, end));
  const size_t body_at = raw.find('\n', end) + 1;
  std::string fmt, ln;
  int64_t n = 0;
  bool in_vertex = false;
  std::vector<std::pair<std::string, std::string>> props;  // (name, type)
  while (std::getline(hs, ln)) {
    const auto t = split_tokens(ln, " \t\r");
    if (t.empty()) continue;
    if (t[0] == "format" && t.size() > 1) fmt = t[1];
    else if (t[0] == "element" && t.size() > 2) {
      in_vertex = t[1] == "vertex";
      if (in_vertex) n = std::strtoll(t[2].c_str(), nullptr, 10);
    } else if (t[0] == "property" && in_vertex && t.size() > 2) {
      if (t[1] == "list") bad(path, "list properties on vertices are not supported");
      props.emplace_back(t[2], t[1]);
    }
  }
  std::vector<std::string> names;
  for (auto& p : props) names.push_back(p.first);
  std::vector<std::vector<float>> cols(props.size(), std::vector<float>((size_t)n));
  if (fmt == "ascii") {
    std::istringstream bs(raw.substr(body_at));
    for (int64_t i = 0; i < n; ++i)
      for (size_t c = 0; c < props.size(); ++c) {
        double v;
        if (!(bs >> v)) bad(path, "truncated vertex data");
        cols[c][i] = (float)v;
      }
  } else if (fmt == "binary_little_endian" || fmt == "binary_big_endian") {
    const bool big = fmt == "binary_big_endian";
    static const std::map<std::string, std::pair<int, char>> types = {
        {"char", {1, 'i'}}, {"uchar", {1, 'u'}}, {"short", {2, 'i'}}, {"ushort", {2, 'u'}}, {"int", {4, 'i'}},
        {"uint", {4, 'u'}}, {"float", {4, 'f'}}, {"double", {8, 'f'}}, {"int8", {1, 'i'}}, {"uint8", {1, 'u'}},
        {"int16", {2, 'i'}}, {"uint16", {2, 'u'}}, {"int32", {4, 'i'}}, {"uint32", {4, 'u'}}, {"float32", {4, 'f'}},
        {"float64", {8, 'f'}}};
    size_t at = body_at;
    for (int64_t i = 0; i < n; ++i)
      for (size_t c = 0; c < props.size(); ++c) {
        auto it = types.find(props[c].second);
        if (it == types.end()) bad(path, "unknown PLY property type " + props[c].second);
        const int sz = it->second.first;
        if (at + sz > raw.size()) bad(path, "truncated vertex data");
        unsigned char b[8];
        for (int k = 0; k < sz; ++k) b[k] = (unsigned char)raw[at + (big ? sz - 1 - k : k)];  // -> little endian
        at += sz;
        double v = 0;
        if (it->second.second == 'f') {
          if (sz == 4) { float x; std::memcpy(&x, b, 4); v = x; }
          else { double x; std::memcpy(&x, b, 8); v = x; }
        } else {
          unsigned long long u = 0;
          for (int k = sz - 1; k >= 0; --k) u = (u << 8) | b[k];
          if (it->second.second == 'i') {
            const unsigned long long sign = 1ull << (8 * sz - 1);
            v = (u & sign) ? -(double)((~u + 1) & ((sign << 1) - 1)) : (double)u;
          } else {
            v = (double)u;
          }
        }
        cols[c][i] = (float)v;
      }
  } else {
    bad(path, "unknown PLY format " + fmt);
  }
  columns_to_cloud(path, names, cols, out);
}

void save_ply(const std::string& path, const HostCloud& c) {
  std::FILE* f = std::fopen(path.c_str(), "w");
  if (!f) bad(path, "cannot create");
  const auto cols = cloud_columns(c);
  std::fprintf(f, "ply\nformat ascii 1.0\ncomment created by pgslam_b200\nelement vertex %lld\n", (long long)c.n);
  for (auto& col : cols) std::fprintf(f, "property float %s\n", col.first.c_str());
  std::fprintf(f, "end_header\n");
  write_rows(f, cols, c.n, ' ');
  std::fclose(f);
}

}  // namespace

void load_cloud_file(const std::string& path, HostCloud& out) {
  out = HostCloud();
  const std::string e = lower_ext(path);
  if (e == ".csv") load_csv(path, out);
  else if (e == ".vtk") load_vtk(path, out);
  else if (e == ".ply") load_ply(path, out);
  else bad(path, "unknown point-cloud file extension '" + e + "' (csv, vtk, ply)");
}

void save_cloud_file(const std::string& path, const HostCloud& c) {
  const std::string e = lower_ext(path);
  if (e == ".csv") save_csv(path, c);
  else if (e == ".vtk") save_vtk(path, c);
  else if (e == ".ply") save_ply(path, c);
  else bad(path, "unknown point-cloud file extension '" + e + "' (csv, vtk, ply)");
}

}  // namespace pgs
