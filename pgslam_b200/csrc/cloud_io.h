// cloud_io.h — host-side parse / write of libpointmatcher's point-cloud text formats (cloud_io.cpp).
#pragma once

#include <string>
#include <vector>

#include "common.cuh"

namespace pgs {

struct HostDesc {
  std::string label;
  int span = 0;
  std::vector<float> data;  // point-major: span floats per point
};
struct HostCloud {
  int64_t n = 0;
  std::vector<float> features;  // n x {x, y, z, 1}
  std::vector<HostDesc> descs;
};

// DataPoints::load / save: dispatch on the extension (.csv, .vtk, .ply); throws pgs::Error
void load_cloud_file(const std::string& path, HostCloud& out);
void save_cloud_file(const std::string& path, const HostCloud& c);

}  // namespace pgs
