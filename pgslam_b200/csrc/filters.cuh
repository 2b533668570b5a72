// filters.cuh — DataPointsFilters / RigidTransformation entry points (filters.cu)
#pragma once

#include "core.cuh"
#include "modules.h"

namespace pgs {

bool is_rigid(const double* T);          // |1 - det(R)| <= 0.001 (A.9, eps to verify upstream)
Xf xf_from_T(const double* T);           // fp32 3x4 of a col-major double 4x4
// RigidTransformation::compute, in place; throws PGS_TRANSFORMATION_ERROR
void rigid_transform_cloud(Cloud& c, const double* T);
// DataPointsFilter::inPlaceFilter on every cloud of the list
void apply_filter(Ctx* ctx, const Module& m, std::vector<Cloud*>& clouds);
void apply_filters(Ctx* ctx, const std::vector<Module>& ms, std::vector<Cloud*>& clouds);

}  // namespace pgs
