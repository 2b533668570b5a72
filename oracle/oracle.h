/*
 * oracle.h — CPU restatement of the scan-registration hot path pgslam drives.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pgslam_b200/ may include, link or
 * call this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs use it, and only as the checker / CPU baseline.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in libpointmatcher /
 * libnabo / Eigen, which pgslam pulls in un-vendored and un-pinned
 * (/root/reference/CMakeLists.txt:19-20) and which are absent from this
 * container; the reference's only test (tests/instantiation.cpp:4-19) pins no
 * number.  This file restates the *published* algorithms
 * along pgslam's own call order:
 *     Localizer.hpp:91-135, 282-348   LoopCloser.hpp:83-110, 343-365
 *     LocalMap.hpp:95-98, 209-224     types.h:19-29
 * and is pinned only by its own cross-checks (brute-force kNN twin, scipy
 * cKDTree, numpy.linalg, analytic ICP cases) — see tests/test_oracle_*.py.
 *
 * Numeric contract (the normative one for this repo, T = float):
 *   - clouds are float; squared distance is ((dx*dx)+dy*dy)+dz*dz in fp32,
 *     no FMA; kNN result is the exact lexicographic (dist, index) minimum.
 *   - rigid transform of points is (((r0*x)+r1*y)+r2*z)+t in fp32, no FMA.
 *   - every reduction / solve / pose composition is fp64.
 */
#ifndef PGSLAM_ORACLE_H
#define PGSLAM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes: 1:1 with libpointmatcher's exception types ---------- */
enum {
  ORC_OK = 0,
  ORC_CONVERGENCE_ERROR = 1,   /* PM::ConvergenceError */
  ORC_TRANSFORMATION_ERROR = 2,/* TransformationError  */
  ORC_INVALID_PARAMETER = 3,
  ORC_INVALID_FIELD = 4
};

/* ---- DataPoints (types.h:20): features 4xN col-major + known descriptors */
typedef struct {
  int64_t n;
  float *feat;     /* 4*n : x,y,z,1 per point                              */
  float *normals;  /* 3*n or NULL                                          */
  float *obsdir;   /* 3*n or NULL  ("observationDirections")               */
  float *noise;    /* n   or NULL  ("simpleSensorNoise")                   */
  float *dens;     /* n   or NULL  ("densities")                           */
  float *eigval;   /* 3*n or NULL  ("eigValues")                           */
  float *eigvec;   /* 9*n or NULL  ("eigVectors", column-major 3x3)        */
  float *meandist; /* n   or NULL  ("meanDists")                           */
  float *matched;  /* matched_span*n or NULL ("matchedIds", ids as floats) */
  int matched_span;
} orc_cloud;

orc_cloud *orc_cloud_new(int64_t n);
orc_cloud *orc_cloud_copy(const orc_cloud *c);
void orc_cloud_free(orc_cloud *c);
/* DP::concatenate (LocalMap.hpp:222): keeps descriptors present in both.  */
void orc_cloud_concatenate(orc_cloud *a, const orc_cloud *b);

/* ---- libnabo kd-tree (A.1/A.2), bucket size 8, sliding midpoint -------- */
typedef struct orc_kdtree orc_kdtree;
orc_kdtree *orc_kdtree_build(const float *feat4, int64_t n);
void orc_kdtree_free(orc_kdtree *t);
/* ids/d2 are k x nq col-major (k contiguous per query).  max_dist = +inf
 * for unbounded.  Unfound: id -1, dist +inf.  Returns total leaf visits.   */
uint64_t orc_kdtree_knn(const orc_kdtree *t, const float *query4, int64_t nq,
                        int k, float max_dist, int allow_self, int32_t *ids,
                        float *d2);
/* O(N*M) twin: exact fp32 argmin, ties -> lower index.                     */
/* Search semantics.  CONTRACT (default): exact, ties -> lower index, far-side bound recomputed
 * (what the CUDA path implements; epsilon is ignored = 0).  NABO: libnabo's own rules, verbatim
 * from SURVEY.md Appendix A.2 - strict '<', first-visited ties, incremental rd, (1+eps)^2
 * pruning - used ONLY to measure how far the contract is from the real library
 * (oracle/README.md).  The mode set here applies to the matcher inside orc_icp_run /
 * orc_icp_seq_run / the probes (with the chain's `epsilon`); filters keep the contract.    */
enum { ORC_SEARCH_CONTRACT = 0, ORC_SEARCH_NABO = 1 };
void orc_set_search_mode(int mode);
int orc_search_mode(void);
uint64_t orc_kdtree_knn_ex(const orc_kdtree *t, const float *query4, int64_t nq, int k, float max_dist,
                           int allow_self, float epsilon, int mode, int32_t *ids, float *d2);
void orc_knn_brute(const float *ref4, int64_t n, const float *query4,
                   int64_t nq, int k, float max_dist, int32_t *ids, float *d2);

/* ---- filters (A.9) ------------------------------------------------------ */
enum {
  ORC_F_RANDOM_SAMPLING = 1, /* p0 = prob, i0 = seed                        */
  ORC_F_VOXEL_GRID = 2,      /* p0,p1,p2 = vSize, i0 = useCentroid, i1 = avgDesc */
  ORC_F_SURFACE_NORMAL = 3,  /* i0 = knn, p0 = maxDist, i1 flags (bit0 normals,
                                bit1 densities, bit2 eigValues, bit3 eigVectors,
                                bit4 matchedIds, bit5 meanDists, bit6 sortEigen) */
  ORC_F_OBSERVATION_DIRECTION = 4, /* p0,p1,p2 = sensor position            */
  ORC_F_ORIENT_NORMALS = 5,  /* i0 = towardCenter                           */
  ORC_F_SIMPLE_SENSOR_NOISE = 6, /* i0 = sensorType, p0 = gain              */
  ORC_F_MAX_DIST = 7,        /* i0 = dim (-1 radial), p0 = maxDist          */
  ORC_F_MIN_DIST = 8,        /* i0 = dim (-1 radial), p0 = minDist          */
  ORC_F_BOUNDING_BOX = 9,    /* box[6] = xMin,xMax,yMin,yMax,zMin,zMax; i0 = removeInside */
  ORC_F_MAX_DENSITY = 10,    /* p0 = maxDensity, i0 = seed; needs `densities` */
  ORC_F_SAMPLING_SURFACE_NORMAL = 11, /* p0 = ratio, p1 = maxBoxDim, p2 = seed, i0 = knn,
                                i1 flags (bit0 normals, bit1 densities, bit2 eigValues,
                                bit3 eigVectors, bit4 samplingMethod = bin,
                                bit5 averageExistingDescriptors)              */
  ORC_F_REMOVE_NAN = 12,      /* drops every point with a NaN coordinate      */
  ORC_F_FIX_STEP_SAMPLING = 13, /* i0 = step, i1 = seed (phase = hash % step) */
  ORC_F_SHADOW = 14,         /* p0 = eps; needs `normals`                    */
  ORC_F_IDENTITY = 15
};
typedef struct {
  int type;
  double p0, p1, p2;
  int64_t i0, i1;
  double box[6];
} orc_filter;

int orc_filter_apply(const orc_filter *f, orc_cloud *c);
int orc_filters_apply(const orc_filter *f, int nf, orc_cloud *c);

/* RigidTransformation::compute (A.9): T is 4x4 col-major double.          */
int orc_rigid_transform(orc_cloud *c, const double *T);

/* ---- outlier filters (A.3) --------------------------------------------- */
enum {
  ORC_O_TRIMMED_DIST = 1, /* p0 = ratio   */
  ORC_O_MAX_DIST = 2,     /* p0 = maxDist */
  ORC_O_MIN_DIST = 3,     /* p0 = minDist */
  ORC_O_MEDIAN_DIST = 4,  /* p0 = factor  */
  ORC_O_SURFACE_NORMAL = 5, /* p0 = maxAngle: needs `normals` on both clouds */
  ORC_O_VAR_TRIMMED_DIST = 6 /* p0 = minRatio, p1 = maxRatio, p2 = lambda   */
};
typedef struct {
  int type;
  double p0, p1, p2;
} orc_outlier;
/* VarTrimmedDistOutlierFilter::optimizeInlierRatio: the ratio minimising the
 * FRMS criterion over the sorted valid distances (as a float).             */
int orc_var_trimmed_ratio(const float *d2, int64_t nk, double min_ratio, double max_ratio,
                          double lambda, float *ratio);
/* weights k x n; returns ORC_CONVERGENCE_ERROR on "no outlier to filter".  */
int orc_outlier_weights(const orc_outlier *o, int no, const float *d2,
                        int64_t nk, float *w);
int orc_dists_quantile(const float *d2, int64_t nk, double q, float *out);
/* same, for filters that also look at the clouds (SurfaceNormalOutlierFilter) */
int orc_outlier_weights_full(const orc_outlier *o, int no, const orc_cloud *reading,
                             const orc_cloud *reference, const int32_t *ids,
                             const float *d2, int k, float *w);

/* ---- error minimizers (A.4 – A.7) -------------------------------------- */
enum {
  ORC_E_POINT_TO_PLANE = 1,
  ORC_E_POINT_TO_PLANE_WITH_COV = 2,
  ORC_E_POINT_TO_POINT = 3
};
typedef struct {
  double T[16];        /* incremental transform, col-major                  */
  double cov[36];      /* WithCov only, col-major, order x,y,z,rx,ry,rz     */
  double A[36], b[6];  /* normal equations (point-to-plane)                 */
  double point_used_ratio, weighted_point_used_ratio;
  double residual;     /* getResidualError on these elements                */
  int64_t kept;
} orc_min_out;
int orc_minimize(int type, double sensor_std_dev, const orc_cloud *reading,
                 const orc_cloud *reference, const int32_t *ids,
                 const float *d2, const float *w, int k, orc_min_out *out);
/* PointToPlaneErrorMinimizer{force2D, force4DOF}                            */
enum { ORC_FORCE_NONE = 0, ORC_FORCE_2D = 1, ORC_FORCE_4DOF = 2 };
int orc_minimize_ex(int type, int force_mode, double sensor_std_dev, const orc_cloud *reading,
                    const orc_cloud *reference, const int32_t *ids, const float *d2,
                    const float *w, int k, orc_min_out *out);
/* getOverlap() on the last error elements (A.5).                           */
double orc_overlap(int type, const orc_cloud *reading, const orc_cloud *reference,
                   const int32_t *ids, const float *d2, const float *w, int k);

/* ---- small dense algebra, exposed for tests ---------------------------- */
void orc_eig3_sym(const double *A /*9 col-major*/, double *w /*3*/, double *V /*9*/);
int orc_solve6(const double *A /*36*/, const double *b /*6*/, double *x /*6*/);
void orc_svd3(const double *M /*9*/, double *U, double *S, double *V);

/* ---- ICP chain (A.8, §3.3) ---------------------------------------------- */
#define ORC_MAX_MODS 8
typedef struct {
  orc_filter reading_filters[ORC_MAX_MODS];       int n_reading_filters;
  orc_filter reading_step_filters[ORC_MAX_MODS];  int n_reading_step_filters;
  orc_filter reference_filters[ORC_MAX_MODS];     int n_reference_filters;
  int knn; double epsilon; double max_dist;       /* KDTreeMatcher          */
  orc_outlier outliers[ORC_MAX_MODS];             int n_outliers;
  int minimizer; double sensor_std_dev;
  int max_iterations;                             /* Counter; <=0: absent   */
  int has_differential; double min_diff_rot, min_diff_trans; int smooth_length;
  int has_bound; double max_rot_norm, max_trans_norm;
  int force_mode;                                 /* ORC_FORCE_*            */
} orc_icp_config;
void orc_icp_config_default(orc_icp_config *cfg);

typedef struct {
  double T[16];        /* T_refIn_dataIn result, col-major                  */
  double cov[36];
  int iterations;
  int max_iter_reached;
  int status;
  double overlap, weighted_ratio, point_used_ratio, residual;
  double last_T_iter[16];
  double time_filters_s, time_index_s, time_loop_s, time_knn_s;
  uint64_t visits;
} orc_icp_result;

/* ICP::operator()(reading, reference, T_init)  (LoopCloser.hpp:98)         */
int orc_icp_run(const orc_icp_config *cfg, const orc_cloud *reading,
                const orc_cloud *reference, const double *T_init,
                orc_icp_result *res);

/* ICPSequence (Localizer.hpp:126,148)                                      */
typedef struct orc_icp_seq orc_icp_seq;
orc_icp_seq *orc_icp_seq_new(const orc_icp_config *cfg);
void orc_icp_seq_free(orc_icp_seq *s);
int orc_icp_seq_set_map(orc_icp_seq *s, const orc_cloud *map);
int orc_icp_seq_run(orc_icp_seq *s, const orc_cloud *reading,
                    const double *T_init, orc_icp_result *res);
const orc_cloud *orc_icp_seq_map(const orc_icp_seq *s);

/* The two probes pgslam hand-rolls from the modules:                       */
/* Localizer::ComputeOverlapWith (Localizer.hpp:282-348)                    */
int orc_probe_overlap(const orc_icp_config *cfg, const orc_cloud *reading,
                      const orc_cloud *reference, const double *T_world_robot,
                      double *weighted_ratio);
/* LoopCloser::ComputeResidualError (LoopCloser.hpp:343-365)                */
int orc_probe_residual(const orc_icp_config *cfg, const orc_cloud *reading,
                       const orc_cloud *reference, const double *T,
                       double *residual);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
