/*
 * pgslam_b200.h — C ABI of libpgslam_b200.so: the B200-native scan-registration
 * hot path behind libpointmatcher's plugin surface, as pgslam binds to it.
 *
 * Drop-in seam: /root/reference/src/pgslam/types.h:19-27 (PM, DP, ICP,
 * ICPSequence, TransformationPtr, DataPointsFilters).  Every entry point below
 * names the reference interface it replaces; the C++ adapter in
 * include/pgslam_b200/pm_adapter.hpp re-creates those class shapes on top of
 * this ABI.
 *
 * Conventions
 *   - plain pointers and sizes only; matrices are COLUMN-MAJOR (Eigen default),
 *     so `features` is an array of N {x,y,z,1} float quadruples (PM layout);
 *   - T = float clouds; poses cross the ABI as double[16] (exact for floats);
 *   - no exceptions cross the ABI: every call returns a pgs_status; the text of
 *     the last failure is pgs_last_error(ctx);
 *   - `on_device == 1` means the pointer is a CUDA device pointer valid on the
 *     context's device; the call is then stream-ordered on the context stream.
 *     For INPUT buffers `on_device == 2` means PINNED host memory copied
 *     asynchronously on a side stream (overlaps with queued kernels); the caller
 *     keeps the buffer alive and unchanged until pgs_ctx_synchronize() or the
 *     next call that returns results to the host;
 *   - a handle is not re-entrant; distinct contexts may be driven from distinct
 *     host threads concurrently (LocalizerMT.hpp:47, LoopCloserMT.hpp:41).
 */
#ifndef PGSLAM_B200_H
#define PGSLAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One code per libpointmatcher exception type (pgslam catches none of them:
 * Localizer.hpp:64, LoopCloser.hpp:67 only throw on config read failure).   */
typedef enum {
  PGS_OK = 0,
  PGS_CONVERGENCE_ERROR = 1,    /* PM::ConvergenceError                      */
  PGS_TRANSFORMATION_ERROR = 2, /* TransformationError (non-rigid matrix)    */
  PGS_INVALID_PARAMETER = 3,    /* Parametrizable::InvalidParameter          */
  PGS_INVALID_FIELD = 4,        /* DataPoints::InvalidField                  */
  PGS_INVALID_MODULE_TYPE = 5,  /* PointMatcherSupport::InvalidModuleType    */
  PGS_INVALID_ELEMENT = 6,      /* Registrar: unknown module name            */
  PGS_CUDA_ERROR = 7,
  PGS_INVALID_ARGUMENT = 8
} pgs_status;

typedef struct pgs_ctx pgs_ctx;
typedef struct pgs_cloud pgs_cloud;       /* PM::DataPoints          types.h:20 */
typedef struct pgs_filters pgs_filters;   /* PM::DataPointsFilters   types.h:27 */
typedef struct pgs_matcher pgs_matcher;   /* PM::Matcher (KDTreeMatcher)        */
typedef struct pgs_outliers pgs_outliers; /* PM::OutlierFilters                 */
typedef struct pgs_minimizer pgs_minimizer; /* PM::ErrorMinimizer               */
typedef struct pgs_icp pgs_icp;           /* PM::ICP / PM::ICPSequence types.h:24-25 */

/* ---- context ------------------------------------------------------------ */
/* `stream` may be NULL (the library creates its own non-blocking stream) or a
 * cudaStream_t the caller owns (e.g. torch's current stream).               */
pgs_status pgs_ctx_create(int device, void *stream, pgs_ctx **out);
void pgs_ctx_destroy(pgs_ctx *ctx);
const char *pgs_last_error(const pgs_ctx *ctx);
pgs_status pgs_ctx_synchronize(pgs_ctx *ctx);
const char *pgs_version(void);

/* ---- DataPoints (types.h:20; LocalMap.hpp:85,214,222) -------------------- */
pgs_status pgs_cloud_create(pgs_ctx *ctx, const float *features4xN, int64_t n,
                            int on_device, pgs_cloud **out);
/* descriptor block `label` of `span` rows: span x N column-major            */
pgs_status pgs_cloud_set_descriptor(pgs_cloud *c, const char *label, int span,
                                    const float *data, int on_device);
pgs_status pgs_cloud_remove_descriptor(pgs_cloud *c, const char *label);
int64_t pgs_cloud_num_points(const pgs_cloud *c);
int pgs_cloud_num_descriptors(const pgs_cloud *c);
/* label is written NUL-terminated into `label` (cap bytes)                  */
pgs_status pgs_cloud_descriptor_info(const pgs_cloud *c, int index, char *label,
                                     int cap, int *span);
pgs_status pgs_cloud_get_features(const pgs_cloud *c, float *out4xN, int on_device);
pgs_status pgs_cloud_get_descriptor(const pgs_cloud *c, const char *label,
                                    float *out, int on_device);
pgs_status pgs_cloud_copy(const pgs_cloud *c, pgs_cloud **out);  /* DP copy-ctor */
/* DP::concatenate (LocalMap.hpp:222): keeps descriptors common to both      */
pgs_status pgs_cloud_concatenate(pgs_cloud *a, const pgs_cloud *b);
void pgs_cloud_destroy(pgs_cloud *c);
/* DataPoints::load / DataPoints::save (static members of PM::DataPoints): libpointmatcher's
 * .csv, legacy ASCII .vtk and .ply (ascii or binary on load) files, by extension.  Columns
 * x y z -> features; nx ny nz -> "normals"; name_x name_y name_z -> span-3 descriptor `name`.
 * pgs_cloud_file_info parses a file on the host only (no device): point and descriptor counts. */
pgs_status pgs_cloud_load(pgs_ctx *ctx, const char *path, pgs_cloud **out);
pgs_status pgs_cloud_save(const pgs_cloud *c, const char *path);
pgs_status pgs_cloud_file_info(const char *path, int64_t *n_points, int *n_descriptors, char *err, int cap);

/* ---- Transformation (Localizer.hpp:20,106; LocalMap.hpp:97,222) ---------- */
/* PM::get().REG(Transformation).create("RigidTransformation")->compute():
 * in-place; PGS_TRANSFORMATION_ERROR if T is not rigid.                     */
pgs_status pgs_rigid_transform(pgs_cloud *c, const double T[16]);
/* LocalMap::BuildCloudFromData (LocalMap.hpp:209-224): out = clouds[0] ++
 * T[1]*clouds[1] ++ ... ; T is n x 16 doubles (T[0] ignored).               */
pgs_status pgs_cloud_assemble(pgs_ctx *ctx, int n, const pgs_cloud *const *clouds,
                              const double *T, pgs_cloud **out);

/* ---- DataPointsFilters (Localizer.hpp:77,103) ---------------------------- */
/* DataPointsFilters(std::istream&): YAML top-level list of modules.         */
pgs_status pgs_filters_create_from_yaml(pgs_ctx *ctx, const char *yaml, size_t len,
                                        pgs_filters **out);
/* Registrar path: REG(DataPointsFilter).create(name, params); params are
 * 2*nkv strings key0,value0,key1,value1,...                                 */
pgs_status pgs_filters_create(pgs_ctx *ctx, pgs_filters **out);
pgs_status pgs_filters_append(pgs_filters *f, const char *name,
                              const char *const *kv, int nkv);
int pgs_filters_count(const pgs_filters *f);
pgs_status pgs_filters_apply(pgs_filters *f, pgs_cloud *c); /* init() + apply() */
void pgs_filters_destroy(pgs_filters *f);

/* ---- Matcher (Localizer.hpp:317,328; LoopCloser.hpp:356,358) ------------- */
pgs_status pgs_matcher_create(pgs_ctx *ctx, const char *name,
                              const char *const *kv, int nkv, pgs_matcher **out);
pgs_status pgs_matcher_init(pgs_matcher *m, const pgs_cloud *reference);
int pgs_matcher_knn(const pgs_matcher *m);
/* Matches{ids,dists}: k x N column-major; dists are SQUARED; unfound -> id -1,
 * dist +inf; exact (eps = 0), ties -> lower index.                          */
pgs_status pgs_matcher_find(pgs_matcher *m, const pgs_cloud *reading,
                            int32_t *ids, float *dists2, int on_device);
/* Queries of the last pgs_matcher_find that the tensor-core distance-tile path
 * (pgs_ctx_set_option "dense_max_ref") could not certify and answered with the
 * exact fallback scan; needs option "dense_count_fallbacks" = 1.               */
unsigned pgs_matcher_dense_fallbacks(const pgs_matcher *m);
void pgs_matcher_destroy(pgs_matcher *m);

/* ---- OutlierFilters (Localizer.hpp:330; LoopCloser.hpp:360) -------------- */
pgs_status pgs_outliers_create(pgs_ctx *ctx, pgs_outliers **out);
pgs_status pgs_outliers_append(pgs_outliers *o, const char *name,
                               const char *const *kv, int nkv);
/* OutlierFilters::compute -> OutlierWeights k x N                           */
pgs_status pgs_outliers_compute(pgs_outliers *o, const pgs_cloud *reading,
                                const pgs_cloud *reference, const int32_t *ids,
                                const float *dists2, int k, float *weights,
                                int on_device);
void pgs_outliers_destroy(pgs_outliers *o);

/* ---- ErrorMinimizer (Localizer.hpp:238,278,332; LoopCloser.hpp:108,331,362) */
typedef struct {
  double T[16];                    /* TransformationParameters, col-major    */
  double covariance[36];           /* getCovariance(): x,y,z,rx,ry,rz        */
  double point_used_ratio;         /* ErrorElements::pointUsedRatio          */
  double weighted_point_used_ratio;/* ErrorElements::weightedPointUsedRatio  */
  double residual;                 /* getResidualError(...)                  */
  double overlap;                  /* getOverlap()                           */
  int64_t kept;                    /* ErrorElements: nb of kept pairs        */
} pgs_min_result;
pgs_status pgs_minimizer_create(pgs_ctx *ctx, const char *name,
                                const char *const *kv, int nkv, pgs_minimizer **out);
/* ErrorElements(reading, reference, weights, matches) + compute()           */
pgs_status pgs_minimizer_compute(pgs_minimizer *e, const pgs_cloud *reading,
                                 const pgs_cloud *reference, const int32_t *ids,
                                 const float *dists2, const float *weights, int k,
                                 int on_device, pgs_min_result *out);
void pgs_minimizer_destroy(pgs_minimizer *e);

/* ---- ICP / ICPSequence (types.h:24-25) ----------------------------------- */
typedef struct {
  double T[16];            /* T_refIn_dataIn, column-major                   */
  double covariance[36];   /* errorMinimizer->getCovariance()                */
  int32_t iterations;
  int32_t max_iterations_reached; /* ICP::getMaxNumIterationsReached() (LoopCloser.hpp:317) */
  int32_t status;          /* pgs_status of this pair                        */
  int32_t reserved;
  double overlap;          /* errorMinimizer->getOverlap() (Localizer.hpp:278) */
  double weighted_point_used_ratio;
  double point_used_ratio;
  double residual;
  int64_t n_reading;       /* reading points after reading filters           */
  int64_t n_reference;     /* reference points after reference filters       */
} pgs_icp_result;

/* ICP::loadFromYaml (Localizer.hpp:70,311; LoopCloser.hpp:73,348)            */
pgs_status pgs_icp_create_from_yaml(pgs_ctx *ctx, const char *yaml, size_t len,
                                    pgs_icp **out);
/* ICPChainBase::setDefault                                                   */
pgs_status pgs_icp_create_default(pgs_ctx *ctx, pgs_icp **out);
void pgs_icp_destroy(pgs_icp *icp);
/* public members of ICPChainBase that pgslam touches (Localizer.hpp:313-330,
 * LoopCloser.hpp:352-362); borrowed handles, owned by the icp object.        */
pgs_filters *pgs_icp_reading_filters(pgs_icp *icp);
pgs_filters *pgs_icp_reading_step_filters(pgs_icp *icp);
pgs_filters *pgs_icp_reference_filters(pgs_icp *icp);
pgs_matcher *pgs_icp_matcher(pgs_icp *icp);
pgs_outliers *pgs_icp_outliers(pgs_icp *icp);
pgs_minimizer *pgs_icp_minimizer(pgs_icp *icp);
/* ICP::operator()(reading, reference, T_init)            (LoopCloser.hpp:98) */
pgs_status pgs_icp_run(pgs_icp *icp, const pgs_cloud *reading,
                       const pgs_cloud *reference, const double T_init[16],
                       pgs_icp_result *out);
/* ICPSequence::setMap / hasMap / operator()(reading, T_init)
 * (Localizer.hpp:126,148,168,254)                                            */
pgs_status pgs_icp_set_map(pgs_icp *icp, const pgs_cloud *map);
int pgs_icp_has_map(const pgs_icp *icp);
pgs_status pgs_icp_run_sequence(pgs_icp *icp, const pgs_cloud *reading,
                                const double T_init[16], pgs_icp_result *out);
/* Batched loop-closure verification: P independent (reading, reference)
 * pairs run concurrently on the context's GPU.  T_inits is
 * P x 16 doubles or NULL (identity).  Returns the first non-OK pair status
 * only if ALL pairs failed; per-pair status is in results[i].status.        */
pgs_status pgs_icp_run_batch(pgs_icp *icp, int n_pairs,
                             const pgs_cloud *const *readings,
                             const pgs_cloud *const *references,
                             const double *T_inits, pgs_icp_result *results);
/* The same for HOST-resident clouds on one or several GPUs of this process:
 * the candidate loop of LoopCloser::AddNewVertex (LoopCloser.hpp:266-297, and
 * the worker thread of LoopCloserMT.hpp:45-67) submits all its candidates at
 * once.  icps[d] are ICP handles with the SAME chain, one per context to use
 * (normally one context per GPU; two contexts of one GPU are allowed).  Pairs
 * are cut into contiguous blocks, pair i -> icps[i / ceil(n_pairs/n_devices)]
 * (the partition of pgslam_b200/dist.py), every device streams its block
 * through chunked uploads that overlap the registrations, and results[i] is
 * bit-identical to what pgs_icp_run gives for pair i on any of the devices.
 * pinned != 0 promises that every buffer is page-locked (cudaHostAlloc /
 * cudaHostRegister): uploads are then asynchronous.  No NCCL: one process.   */
typedef struct {
  const float *features4xN;        /* PM features, column-major 4 x N          */
  int64_t n;
  int n_descriptors;               /* optional descriptor blocks, span x N     */
  const char *const *labels;
  const int *spans;
  const float *const *data;
} pgs_host_cloud;
pgs_status pgs_icp_run_batch_multi(pgs_icp *const *icps, int n_devices, int n_pairs,
                                   const pgs_host_cloud *readings,
                                   const pgs_host_cloud *references,
                                   const double *T_inits, int pinned,
                                   pgs_icp_result *results);
/* Localizer::ComputeOverlapWith (Localizer.hpp:282-348) as one fused call.   */
pgs_status pgs_icp_probe_overlap(pgs_icp *icp, const pgs_cloud *reading,
                                 const pgs_cloud *reference,
                                 const double T_world_robot[16],
                                 double *weighted_point_used_ratio);
/* LoopCloser::ComputeResidualError (LoopCloser.hpp:343-365) as one call.     */
pgs_status pgs_icp_probe_residual(pgs_icp *icp, const pgs_cloud *reading,
                                  const pgs_cloud *reference, const double T[16],
                                  double *residual);

/* ---- configuration check (host only, no device needed) ---------------------- */
/* Parses and validates a libpointmatcher YAML exactly as the loaders above do
 * (ICPChainBase::loadFromYaml when is_chain != 0, DataPointsFilters(istream&)
 * otherwise) without touching the GPU.  On failure the message is written
 * NUL-terminated into err (cap bytes).  On success *n_modules receives the
 * number of modules the configuration instantiates.                          */
pgs_status pgs_config_check(const char *yaml, size_t len, int is_chain,
                            int *n_modules, char *err, int cap);
/* Where a VALID configuration still departs from what libpointmatcher would do with it, in
 * words (one line per item, NUL-terminated into `out`; returns the number of items):
 * stochastic filters in readingStepDataPointsFilters (drawn once per registration here, once per
 * iteration upstream), `epsilon` accepted under PGS_EPSILON_POLICY=exact, seeds of the
 * counter-based generator.  Host only.                                                         */
int pgs_config_warnings(const char *yaml, size_t len, int is_chain, char *out, int cap);
/* Registrar introspection: number of registered modules of a kind
 * (0 DataPointsFilter, 1 Matcher, 2 OutlierFilter, 3 ErrorMinimizer,
 * 4 TransformationChecker, 5 Inspector, 6 Logger, 7 Transformation) and the
 * name of the i-th one.                                                       */
int pgs_registrar_count(int kind);
const char *pgs_registrar_name(int kind, int index);
/* Parametrizable::availableParameters(): number of documented parameters of a
 * registered module (-1: unknown module) and the i-th one; every returned
 * string is static ("" = unbounded).  Types: 'i' integer, 'u' bool/unsigned,
 * 'f' real, 's' string.                                                       */
int pgs_registrar_param_count(int kind, const char *name);
pgs_status pgs_registrar_param(int kind, const char *name, int index, const char **key,
                               const char **doc, const char **default_value,
                               const char **min_value, const char **max_value, char *type);
/* REG(kind).create(name, params) without a device: validates the name and the
 * parameters exactly as the creators above do (TransformationChecker,
 * Inspector, Logger and Transformation modules have no device object).       */
pgs_status pgs_module_validate(int kind, const char *name, const char *const *kv, int nkv,
                               char *err, int cap);

/* ---- instrumentation ------------------------------------------------------ */
/* Number of kernels this context launched since creation (bench.py's
 * gpu_launches) and device milliseconds of the last ICP stage breakdown.     */
uint64_t pgs_ctx_launch_count(const pgs_ctx *ctx);
typedef struct {
  float filters_ms, index_ms, loop_ms, total_ms;
  float match_ms, select_ms, accumulate_ms; /* filled when profiling is on   */
  int32_t iterations_launched;
} pgs_stage_times;
pgs_status pgs_ctx_set_profiling(pgs_ctx *ctx, int enabled);
/* pgs_icp_run_batch splits a batch of independent pairs over this many worker
 * streams / host threads of the context (default 4; 1 = everything on the
 * context's own stream).  Results are bit-identical whatever the split.       */
pgs_status pgs_ctx_set_batch_streams(pgs_ctx *ctx, int n_streams);
pgs_status pgs_ctx_last_stage_times(const pgs_ctx *ctx, pgs_stage_times *out);
/* Scheduling knobs of the hot kernels (never change a result): "match_mode"
 * 0..4, "pm_blocks", "pm_refill", "pm_pair_w", "pm_leaf_w", "mq_batches",
 * "mq_blocks", "resort_it", "batch_chunk", "dense_max_ref",
 * "dense_count_fallbacks" (DESIGN.md §6).                                      */
pgs_status pgs_ctx_set_option(pgs_ctx *ctx, const char *key, double value);

#ifdef __cplusplus
}
#endif
#endif /* PGSLAM_B200_H */
