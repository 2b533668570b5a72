// api.cu — the extern "C" boundary declared in include/pgslam_b200.h.
// Each function names, in the header, the pgslam call site it serves.
#include <cstring>
#include <mutex>
#include <thread>

#include "cloud_io.h"
#include "dense.cuh"
#include "filters.cuh"
#include "icp.cuh"
#include "modules.h"

using namespace pgs;

struct pgs_ctx {
  Ctx c;
};
struct pgs_cloud {
  std::unique_ptr<Cloud> c;
};
struct pgs_filters {
  Ctx* ctx;
  std::vector<Module> mods;
};
struct pgs_matcher {
  Ctx* ctx;
  Module mod;
  std::unique_ptr<Index> index;
  std::unique_ptr<DenseRef> dense;  // tensor-core tiles, built on first dense search
  unsigned dense_fallbacks = 0;     // of the last dense search
};
struct pgs_outliers {
  Ctx* ctx;
  std::vector<Module> mods;
};
struct pgs_minimizer {
  Ctx* ctx;
  Module mod;
};
struct pgs_icp {
  Ctx* ctx;
  ChainConfig cfg;
  std::unique_ptr<IcpEngine> engine;  // created on first fused use
  pgs_filters reading_filters, reading_step_filters, reference_filters;
  pgs_matcher matcher;
  pgs_outliers outliers;
  pgs_minimizer minimizer;
};

namespace {

thread_local std::string g_no_ctx_error;

pgs_status fail(Ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->last_error = msg;
  else g_no_ctx_error = msg;
  return (pgs_status)code;
}

// Every entry point runs with its context's device current (and puts the caller's device back):
// a handle may be used from any host thread, whatever device that thread last selected.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(const Ctx* c) {
    if (!c) return;
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != c->device) {
      PGS_CUDA(cudaSetDevice(c->device));
      switched = true;
    }
  }
  ~DeviceGuard() {
    if (switched && prev >= 0) cudaSetDevice(prev);
  }
};

// A cloud that belongs to another context of the same device is joined to the consumer's
// stream for the duration of a call: the consumer waits for everything queued on the owner's
// stream so far (the upload, filters ...), and when the call has been enqueued the owner's
// stream waits for the consumer, so a later free / overwrite on the owner's stream cannot
// overtake the consumer's reads.  A cloud on another DEVICE is rejected.
struct Borrow {
  Ctx* consumer;
  std::vector<Ctx*> owners;
  explicit Borrow(Ctx* c) : consumer(c) {}
  void use(const Cloud* cl) {
    if (!cl) return;
    cl->wait_ready();
    Ctx* o = cl->ctx;
    if (o == consumer) return;
    if (o->device != consumer->device)
      throw Error(PGS_INVALID_ARGUMENT, "cloud lives on device " + std::to_string(o->device) + ", the handle on device " +
                                            std::to_string(consumer->device));
    for (Ctx* seen : owners)
      if (seen == o) return;
    if (!o->cross_ev) PGS_CUDA(cudaEventCreateWithFlags(&o->cross_ev, cudaEventDisableTiming));
    PGS_CUDA(cudaEventRecord(o->cross_ev, o->stream));
    PGS_CUDA(cudaStreamWaitEvent(consumer->stream, o->cross_ev, 0));
    owners.push_back(o);
  }
  void use(const pgs_cloud* c) { use(c ? c->c.get() : nullptr); }
  ~Borrow() {
    if (owners.empty()) return;
    if (!consumer->cross_ev && cudaEventCreateWithFlags(&consumer->cross_ev, cudaEventDisableTiming) != cudaSuccess) return;
    cudaEventRecord(consumer->cross_ev, consumer->stream);
    for (Ctx* o : owners) cudaStreamWaitEvent(o->stream, consumer->cross_ev, 0);
  }
};

#define PGS_API_BEGIN(ctxptr) try { DeviceGuard _dg(ctxptr);
#define PGS_API_END(ctxptr)                                              \
  }                                                                      \
  catch (const pgs::Error& _ex) { return fail((ctxptr), _ex.code, _ex.what()); } \
  catch (const std::exception& _ex) { return fail((ctxptr), PGS_CUDA_ERROR, _ex.what()); } \
  return PGS_OK;

Params kv_params(const char* const* kv, int nkv) {
  Params p;
  for (int i = 0; i < nkv; ++i) p[kv[2 * i]] = kv[2 * i + 1];
  return p;
}

void ensure_copy_stream(Ctx* ctx) {
  if (ctx->copy_stream) return;
  PGS_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (auto& e : ctx->copy_ev) PGS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
}

// Upload from PINNED host memory on the side stream into a block allocated ON
// that stream, so the copy depends on nothing queued on the compute stream and
// overlaps with it; the compute stream joins the copy at this point of its order.
void* upload_async(Ctx* ctx, const void* src, size_t bytes, cudaEvent_t* ready) {
  ensure_copy_stream(ctx);
  void* dst = nullptr;
  PGS_CUDA(cudaMallocAsync(&dst, bytes ? bytes : 16, ctx->copy_stream));
  if (bytes) PGS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
  PGS_CUDA(cudaEventCreateWithFlags(ready, cudaEventDisableTiming));
  PGS_CUDA(cudaEventRecord(*ready, ctx->copy_stream));
  return dst;
}

void copy_in(Ctx* ctx, void* dst, const void* src, size_t bytes, int on_device) {
  if (!bytes) return;
  if (on_device == 2) {
    // pinned host memory into an EXISTING compute-stream allocation: order the
    // side stream after the allocation, the compute stream after the copy
    ensure_copy_stream(ctx);
    PGS_CUDA(cudaEventRecord(ctx->copy_ev[0], ctx->stream));
    PGS_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[0], 0));
    PGS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    PGS_CUDA(cudaEventRecord(ctx->copy_ev[1], ctx->copy_stream));
    PGS_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[1], 0));
    return;
  }
  PGS_CUDA(cudaMemcpyAsync(dst, src, bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
  if (!on_device) ctx->sync();  // the caller may reuse its host buffer right after the call
}
void copy_out(Ctx* ctx, void* dst, const void* src, size_t bytes, int on_device) {
  if (!bytes) return;
  PGS_CUDA(cudaMemcpyAsync(dst, src, bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
  if (!on_device) ctx->sync();
}

IcpEngine& engine_of(pgs_icp* icp) {
  if (!icp->engine) icp->engine = std::make_unique<IcpEngine>(icp->ctx, icp->cfg);
  return *icp->engine;
}

void wire_icp(pgs_icp* icp) {
  Ctx* ctx = icp->ctx;
  icp->reading_filters = pgs_filters{ctx, icp->cfg.reading_filters};
  icp->reading_step_filters = pgs_filters{ctx, icp->cfg.reading_step_filters};
  icp->reference_filters = pgs_filters{ctx, icp->cfg.reference_filters};
  icp->matcher.ctx = ctx;
  icp->matcher.mod = icp->cfg.matcher;
  icp->outliers = pgs_outliers{ctx, icp->cfg.outlier_filters};
  icp->minimizer = pgs_minimizer{ctx, icp->cfg.minimizer};
}

void matcher_find(pgs_matcher* m, const Cloud& reading, int32_t* d_ids, float* d_d2) {
  if (!m->index) throw Error(PGS_INVALID_FIELD, "KDTreeMatcher: init() must be called before findClosests()");
  const int k = (int)m->mod.integer("knn");
  // small reference clouds may take the tensor-core distance-tile path (option "dense_max_ref")
  if (k == 1 && m->index->n > 0 && m->index->n <= std::min(m->ctx->tune.dense_max_ref, kDenseMaxRef)) {
    if (!m->dense) {
      m->dense = std::make_unique<DenseRef>();
      dense_prepare(m->ctx, *m->index, *m->dense);
    }
    DenseQuery q{m->dense.get(), m->index->view(), reading.feat.p, (int)reading.n, d_ids, d_d2, nullptr, nullptr, nullptr};
    m->dense_fallbacks = dense_knn1(m->ctx, {q}, (float)m->mod.real("maxDist"), m->ctx->tune.dense_count_fallbacks != 0);
    return;
  }
  KnnJob job{m->index->view(), reading.feat.p, nullptr, (int)reading.n, d_ids, d_d2};
  knn_batched(m->ctx, {job}, k, (float)m->mod.real("maxDist"));
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

const char* pgs_version(void) { return "pgslam_b200 0.1 (sm_100a)"; }

pgs_status pgs_ctx_create(int device, void* stream, pgs_ctx** out) {
  Ctx* cp = nullptr;
  PGS_API_BEGIN(cp)
  if (!out) throw Error(PGS_INVALID_ARGUMENT, "out is NULL");
  int count = 0;
  struct Restore {
    int prev = -1;
    Restore() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
    ~Restore() { if (prev >= 0) cudaSetDevice(prev); }
  } restore;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    throw Error(PGS_CUDA_ERROR, std::string("no CUDA device available: ") + cudaGetErrorString(e) +
                                    " (libpgslam_b200 has no CPU fallback)");
  if (device < 0 || device >= count) throw Error(PGS_INVALID_ARGUMENT, "device index out of range");
  PGS_CUDA(cudaSetDevice(device));
  auto h = std::make_unique<pgs_ctx>();
  h->c.device = device;
  if (stream) {
    h->c.stream = static_cast<cudaStream_t>(stream);
    h->c.own_stream = false;
  } else {
    PGS_CUDA(cudaStreamCreateWithFlags(&h->c.stream, cudaStreamNonBlocking));
    h->c.own_stream = true;
  }
  cudaDeviceProp prop;
  PGS_CUDA(cudaGetDeviceProperties(&prop, device));
  h->c.num_sms = prop.multiProcessorCount;
  // keep freed blocks in the stream-ordered pool: a registration allocates and
  // frees dozens of temporaries and must not hit the OS allocator each time
  cudaMemPool_t pool;
  PGS_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t thr = UINT64_MAX;
  PGS_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  *out = h.release();
  PGS_API_END(cp)
}

void pgs_ctx_destroy(pgs_ctx* ctx) {
  if (!ctx) return;
  try {
    DeviceGuard dg(&ctx->c);
    ctx->c.destroy_resources();
    delete ctx;
  } catch (...) {
  }
}

const char* pgs_last_error(const pgs_ctx* ctx) { return ctx ? ctx->c.last_error.c_str() : g_no_ctx_error.c_str(); }

pgs_status pgs_ctx_synchronize(pgs_ctx* ctx) {
  PGS_API_BEGIN(&ctx->c)
  ctx->c.sync();
  PGS_API_END(&ctx->c)
}

pgs_status pgs_config_check(const char* yaml, size_t len, int is_chain, int* n_modules, char* err, int cap) {
  int code = PGS_OK;
  std::string msg;
  try {
    int n = 0;
    if (is_chain) {
      ChainConfig c = chain_from_yaml(std::string(yaml, len));
      params_from_chain(c);  // also rejects combinations the device path does not implement
      n = (int)(c.reading_filters.size() + c.reading_step_filters.size() + c.reference_filters.size() +
                c.outlier_filters.size() + c.checkers.size()) + 4;
    } else {
      n = (int)module_list_from_yaml(Kind::DataPointsFilter, parse_yaml(std::string(yaml, len))).size();
    }
    if (n_modules) *n_modules = n;
  } catch (const pgs::Error& ex) {
    code = ex.code;
    msg = ex.what();
  } catch (const std::exception& ex) {
    code = PGS_INVALID_ARGUMENT;
    msg = ex.what();
  }
  if (err && cap > 0) {
    std::strncpy(err, msg.c_str(), cap - 1);
    err[cap - 1] = '\0';
  }
  return (pgs_status)code;
}

int pgs_config_warnings(const char* yaml, size_t len, int is_chain, char* out, int cap) {
  std::vector<std::string> w;
  try {
    auto stochastic = [](const Module& m) {
      return m.name == "RandomSamplingDataPointsFilter" || m.name == "MaxDensityDataPointsFilter" ||
             (m.name == "SamplingSurfaceNormalDataPointsFilter" && m.integer("samplingMethod") == 0);
    };
    auto scan = [&](const std::vector<Module>& mods, const char* where, bool step) {
      for (auto& m : mods) {
        if (stochastic(m)) {
          w.push_back(std::string(where) + ": " + m.name + " draws from a counter-based hash of (seed, point index), not from rand(): "
                      "same distribution, reproducible, different points than libpointmatcher keeps");
          if (step)
            w.push_back(std::string(where) + ": " + m.name + " is applied ONCE per registration; libpointmatcher re-applies the step "
                        "filters to a fresh copy of the reading every iteration, so its random subset changes per iteration");
        }
        auto e = m.params.find("epsilon");
        if (e != m.params.end() && m.name != "ShadowDataPointsFilter" && std::atof(e->second.c_str()) != 0.0)
          w.push_back(std::string(where) + ": " + m.name + " epsilon = " + e->second + " accepted under PGS_EPSILON_POLICY=exact: the "
                      "search is exact, libnabo's would be approximate");
      }
    };
    const YamlNode root = parse_yaml(std::string(yaml, len));
    if (is_chain) {
      ChainConfig c = chain_from_yaml(std::string(yaml, len));
      scan(c.reading_filters, "readingDataPointsFilters", false);
      scan(c.reading_step_filters, "readingStepDataPointsFilters", true);
      scan(c.reference_filters, "referenceDataPointsFilters", false);
      scan({c.matcher}, "matcher", false);
    } else {
      scan(module_list_from_yaml(Kind::DataPointsFilter, root), "filters", false);
    }
  } catch (const std::exception& ex) {
    w.push_back(std::string("configuration rejected: ") + ex.what());
  }
  std::string text;
  for (auto& s : w) text += s + "\n";
  if (out && cap > 0) {
    std::strncpy(out, text.c_str(), cap - 1);
    out[cap - 1] = '\0';
  }
  return (int)w.size();
}

int pgs_registrar_count(int kind) {
  if (kind < 0 || kind > 7) return 0;
  return (int)registered_modules((Kind)kind).size();
}

const char* pgs_registrar_name(int kind, int index) {
  static thread_local std::string name;
  if (kind < 0 || kind > 7) return "";
  auto v = registered_modules((Kind)kind);
  if (index < 0 || index >= (int)v.size()) return "";
  name = v[index];
  return name.c_str();
}

int pgs_registrar_param_count(int kind, const char* name) {
  if (kind < 0 || kind > 7 || !name) return -1;
  const std::vector<ParamDoc>* p = module_params((Kind)kind, name);
  return p ? (int)p->size() : -1;
}

pgs_status pgs_registrar_param(int kind, const char* name, int index, const char** key, const char** doc,
                               const char** default_value, const char** min_value, const char** max_value, char* type) {
  if (kind < 0 || kind > 7 || !name) return PGS_INVALID_ARGUMENT;
  const std::vector<ParamDoc>* p = module_params((Kind)kind, name);
  if (!p) return PGS_INVALID_ELEMENT;
  if (index < 0 || index >= (int)p->size()) return PGS_INVALID_ARGUMENT;
  const ParamDoc& d = (*p)[index];
  if (key) *key = d.name;
  if (doc) *doc = d.doc;
  if (default_value) *default_value = d.def;
  if (min_value) *min_value = d.min;
  if (max_value) *max_value = d.max;
  if (type) *type = d.type;
  return PGS_OK;
}

pgs_status pgs_module_validate(int kind, const char* name, const char* const* kv, int nkv, char* err, int cap) {
  int code = PGS_OK;
  std::string msg;
  try {
    if (kind < 0 || kind > 7 || !name) throw Error(PGS_INVALID_ARGUMENT, "pgs_module_validate: bad arguments");
    create_module((Kind)kind, name, kv_params(kv, nkv));
  } catch (const pgs::Error& ex) {
    code = ex.code;
    msg = ex.what();
  } catch (const std::exception& ex) {
    code = PGS_INVALID_ARGUMENT;
    msg = ex.what();
  }
  if (err && cap > 0) {
    std::strncpy(err, msg.c_str(), cap - 1);
    err[cap - 1] = '\0';
  }
  return (pgs_status)code;
}

uint64_t pgs_ctx_launch_count(const pgs_ctx* ctx) { return ctx->c.launches; }

pgs_status pgs_ctx_set_profiling(pgs_ctx* ctx, int enabled) {
  ctx->c.profiling = enabled != 0;
  return PGS_OK;
}

pgs_status pgs_ctx_set_batch_streams(pgs_ctx* ctx, int n_streams) {
  if (n_streams < 1 || n_streams > 32) return fail(&ctx->c, PGS_INVALID_ARGUMENT, "batch streams must be in [1, 32]");
  ctx->c.batch_streams = n_streams;
  return PGS_OK;
}

pgs_status pgs_ctx_set_option(pgs_ctx* ctx, const char* key, double value) {
  PGS_API_BEGIN(&ctx->c)
  if (!key) throw Error(PGS_INVALID_ARGUMENT, "pgs_ctx_set_option: key is NULL");
  const std::string k(key);
  Tuning& t = ctx->c.tune;
  const int v = (int)value;
  if (k == "match_mode") { if (v < 0 || v > 5) throw Error(PGS_INVALID_ARGUMENT, "match_mode must be 0..5"); t.match_mode = v; }
  else if (k == "pm_blocks") { if (v < 1 || v > 32) throw Error(PGS_INVALID_ARGUMENT, "pm_blocks must be 1..32"); t.pm_blocks = v; }
  else if (k == "pm_refill") { if (v < 1 || v > 32) throw Error(PGS_INVALID_ARGUMENT, "pm_refill must be 1..32"); t.pm_refill = v; }
  else if (k == "pm_pair_w") { if (v < 1) throw Error(PGS_INVALID_ARGUMENT, "pm_pair_w must be >= 1"); t.pm_pair_w = v; }
  else if (k == "pm_leaf_w") { if (v < 1) throw Error(PGS_INVALID_ARGUMENT, "pm_leaf_w must be >= 1"); t.pm_leaf_w = v; }
  else if (k == "mq_batches") { if (v < 1 || v > 1024) throw Error(PGS_INVALID_ARGUMENT, "mq_batches must be 1..1024"); t.mq_batches = v; }
  else if (k == "mq_blocks") { t.mq_blocks = v; }
  else if (k == "resort_it") { t.resort_it = v; }
  else if (k == "dense_max_ref") { t.dense_max_ref = v; }
  else if (k == "dense_count_fallbacks") { t.dense_count_fallbacks = v; }
  else if (k == "batch_chunk") { if (v < 1 || v > 4096) throw Error(PGS_INVALID_ARGUMENT, "batch_chunk must be 1..4096"); t.batch_chunk = v; }
  else throw Error(PGS_INVALID_ARGUMENT, "pgs_ctx_set_option: unknown key " + k);
  PGS_API_END(&ctx->c)
}

pgs_status pgs_ctx_last_stage_times(const pgs_ctx* ctx, pgs_stage_times* out) {
  *out = ctx->c.times;
  return PGS_OK;
}

// ---- DataPoints ---------------------------------------------------------------
pgs_status pgs_cloud_create(pgs_ctx* ctx, const float* features4xN, int64_t n, int on_device, pgs_cloud** out) {
  PGS_API_BEGIN(&ctx->c)
  if (!out || n < 0 || (n > 0 && !features4xN)) throw Error(PGS_INVALID_ARGUMENT, "pgs_cloud_create: bad arguments");
  if (n > 0x7fffffff - 1024) throw Error(PGS_INVALID_ARGUMENT, "pgs_cloud_create: too many points");
  auto h = std::make_unique<pgs_cloud>();
  h->c = std::make_unique<Cloud>(&ctx->c);
  h->c->n = n;
  if (on_device == 2) {
    h->c->feat.adopt(&ctx->c, static_cast<float4*>(upload_async(&ctx->c, features4xN, (size_t)n * sizeof(float4), &h->c->ready)),
                     (size_t)n);
  } else {
    h->c->feat.reset(&ctx->c, (size_t)n);
    copy_in(&ctx->c, h->c->feat.p, features4xN, (size_t)n * sizeof(float4), on_device);
  }
  *out = h.release();
  PGS_API_END(&ctx->c)
}

pgs_status pgs_cloud_set_descriptor(pgs_cloud* c, const char* label, int span, const float* data, int on_device) {
  Ctx* ctx = c->c->ctx;
  PGS_API_BEGIN(ctx)
  c->c->wait_ready();
  if (span <= 0 || !label) throw Error(PGS_INVALID_ARGUMENT, "pgs_cloud_set_descriptor: bad arguments");
  Desc& d = c->c->add(label, span);
  copy_in(ctx, d.data.p, data, (size_t)c->c->n * span * sizeof(float), on_device);
  PGS_API_END(ctx)
}

pgs_status pgs_cloud_remove_descriptor(pgs_cloud* c, const char* label) {
  Ctx* ctx = c->c->ctx;
  PGS_API_BEGIN(ctx)
  if (!c->c->find(label)) throw Error(PGS_INVALID_FIELD, std::string("Cannot find descriptor ") + label);
  c->c->remove(label);
  PGS_API_END(ctx)
}

int64_t pgs_cloud_num_points(const pgs_cloud* c) { return c->c->n; }
int pgs_cloud_num_descriptors(const pgs_cloud* c) { return (int)c->c->descs.size(); }

pgs_status pgs_cloud_descriptor_info(const pgs_cloud* c, int index, char* label, int cap, int* span) {
  Ctx* ctx = c->c->ctx;
  PGS_API_BEGIN(ctx)
  if (index < 0 || index >= (int)c->c->descs.size()) throw Error(PGS_INVALID_FIELD, "descriptor index out of range");
  const Desc& d = c->c->descs[index];
  if (label && cap > 0) {
    std::strncpy(label, d.label.c_str(), cap - 1);
    label[cap - 1] = '\0';
  }
  if (span) *span = d.span;
  PGS_API_END(ctx)
}

pgs_status pgs_cloud_get_features(const pgs_cloud* c, float* out4xN, int on_device) {
  Ctx* ctx = c->c->ctx;
  PGS_API_BEGIN(ctx)
  c->c->wait_ready();
  copy_out(ctx, out4xN, c->c->feat.p, (size_t)c->c->n * sizeof(float4), on_device);
  PGS_API_END(ctx)
}

pgs_status pgs_cloud_get_descriptor(const pgs_cloud* c, const char* label, float* out, int on_device) {
  Ctx* ctx = c->c->ctx;
  PGS_API_BEGIN(ctx)
  c->c->wait_ready();
  const Desc* d = c->c->find(label);
  if (!d) throw Error(PGS_INVALID_FIELD, std::string("Cannot find descriptor ") + label);
  copy_out(ctx, out, d->data.p, (size_t)c->c->n * d->span * sizeof(float), on_device);
  PGS_API_END(ctx)
}

pgs_status pgs_cloud_copy(const pgs_cloud* c, pgs_cloud** out) {
  Ctx* ctx = c->c->ctx;
  PGS_API_BEGIN(ctx)
  c->c->wait_ready();
  auto h = std::make_unique<pgs_cloud>();
  h->c = c->c->clone();
  *out = h.release();
  PGS_API_END(ctx)
}

pgs_status pgs_cloud_concatenate(pgs_cloud* a, const pgs_cloud* b) {
  Ctx* ctx = a->c->ctx;
  PGS_API_BEGIN(ctx)
  Borrow bw(ctx);
  bw.use(a); bw.use(b);
  concatenate_cloud(*a->c, *b->c);
  PGS_API_END(ctx)
}

void pgs_cloud_destroy(pgs_cloud* c) {
  if (!c) return;
  try {
    DeviceGuard dg(c->c ? c->c->ctx : nullptr);
    delete c;
  } catch (...) {
  }
}

pgs_status pgs_cloud_load(pgs_ctx* ctx, const char* path, pgs_cloud** out) {
  PGS_API_BEGIN(&ctx->c)
  if (!path || !out) throw Error(PGS_INVALID_ARGUMENT, "pgs_cloud_load: bad arguments");
  HostCloud hc;
  load_cloud_file(path, hc);
  auto h = std::make_unique<pgs_cloud>();
  h->c = std::make_unique<Cloud>(&ctx->c);
  h->c->n = hc.n;
  h->c->feat.reset(&ctx->c, (size_t)hc.n);
  copy_in(&ctx->c, h->c->feat.p, hc.features.data(), (size_t)hc.n * sizeof(float4), 0);
  for (auto& d : hc.descs) {
    Desc& dd = h->c->add(d.label, d.span);
    copy_in(&ctx->c, dd.data.p, d.data.data(), d.data.size() * sizeof(float), 0);
  }
  *out = h.release();
  PGS_API_END(&ctx->c)
}

pgs_status pgs_cloud_save(const pgs_cloud* c, const char* path) {
  Ctx* ctx = c->c->ctx;
  PGS_API_BEGIN(ctx)
  if (!path) throw Error(PGS_INVALID_ARGUMENT, "pgs_cloud_save: path is NULL");
  c->c->wait_ready();
  HostCloud hc;
  hc.n = c->c->n;
  hc.features.resize((size_t)hc.n * 4);
  copy_out(ctx, hc.features.data(), c->c->feat.p, (size_t)hc.n * sizeof(float4), 0);
  for (auto& d : c->c->descs) {
    HostDesc hd;
    hd.label = d.label;
    hd.span = d.span;
    hd.data.resize((size_t)hc.n * d.span);
    copy_out(ctx, hd.data.data(), d.data.p, hd.data.size() * sizeof(float), 0);
    hc.descs.push_back(std::move(hd));
  }
  save_cloud_file(path, hc);
  PGS_API_END(ctx)
}

pgs_status pgs_cloud_file_info(const char* path, int64_t* n_points, int* n_descriptors, char* err, int cap) {
  int code = PGS_OK;
  std::string msg;
  try {
    if (!path) throw Error(PGS_INVALID_ARGUMENT, "pgs_cloud_file_info: path is NULL");
    HostCloud hc;
    load_cloud_file(path, hc);
    if (n_points) *n_points = hc.n;
    if (n_descriptors) *n_descriptors = (int)hc.descs.size();
  } catch (const pgs::Error& ex) {
    code = ex.code;
    msg = ex.what();
  } catch (const std::exception& ex) {
    code = PGS_INVALID_ARGUMENT;
    msg = ex.what();
  }
  if (err && cap > 0) {
    std::strncpy(err, msg.c_str(), cap - 1);
    err[cap - 1] = '\0';
  }
  return (pgs_status)code;
}

// ---- Transformation -------------------------------------------------------------
pgs_status pgs_rigid_transform(pgs_cloud* c, const double T[16]) {
  Ctx* ctx = c->c->ctx;
  PGS_API_BEGIN(ctx)
  c->c->wait_ready();
  rigid_transform_cloud(*c->c, T);
  PGS_API_END(ctx)
}

pgs_status pgs_cloud_assemble(pgs_ctx* ctx, int n, const pgs_cloud* const* clouds, const double* T, pgs_cloud** out) {
  PGS_API_BEGIN(&ctx->c)
  if (n < 1) throw Error(PGS_INVALID_ARGUMENT, "pgs_cloud_assemble: need at least one cloud");
  Borrow bw(&ctx->c);
  for (int i = 0; i < n; ++i) bw.use(clouds[i]);
  auto h = std::make_unique<pgs_cloud>();
  h->c = clouds[0]->c->clone(&ctx->c);
  for (int i = 1; i < n; ++i) {
    auto tmp = clouds[i]->c->clone(&ctx->c);
    rigid_transform_cloud(*tmp, T + 16 * i);
    concatenate_cloud(*h->c, *tmp);
  }
  *out = h.release();
  PGS_API_END(&ctx->c)
}

// ---- DataPointsFilters ------------------------------------------------------------
pgs_status pgs_filters_create_from_yaml(pgs_ctx* ctx, const char* yaml, size_t len, pgs_filters** out) {
  PGS_API_BEGIN(&ctx->c)
  YamlNode root = parse_yaml(std::string(yaml, len));
  auto h = std::make_unique<pgs_filters>();
  h->ctx = &ctx->c;
  h->mods = module_list_from_yaml(Kind::DataPointsFilter, root);
  *out = h.release();
  PGS_API_END(&ctx->c)
}

pgs_status pgs_filters_create(pgs_ctx* ctx, pgs_filters** out) {
  PGS_API_BEGIN(&ctx->c)
  auto h = std::make_unique<pgs_filters>();
  h->ctx = &ctx->c;
  *out = h.release();
  PGS_API_END(&ctx->c)
}

pgs_status pgs_filters_append(pgs_filters* f, const char* name, const char* const* kv, int nkv) {
  PGS_API_BEGIN(f->ctx)
  f->mods.push_back(create_module(Kind::DataPointsFilter, name, kv_params(kv, nkv)));
  PGS_API_END(f->ctx)
}

int pgs_filters_count(const pgs_filters* f) { return (int)f->mods.size(); }

pgs_status pgs_filters_apply(pgs_filters* f, pgs_cloud* c) {
  PGS_API_BEGIN(f->ctx)
  // the filters run on the CLOUD's context: its buffers are reallocated on that stream
  c->c->wait_ready();
  if (c->c->ctx->device != f->ctx->device) throw Error(PGS_INVALID_ARGUMENT, "cloud and filters live on different devices");
  std::vector<Cloud*> cl{c->c.get()};
  apply_filters(c->c->ctx, f->mods, cl);
  PGS_API_END(f->ctx)
}

void pgs_filters_destroy(pgs_filters* f) { delete f; }

// ---- Matcher -----------------------------------------------------------------------
pgs_status pgs_matcher_create(pgs_ctx* ctx, const char* name, const char* const* kv, int nkv, pgs_matcher** out) {
  PGS_API_BEGIN(&ctx->c)
  auto h = std::make_unique<pgs_matcher>();
  h->ctx = &ctx->c;
  h->mod = create_module(Kind::Matcher, name, kv_params(kv, nkv));
  if (h->mod.integer("knn") > 256) throw Error(PGS_INVALID_PARAMETER, "KDTreeMatcher: knn > 256 is not supported");
  *out = h.release();
  PGS_API_END(&ctx->c)
}

pgs_status pgs_matcher_init(pgs_matcher* m, const pgs_cloud* reference) {
  PGS_API_BEGIN(m->ctx)
  Borrow bw(m->ctx);
  bw.use(reference);
  std::vector<std::unique_ptr<Index>> idx;
  build_indices(m->ctx, {reference->c->feat.p}, {(int)reference->c->n}, nullptr, idx);
  m->index = std::move(idx[0]);
  m->dense.reset();
  PGS_API_END(m->ctx)
}

unsigned pgs_matcher_dense_fallbacks(const pgs_matcher* m) { return m->dense_fallbacks; }

int pgs_matcher_knn(const pgs_matcher* m) { return (int)m->mod.integer("knn"); }

pgs_status pgs_matcher_find(pgs_matcher* m, const pgs_cloud* reading, int32_t* ids, float* dists2, int on_device) {
  PGS_API_BEGIN(m->ctx)
  Borrow bw(m->ctx);
  bw.use(reading);
  const int k = (int)m->mod.integer("knn");
  const size_t cnt = (size_t)reading->c->n * k;
  if (on_device) {
    matcher_find(m, *reading->c, ids, dists2);
  } else {
    DBuf<int32_t> di(m->ctx, cnt);
    DBuf<float> dd(m->ctx, cnt);
    matcher_find(m, *reading->c, di.p, dd.p);
    di.download(ids, cnt);
    dd.download(dists2, cnt);
    m->ctx->sync();
  }
  PGS_API_END(m->ctx)
}

void pgs_matcher_destroy(pgs_matcher* m) { delete m; }

// ---- OutlierFilters ------------------------------------------------------------------
pgs_status pgs_outliers_create(pgs_ctx* ctx, pgs_outliers** out) {
  PGS_API_BEGIN(&ctx->c)
  auto h = std::make_unique<pgs_outliers>();
  h->ctx = &ctx->c;
  *out = h.release();
  PGS_API_END(&ctx->c)
}

pgs_status pgs_outliers_append(pgs_outliers* o, const char* name, const char* const* kv, int nkv) {
  PGS_API_BEGIN(o->ctx)
  o->mods.push_back(create_module(Kind::OutlierFilter, name, kv_params(kv, nkv)));
  PGS_API_END(o->ctx)
}

pgs_status pgs_outliers_compute(pgs_outliers* o, const pgs_cloud* reading, const pgs_cloud* reference,
                                const int32_t* ids, const float* dists2, int k, float* weights, int on_device) {
  PGS_API_BEGIN(o->ctx)
  Borrow bw(o->ctx);
  bw.use(reading); bw.use(reference);
  const int64_t nk = reading->c->n * k;
  if (on_device) {
    outlier_weights_device(o->ctx, o->mods, dists2, nk, weights, reading->c.get(), reference->c.get(), ids, k);
  } else {
    DBuf<float> dd(o->ctx, (size_t)nk), dw(o->ctx, (size_t)nk);
    DBuf<int32_t> di(o->ctx, (size_t)nk);
    copy_in(o->ctx, dd.p, dists2, (size_t)nk * sizeof(float), 0);
    copy_in(o->ctx, di.p, ids, (size_t)nk * sizeof(int32_t), 0);
    outlier_weights_device(o->ctx, o->mods, dd.p, nk, dw.p, reading->c.get(), reference->c.get(), di.p, k);
    copy_out(o->ctx, weights, dw.p, (size_t)nk * sizeof(float), 0);
  }
  PGS_API_END(o->ctx)
}

void pgs_outliers_destroy(pgs_outliers* o) { delete o; }

// ---- ErrorMinimizer --------------------------------------------------------------------
pgs_status pgs_minimizer_create(pgs_ctx* ctx, const char* name, const char* const* kv, int nkv, pgs_minimizer** out) {
  PGS_API_BEGIN(&ctx->c)
  auto h = std::make_unique<pgs_minimizer>();
  h->ctx = &ctx->c;
  h->mod = create_module(Kind::ErrorMinimizer, name, kv_params(kv, nkv));
  *out = h.release();
  PGS_API_END(&ctx->c)
}

pgs_status pgs_minimizer_compute(pgs_minimizer* e, const pgs_cloud* reading, const pgs_cloud* reference,
                                 const int32_t* ids, const float* dists2, const float* weights, int k, int on_device,
                                 pgs_min_result* out) {
  PGS_API_BEGIN(e->ctx)
  Borrow bw(e->ctx);
  bw.use(reading); bw.use(reference);
  const size_t nk = (size_t)reading->c->n * k;
  if (on_device) {
    minimize_device(e->ctx, e->mod, *reading->c, *reference->c, ids, dists2, weights, k, out);
  } else {
    DBuf<int32_t> di(e->ctx, nk);
    DBuf<float> dd(e->ctx, nk), dw(e->ctx, nk);
    copy_in(e->ctx, di.p, ids, nk * sizeof(int32_t), 0);
    copy_in(e->ctx, dd.p, dists2, nk * sizeof(float), 0);
    copy_in(e->ctx, dw.p, weights, nk * sizeof(float), 0);
    minimize_device(e->ctx, e->mod, *reading->c, *reference->c, di.p, dd.p, dw.p, k, out);
  }
  PGS_API_END(e->ctx)
}

void pgs_minimizer_destroy(pgs_minimizer* e) { delete e; }

// ---- ICP / ICPSequence ---------------------------------------------------------------------
pgs_status pgs_icp_create_from_yaml(pgs_ctx* ctx, const char* yaml, size_t len, pgs_icp** out) {
  PGS_API_BEGIN(&ctx->c)
  auto h = std::make_unique<pgs_icp>();
  h->ctx = &ctx->c;
  h->cfg = chain_from_yaml(std::string(yaml, len));
  wire_icp(h.get());
  *out = h.release();
  PGS_API_END(&ctx->c)
}

pgs_status pgs_icp_create_default(pgs_ctx* ctx, pgs_icp** out) {
  PGS_API_BEGIN(&ctx->c)
  auto h = std::make_unique<pgs_icp>();
  h->ctx = &ctx->c;
  h->cfg = chain_default();
  wire_icp(h.get());
  *out = h.release();
  PGS_API_END(&ctx->c)
}

void pgs_icp_destroy(pgs_icp* icp) { delete icp; }

pgs_filters* pgs_icp_reading_filters(pgs_icp* icp) { return &icp->reading_filters; }
pgs_filters* pgs_icp_reading_step_filters(pgs_icp* icp) { return &icp->reading_step_filters; }
pgs_filters* pgs_icp_reference_filters(pgs_icp* icp) { return &icp->reference_filters; }
pgs_matcher* pgs_icp_matcher(pgs_icp* icp) { return &icp->matcher; }
pgs_outliers* pgs_icp_outliers(pgs_icp* icp) { return &icp->outliers; }
pgs_minimizer* pgs_icp_minimizer(pgs_icp* icp) { return &icp->minimizer; }

pgs_status pgs_icp_run(pgs_icp* icp, const pgs_cloud* reading, const pgs_cloud* reference, const double T_init[16],
                       pgs_icp_result* out) {
  PGS_API_BEGIN(icp->ctx)
  Borrow bw(icp->ctx);
  bw.use(reading); bw.use(reference);
  std::vector<const Cloud*> rd{reading->c.get()}, rf{reference->c.get()};
  engine_of(icp).run_batch(rd, rf, T_init, out);
  if (out->status != PGS_OK) {
    static const char* why[] = {"", "ICP failed to converge (no outlier to filter / no point to minimize / bound exceeded)",
                                "RigidTransformation: Error, rotation matrix is not orthogonal.", "invalid parameter",
                                "reference cloud has no 'normals' descriptor (point-to-plane)"};
    return fail(icp->ctx, out->status, out->status < 5 ? why[out->status] : "ICP failed");
  }
  PGS_API_END(icp->ctx)
}

pgs_status pgs_icp_set_map(pgs_icp* icp, const pgs_cloud* map) {
  PGS_API_BEGIN(icp->ctx)
  Borrow bw(icp->ctx);
  bw.use(map);
  engine_of(icp).set_map(*map->c);
  PGS_API_END(icp->ctx)
}

int pgs_icp_has_map(const pgs_icp* icp) { return icp->engine && icp->engine->has_map(); }

pgs_status pgs_icp_run_sequence(pgs_icp* icp, const pgs_cloud* reading, const double T_init[16], pgs_icp_result* out) {
  PGS_API_BEGIN(icp->ctx)
  Borrow bw(icp->ctx);
  bw.use(reading);
  engine_of(icp).run_sequence(*reading->c, T_init, out);
  if (out->status != PGS_OK) return fail(icp->ctx, out->status, "ICPSequence failed for this reading");
  PGS_API_END(icp->ctx)
}

pgs_status pgs_icp_run_batch(pgs_icp* icp, int n_pairs, const pgs_cloud* const* readings,
                             const pgs_cloud* const* references, const double* T_inits, pgs_icp_result* results) {
  PGS_API_BEGIN(icp->ctx)
  if (n_pairs <= 0) return PGS_OK;
  std::vector<const Cloud*> rd(n_pairs), rf(n_pairs);
  Borrow bw(icp->ctx);
  for (int i = 0; i < n_pairs; ++i) {
    bw.use(readings[i]);
    bw.use(references[i]);
    rd[i] = readings[i]->c.get();
    rf[i] = references[i]->c.get();
  }
  engine_of(icp).run_batch(rd, rf, T_inits, results);
  bool any_ok = false;
  int first_bad = PGS_OK;
  for (int i = 0; i < n_pairs; ++i) {
    if (results[i].status == PGS_OK) any_ok = true;
    else if (first_bad == PGS_OK) first_bad = results[i].status;
  }
  if (!any_ok) return fail(icp->ctx, first_bad, "every pair of the batch failed");
  PGS_API_END(icp->ctx)
}

namespace {
// host-resident pairs: every chunk is uploaded into the worker context that registers it
struct HostSource : PairSource {
  const pgs_host_cloud* rd;
  const pgs_host_cloud* rf;
  int base;  // first pair of this device's block
  int mode;  // on_device mode of the uploads: 2 pinned (asynchronous), 0 pageable
  std::unique_ptr<Cloud> upload(Ctx* ctx, const pgs_host_cloud& h) {
    if (h.n < 0 || (h.n > 0 && !h.features4xN)) throw Error(PGS_INVALID_ARGUMENT, "pgs_host_cloud: bad features");
    if (h.n > 0x7fffffff - 1024) throw Error(PGS_INVALID_ARGUMENT, "pgs_host_cloud: too many points");
    auto c = std::make_unique<Cloud>(ctx);
    c->n = h.n;
    const size_t bytes = (size_t)h.n * sizeof(float4);
    if (mode == 2) {
      c->feat.adopt(ctx, static_cast<float4*>(upload_async(ctx, h.features4xN, bytes, &c->ready)), (size_t)h.n);
    } else {
      c->feat.reset(ctx, (size_t)h.n);
      if (bytes) PGS_CUDA(cudaMemcpyAsync(c->feat.p, h.features4xN, bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    for (int d = 0; d < h.n_descriptors; ++d) {
      if (!h.labels || !h.spans || !h.data || !h.labels[d] || h.spans[d] <= 0 || !h.data[d])
        throw Error(PGS_INVALID_ARGUMENT, "pgs_host_cloud: bad descriptor block");
      Desc& desc = c->add(h.labels[d], h.spans[d]);
      const size_t db = (size_t)h.n * h.spans[d] * sizeof(float);
      if (mode == 2) copy_in(ctx, desc.data.p, h.data[d], db, 2);
      else if (db) PGS_CUDA(cudaMemcpyAsync(desc.data.p, h.data[d], db, cudaMemcpyHostToDevice, ctx->stream));
    }
    return c;
  }
  void fetch(Ctx* ctx, int lo, int hi, FetchedPairs& out) override {
    for (int i = lo; i < hi; ++i) {
      out.owned.push_back(upload(ctx, rd[base + i]));
      out.readings.push_back(out.owned.back().get());
      out.owned.push_back(upload(ctx, rf[base + i]));
      out.references.push_back(out.owned.back().get());
    }
  }
  void before_run(Ctx*, FetchedPairs& f) override {
    for (auto& c : f.owned) c->wait_ready();  // the compute stream joins the uploads here, not earlier
  }
};
}  // namespace

pgs_status pgs_icp_run_batch_multi(pgs_icp* const* icps, int n_devices, int n_pairs, const pgs_host_cloud* readings,
                                   const pgs_host_cloud* references, const double* T_inits, int pinned,
                                   pgs_icp_result* results) {
  Ctx* ctx0 = (icps && n_devices > 0 && icps[0]) ? icps[0]->ctx : nullptr;
  PGS_API_BEGIN(ctx0)
  if (!icps || n_devices < 1 || n_pairs < 0 || (n_pairs > 0 && (!readings || !references || !results)))
    throw Error(PGS_INVALID_ARGUMENT, "pgs_icp_run_batch_multi: bad arguments");
  for (int d = 0; d < n_devices; ++d) {
    if (!icps[d]) throw Error(PGS_INVALID_ARGUMENT, "pgs_icp_run_batch_multi: NULL ICP handle");
    for (int e = 0; e < d; ++e)
      if (icps[e]->ctx == icps[d]->ctx) throw Error(PGS_INVALID_ARGUMENT, "pgs_icp_run_batch_multi: two handles share a context");
  }
  if (n_pairs == 0) return PGS_OK;
  const int per = (n_pairs + n_devices - 1) / n_devices;
  std::vector<std::thread> threads;
  std::vector<std::unique_ptr<Error>> errors(n_devices);
  for (int d = 0; d < n_devices; ++d) {
    const int lo = std::min(d * per, n_pairs), hi = std::min((d + 1) * per, n_pairs);
    if (lo >= hi) continue;
    threads.emplace_back([=, &errors]() {
      try {
        DeviceGuard dg(icps[d]->ctx);
        HostSource src;
        src.rd = readings;
        src.rf = references;
        src.base = lo;
        src.mode = pinned ? 2 : 0;
        engine_of(icps[d]).run_batch_source(hi - lo, src, T_inits ? T_inits + (size_t)16 * lo : nullptr, results + lo);
      } catch (const Error& e) {
        errors[d] = std::make_unique<Error>(e);
      } catch (const std::exception& e) {
        errors[d] = std::make_unique<Error>(PGS_CUDA_ERROR, e.what());
      }
    });
  }
  for (auto& t : threads) t.join();
  for (int d = 0; d < n_devices; ++d)
    if (errors[d]) {
      icps[d]->ctx->last_error = errors[d]->what();
      throw *errors[d];
    }
  bool any_ok = false;
  int first_bad = PGS_OK;
  for (int i = 0; i < n_pairs; ++i) {
    if (results[i].status == PGS_OK) any_ok = true;
    else if (first_bad == PGS_OK) first_bad = results[i].status;
  }
  if (!any_ok) return fail(ctx0, first_bad, "every pair of the batch failed");
  PGS_API_END(ctx0)
}

pgs_status pgs_icp_probe_overlap(pgs_icp* icp, const pgs_cloud* reading, const pgs_cloud* reference,
                                 const double T_world_robot[16], double* weighted_point_used_ratio) {
  PGS_API_BEGIN(icp->ctx)
  Ctx* ctx = icp->ctx;
  Borrow bw(ctx);
  bw.use(reading); bw.use(reference);
  // Localizer.hpp:309-347, module by module, without leaving the device
  auto ref = reference->c->clone(ctx);
  std::vector<Cloud*> rl{ref.get()};
  apply_filters(ctx, icp->cfg.reference_filters, rl);
  pgs_matcher m;
  m.ctx = ctx;
  m.mod = icp->cfg.matcher;
  std::vector<std::unique_ptr<Index>> idx;
  build_indices(ctx, {ref->feat.p}, {(int)ref->n}, nullptr, idx);
  m.index = std::move(idx[0]);
  auto rd = reading->c->clone(ctx);
  std::vector<Cloud*> dl{rd.get()};
  apply_filters(ctx, icp->cfg.reading_filters, dl);
  rigid_transform_cloud(*rd, T_world_robot);
  apply_filters(ctx, icp->cfg.reading_step_filters, dl);
  const int k = (int)m.mod.integer("knn");
  const size_t nk = (size_t)rd->n * k;
  DBuf<int32_t> ids(ctx, nk);
  DBuf<float> d2(ctx, nk), w(ctx, nk);
  matcher_find(&m, *rd, ids.p, d2.p);
  outlier_weights_device(ctx, icp->cfg.outlier_filters, d2.p, (int64_t)nk, w.p, rd.get(), ref.get(), ids.p, k);
  double kept = 0, wsum = 0;
  weights_ratio_device(ctx, d2.p, w.p, (int64_t)nk, &kept, &wsum);
  if (!(kept > 0)) throw Error(PGS_CONVERGENCE_ERROR, "no point to minimize");
  *weighted_point_used_ratio = wsum / (double)nk;
  PGS_API_END(icp->ctx)
}

pgs_status pgs_icp_probe_residual(pgs_icp* icp, const pgs_cloud* reading, const pgs_cloud* reference,
                                  const double T[16], double* residual) {
  PGS_API_BEGIN(icp->ctx)
  Ctx* ctx = icp->ctx;
  Borrow bw(ctx);
  bw.use(reading); bw.use(reference);
  // LoopCloser.hpp:346-362: raw candidate cloud, un-centred, unfiltered
  auto rd = reading->c->clone(ctx);
  rigid_transform_cloud(*rd, T);
  pgs_matcher m;
  m.ctx = ctx;
  m.mod = icp->cfg.matcher;
  std::vector<std::unique_ptr<Index>> idx;
  build_indices(ctx, {reference->c->feat.p}, {(int)reference->c->n}, nullptr, idx);
  m.index = std::move(idx[0]);
  const int k = (int)m.mod.integer("knn");
  const size_t nk = (size_t)rd->n * k;
  DBuf<int32_t> ids(ctx, nk);
  DBuf<float> d2(ctx, nk), w(ctx, nk);
  matcher_find(&m, *rd, ids.p, d2.p);
  outlier_weights_device(ctx, icp->cfg.outlier_filters, d2.p, (int64_t)nk, w.p, rd.get(), reference->c.get(), ids.p, k);
  pgs_min_result mr;
  minimize_device(ctx, icp->cfg.minimizer, *rd, *reference->c, ids.p, d2.p, w.p, k, &mr);
  *residual = mr.residual;
  PGS_API_END(icp->ctx)
}

}  // extern "C"
#pragma GCC visibility pop
