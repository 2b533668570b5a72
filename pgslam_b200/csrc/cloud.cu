// cloud.cu — context (stream, stream-ordered memory pool, pinned staging) and
// the device-resident DataPoints container (types.h:20; LocalMap.hpp:85,214,222).
#include <cstring>

#include "core.cuh"

namespace pgs {

void* Ctx::alloc(size_t bytes) {
  void* p = nullptr;
  PGS_CUDA(cudaMallocAsync(&p, bytes ? bytes : 16, stream));
  return p;
}

void Ctx::free(void* p) {
  if (p) cudaFreeAsync(p, stream);
}

// Small host->device uploads (job tables, a few KB) are staged in a pinned, device-mapped ring
// and moved by a one-block KERNEL, not by the copy engine: the engine also carries the bulk
// cloud uploads of the next batch (hundreds of MB queued one step ahead), and a 2 KB job table
// queued behind them held back the first kernels of every stage (+1.6 ms per step, measured).
namespace {
__global__ void stage_copy_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int words) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x) dst[i] = src[i];
}
}  // namespace

void Ctx::upload_small(void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return;
  const size_t kRing = 4u << 20;
  if (!pinned) {
    PGS_CUDA(cudaHostAlloc(&pinned, kRing, cudaHostAllocMapped));
    PGS_CUDA(cudaHostGetDevicePointer(&pinned_dev, pinned, 0));
    pinned_bytes = kRing;
    pinned_head = 0;
  }
  if (bytes > pinned_bytes / 2 || (bytes & 3) != 0) {
    PGS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
    PGS_CUDA(cudaStreamSynchronize(stream));
    return;
  }
  size_t aligned = (bytes + 255) & ~size_t(255);
  if (pinned_head + aligned > pinned_bytes) {
    PGS_CUDA(cudaStreamSynchronize(stream));  // ring wrapped: wait for in-flight copies
    pinned_head = 0;
  }
  char* slot = static_cast<char*>(pinned) + pinned_head;
  const char* slot_dev = static_cast<const char*>(pinned_dev) + pinned_head;
  pinned_head += aligned;
  std::memcpy(slot, src, bytes);
  const int words = (int)(bytes / 4);
  stage_copy_kernel<<<std::min((words + 255) / 256, 64), 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(slot_dev),
                                                                           static_cast<uint32_t*>(dst), words);
  ++launches;
}

static int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v && *v ? std::atoi(v) : dflt;
}

#ifndef PGS_MATCH_MODE_DEFAULT
#define PGS_MATCH_MODE_DEFAULT 0
#endif
#ifndef PGS_PM_MIN_BLOCKS
#define PGS_PM_MIN_BLOCKS 10
#endif
Tuning::Tuning()
    : match_mode(env_int("PGS_MATCH_MODE", PGS_MATCH_MODE_DEFAULT)),
      pm_blocks(env_int("PGS_PM_BLOCKS", PGS_PM_MIN_BLOCKS)),
      pm_refill(env_int("PGS_PM_REFILL", 4)),
      pm_pair_w(env_int("PGS_PM_PAIR_W", 5)),
      pm_leaf_w(env_int("PGS_PM_LEAF_W", 3)),
      mq_batches(env_int("PGS_MQ_BATCHES", 8)),
      mq_blocks(env_int("PGS_MQ_BLOCKS", 12)),
      resort_it(env_int("PGS_RESORT_IT", -1)),
      dense_max_ref(env_int("PGS_DENSE_MAX_REF", 0)),
      dense_count_fallbacks(env_int("PGS_DENSE_COUNT_FALLBACKS", 0)),
      batch_chunk(env_int("PGS_BATCH_CHUNK", 24)) {}

Ctx* Ctx::worker(int i) {
  while ((int)workers.size() <= i) {
    Ctx* w = new Ctx();
    w->device = device;
    w->num_sms = num_sms;
    w->batch_streams = 1;
    w->blocking_waits = true;
    PGS_CUDA(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
    w->own_stream = true;
    workers.push_back(w);
  }
  workers[i]->tune = tune;
  return workers[i];
}

void Ctx::destroy_resources() {
  for (Ctx* w : workers) {
    w->destroy_resources();
    delete w;
  }
  workers.clear();
  cudaStreamSynchronize(stream);
  if (pinned) cudaFreeHost(pinned);
  if (h_progress) cudaFreeHost((void*)h_progress);
  for (auto& e : loop_ev)
    if (e) cudaEventDestroy(e);
  for (auto& e : copy_ev)
    if (e) cudaEventDestroy(e);
  if (fork_ev) cudaEventDestroy(fork_ev);
  if (sync_ev) cudaEventDestroy(sync_ev);
  if (cross_ev) cudaEventDestroy(cross_ev);
  if (copy_stream) cudaStreamDestroy(copy_stream);
  if (own_stream) cudaStreamDestroy(stream);
  pinned = nullptr;
  h_progress = nullptr;
}

void Ctx::ensure_progress() {
  if (h_progress) return;
  void* h = nullptr;
  PGS_CUDA(cudaHostAlloc(&h, 64, cudaHostAllocMapped));
  void* d = nullptr;
  PGS_CUDA(cudaHostGetDevicePointer(&d, h, 0));
  h_progress = static_cast<volatile int*>(h);
  d_progress = static_cast<volatile int*>(d);
  *h_progress = 0;
  for (auto& e : loop_ev)
    PGS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | (blocking_waits ? cudaEventBlockingSync : 0)));
}

Desc& Cloud::add(const std::string& label, int span) {
  Desc* d = find(label);
  if (!d) {
    descs.emplace_back();
    d = &descs.back();
    d->label = label;
  }
  d->span = span;
  d->data.reset(ctx, (size_t)n * span);
  d->data.zero();
  return *d;
}

void Cloud::remove(const std::string& label) {
  for (size_t i = 0; i < descs.size(); ++i)
    if (descs[i].label == label) {
      descs.erase(descs.begin() + i);
      return;
    }
}

std::unique_ptr<Cloud> Cloud::clone(Ctx* target) const {
  Ctx* ctx = target ? target : this->ctx;
  auto c = std::make_unique<Cloud>(ctx);
  c->n = n;
  c->kd_order = kd_order;  // same points, same order, same index
  c->kd_order_n = kd_order_n;
  c->index_cache = index_cache;
  c->feat.reset(ctx, (size_t)n);
  if (n) PGS_CUDA(cudaMemcpyAsync(c->feat.p, feat.p, (size_t)n * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
  for (auto& d : descs) {
    c->descs.emplace_back();
    Desc& o = c->descs.back();
    o.label = d.label;
    o.span = d.span;
    o.data.reset(ctx, (size_t)n * d.span);
    if (n) PGS_CUDA(cudaMemcpyAsync(o.data.p, d.data.p, (size_t)n * d.span * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  return c;
}

// DataPoints::concatenate (A.9): append columns; keep only descriptors present
// in both clouds with equal span.
void concatenate_cloud(Cloud& a, const Cloud& b) {
  Ctx* ctx = a.ctx;
  a.touch();
  const int64_t na = a.n, nb = b.n;
  DBuf<float4> nf(ctx, (size_t)(na + nb));
  if (na) PGS_CUDA(cudaMemcpyAsync(nf.p, a.feat.p, (size_t)na * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
  if (nb) PGS_CUDA(cudaMemcpyAsync(nf.p + na, b.feat.p, (size_t)nb * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
  a.feat = std::move(nf);
  std::vector<Desc> kept;
  for (auto& d : a.descs) {
    const Desc* o = b.find(d.label);
    if (!o || o->span != d.span) continue;
    Desc nd;
    nd.label = d.label;
    nd.span = d.span;
    nd.data.reset(ctx, (size_t)(na + nb) * d.span);
    if (na) PGS_CUDA(cudaMemcpyAsync(nd.data.p, d.data.p, (size_t)na * d.span * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    if (nb) PGS_CUDA(cudaMemcpyAsync(nd.data.p + (size_t)na * d.span, o->data.p, (size_t)nb * d.span * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    kept.push_back(std::move(nd));
  }
  a.descs = std::move(kept);
  a.n = na + nb;
}

}  // namespace pgs
