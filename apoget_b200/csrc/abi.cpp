// extern "C" boundary of libapt_b200.so (declarations and contracts: include/apt_b200.h).
#include <array>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <memory>

#include "engine.h"

using namespace apt;

namespace {
thread_local std::string g_error;

template <typename F>
int guarded (F f) {
  try { f(); g_error.clear(); return APT_OK; }
  catch (const ApiError &e) { g_error = e.what(); return e.status; }
  catch (const std::bad_alloc &) { g_error = "out of host memory"; return APT_ERR_NOMEM; }
  catch (const std::exception &e) { g_error = e.what(); return APT_ERR_STATE; }
}

void need (bool ok, const char *what) { if (!ok) throw ApiError(APT_ERR_ARGUMENT, what); }

void requireDevice (int device) {
  int count = 0;
  cudaCheck(cudaGetDeviceCount(&count), "cudaGetDeviceCount");
  if (device < 0 || device >= count) throw ApiError(APT_ERR_CUDA, "no such CUDA device");
  cudaCheck(cudaSetDevice(device), "cudaSetDevice");
  cudaDeviceProp prop;
  cudaCheck(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
  if (prop.major != 10) throw ApiError(APT_ERR_CUDA, "libapt_b200 holds sm_100a code only");
}

template <typename T>
struct Tmp {             // device scratch for the host-pointer variants of the stateless kernels
  T *p = nullptr;
  explicit Tmp (size_t n) { if (n) cudaCheck(cudaMalloc(&p, n * sizeof(T)), "cudaMalloc"); }
  ~Tmp (void) { if (p) cudaFree(p); }
  void up (const T *h, size_t n) { if (n) cudaCheck(cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice), "H2D"); }
  void down (T *h, size_t n) { if (n) cudaCheck(cudaMemcpy(h, p, n * sizeof(T), cudaMemcpyDeviceToHost), "D2H"); }
};
}

struct apt_tree { Engine engine; explicit apt_tree (const apt_params &p) : engine(p) {} };

/// Stand-alone enveloppe pool for the stateless kernels
struct apt_pool {
  apt_params prm;
  uint32_t K, W, wp, nSpecies, lutStride;
  bool lut;
  uint32_t *payload = nullptr;  float4 *meta = nullptr;  uint32_t *fill = nullptr;  double *dlut = nullptr;
  std::vector<float> classes;       // triples, class id = index
  // tensor-core sweep (tc_dense.cu): the pool expanded into tile images + per-tile metadata, built on first use
  uint8_t *tcImg = nullptr;  uint32_t *tcMeta = nullptr;
  // ... and the expanded births of the sweep in flight: kept between calls (a stream-ordered allocation of 2 GiB per call
  // went back to the driver at every synchronisation; eight ranks allocating at once queue up behind each other)
  uint8_t *tcBirths = nullptr;  size_t tcBirthsBytes = 0;
  ~apt_pool (void) {
    cudaFree(payload); cudaFree(meta); cudaFree(fill); cudaFree(dlut); cudaFree(tcImg); cudaFree(tcMeta); cudaFree(tcBirths);
  }
  /// long bit-string genomes with the exact table go through the tensor cores (APT_DENSE_TC=0: popcount sweep)
  bool tcWanted (void) const {
    const char *e = getenv("APT_DENSE_TC");
    const bool off = e && e[0] == '0';      // read per call: tests switch between the two sweeps
    return !off && nSpecies > 0 && dense_tc_supports(cfg(), wp, K);
  }
  void tcEnsure (cudaStream_t st) {
    if (tcImg) return;
    cudaCheck(cudaMalloc(&tcImg, tc_pool_image_bytes(nSpecies, K, wp)), "cudaMalloc (expanded pool)");
    cudaCheck(cudaMalloc(&tcMeta, tc_pool_meta_bytes(nSpecies, K)), "cudaMalloc (tile metadata)");
    launch_tc_expand_pool(view(), tcImg, tcMeta, st);
    cudaCheck(cudaGetLastError(), "pool expansion launch");
  }
  /// dense sweep of n padded births (dB: n x wp, dM: metadata) on the tensor cores, in slabs of births whose expanded
  /// image stays below 2 GiB; `matrix` != null: distance-matrix mode
  void tcSweep (const uint32_t *dB, const float4 *dM, uint32_t n, uint32_t *best, float *score, uint32_t *first, uint32_t *count,
                uint32_t *matrix, size_t matrixLd, cudaStream_t st, const ShardSink *sink) {
    tcEnsure(st);
    const size_t perBirth = tc_births_image_bytes(128, wp) / 128;
    uint32_t slab = uint32_t(std::min<size_t>(n, std::max<size_t>(128, ((size_t(2) << 30) / perBirth) / (148 * 128) * (148 * 128))));
    const size_t imgBytes = tc_births_image_bytes(slab, wp);
    if (imgBytes > tcBirthsBytes) {
      // (cudaFree waits for the device: no sweep of an earlier call still reads the old image)
      cudaFree(tcBirths);  tcBirths = nullptr;  tcBirthsBytes = 0;
      cudaCheck(cudaMalloc(&tcBirths, imgBytes), "cudaMalloc (expanded births)");
      tcBirthsBytes = imgBytes;
    }
    uint8_t *img = tcBirths;
    for (uint32_t b0 = 0; b0 < n; b0 += slab) {
      const uint32_t m = std::min(slab, n - b0);
      launch_tc_expand_births(dB + size_t(b0) * wp, m, wp, img, st);
      launch_classify_dense_tc(view(), cfg(), img, tcImg, tcMeta, dM + b0, m, best ? best + b0 : nullptr, score ? score + b0 : nullptr,
                               first ? first + b0 : nullptr, count ? count + b0 : nullptr,
                               matrix ? matrix + size_t(b0) * matrixLd : nullptr, st, sink, n, b0);
    }
    cudaCheck(cudaGetLastError(), "tensor-core sweep launch");
  }
  std::map<std::array<float, 3>, uint32_t> classIds;
  uint32_t classOf (float mu, float sin, float sout) {
    const std::array<float, 3> key { mu, sin, sout };
    auto it = classIds.find(key);
    if (it != classIds.end()) return it->second;
    const uint32_t id = uint32_t(classes.size() / 3);
    classIds.emplace(key, id);
    classes.push_back(mu);  classes.push_back(sin);  classes.push_back(sout);
    return id;
  }
  void uploadLut (void) {
    if (!lut) return;
    const size_t n = classes.size() / 3;
    // same bound as the tree engine: beyond 256 MiB of table the caller has to ask for the generic path
    if (n * lutStride * sizeof(double) > (size_t(256) << 20))
      throw ApiError(APT_ERR_ARGUMENT, "too many distinct compatibility triples for the table (use APT_FLAG_FORCE_GENERIC)");
    std::vector<double> h (n * lutStride);
    for (size_t c = 0; c < n; c++)
      for (uint32_t d = 0; d < lutStride; d++) {
        const float mu = classes[3 * c];
        const double dd = d, sigma = dd < mu ? classes[3 * c + 1] : classes[3 * c + 2];
        h[c * lutStride + d] = std::exp(-((dd - mu) * (dd - mu)) / (2. * sigma * sigma));   // crossover.h:89-92
      }
    cudaFree(dlut);  dlut = nullptr;
    cudaCheck(cudaMalloc(&dlut, h.size() * sizeof(double)), "cudaMalloc");
    cudaCheck(cudaMemcpy(dlut, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice), "LUT upload");
  }
  PoolView view (void) const { return PoolView { payload, meta, fill, K, wp, nSpecies }; }
  ScoreCfg cfg (void) const {
    return ScoreCfg { int(prm.kind), prm.full_continuous, lut ? 1 : 0, prm.compat_thr, prm.sim_thr, prm.avg_thr, dlut, lutStride, uint32_t(classes.size() / 3) };
  }
  /// pads + packs host births into device buffers
  void packBirths (uint32_t n, const uint32_t *pl, const float *compat, Tmp<uint32_t> &dB, Tmp<float4> &dM) {
    std::vector<uint32_t> hb (size_t(n) * wp, 0u);
    std::vector<float4> hm (n);
    const size_t before = classes.size();
    for (uint32_t i = 0; i < n; i++) {
      memcpy(&hb[size_t(i) * wp], pl + size_t(i) * W, W * 4);
      const uint32_t cls = lut ? classOf(compat[3 * i], compat[3 * i + 1], compat[3 * i + 2]) : 0;
      float fc;  memcpy(&fc, &cls, 4);
      hm[i] = make_float4(compat[3 * i], compat[3 * i + 1], compat[3 * i + 2], fc);
    }
    if (classes.size() != before) uploadLut();
    dB.up(hb.data(), hb.size());
    dM.up(hm.data(), n);
  }
};

extern "C" {

void apt_params_default (apt_params *p) {
  memset(p, 0, sizeof *p);
  p->kind = APT_BITS;  p->words = 8;
  p->rset_size = 5;  p->full_continuous = 1;  p->env_crit = 1;         // ptreeconfig.cpp:10, 18, 19
  p->compat_thr = .1;  p->sim_thr = .5;  p->avg_thr = .25;             // ptreeconfig.cpp:7-9
  p->trim_period = 100;  p->trim_thr = .25f;  p->trim_delay = 4.f;  p->trim_min_delay = 200;   // :13-16
}

const char* apt_last_error (void) { return g_error.c_str(); }

int apt_device_available (void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
  for (int d = 0; d < count; d++) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, d) == cudaSuccess && prop.major == 10) return 1;
  }
  return 0;
}

int apt_tree_create (const apt_params *params, apt_tree **out) {
  return guarded([&] {
    need(params && out, "null argument");
    *out = new apt_tree(*params);
  });
}

void apt_tree_destroy (apt_tree *t) { delete t; }

int apt_tree_add_genomes (apt_tree *t, uint32_t n, const uint32_t *payload, const float *compat,
                          const uint32_t *gid, const uint32_t *mgid, const uint32_t *msid,
                          const uint32_t *fgid, const uint32_t *fsid,
                          uint32_t *out_sid, int32_t *out_slot, uint32_t *out_replaced, uint32_t *n_done) {
  return guarded([&] {
    need(t && (n == 0 || (payload && compat && gid && mgid && msid && fgid && fsid)), "null argument");
    t->engine.addGenomes(n, payload, compat, false, nullptr, gid, mgid, msid, fgid, fsid,
                         out_sid, out_slot, out_replaced, n_done);
  });
}

int apt_tree_add_genomes_dev (apt_tree *t, uint32_t n, const uint32_t *d_payload, const float *compat, void *stream,
                              const uint32_t *gid, const uint32_t *mgid, const uint32_t *msid,
                              const uint32_t *fgid, const uint32_t *fsid,
                              uint32_t *out_sid, int32_t *out_slot, uint32_t *out_replaced, uint32_t *n_done) {
  return guarded([&] {
    need(t && (n == 0 || (d_payload && compat && gid && mgid && msid && fgid && fsid)), "null argument");
    t->engine.addGenomes(n, d_payload, compat, true, static_cast<cudaStream_t>(stream), gid, mgid, msid, fgid, fsid,
                         out_sid, out_slot, out_replaced, n_done);
  });
}

int apt_tree_parallel_for (apt_tree *t, uint32_t n, void (*fn)(void *ctx, uint32_t lo, uint32_t hi), void *ctx) {
  // (an exception escaping fn on a worker thread terminates the process, as it would out of any C callback)
  return guarded([&] { need(t && fn, "null argument"); t->engine.parallelFor(n, fn, ctx); });
}

int apt_host_alloc (uint64_t bytes, void **out) {
  return guarded([&] {
    need(out, "null argument");
    *out = nullptr;
    if (bytes) apt::cudaCheck(cudaMallocHost(out, size_t(bytes)), "cudaMallocHost");
  });
}

void apt_host_free (void *p) { if (p) cudaFreeHost(p); }

int apt_tree_del_genome (apt_tree *t, uint32_t sid) { return guarded([&] { need(t, "null tree"); t->engine.delGenome(sid); }); }

int apt_tree_del_genomes (apt_tree *t, uint32_t n, const uint32_t *sids) {
  return guarded([&] {
    need(t && (n == 0 || sids), "null argument");
    for (uint32_t i = 0; i < n; i++) t->engine.delGenome(sids[i]);
  });
}

int apt_tree_register_candidate (apt_tree *t, uint32_t msid, uint32_t fsid, int dir) {
  return guarded([&] { need(t && (dir == 1 || dir == -1), "bad argument"); t->engine.candidacy(msid, fsid, dir); });
}

int apt_tree_set_step (apt_tree *t, uint32_t step) { return guarded([&] { need(t, "null tree"); t->engine.setStep(step); }); }

int apt_tree_step (apt_tree *t, uint32_t step, const uint32_t *alive, uint32_t n) {
  return guarded([&] { need(t && (n == 0 || alive), "null argument"); t->engine.doStep(step, alive, n); });
}

int apt_tree_stats (apt_tree *t, uint32_t out[8]) { return guarded([&] { need(t && out, "null argument"); t->engine.stats(out); }); }
int apt_tree_reset_stats (apt_tree *t) { return guarded([&] { need(t, "null tree"); t->engine.resetStats(); }); }

uint32_t apt_tree_event_count (apt_tree *t) { return t ? uint32_t(t->engine.events().size()) : 0; }

uint32_t apt_tree_events (apt_tree *t, apt_event *out, uint32_t cap, int drain) {
  if (!t) return 0;
  auto &ev = t->engine.events();
  const uint32_t n = uint32_t(std::min<size_t>(cap, ev.size()));
  if (out && n) memcpy(out, ev.data(), n * sizeof(apt_event));
  if (drain) ev.erase(ev.begin(), ev.begin() + n);        // (moves trivially copyable elements down: cannot throw)
  return n;
}

uint64_t apt_tree_to_json (apt_tree *t, char *out, uint64_t cap) {
  // 0 = failure (an empty tree still dumps a non-empty object); the message is in apt_last_error()
  uint64_t size = 0;
  guarded([&] {
    need(t, "null tree");
    const std::string s = t->engine.toJson();
    if (out && cap) memcpy(out, s.data(), std::min<uint64_t>(cap, s.size()));
    size = s.size();
  });
  return size;
}

int apt_tree_from_json (apt_tree *t, const char *text, uint64_t len) {
  return guarded([&] {
    need(t && text, "null argument");
    t->engine.fromJson(std::string(text, size_t(len)));
  });
}

int apt_tree_clone (apt_tree *src, apt_tree **out) {
  return guarded([&] {
    need(src && out, "null argument");
    std::unique_ptr<apt_tree> t (new apt_tree(src->engine.params()));
    t->engine.cloneFrom(src->engine);
    *out = t.release();
  });
}

uint32_t apt_tree_trimmed (apt_tree *t, uint32_t *out, uint32_t cap, int drain) {
  if (!t) return 0;
  auto &v = t->engine.trimmedSpecies();
  const uint32_t n = uint32_t(std::min<size_t>(cap, v.size()));
  if (out && n) memcpy(out, v.data(), n * sizeof(uint32_t));
  if (drain) v.erase(v.begin(), v.begin() + n);
  return out ? n : uint32_t(v.size());
}

int apt_tree_node_info (apt_tree *t, uint32_t sid, apt_node_info *out) {
  return guarded([&] {
    need(t && out, "null argument");
    const Species *s = t->engine.nodeAt(sid);
    out->sid = s->sid;  out->parent_sid = s->parent ? s->parent->sid : APT_INVALID;
    out->data[0] = s->first; out->data[1] = s->last; out->data[2] = s->count; out->data[3] = s->alive; out->data[4] = s->pending;
    out->fill = s->k;  out->n_children = uint32_t(s->children.size());  out->n_contributors = uint32_t(s->contribs.size());
  });
}

int apt_tree_node_children (apt_tree *t, uint32_t sid, uint32_t *out, uint32_t cap) {
  return guarded([&] {
    need(t && out, "null argument");
    const Species *s = t->engine.nodeAt(sid);
    for (size_t i = 0; i < s->children.size() && i < cap; i++) out[i] = s->children[i]->sid;
  });
}

int apt_tree_node_contributors (apt_tree *t, uint32_t sid, uint32_t *out, uint32_t cap) {
  return guarded([&] {
    need(t && out, "null argument");
    const Species *s = t->engine.nodeAt(sid);
    for (size_t i = 0; i < s->contribs.size() && i < cap; i++) {
      out[3 * i] = s->contribs[i].sid;  out[3 * i + 1] = s->contribs[i].count;  out[3 * i + 2] = s->contribs[i].elligible;
    }
  });
}

int apt_tree_node_rep_gids (apt_tree *t, uint32_t sid, uint32_t *out, uint32_t cap) {
  return guarded([&] {
    need(t && out, "null argument");
    const Species *s = t->engine.nodeAt(sid);
    for (uint32_t r = 0; r < s->k && r < cap; r++) out[r] = s->reps[r].gid;
  });
}

int apt_tree_node_distances (apt_tree *t, uint32_t sid, float *out) {
  return guarded([&] {
    need(t && out, "null argument");
    const Species *s = t->engine.nodeAt(sid);
    const uint32_t K = t->engine.params().rset_size;
    for (uint32_t i = 0; i < K; i++)
      for (uint32_t j = 0; j < K; j++)
        out[i * K + j] = (i != j && i < s->k && j < s->k) ? s->dist[i * K + j] : std::numeric_limits<float>::quiet_NaN();
  });
}

int apt_tree_root (apt_tree *t, uint32_t *out) {
  return guarded([&] { need(t && out, "null argument"); *out = t->engine.root() ? t->engine.root()->sid : APT_INVALID; });
}

int apt_tree_alive_species (apt_tree *t, uint32_t *out, uint32_t cap, uint32_t *n) {
  return guarded([&] {
    need(t && n, "null argument");
    const auto &a = t->engine.aliveSpecies();
    *n = uint32_t(a.size());
    uint32_t i = 0;
    for (uint32_t s: a) { if (i >= cap || !out) break; out[i++] = s; }
  });
}

int apt_tree_counters (apt_tree *t, uint64_t out[16]) {
  return guarded([&] { need(t && out, "null argument"); memcpy(out, t->engine.counters, sizeof t->engine.counters); });
}

int apt_tree_timers (apt_tree *t, double out[8]) {
  return guarded([&] { need(t && out, "null argument"); memcpy(out, t->engine.timers, sizeof t->engine.timers); });
}

int apt_tree_kernel_span (apt_tree *t, double out[2]) {
  return guarded([&] { need(t && out, "null argument"); out[0] = t->engine.streamSpanMs_; out[1] = double(t->engine.streamSpanN_); });
}

// ---- stateless kernels ----------------------------------------------------------------------------

int apt_pool_create (const apt_params *params, uint32_t n_species, const uint32_t *rep_payload,
                     const float *rep_compat, const uint32_t *fill, apt_pool **out) {
  return guarded([&] {
    need(params && out && (n_species == 0 || (rep_payload && rep_compat && fill)), "null argument");
    need(params->kind <= APT_F32 && params->words >= 1 && params->words <= 1024, "bad genome shape");
    need(params->rset_size >= 1 && params->rset_size <= KMAX, "rset_size must be in 1..32");
    requireDevice(params->device);
    std::unique_ptr<apt_pool> p (new apt_pool);
    p->prm = *params;  p->K = params->rset_size;  p->W = params->words;  p->wp = (p->W + 3u) & ~3u;
    p->nSpecies = n_species;  p->lutStride = 32u * p->W + 1u;
    p->lut = params->kind == APT_BITS && !(params->flags & APT_FLAG_FORCE_GENERIC);
    const size_t rows = size_t(n_species) * p->K;
    std::vector<uint32_t> hp (rows * p->wp, 0u);
    std::vector<float4> hm (rows);
    for (size_t r = 0; r < rows; r++) {
      memcpy(&hp[r * p->wp], rep_payload + r * p->W, p->W * 4);
      const bool used = (r % p->K) < fill[r / p->K];
      const uint32_t cls = (p->lut && used) ? p->classOf(rep_compat[3 * r], rep_compat[3 * r + 1], rep_compat[3 * r + 2]) : 0;
      float fc;  memcpy(&fc, &cls, 4);
      hm[r] = make_float4(rep_compat[3 * r], rep_compat[3 * r + 1], rep_compat[3 * r + 2], fc);
    }
    cudaCheck(cudaMalloc(&p->payload, std::max<size_t>(hp.size(), 4) * 4), "cudaMalloc");
    cudaCheck(cudaMalloc(&p->meta, std::max<size_t>(rows, 1) * sizeof(float4)), "cudaMalloc");
    cudaCheck(cudaMalloc(&p->fill, std::max<size_t>(n_species, 1) * 4), "cudaMalloc");
    if (rows) {
      cudaCheck(cudaMemcpy(p->payload, hp.data(), hp.size() * 4, cudaMemcpyHostToDevice), "pool upload");
      cudaCheck(cudaMemcpy(p->meta, hm.data(), rows * sizeof(float4), cudaMemcpyHostToDevice), "pool upload");
      cudaCheck(cudaMemcpy(p->fill, fill, size_t(n_species) * 4, cudaMemcpyHostToDevice), "pool upload");
    }
    p->uploadLut();
    *out = p.release();
  });
}

void apt_pool_destroy (apt_pool *p) { delete p; }

int apt_score_pairs (apt_pool *p, uint32_t n_births, const uint32_t *payload, const float *compat,
                     uint32_t n_pairs, const uint32_t *pair_birth, const uint32_t *pair_species,
                     float *out_score, float *out_dist) {
  return guarded([&] {
    need(p && out_score && out_dist && (n_pairs == 0 || (payload && compat && pair_birth && pair_species)), "null argument");
    for (uint32_t i = 0; i < n_pairs; i++)
      need(pair_birth[i] < n_births && pair_species[i] < p->nSpecies, "pair index out of range");
    cudaCheck(cudaSetDevice(p->prm.device), "cudaSetDevice");
    Tmp<uint32_t> dB (size_t(n_births) * p->wp + 4);  Tmp<float4> dM (n_births + 1);
    p->packBirths(n_births, payload, compat, dB, dM);
    Tmp<uint32_t> dPb (n_pairs), dPs (n_pairs);  Tmp<float> dS (n_pairs), dD (size_t(n_pairs) * p->K);
    dPb.up(pair_birth, n_pairs);  dPs.up(pair_species, n_pairs);
    launch_score_pairs(p->view(), p->cfg(), dB.p, dM.p, n_pairs, dPb.p, dPs.p, dS.p, dD.p, nullptr);
    cudaCheck(cudaGetLastError(), "score_pairs launch");
    cudaCheck(cudaDeviceSynchronize(), "score_pairs");
    dS.down(out_score, n_pairs);  dD.down(out_dist, size_t(n_pairs) * p->K);
  });
}

int apt_classify_dense_dev (apt_pool *p, uint32_t n_births, const uint32_t *d_payload, const float *d_compat,
                            uint32_t *d_best, float *d_best_score, uint32_t *d_first, uint32_t *d_count, void *stream) {
  return guarded([&] {
    need(p && d_payload && d_compat && d_best && d_best_score && d_first && d_count, "null argument");
    need(!p->lut || p->classes.size() <= 3, "device-resident births need a single compatibility class (or APT_FLAG_FORCE_GENERIC)");
    cudaCheck(cudaSetDevice(p->prm.device), "cudaSetDevice");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint32_t *dB = nullptr;  float4 *dM = nullptr;
    cudaCheck(cudaMallocAsync(&dB, (size_t(n_births) * p->wp + 4) * 4, st), "cudaMallocAsync");
    cudaCheck(cudaMallocAsync(&dM, (size_t(n_births) + 1) * sizeof(float4), st), "cudaMallocAsync");
    launch_pack_births(d_payload, d_compat, nullptr, n_births, p->W, p->wp, dB, dM, st);
    if (p->tcWanted()) p->tcSweep(dB, dM, n_births, d_best, d_best_score, d_first, d_count, nullptr, 0, st, nullptr);
    else launch_classify_dense(p->view(), p->cfg(), dB, dM, n_births, d_best, d_best_score, d_first, d_count, st);
    cudaCheck(cudaGetLastError(), "classify_dense launch");
    cudaCheck(cudaFreeAsync(dB, st), "cudaFreeAsync");
    cudaCheck(cudaFreeAsync(dM, st), "cudaFreeAsync");
  });
}

int apt_classify_dense (apt_pool *p, uint32_t n_births, const uint32_t *payload, const float *compat,
                        uint32_t *out_best, float *out_best_score, uint32_t *out_first, uint32_t *out_count) {
  return guarded([&] {
    need(p && (n_births == 0 || (payload && compat && out_best && out_best_score && out_first && out_count)), "null argument");
    cudaCheck(cudaSetDevice(p->prm.device), "cudaSetDevice");
    Tmp<uint32_t> dB (size_t(n_births) * p->wp + 4);  Tmp<float4> dM (n_births + 1);
    p->packBirths(n_births, payload, compat, dB, dM);
    Tmp<uint32_t> dBest (n_births), dFirst (n_births), dCount (n_births);  Tmp<float> dScore (n_births);
    if (p->tcWanted()) p->tcSweep(dB.p, dM.p, n_births, dBest.p, dScore.p, dFirst.p, dCount.p, nullptr, 0, nullptr, nullptr);
    else launch_classify_dense(p->view(), p->cfg(), dB.p, dM.p, n_births, dBest.p, dScore.p, dFirst.p, dCount.p, nullptr);
    cudaCheck(cudaGetLastError(), "classify_dense launch");
    cudaCheck(cudaDeviceSynchronize(), "classify_dense");
    dBest.down(out_best, n_births);  dScore.down(out_best_score, n_births);
    dFirst.down(out_first, n_births);  dCount.down(out_count, n_births);
  });
}

int apt_pool_uses_tensor_cores (apt_pool *p) { return p && p->tcWanted() ? 1 : 0; }

int apt_classify_dense_shard_dev (apt_pool *p, uint32_t n_births, const uint32_t *d_payload, const float *d_compat,
                                  uint32_t species_offset, uint32_t rank, uint32_t world,
                                  const uint64_t *peer_records, void *stream) {
  return guarded([&] {
    need(p && d_payload && d_compat && peer_records, "null argument");
    need(world >= 1 && world <= MAX_PEERS && rank < world, "bad shard layout (at most 8 shards)");
    need(!p->lut || p->classes.size() <= 3, "device-resident births need a single compatibility class (or APT_FLAG_FORCE_GENERIC)");
    cudaCheck(cudaSetDevice(p->prm.device), "cudaSetDevice");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ShardSink sink;
    for (uint32_t q = 0; q < world; q++) {
      need(peer_records[q] != 0, "null peer buffer");
      sink.peer[q] = reinterpret_cast<uint4*>(uintptr_t(peer_records[q]));
    }
    sink.world = world;  sink.rank = rank;  sink.species_offset = species_offset;
    uint32_t *dB = nullptr;  float4 *dM = nullptr;
    cudaCheck(cudaMallocAsync(&dB, (size_t(n_births) * p->wp + 4) * 4, st), "cudaMallocAsync");
    cudaCheck(cudaMallocAsync(&dM, (size_t(n_births) + 1) * sizeof(float4), st), "cudaMallocAsync");
    launch_pack_births(d_payload, d_compat, nullptr, n_births, p->W, p->wp, dB, dM, st);
    if (p->tcWanted()) {
      // tensor-core sweep: its epilogue stores the records into the peers' buffers
      p->tcSweep(dB, dM, n_births, nullptr, nullptr, nullptr, nullptr, nullptr, 0, st, &sink);
    } else if (dense_tiled_supports(p->wp, p->K)) {
      // the sweep kernel itself stores the records into the peers' buffers
      launch_classify_dense_tiled(p->view(), p->cfg(), dB, dM, n_births, nullptr, nullptr, nullptr, nullptr, st, &sink);
    } else {
      uint32_t *dW = nullptr;
      cudaCheck(cudaMallocAsync(&dW, size_t(n_births) * 16 + 16, st), "cudaMallocAsync");
      uint32_t *best = dW, *first = dW + n_births, *count = dW + 2 * size_t(n_births);
      float *score = reinterpret_cast<float*>(dW + 3 * size_t(n_births));
      launch_classify_dense(p->view(), p->cfg(), dB, dM, n_births, best, score, first, count, st);
      launch_publish_records(best, score, first, count, n_births, sink, st);
      cudaCheck(cudaFreeAsync(dW, st), "cudaFreeAsync");
    }
    cudaCheck(cudaGetLastError(), "classify_dense launch");
    cudaCheck(cudaFreeAsync(dB, st), "cudaFreeAsync");
    cudaCheck(cudaFreeAsync(dM, st), "cudaFreeAsync");
  });
}

int apt_combine_shards_dev (int device, uint32_t n_births, uint32_t world, const void *d_records,
                            uint32_t *d_best, float *d_best_score, uint32_t *d_first, uint32_t *d_count, void *stream) {
  return guarded([&] {
    need(world >= 1 && (n_births == 0 || (d_records && d_best && d_best_score && d_first && d_count)), "null argument");
    requireDevice(device);
    launch_combine_shards(static_cast<const uint4*>(d_records), n_births, world, d_best, d_best_score, d_first, d_count,
                          static_cast<cudaStream_t>(stream));
    cudaCheck(cudaGetLastError(), "combine_shards launch");
  });
}

int apt_distance_matrix (const apt_params *params, uint32_t n_a, const uint32_t *payload_a, uint32_t n_b,
                         const uint32_t *payload_b, uint32_t *out) {
  return guarded([&] {
    need(params && (n_a == 0 || n_b == 0 || (payload_a && payload_b && out)), "null argument");
    need(params->kind == APT_BITS && params->words >= 1 && params->words <= 1024, "bit-string genomes only");
    if (!n_a || !n_b) return;
    requireDevice(params->device);
    // a pool of n_b one-representative species: a tile of the expanded pool is then 256 consecutive rows of b
    apt_params prm = *params;
    prm.rset_size = 1;  prm.flags &= ~uint32_t(APT_FLAG_FORCE_GENERIC);
    std::vector<float> compat (size_t(std::max(n_a, n_b)) * 3, 1.f);
    std::vector<uint32_t> fill (n_b, 1u);
    apt_pool *raw = nullptr;
    if (apt_pool_create(&prm, n_b, payload_b, compat.data(), fill.data(), &raw) != APT_OK) throw ApiError(APT_ERR_CUDA, g_error);
    std::unique_ptr<apt_pool> p (raw);
    need(dense_tc_supports(p->cfg(), p->wp, 1), "genomes of 1024 to 32768 bits only");
    Tmp<uint32_t> dB (size_t(n_a) * p->wp + 4);  Tmp<float4> dM (n_a + 1);
    p->packBirths(n_a, payload_a, compat.data(), dB, dM);
    const size_t ld = (size_t(n_b) + 255) / 256 * 256;
    Tmp<uint32_t> dOut (size_t(n_a) * ld);
    p->tcSweep(dB.p, dM.p, n_a, nullptr, nullptr, nullptr, nullptr, dOut.p, ld, nullptr, nullptr);
    cudaCheck(cudaDeviceSynchronize(), "distance matrix");
    std::vector<uint32_t> h (size_t(n_a) * ld);
    dOut.down(h.data(), h.size());
    for (uint32_t i = 0; i < n_a; i++) memcpy(out + size_t(i) * n_b, h.data() + size_t(i) * ld, size_t(n_b) * 4);
  });
}

int apt_env_contribution (const apt_params *params, uint32_t n, const float *edist, const float *gdist,
                          uint8_t *out_better, uint32_t *out_than, float *out_value) {
  return guarded([&] {
    need(params && (n == 0 || (edist && gdist && out_better && out_than && out_value)), "null argument");
    need(params->env_crit >= 0 && params->env_crit <= 3, "No function for this use case");
    need(params->rset_size >= 1 && params->rset_size <= KMAX, "rset_size must be in 1..32");
    requireDevice(params->device);
    const uint32_t K = params->rset_size;
    Tmp<float> dE (size_t(n) * K * K), dG (size_t(n) * K), dV (n);
    Tmp<uint8_t> dBt (n);  Tmp<uint32_t> dT (n);
    dE.up(edist, size_t(n) * K * K);  dG.up(gdist, size_t(n) * K);
    launch_env_contribution(params->env_crit, K, n, dE.p, dG.p, dBt.p, dT.p, dV.p, nullptr);
    cudaCheck(cudaGetLastError(), "env_contribution launch");
    cudaCheck(cudaDeviceSynchronize(), "env_contribution");
    dBt.down(out_better, n);  dT.down(out_than, n);  dV.down(out_value, n);
  });
}

int apt_pairwise (const apt_params *params, uint32_t n_genomes, const uint32_t *payload, uint32_t n_pairs,
                  const uint32_t *idx_a, const uint32_t *idx_b, float *out) {
  return guarded([&] {
    need(params && (n_pairs == 0 || (payload && idx_a && idx_b && out)), "null argument");
    need(params->kind <= APT_F32 && params->words >= 1 && params->words <= 1024, "bad genome shape");
    for (uint32_t i = 0; i < n_pairs; i++) need(idx_a[i] < n_genomes && idx_b[i] < n_genomes, "pair index out of range");
    requireDevice(params->device);
    const uint32_t W = params->words, wp = (W + 3u) & ~3u;
    std::vector<uint32_t> hp (size_t(n_genomes) * wp, 0u);
    for (uint32_t i = 0; i < n_genomes; i++) memcpy(&hp[size_t(i) * wp], payload + size_t(i) * W, W * 4);
    Tmp<uint32_t> dP (hp.size() + 4), dA (n_pairs), dB (n_pairs);  Tmp<float> dO (n_pairs);
    dP.up(hp.data(), hp.size());  dA.up(idx_a, n_pairs);  dB.up(idx_b, n_pairs);
    launch_pairwise(int(params->kind), wp, dP.p, n_pairs, dA.p, dB.p, dO.p, nullptr);
    cudaCheck(cudaGetLastError(), "pairwise launch");
    cudaCheck(cudaDeviceSynchronize(), "pairwise");
    dO.down(out, n_pairs);
  });
}

int apt_mate_compat (const apt_params *params, uint32_t n_genomes, const uint32_t *payload, const float *compat,
                     uint32_t n_pairs, const uint32_t *idx_m, const uint32_t *idx_f, float *out_dist, double *out_compat) {
  return guarded([&] {
    need(params && (n_pairs == 0 || (payload && compat && idx_m && idx_f && out_dist && out_compat)), "null argument");
    need(params->kind <= APT_F32 && params->words >= 1 && params->words <= 1024, "bad genome shape");
    for (uint32_t i = 0; i < n_pairs; i++) need(idx_m[i] < n_genomes && idx_f[i] < n_genomes, "pair index out of range");
    requireDevice(params->device);
    const uint32_t W = params->words, wp = (W + 3u) & ~3u;
    std::vector<uint32_t> hp (size_t(n_genomes) * wp, 0u);
    std::vector<float4> hm (n_genomes);
    for (uint32_t i = 0; i < n_genomes; i++) {
      memcpy(&hp[size_t(i) * wp], payload + size_t(i) * W, W * 4);
      hm[i] = make_float4(compat[3 * i], compat[3 * i + 1], compat[3 * i + 2], 0.f);
    }
    Tmp<uint32_t> dP (hp.size() + 4), dA (n_pairs), dB (n_pairs);  Tmp<float4> dM (n_genomes + 1);
    Tmp<float> dD (n_pairs);  Tmp<double> dC (n_pairs);
    dP.up(hp.data(), hp.size());  dM.up(hm.data(), n_genomes);  dA.up(idx_m, n_pairs);  dB.up(idx_f, n_pairs);
    launch_mate_compat(int(params->kind), wp, dP.p, dM.p, n_pairs, dA.p, dB.p, dD.p, dC.p, nullptr);
    cudaCheck(cudaGetLastError(), "mate_compat launch");
    cudaCheck(cudaDeviceSynchronize(), "mate_compat");
    dD.down(out_dist, n_pairs);  dC.down(out_compat, n_pairs);
  });
}

} // extern "C"
