// Tree engine: hierarchy bookkeeping + ordered resolution pass (see engine.h).
// Reference citations are relative to /root/reference/src/core/tree unless a path is given.
#include "engine.h"

#include <sched.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <exception>
#include <limits>

namespace apt {

namespace {
constexpr uint32_t INV = APT_INVALID;
// Scores that were not produced with reference-identical arithmetic (device exp(), tree-order
// float-vector distances, float-rounded distances) are only acted upon outside this band; inside
// it the host recomputes them exactly.  The error of those scores is bounded by a few float ulps
// of a value <= 1 (~1e-7), two orders of magnitude below the band.
constexpr float GUARD = 1e-5f;

double now (void) {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}

WorkerPool::WorkerPool (unsigned n) : n_(n ? n : 1) {
  for (unsigned t = 1; t < n_; t++) threads_.emplace_back([this, t] { loop(t); });
}

WorkerPool::~WorkerPool (void) {
  { std::lock_guard<std::mutex> lk (m_); stop_ = true; }
  wake_.notify_all();
  for (auto &th: threads_) th.join();
}

void WorkerPool::loop (unsigned t) {
  uint64_t seen = 0, seenHint = 0;
  for (;;) {
    // phases of one chunk follow each other within microseconds: spin briefly before sleeping
    bool got = false;
    // ... a few microseconds of pure spinning, then yielding between polls: a worker that only spins keeps its core
    // from the driver's and NCCL's threads when every core of the host has a spinner on it (several ranks per host),
    // and the thread that loses its time slice (1.5-4 ms) may be the one everybody waits for
    for (int spin = 0; spin < 6000; spin++) {
      if (gen_.load(std::memory_order_acquire) != seen) { got = true; break; }
      if (spin < 512) __builtin_ia32_pause(); else sched_yield();
    }
    if (!got) {
      std::unique_lock<std::mutex> lk (m_);
      wake_.wait(lk, [&] { return stop_ || gen_.load(std::memory_order_acquire) != seen || hint_ != seenHint; });
      if (stop_) return;
      seenHint = hint_;
      if (gen_.load(std::memory_order_acquire) == seen) continue;     // woken ahead of the work (wake()): spin for it
    }
    seen = gen_.load(std::memory_order_acquire);
    try { (*fn_)(t); } catch (...) { workerFailed_.store(true, std::memory_order_release); }
    if (pending_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
      std::lock_guard<std::mutex> lk (m_);
      done_.notify_one();
    }
  }
}

/// Work is coming within microseconds: sleeping workers get up now and spin, so that their wake-up latency
/// (tens of microseconds) overlaps whatever the caller still does before the first run()
void WorkerPool::wake (void) {
  if (n_ == 1) return;
  { std::lock_guard<std::mutex> lk (m_); hint_++; }
  wake_.notify_all();
}

void WorkerPool::run (const std::function<void(unsigned)> &fn) {
  if (n_ == 1) { fn(0); return; }
  {
    std::lock_guard<std::mutex> lk (m_);
    fn_ = &fn;
    pending_.store(n_ - 1, std::memory_order_release);
    gen_.fetch_add(1, std::memory_order_acq_rel);
  }
  wake_.notify_all();
  // the workers run *fn_ = the caller's function object: it must outlive them even when the caller's own share throws
  std::exception_ptr failed;
  try { fn(0); } catch (...) { failed = std::current_exception(); }
  bool done = false;
  for (int spin = 0; spin < 20000 && !done; spin++) {
    done = pending_.load(std::memory_order_acquire) == 0;
    if (!done) { if (spin < 4096) __builtin_ia32_pause(); else sched_yield(); }
  }
  if (!done) {
    std::unique_lock<std::mutex> lk (m_);
    done_.wait(lk, [&] { return pending_.load(std::memory_order_acquire) == 0; });
  }
  if (failed) std::rethrow_exception(failed);
  if (workerFailed_.exchange(false)) throw ApiError(APT_ERR_STATE, "a worker thread failed");
}

double allocClock (void) { return now(); }

static std::atomic<double> g_allocSeconds { 0. };
double allocSeconds (void) { return g_allocSeconds.load(std::memory_order_relaxed); }

void noteAlloc (const char *kind, size_t bytes, double t0) {
  static const bool on = getenv("APT_PROFILE") != nullptr;
  const double dt = now() - t0;
  double cur = g_allocSeconds.load(std::memory_order_relaxed);
  while (!g_allocSeconds.compare_exchange_weak(cur, cur + dt, std::memory_order_relaxed)) {}
  if (on) fprintf(stderr, "alloc: %s %zu bytes in %.3f ms\n", kind, bytes, 1e3 * dt);
}

void cudaCheck (cudaError_t e, const char *what) {
  if (e != cudaSuccess)
    throw ApiError(APT_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

// ---------------------------------------------------------------------------------------------
Engine::Engine (const apt_params &p) : p_(p) {
  if (p.kind > APT_F32) throw ApiError(APT_ERR_ARGUMENT, "unknown genome kind");
  if (p.words == 0 || p.words > 1024) throw ApiError(APT_ERR_ARGUMENT, "words must be in 1..1024");
  if (p.rset_size == 0 || p.rset_size > KMAX) throw ApiError(APT_ERR_ARGUMENT, "rset_size must be in 1..32");
  if (p.env_crit < 0 || p.env_crit > 3) throw ApiError(APT_ERR_ARGUMENT, "No function for this use case");
  K_ = p.rset_size;  W_ = p.words;  wp_ = (W_ + 3u) & ~3u;
  lutMode_ = (p.kind == APT_BITS) && !(p.flags & APT_FLAG_FORCE_GENERIC);
  lutStride_ = 32u * W_ + 1u;

  int count = 0;
  cudaCheck(cudaGetDeviceCount(&count), "cudaGetDeviceCount");
  if (p.device < 0 || p.device >= count) throw ApiError(APT_ERR_CUDA, "no such CUDA device");
  cudaCheck(cudaSetDevice(p.device), "cudaSetDevice");
  cudaDeviceProp prop;
  cudaCheck(cudaGetDeviceProperties(&prop, p.device), "cudaGetDeviceProperties");
  if (prop.major != 10)
    throw ApiError(APT_ERR_CUDA, "libapt_b200 holds sm_100a code only; device is sm_" +
                   std::to_string(prop.major) + std::to_string(prop.minor));
  cudaCheck(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking), "cudaStreamCreate");
  cudaCheck(cudaEventCreateWithFlags(&evUser_, cudaEventDisableTiming), "cudaEventCreate");
  cudaCheck(cudaEventCreate(&evK0_), "cudaEventCreate");
  cudaCheck(cudaEventCreate(&evK1_), "cudaEventCreate");
  cudaCheck(cudaEventCreate(&evS0_), "cudaEventCreate");
  cudaCheck(cudaEventCreate(&evS1_), "cudaEventCreate");
  cudaCheck(cudaEventCreate(&evF0_), "cudaEventCreate");
  cudaCheck(cudaEventCreate(&evF1_), "cudaEventCreate");
  cudaCheck(cudaEventCreateWithFlags(&evG_, cudaEventDisableTiming), "cudaEventCreate");
  cudaCheck(cudaStreamCreateWithFlags(&streamG_, cudaStreamNonBlocking), "cudaStreamCreate");
  // the device resolves streams for the default criteria; vote-count scoring needs the exact table
  // (a flipped vote cannot be seen from the score, so no guard band protects it)
  childKernel_ = !(p.flags & APT_FLAG_NO_CHILD_KERNEL);
  {
    unsigned hw = std::thread::hardware_concurrency();
    // measured on a 16-core B200 host (bench.py, 100 000 births per call): 4 threads 1.19 ms/step, 8: 1.05, 12: 0.95, 16: 1.18
    unsigned nt = hw >= 16 ? 12 : (hw >= 8 ? hw / 2 : (hw >= 4 ? 2 : 1));
    if (const char *e = getenv("APT_HOST_THREADS")) nt = unsigned(std::max(1, atoi(e)));
    nt = std::min(nt, 64u);
    if (const char *e = getenv("APT_PAR_MIN")) parMin_ = uint32_t(std::max(1, atoi(e)));
    if (const char *e = getenv("APT_STAGE_MIN")) stageMin_ = size_t(std::max(1, atoi(e)));
    profile_ = getenv("APT_PROFILE") != nullptr;
    if (getenv("APT_SUBTREE_WALK")) ageOrdered_ = false;     // (tests: keep the reference's walk in elligibile())
    if (const char *e = getenv("APT_STREAM_PROF")) streamProf_ = e;
    if (const char *e = getenv("APT_DEVLIST_MIN")) devListMin_ = uint32_t(std::max(1, atoi(e)));
    if (const char *e = getenv("APT_DEVGROUP_MIN")) devGroupMin_ = uint32_t(std::max(1, atoi(e)));
    devGroupPageable_ = getenv("APT_DEVGROUP_PAGEABLE") != nullptr;
    if (const char *e = getenv("APT_LUT_MAX_BYTES")) lutMaxBytes_ = size_t(std::max(0ll, atoll(e)));
    if (nt > 1) pool_.reset(new WorkerPool(nt));
  }
  streamOk_ = !(p.flags & APT_FLAG_NO_STREAM_KERNEL) && stream_kernel_supports(wp_, K_, p.env_crit)
              && (p.full_continuous || lutMode_);
  famOk_ = streamOk_ && childKernel_ && !(p.flags & APT_FLAG_NO_FAMILY_KERNEL) && family_kernel_supports(wp_, K_, p.env_crit);
  // The runtime loads a kernel when it is first used (milliseconds for the big ones): the variants this tree's
  // configuration selects are loaded now, not inside the first generations' passes.
  if (streamOk_) {
    StreamArgs a {};
    a.pool = poolView();  a.cfg = scoreCfg();  a.crit = p.env_crit;  a.preload = 1;
    a.max_stream = 0;     launch_resolve_streams(a, stream_);
    a.max_stream = 4096;  launch_resolve_streams(a, stream_);          // (the wide variant of long streams)
  }
  if (famOk_) {
    FamilyArgs f {};
    f.pool = poolView();  f.cfg = scoreCfg();  f.crit = p.env_crit;  f.preload = 1;  f.cap = 12;
    f.n_fam = 1;     launch_resolve_families(f, stream_);               // (few families: 256 threads, clusters)
    f.n_fam = 1000;  launch_resolve_families(f, stream_);
  }
  cudaGetLastError();
}

Engine::~Engine (void) {
  if (profile_)
    fprintf(stderr, "serial commit: plain entries %llu in %.3f s, rejected at top level %llu in %.3f s, other %llu in %.3f s\n",
            (unsigned long long) profWalkN_[0], profWalk_[0], (unsigned long long) profWalkN_[1], profWalk_[1],
            (unsigned long long) profWalkN_[2], profWalk_[2]);
  for (Species *s: nodes_) delete s;
  for (Species *s: graveyard_) delete s;
  if (evUser_) cudaEventDestroy(evUser_);
  if (evK0_) cudaEventDestroy(evK0_);
  if (evK1_) cudaEventDestroy(evK1_);
  if (evS0_) cudaEventDestroy(evS0_);
  if (evS1_) cudaEventDestroy(evS1_);
  if (evF0_) cudaEventDestroy(evF0_);
  if (evF1_) cudaEventDestroy(evF1_);
  if (evG_) cudaEventDestroy(evG_);
  if (streamG_) cudaStreamDestroy(streamG_);
  if (stream_) cudaStreamDestroy(stream_);
}

// phylogenetictree.hpp:168-177
Species* Engine::nodeAt (uint32_t sid) const {
  if (sid == INV) throw ApiError(APT_ERR_ARGUMENT, "SID::INVALID is (by definition) invalid");
  if (sid >= nodes_.size() || !nodes_[sid])
    throw ApiError(APT_ERR_ARGUMENT, "No node found for species " + std::to_string(sid));
  return nodes_[sid];
}

// ---- contributors / parents -------------------------------------------------------------------

// speciescontributors.h:177-197
bool Engine::elligibile (uint32_t lhs, uint32_t rhs) const {
  const Species *n = nodes_.at(lhs);
  if (rhs >= nodes_.size() || !nodes_[rhs]) return false;      // removed species are ignored
  const Species *q = nodes_[rhs];
  if (n->first <= q->first) return false;                       // no younger (or same-age) parent
  // The walk below cannot find n while every contributor flagged eligible is older than its species (ageOrdered_):
  // parents are picked among those (currentMain), so `first` strictly decreases along q's ancestors, all of them older
  // than q, which is older than n.  Flags come from this function or from a JSON dump (checked on load), `first` never
  // changes.  A species that is being created has no subtree either way.  The walk is a chain of cold cache lines as long
  // as the lineage is deep -- the cost of founding a species grew with the age of the simulation.
  if (ageOrdered_ || n == freshNode_) return true;
  while (q && q != n) q = q->parent;                            // not inside n's subtree
  return q != n;
}

// speciescontributors.cpp:67-83
uint32_t Engine::currentMain (Species *n) const {
  const auto &v = n->contribs;
  if (v.empty()) return INV;
  size_t i = 1;
  Contributor mc = v[0];
  while (!mc.elligible && i < v.size()) mc = v[i++];
  return mc.elligible ? mc.sid : INV;
}

// speciescontributors.cpp:15-65
uint32_t Engine::contributorsUpdate (Species *n, const Contributions &ctbs) {
  // Same result as the reference's copy / erase / stable_sort, without touching the heap: a contribution list has one
  // or two entries, a contributor list a handful.  `done` marks the contributions already added to an existing entry
  // (the reference erases them from its copy); INVALID species are dropped first (:27-28).
  auto &v = n->contribs;
  uint64_t done = 0;
  size_t unprocessed = ctbs.size();                      // (counted before the INVALID entries are dropped, as in :19-20)
  if (ctbs.size() > 64) throw ApiError(APT_ERR_STATE, "contribution list too long");
  for (size_t c = 0; c < ctbs.size(); c++) if (ctbs[c].species == INV) done |= (uint64_t(1) << c);
  for (size_t i = 0; i < v.size() && unprocessed > 0; i++) {
    const uint32_t sid = v[i].sid;
    for (size_t c = 0; c < ctbs.size(); c++)
      if (!((done >> c) & 1u) && ctbs[c].species == sid) {
        v[i].count += ctbs[c].count;
        unprocessed--;
        done |= (uint64_t(1) << c);
        break;
      }
  }
  for (size_t c = 0; c < ctbs.size(); c++)
    if (!((done >> c) & 1u)) v.push_back({ctbs[c].species, ctbs[c].count, elligibile(n->sid, ctbs[c].species)});
  // stable sort by decreasing count (:62): insertion sort keeps equal counts in their order
  for (size_t i = 1; i < v.size(); i++) {
    const Contributor x = v[i];
    size_t j = i;
    while (j > 0 && v[j - 1].count < x.count) { v[j] = v[j - 1]; j--; }
    v[j] = x;
  }
  return currentMain(n);
}

// node.hpp:239-241
Species* Engine::updateParent (Species *n, uint32_t sid) {
  if (sid == INV) return n->parent = nullptr;
  if (sid >= nodes_.size() || !nodes_[sid])     // std::map::at would throw in the reference
    throw ApiError(APT_ERR_STATE, "major contributor " + std::to_string(sid) + " no longer exists");
  return n->parent = nodes_[sid];
}

// node.hpp:182-185
Species* Engine::nodeUpdate (Species *n, const Contributions &c) { return updateParent(n, contributorsUpdate(n, c)); }

// phylogenetictree.hpp:773-783 with node.hpp:191-194 and speciescontributors.cpp:85-90
void Engine::updateElligibilities (void) {
  // With ageOrdered_ (see elligibile) a flag is "the contributor exists and is older than the species": the ages never
  // change, every flag was computed by that rule when its entry was made, and every contributors update re-derives the
  // parent (nodeUpdate).  Until a species is REMOVED (stillborn trimming leaves flags naming it behind, as in the
  // reference) or flags arrive from a dump, this pass over every species of the tree would rewrite what is there -- at
  // a million species it was 75 ms per change of a major contributor.
  if (ageOrdered_ && !eligStale_) return;
  for (Species *n: nodes_) {
    if (!n) continue;
    for (Contributor &c: n->contribs) c.elligible = elligibile(n->sid, c.sid);
    updateParent(n, currentMain(n));
  }
  eligStale_ = false;
}

/// sidSlot_ mirrors "the device holds a current copy of this species" (dev_fill > 0)
void Engine::setStreamable (const Species *s, bool on) {
  if (sidSlot_.size() <= s->sid) sidSlot_.resize(std::max<size_t>(s->sid + 1, sidSlot_.size() * 2), INV);
  const uint32_t v = on ? s->slot : INV;
  if (sidSlot_[s->sid] != v) { sidSlot_[s->sid] = v;  sidSlotDirty_ = true; }
}

uint32_t Engine::allocSlot (void) {
  uint32_t slot;
  if (forcedSlot_ != INV) { slot = forcedSlot_; forcedSlot_ = INV; return slot; }      // (a slot of the block the family pass reserved)
  if (!freeSlots_.empty()) { slot = freeSlots_.back(); freeSlots_.pop_back(); }
  else {
    slot = nSlots_++;
    hPayload_.resize(size_t(nSlots_) * K_ * wp_, true);
  }
  return slot;
}

// phylogenetictree.hpp:398-426
Species* Engine::makeNode (const Contributions &contrib) {
  Species *n = new Species;
  n->sid = nextNodeID_++;
  n->slot = allocSlot();
  n->reps.resize(K_);
  n->dist.assign(size_t(K_) * K_, 0.f);
  n->first = n->last = step_;
  touchSnap(n);                                              // k = 0: nothing of it is on the device
  if (nodes_.size() <= n->sid) nodes_.resize(std::max<size_t>(n->sid + 1, nodes_.size() * 2), nullptr);
  nodes_[n->sid] = n;
  width_++;
  freshNode_ = n;
  nodeUpdate(n, contrib);
  freshNode_ = nullptr;
  if (n->parent) n->parent->children.push_back(n);
  emit(APT_EV_NEW_SPECIES, n->parent ? n->parent->sid : INV, n->sid, 0);
  return n;
}

// phylogenetictree.hpp:741-769
void Engine::updateContributions (Species *s, const Contributions &contrib) {
  Species *oldMC = s->parent, *newMC = nodeUpdate(s, contrib);
  if (oldMC != newMC) {
    if (!newMC) throw ApiError(APT_ERR_STATE, "species lost its major contributor");   // assert(newMC), :751
    if (oldMC) {
      auto &c = oldMC->children;
      c.erase(std::remove(c.begin(), c.end(), s), c.end());
      oldMC->child_change_epoch = chunkEpoch_;
    }
    newMC->children.push_back(s);
    newMC->child_change_epoch = chunkEpoch_;
    updateElligibilities();
    // the reference dereferences a null oldMC here (:765-766); INVALID is reported instead
    emit(APT_EV_MAJOR_CONTRIBUTOR, s->sid, oldMC ? oldMC->sid : INV, newMC->sid);
  }
}

// phylogenetictree.hpp:816-860
void Engine::performStillbornTrimming (void) {
  const float T = p_.trim_thr, D = p_.trim_delay, MD = float(p_.trim_min_delay);
  for (size_t i = 0; i < nodes_.size(); i++) {
    Species *s = nodes_[i];
    if (!s || !s->children.empty() || !s->extinct()) continue;
    const bool underfilled = (s->k < T * K_);
    const uint32_t liveTime = s->last - s->first, deadTime = step_ - s->last;
    if (underfilled && std::max(MD, liveTime * D) < deadTime) {
      if (s->parent) {
        auto &c = s->parent->children;
        c.erase(std::remove(c.begin(), c.end(), s), c.end());
      }
      stillborns_++;
      nodes_[i] = nullptr;
      eligStale_ = true;
      setStreamable(s, false);
      width_--;
      freeSlots_.push_back(s->slot);
      fillUpdates_.push_back(s->slot);  fillUpdates_.push_back(0);
      trimmedSids_.push_back(s->sid);
      s->trimmed = true;  s->stale_mask = 0;                // nothing left to refresh
      graveyard_.push_back(s);                              // `root_` may still point at it, as in the reference
    }
  }
}

// phylogenetictree.hpp:245-262
void Engine::doStep (uint32_t s, const uint32_t *sids, uint32_t n) {
  aliveSpecies_.clear();
  for (uint32_t i = 0; i < n; i++) aliveSpecies_.insert(sids[i]);
  for (uint32_t sid: aliveSpecies_) nodeAt(sid)->last = s;
  step_ = s;
  if (p_.trim_period > 0 && (step_ % p_.trim_period) == 0) performStillbornTrimming();
  emit(APT_EV_STEPPED, s, uint32_t(aliveSpecies_.size()), 0);
}

// phylogenetictree.hpp:305-317
void Engine::parallelFor (uint32_t n, void (*fn)(void *ctx, uint32_t lo, uint32_t hi), void *ctx) {
  if (!n) return;
  if (!pool_ || n < parMin_) { fn(ctx, 0, n); return; }
  pool_->wake();
  const unsigned T = pool_->size();
  const uint32_t per = (n + T - 1) / T;
  pool_->run([&] (unsigned t) {
    const uint32_t lo = std::min(n, t * per), hi = std::min(n, lo + per);
    if (hi > lo) fn(ctx, lo, hi);
  });
}

void Engine::delGenome (uint32_t sid) {
  Species *n = nodeAt(sid);
  n->last = step_;
  n->alive--;
  deletions_++;
}

// phylogenetictree.hpp:786-792
void Engine::candidacy (uint32_t msid, uint32_t fsid, int dir) {
  if (msid != INV) nodeAt(msid)->pending += dir;
  if (msid != fsid && fsid != INV) nodeAt(fsid)->pending += dir;
}

void Engine::stats (uint32_t out[8]) const {
  out[0] = insertions_; out[1] = deletions_; out[2] = comparisons_; out[3] = branching_;
  out[4] = nextNodeID_; out[5] = width_; out[6] = step_; out[7] = stillborns_;
}

// ---- exact host arithmetic ----------------------------------------------------------------------

uint32_t Engine::classOf (float mu, float sin, float sout) {
  if (lastClass_ != INV && lastTriple_.mu == mu && lastTriple_.sin == sin && lastTriple_.sout == sout)
    return lastClass_;
  const Triple t { mu, sin, sout };
  auto it = classIds_.find(t);
  uint32_t id;
  if (it != classIds_.end()) id = it->second;
  else {
    id = uint32_t(classes_.size());
    classIds_.emplace(t, id);
    classes_.push_back(t);
    if (lutMode_) {
      // one table row per class, filled by the HOST libm: the device reads back the very doubles
      // the reference would compute (crossover.h:89-92)
      if (size_t(id + 1) * lutStride_ * sizeof(double) > lutMaxBytes_) {
        // too many distinct triples for a table: device exp() + guard band from here on.  Vote-count scoring has no
        // guard band (a flipped vote cannot be seen from the score): without the exact table neither the stream
        // kernel nor the sub-species kernel may decide such births any more -- the host does, with libm.
        lutMode_ = false;
        hLut_.clear();  hLut_.shrink_to_fit();
        streamOk_ = streamOk_ && p_.full_continuous;
        childKernel_ = childKernel_ && p_.full_continuous;
        famOk_ = famOk_ && p_.full_continuous;
      } else {
        hLut_.resize(size_t(id + 1) * lutStride_);
        for (uint32_t d = 0; d < lutStride_; d++) {
          const double dd = d, sigma = dd < mu ? sin : sout;
          hLut_[size_t(id) * lutStride_ + d] = std::exp(-((dd - mu) * (dd - mu)) / (2. * sigma * sigma));
        }
      }
    }
  }
  lastTriple_ = t;  lastClass_ = id;
  return id;
}

// crossover.h:89-92
double Engine::compatOf (uint32_t cls, float mu, float sin, float sout, double d) const {
  if (lutMode_) return hLut_[size_t(cls) * lutStride_ + uint32_t(d)];
  const double sigma = d < mu ? sin : sout;
  return std::exp(-((d - mu) * (d - mu)) / (2. * sigma * sigma));
}

/// The toy genomes' distance(): Hamming, or L2 accumulated in double in index order
double Engine::distanceExact (const uint32_t *a, const uint32_t *b) const {
  if (p_.kind == APT_F32) {
    double s = 0;
    for (uint32_t i = 0; i < W_; i++) {
      float fa, fb;
      memcpy(&fa, a + i, 4);  memcpy(&fb, b + i, 4);
      const double t = double(fa) - double(fb);
      s += t * t;
    }
    return std::sqrt(s);
  }
  uint32_t s = 0, i = 0;
  for (; i + 2 <= W_; i += 2) {
    uint64_t x, y;
    memcpy(&x, a + i, 8);  memcpy(&y, b + i, 8);
    s += uint32_t(__builtin_popcountll(x ^ y));
  }
  if (i < W_) s += uint32_t(__builtin_popcount(a[i] ^ b[i]));
  return double(s);
}

// phylogenetictree.hpp:465-477 / :439-451
float Engine::scoreFrom (const double *c, uint32_t k) const {
  if (p_.full_continuous) {
    float avg = 0;
    for (uint32_t r = 0; r < k; r++) avg += c[r];
    return avg / float(k) - p_.avg_thr;
  }
  uint32_t matable = 0;
  for (uint32_t r = 0; r < k; r++) if (c[r] >= p_.compat_thr) matable++;
  return matable - p_.sim_thr * k;
}

bool Engine::trusted (float score) const {
  if (lutMode_) return true;
  if (!p_.full_continuous) return false;          // vote counts can flip without moving the score
  return std::fabs(score) > GUARD;
}

/// speciesMatchingScore (:430-490) entirely on the host, reference-identical arithmetic
void Engine::evalExact (const uint32_t *w, const RepMeta &g, Species *s, Eval &e) {
  double c[KMAX];
  for (uint32_t r = 0; r < s->k; r++) {
    const RepMeta &m = s->reps[r];
    const double d = distanceExact(w, repPayload(s, r));
    c[r] = std::min(compatOf(g.cls, g.mu, g.sin, g.sout, d), compatOf(m.cls, m.mu, m.sin, m.sout, d));
    e.d[r] = float(d);
  }
  counters[5] += s->k;
  e.score = scoreFrom(c, s->k);
  e.exact = true;
}

/// Same score, using what the device computed for this (birth, species) pair wherever the
/// enveloppe still is what the device saw when the chunk was submitted.
void Engine::evalSpeculated (const uint32_t *w, const RepMeta &g, Species *s, int64_t pair, Eval &e) {
  comparisons_ += s->k;                                     // stats.comparisons++, :444 / :470
  touchSnap(s);
  if (pair < 0) { counters[3]++; return evalExact(w, g, s, e); }
  const float *dd = hDist_.p + size_t(pair) * K_;
  const bool bits = (p_.kind == APT_BITS);
  if (s->snap_mask == 0 && s->snap_fill == s->k) {
    const float sc = hScore_.p[pair];
    if (trusted(sc)) {
      memcpy(e.d, dd, s->k * sizeof(float));
      e.score = sc;  e.exact = lutMode_;
      counters[1]++;
      return;
    }
    counters[6]++;
    if (!bits) { counters[3]++; return evalExact(w, g, s, e); }
  } else if (!bits) {
    // float-vector distances cannot be patched exactly from their float roundings
    counters[3]++;
    return evalExact(w, g, s, e);
  }
  // bit-strings: device distances are exact integers; re-measure only the changed representatives
  double c[KMAX];
  uint32_t measured = 0;
  for (uint32_t r = 0; r < s->k; r++) {
    const RepMeta &m = s->reps[r];
    double d;
    if (r < s->snap_fill && !((s->snap_mask >> r) & 1u)) d = dd[r];
    else { d = distanceExact(w, repPayload(s, r)); measured++; }
    c[r] = std::min(compatOf(g.cls, g.mu, g.sin, g.sout, d), compatOf(m.cls, m.mu, m.sin, m.sout, d));
    e.d[r] = float(d);
  }
  counters[5] += measured;
  counters[2]++;
  e.score = scoreFrom(c, s->k);
  e.exact = true;
}

/// First use of a species in the current chunk: remember what the device snapshot holds of it
void Engine::touchSnap (Species *s) {
  if (s->snap_epoch != chunkEpoch_) {
    s->snap_epoch = chunkEpoch_;
    s->snap_mask = 0;
    s->snap_fill = s->k;
  }
}

/// The host changed representative r: the device row is out of date, and a device-resolved stream
/// of this species (if any) no longer describes what happens to its later births.
void Engine::markStale (Species *s, uint32_t r) {
  touchSnap(s);
  s->snap_mask |= (1u << r);
  s->stale_mask |= (1u << r);
  queueRefresh(s);
  if (s->grp_epoch == chunkEpoch_) diverge(s);
}

/// The host takes the stream of s over: device results for its remaining births are discarded and
/// what the kernel wrote back for it is overwritten at the next refresh.
void Engine::diverge (Species *s) {
  if (s->div_epoch == chunkEpoch_) return;
  s->div_epoch = chunkEpoch_;
  s->full_upload = true;
  queueRefresh(s);
  counters[11]++;
  if (s->grp_epoch == chunkEpoch_ && !grpState_.empty()) settleStream(s->grp_index, true, curBirth_, &requeue_);
}

/// Credits to s's own contributor entry what its accepted-unchanged stream births with index
/// < `before` contribute to it.  Needed right before ANOTHER species' contribution to s is
/// recorded: the stable sort of Contributors::update (speciescontributors.cpp:62) orders equal
/// counts by history, so the own entry must hold exactly the count it would have at that time.
void Engine::catchUpSelf (Species *s, uint32_t before) {
  if (s->grp_epoch != chunkEpoch_ || s->div_epoch == chunkEpoch_ || grpState_.empty()) return;
  const uint32_t g = s->grp_index;
  GroupState &st = grpState_[g];
  const uint32_t end = hGrpOff_.p[g + 1];
  uint32_t pos = hGrpOff_.p[g] + st.cursor;
  for (; pos < end; pos++) {
    const uint32_t ord = hOrder_.p[pos], b = ord & ORDER_INDEX_MASK;
    if (b >= before) break;
    if (hAct_.p[b] == ACT_ACCEPTED) st.scanned[ord >> ORDER_CLASS_SHIFT]++;
  }
  st.cursor = pos - hGrpOff_.p[g];
  const uint32_t due = st.scanned[1] + 2 * st.scanned[2];
  if (due > st.selfApplied) { selfContribution(s, due - st.selfApplied); st.selfApplied = due; }
}

/// Books the accepted-unchanged births of stream g that the host never looks at individually
/// (:687-691 nothing to do for the enveloppe; :721-723 counters; :741-769 own contribution).
/// At the end of a chunk: all of them (totals from the kernel).  When the host takes the stream
/// over at birth `before`: only those before it; the later ones are handed back through `requeue`.
void Engine::settleStream (uint32_t g, bool diverging, uint32_t before, std::vector<uint32_t> *requeue) {
  Species *s = grpSpecies_[g];
  GroupState &st = grpState_[g];
  uint32_t n[3];
  if (diverging) {
    const uint32_t end = hGrpOff_.p[g + 1];
    uint32_t pos = hGrpOff_.p[g] + st.cursor;
    for (; pos < end; pos++) {
      const uint32_t ord = hOrder_.p[pos], b = ord & ORDER_INDEX_MASK;
      if (hAct_.p[b] != ACT_ACCEPTED) continue;
      if (b < before) st.scanned[ord >> ORDER_CLASS_SHIFT]++;
      else if (requeue) requeue->push_back(b);
    }
    st.cursor = end - hGrpOff_.p[g];
    n[0] = st.scanned[0];  n[1] = st.scanned[1];  n[2] = st.scanned[2];
  } else {
    const uint4 sum = hGrpSum_.p[g];
    n[0] = sum.x;  n[1] = sum.y;  n[2] = sum.z;
  }
  const uint32_t total = n[0] + n[1] + n[2];
  if (!total) return;
  s->pending = s->pending > total ? s->pending - total : 0;          // :293
  s->count += total;  s->alive += total;  s->last = step_;            // :721-723
  comparisons_ += total * K_;                   // an accepted birth leaves the enveloppe alone only when it is full
  insertions_ += total;
  counters[0] += total;  counters[1] += total;
  const uint32_t due = n[1] + 2 * n[2];
  if (due > st.selfApplied) { selfContribution(s, due - st.selfApplied); st.selfApplied = due; }
}

// ---- enveloppe criteria (src/core/tree/enveloppecriteria.cpp) on the K x K matrix ---------------
namespace {
struct EC { bool better; uint32_t than; float value; };

inline float ninf (void) { return -std::numeric_limits<float>::infinity(); }   // -DBL_MAX into a float, :46/:93

// :87-133
EC maxMinDist (const float *e, const float *g, uint32_t k, uint32_t K) {
  EC ec { false, uint32_t(-1), ninf() };
  for (uint32_t i = 0; i < k; i++) {
    float minBase = std::numeric_limits<float>::max(), minNew = std::numeric_limits<float>::max();
    for (uint32_t j = 0; j < k; j++) {
      if (i == j) continue;
      minBase = std::min(minBase, e[i * K + j]);
      minNew = std::min(minNew, g[j]);
    }
    const double c = - minBase + minNew;
    if (ec.value < c) { ec.value = float(c); ec.than = i; }
  }
  ec.better = (ec.value > 0);
  return ec;
}

// :40-84
EC maxAverage (const float *e, const float *g, uint32_t k, uint32_t K) {
  EC ec { false, uint32_t(-1), ninf() };
  for (uint32_t i = 0; i < k; i++) {
    double c = 0;
    for (uint32_t j = 0; j < k; j++) {
      if (i == j) continue;
      c += - e[i * K + j] + g[j];
    }
    if (ec.value < c) { ec.value = float(c); ec.than = i; }
  }
  ec.better = (ec.value > 0);
  return ec;
}

// :29-37 over the map's iteration order (keys (i,j), i<j, lexicographic)
void avgAndStdDev (const float *m, uint32_t k, uint32_t K, double &avg, double &stdDev) {
  const double n = double(k) * (k - 1) / 2.;
  avg = 0;
  for (uint32_t i = 0; i < k; i++) for (uint32_t j = i + 1; j < k; j++) avg += m[i * K + j];
  avg /= n;
  stdDev = 0;
  for (uint32_t i = 0; i < k; i++) for (uint32_t j = i + 1; j < k; j++) stdDev += std::pow(avg - m[i * K + j], 2);
  stdDev = std::sqrt(stdDev / n);
}

// :136-202
EC maxAvgMinStdDev (const float *e, const float *g, uint32_t k, uint32_t K) {
  EC ec { false, uint32_t(-1), ninf() };
  double baseAVG, baseStdDev;
  avgAndStdDev(e, k, K, baseAVG, baseStdDev);
  std::vector<float> n (size_t(K) * K);
  for (uint32_t i = 0; i < k; i++) {
    std::copy(e, e + size_t(K) * K, n.begin());
    for (uint32_t j = 0; j < k; j++) if (i != j) n[i * K + j] = n[j * K + i] = g[j];
    double newAVG, newStdDev;
    avgAndStdDev(n.data(), k, K, newAVG, newStdDev);
    const double c = - baseAVG + newAVG + baseStdDev - newStdDev;
    if (ec.value < c) { ec.value = float(c); ec.than = i; }
  }
  ec.better = (ec.value > 0);
  return ec;
}

// :205-272 with its sizing quirk (vectors built with k zeros, then k-1 values appended, :231, :238-239)
EC maxWeightedDist2Avg (const float *e, const float *g, uint32_t k, uint32_t K) {
  EC ec { false, uint32_t(-1), ninf() };
  double A = 0;
  for (uint32_t i = 0; i < k; i++) for (uint32_t j = i + 1; j < k; j++) A += e[i * K + j];
  A /= double(k) * (k - 1) / 2.;
  auto weight = [A] (double d) { return 1 - exp(- (d-A)*(d-A) / (2. * A * A / 16.)); };
  for (uint32_t i = 0; i < k; i++) {
    std::vector<double> d_i (k), d_g (k);
    for (uint32_t j = 0; j < k; j++) {
      if (i == j) continue;
      d_i.push_back(e[i * K + j]);
      d_g.push_back(g[j]);
    }
    std::sort(d_i.begin(), d_i.end(), [] (double a, double b) { return a > b; });
    std::sort(d_g.begin(), d_g.end(), [] (double a, double b) { return a > b; });
    double c = 0;
    for (uint32_t j = 0; j + 1 < k; j++) {
      const double nc = - d_i[j], pc = + d_g[j];
      c += weight(pc) * (nc + pc);
    }
    if (ec.value < c) { ec.value = float(c); ec.than = i; }
  }
  ec.better = (ec.value > 0);
  return ec;
}

// :275-289
EC computeContribution (int crit, const float *e, const float *g, uint32_t k, uint32_t K) {
  switch (crit) {
  case 0: return maxAverage(e, g, k, K);
  case 1: return maxMinDist(e, g, k, K);
  case 2: return maxAvgMinStdDev(e, g, k, K);
  default: return maxWeightedDist2Avg(e, g, k, K);
  }
}
} // namespace

// phylogenetictree.hpp:657-726.  Returns the slot the genome now occupies, or -1.
int32_t Engine::insertInto (const uint32_t *w, const RepMeta &g, Species *s, const Eval &e, bool haveCache,
                            uint32_t &replaced) {
  const uint32_t k = s->k;
  int32_t slot = -1;
  replaced = INV;
  auto place = [&] (uint32_t r, uint32_t udata) {
    RepMeta m = g;
    m.timestamp = step_;
    m.udata = udata;
    s->reps[r] = m;
    uint32_t *dst = &hPayload_[(size_t(s->slot) * K_ + r) * wp_];
    memcpy(dst, w, W_ * sizeof(uint32_t));
    markStale(s, r);
  };
  if (k < K_) {
    if (k && !haveCache) throw ApiError(APT_ERR_STATE, "append without cached distances");
    place(k, g.gid);
    s->k = k + 1;
    emit(APT_EV_ENTERS_ENVELOPPE, s->sid, g.gid, 0);
    for (uint32_t i = 0; i < k; i++) s->dist[i * K_ + k] = s->dist[k * K_ + i] = e.d[i];
    slot = int32_t(k);
  } else {
    if (K_ == 1) {
      // a single representative has no pairwise distance: every criterion yields a value <= 0 or
      // NaN (minima stay at FLT_MAX -> c = 0, :108-117; empty sums; 0/0 averages) -> never better
    } else {
      const EC ec = computeContribution(p_.env_crit, s->dist.data(), e.d, k, K_);
      if (ec.better) {
        const uint32_t than = ec.than, epId = s->reps[than].gid;
        emit(APT_EV_LEAVES_ENVELOPPE, s->sid, epId, 0);
        emit(APT_EV_ENTERS_ENVELOPPE, s->sid, g.gid, 0);
        place(than, epId);                                   // *ep.userData = UserData(ep_id), :710
        for (uint32_t i = 0; i < k; i++)
          if (i != than) s->dist[i * K_ + than] = s->dist[than * K_ + i] = e.d[i];
        slot = int32_t(than);
        replaced = epId;
      }
    }
  }
  s->count++;
  s->alive++;
  s->last = step_;
  return slot;
}

/// findBestDerived (:493-548): direct sub-species, newest first, alternating between the lists;
/// the first score > 0 wins.
void Engine::resolveBelow (const uint32_t *w, const RepMeta &g, Species *s0, Species *s1, Species *&best,
                           float &bestScore, Eval &bestEval) {
  Species *lists[2] = { s0, s1 };
  const uint32_t S = s1 ? 2 : 1;
  size_t left[2] = { s0->children.size(), s1 ? s1->children.size() : 0 };
  uint32_t done = 0;
  const uint32_t allDone = (1u << S) - 1;
  uint32_t k = 0;
  Eval cur;
  while (done != allDone) {
    if (left[k] == 0) done |= (1u << k);
    else {
      branching_++;
      Species *sub = lists[k]->children[left[k] - 1];
      evalSpeculated(w, g, sub, -1, cur);
      if (bestScore < cur.score) { best = sub; bestScore = cur.score; bestEval = cur; }
      if (bestScore > 0) return;
      left[k]--;
    }
    k = (k + 1) % S;
  }
}

/// findBestDerived for a birth a device stream rejected at top level: the sub-species kernel already
/// walked the sub-species its species had when the chunk was submitted (newest first, stopping at
/// the first acceptance).  Its verdict on a sub-species is used when that enveloppe is, at this
/// point of the ordered commit, exactly what the kernel saw: untouched by the host, and not entered
/// by any later birth of its own stream.  Sub-species founded inside the chunk come first (:506) and
/// are scored here, as are those the kernel's view is not valid for.
void Engine::resolveBelowStream (const uint32_t *w, const RepMeta &g, Species *s0, int64_t ft, Species *&best,
                                 float &bestScore, Eval &bestEval) {
  const GroupState &st = grpState_[s0->grp_index];
  const bool dev = ft >= 0 && s0->child_change_epoch != chunkEpoch_;
  const size_t snapN = st.snapChildren;
  const uint32_t firstRank = dev ? hFtFirst_.p[ft] : 0u;
  const bool firstUnsure = dev && (hFtFlags_.p[ft] & 1u);
  Eval cur;
  size_t idx = s0->children.size();
  while (idx-- > 0) {
    Species *sub = s0->children[idx];
    // (the walk is a chain of cold records: the next two sub-species' are requested while this one is looked at, and
    // the stream state of the next one)
    if (idx >= 2) {
      const char *cp = reinterpret_cast<const char*>(s0->children[idx - 2]);
      __builtin_prefetch(cp);  __builtin_prefetch(cp + 64);  __builtin_prefetch(cp + 128);
    }
    if (idx >= 1) {
      const Species *nx = s0->children[idx - 1];
      if (nx->grp_epoch == chunkEpoch_ && nx->grp_index < grpState_.size()) {
        __builtin_prefetch(&grpState_[nx->grp_index]);  __builtin_prefetch(&hGrpSum_.p[nx->grp_index]);
      }
    }
    branching_++;
    bool decided = false;
    if (dev && idx < snapN) {
      const uint32_t rank = uint32_t(snapN - 1 - idx);
      const bool known = rank < firstRank || (rank == firstRank && !firstUnsure);
      // (a sub-species with a stream of its own: the kernel saw it as that stream left it -- all its entries applied.  A
      // stream the device handed back in the middle never wrote its enveloppe back: the verdict then rests on the state
      // BEFORE the pass although the host has committed that stream's earlier entries, and must not be used)
      if (known && sub->stale_mask == 0 && !sub->full_upload && sub->div_epoch != chunkEpoch_ && sub->fam_epoch != chunkEpoch_ &&
          (sub->grp_epoch != chunkEpoch_ || (grpState_[sub->grp_index].lastEnter < int64_t(curBirth_) && !hGrpSum_.p[sub->grp_index].w))) {
        touchSnap(sub);
        comparisons_ += sub->k;
        counters[1]++;
        decided = true;
        if (rank == firstRank) {                            // first acceptance: done (:540-541)
          best = sub;  bestScore = hFtScore_.p[ft];
          memcpy(bestEval.d, hFtDist_.p + size_t(ft) * K_, sub->k * sizeof(float));
          bestEval.score = bestScore;  bestEval.exact = lutMode_;
          return;
        }
      }
    }
    if (!decided) {
      evalSpeculated(w, g, sub, -1, cur);
      if (bestScore < cur.score) { best = sub; bestScore = cur.score; bestEval = cur; }
      if (bestScore > 0) return;
    }
  }
}

/// :612-643: insert into the best species, or found a new one
void Engine::finishBirth (const uint32_t *w, RepMeta &g, Species *best, float bestScore, const Eval &bestEval,
                          const Contributions &contrib, uint32_t &sid, int32_t &slot, uint32_t &replaced) {
  if (bestScore > 0) {
    slot = insertInto(w, g, best, bestEval, true, replaced);
    catchUpSelf(best, curBirth_);
    updateContributions(best, contrib);
    sid = best->sid;
  } else {
    Species *sub = makeNode(contrib);                       // simpleNewSpecies, :637-643
    Eval none;
    slot = insertInto(w, g, sub, none, false, replaced);
    sid = sub->sid;
  }
  insertions_++;
}

// One birth on the host, in order: public addGenome (:268-302) + private addGenome (:553-650) +
// findBestDerived (:493-548) + updateSpeciesContents (:730-738).
void Engine::walkOne (const uint32_t *w, RepMeta &g, int64_t p0, int64_t p1, uint32_t &sid, int32_t &slot,
                      uint32_t &replaced) {
  const uint32_t mSID = g.msid, fSID = g.fsid;
  Species *s0 = nullptr, *s1 = nullptr;
  if (mSID == INV && fSID == INV) s0 = root_;
  else if (fSID == INV || mSID == fSID) s0 = nodeAt(mSID);
  else { s0 = nodeAt(mSID); s1 = nodeAt(fSID); }
  if (s0->pending > 0) s0->pending--;
  if (s1 && s1->pending > 0) s1->pending--;

  Eval ev0, ev1, bestEval;
  Species *best = nullptr;
  float bestScore = -std::numeric_limits<float>::max();

  Contributions contrib { { mSID, 1u + (mSID == fSID) } };
  if (s1) contrib.push_back({ fSID, 1 });

  evalSpeculated(w, g, s0, p0, ev0);
  if (s1) {
    evalSpeculated(w, g, s1, p1, ev1);
    // decisions that compare the two scores (:590, :599-602) need them exact when they are close
    if ((!ev0.exact || !ev1.exact) && std::fabs(ev0.score - ev1.score) <= GUARD) {
      if (!ev0.exact) evalExact(w, g, s0, ev0);
      if (!ev1.exact) evalExact(w, g, s1, ev1);
    }
  }
  if (bestScore < ev0.score) { best = s0; bestScore = ev0.score; }
  if (s1) {
    if (bestScore < ev1.score) { best = s1; bestScore = ev1.score; }
    // std::sort with `>=` on two elements (:599-602): the second goes first iff score1 >= score0
    if (ev1.score >= ev0.score) std::swap(contrib[0], contrib[1]);
  }
  if (best) bestEval = (best == s1 && s1) ? ev1 : ev0;

  if (!(bestScore > 0)) {
    // a single-parent birth whose stream was handed back may still have a sub-species verdict from the
    // device (it depends on the birth and on the sub-species only, not on what became of the stream)
    if (!s1 && curFt_ >= 0 && s0->grp_epoch == chunkEpoch_ && !grpState_.empty())
      resolveBelowStream(w, g, s0, curFt_, best, bestScore, bestEval);
    else resolveBelow(w, g, s0, s1, best, bestScore, bestEval);
  }
  finishBirth(w, g, best, bestScore, bestEval, contrib, sid, slot, replaced);
}

/// Contributors::update (speciescontributors.cpp:15-65) for the commonest case, a species
/// contributing `count` to ITSELF: its own entry is never eligible (:186-187), so moving it cannot
/// change the major contributor; stable_sort after one increment = bubble past strictly smaller counts.
void Engine::selfContribution (Species *s, uint32_t count) {
  auto &v = s->contribs;
  size_t i = s->self_idx;
  if (i >= v.size() || v[i].sid != s->sid) {
    i = 0;
    while (i < v.size() && v[i].sid != s->sid) i++;
    if (i == v.size()) { updateContributions(s, { { s->sid, count } }); return; }
  }
  v[i].count += count;
  while (i > 0 && v[i - 1].count < v[i].count) { std::swap(v[i - 1], v[i]); i--; }
  s->self_idx = uint32_t(i);
}

/// Mirrors one enveloppe entry decided by the device (append :668-677 / replacement :694-718) into
/// the host copy of species s0.  Touches s0 only (safe to run for different species in parallel);
/// the callbacks to fire are returned in `ev` (at most two, in the reference's order).
int Engine::mirrorEntry (const uint32_t *w, const RepMeta &g, Species *s0, uint32_t act, int32_t &slot,
                         uint32_t &replaced, apt_event ev[2]) {
  int ne = 0;
  const uint32_t r = (act >> ACT_SLOT_SHIFT) & 31u, k = s0->k;
  const float *d = hRows_.p + size_t(act >> ACT_ROW_SHIFT) * K_;
  RepMeta m = g;
  m.timestamp = step_;
  if (act & ACT_REPLACED) {
    const uint32_t epId = s0->reps[r].gid;
    ev[ne++] = { APT_EV_LEAVES_ENVELOPPE, s0->sid, epId, 0 };
    ev[ne++] = { APT_EV_ENTERS_ENVELOPPE, s0->sid, g.gid, 0 };
    m.udata = epId;                                       // :710
    replaced = epId;
  } else {
    if (r != k) throw ApiError(APT_ERR_STATE, "device/host enveloppe fill mismatch");
    ev[ne++] = { APT_EV_ENTERS_ENVELOPPE, s0->sid, g.gid, 0 };
    m.udata = g.gid;
    s0->k = k + 1;
  }
  s0->reps[r] = m;
  memcpy(&hPayload_[(size_t(s0->slot) * K_ + r) * wp_], w, W_ * sizeof(uint32_t));
  for (uint32_t i = 0; i < k; i++)
    if (i != r) s0->dist[i * K_ + r] = s0->dist[r * K_ + i] = d[i];
  s0->snap_mask |= (1u << r);
  s0->dev_fill = s0->k;                                   // the kernel wrote the same change to the pool
  slot = int32_t(r);
  return ne;
}

// One birth whose top-level decision was taken by the device (resolve_streams_kernel): the host
// mirrors the decision into the tree (counters, representative metadata, distance rows,
// contributions, events), exactly as insertInto (:657-726) + updateContributions (:741-769) would.
void Engine::walkStream (const uint32_t *w, RepMeta &g, Species *s0, uint32_t act, uint32_t &sid, int32_t &slot,
                         uint32_t &replaced) {
  if (s0->pending > 0) s0->pending--;
  touchSnap(s0);
  comparisons_ += s0->k;
  const uint32_t code = act & ACT_CODE_MASK;
  if (code == ACT_FALLTHROUGH) {
    Species *best = nullptr;
    Eval bestEval;
    float bestScore = 0.f;          // any value <= 0: the actual top-level score only matters by its sign (:612, :534)
    resolveBelowStream(w, g, s0, curFt_, best, bestScore, bestEval);
    Contributions contrib { { g.msid, 1u + (g.msid == g.fsid) } };
    finishBirth(w, g, best, bestScore, bestEval, contrib, sid, slot, replaced);
    return;
  }
  counters[1]++;
  slot = -1;  replaced = INV;
  if (code == ACT_ENTERED) {
    counters[14]++;
    apt_event ev[2];
    const int ne = mirrorEntry(w, g, s0, act, slot, replaced, ev);
    for (int e = 0; e < ne; e++) emit(ev[e].type, ev[e].a, ev[e].b, ev[e].c);
  }
  s0->count++;
  s0->alive++;
  s0->last = step_;
  if (g.msid != INV) selfContribution(s0, 1u + (g.msid == g.fsid));
  // (a primordial birth contributes (INVALID, 2), which Contributors::update discards, :27-28)
  sid = s0->sid;
  insertions_++;
}

/// Contributors::update (speciescontributors.cpp:15-65) for n1 births contributing (from, 1) and n2 contributing
/// (from, 2) to species c, none of which changes its enveloppe: the same entry grows by n1 + 2 n2, and a stable sort
/// after every single increment leaves it where one sort after the sum does (it only ever passes strictly smaller
/// counts).  `from` is c's major contributor (c is one of its sub-species), so the parent cannot change.
void Engine::bulkContribution (Species *c, Species *from, uint32_t n1, uint32_t n2) {
  auto &v = c->contribs;
  size_t i = 0;
  while (i < v.size() && v[i].sid != from->sid) i++;
  if (i == v.size() || c->parent != from) {           // not the situation described above: one update per birth, as the reference
    for (uint32_t q = 0; q < n1; q++) updateContributions(c, { { from->sid, 1u } });
    for (uint32_t q = 0; q < n2; q++) updateContributions(c, { { from->sid, 2u } });
    return;
  }
  v[i].count += n1 + 2 * n2;
  while (i > 0 && v[i - 1].count < v[i].count) { std::swap(v[i - 1], v[i]); i--; }
}

/// One birth the family pass (K3) decided with an effect on the tree: it entered a sub-species' enveloppe (append
/// :668-677 / replacement :694-718) or founded a species (:637-643, :398-426).  The host mirrors the decision --
/// the new species gets its id here, in birth order (:225-230) -- exactly as insertInto + updateContributions would;
/// counters shared by all the births the family decided are booked per family (settleFamilies).
void Engine::walkFamily (uint32_t i, const uint32_t *w, RepMeta &g, Species *s0, uint32_t fact, uint32_t &sid, int32_t &slot,
                         uint32_t &replaced) {
  const Family &F = fams_[size_t(grpFam_[s0->grp_index])];
  std::vector<Species*> &founded = famFounded_[size_t(grpFam_[s0->grp_index])];
  const uint32_t code = fact & FAM_CODE_MASK, child = fact >> FAM_CHILD_SHIFT, r = (fact >> FAM_SLOT_SHIFT) & 31u;
  const Contributions contrib { { g.msid, 1u + (g.msid == g.fsid) } };
  replaced = INV;
  Species *c;
  RepMeta m = g;
  m.timestamp = step_;
  if (code == FAM_FOUNDED) {
    const uint32_t f = uint32_t(grpFam_[s0->grp_index]), j = child - F.snapN;
    if (j != founded.size()) throw ApiError(APT_ERR_STATE, "family pass: species founded out of order");
    forcedSlot_ = hNewFounder_.p[size_t(fams_.size()) * famMaxNew_ + size_t(f) * famMaxNew_ + j];
    c = makeNode(contrib);                                    // simpleNewSpecies, :637-643
    if (c->parent != s0) throw ApiError(APT_ERR_STATE, "family pass: a founded species did not become a sub-species");
    founded.push_back(c);
    m.udata = g.gid;
    c->reps[0] = m;
    memcpy(&hPayload_[(size_t(c->slot) * K_) * wp_], w, W_ * sizeof(uint32_t));
    c->k = 1;
    emit(APT_EV_ENTERS_ENVELOPPE, c->sid, g.gid, 0);
    c->count++;  c->alive++;  c->last = step_;
    // (:639-643: the founder is inserted with an empty contribution)
  } else {
    c = child < F.snapN ? s0->children[F.snapN - 1 - child] : founded.at(child - F.snapN);
    const uint32_t k = c->k;
    const float *d = hRows_.p + size_t(hFamRow_.p[i]) * K_;
    if (fact & FAM_REPLACED) {
      const uint32_t epId = c->reps[r].gid;
      emit(APT_EV_LEAVES_ENVELOPPE, c->sid, epId, 0);
      emit(APT_EV_ENTERS_ENVELOPPE, c->sid, g.gid, 0);
      m.udata = epId;                                         // :710
      replaced = epId;
    } else {
      if (r != k) throw ApiError(APT_ERR_STATE, "device/host enveloppe fill mismatch (family pass)");
      emit(APT_EV_ENTERS_ENVELOPPE, c->sid, g.gid, 0);
      m.udata = g.gid;
      c->k = k + 1;
    }
    c->reps[r] = m;
    memcpy(&hPayload_[(size_t(c->slot) * K_ + r) * wp_], w, W_ * sizeof(uint32_t));
    for (uint32_t q = 0; q < k; q++)
      if (q != r) c->dist[q * K_ + r] = c->dist[r * K_ + q] = d[q];
    c->count++;  c->alive++;  c->last = step_;
    bulkContribution(c, s0, g.msid == g.fsid ? 0u : 1u, g.msid == g.fsid ? 1u : 0u);
  }
  // the kernel wrote the same change to the pool
  c->fam_epoch = chunkEpoch_;
  touchSnap(c);
  c->snap_mask |= (1u << r);
  c->dev_fill = c->k;
  setStreamable(c, true);
  sid = c->sid;
  slot = int32_t(r);
  counters[14]++;
}

/// Books what the family pass decided without the host looking at the births one by one: per family the births'
/// share of the statistics and of their species' pending counter (:293, :298, :444, :470, :526), per sub-species the
/// births that joined it leaving its enveloppe alone (:721-723 counters, :741-769 contribution of the parent species).
void Engine::settleFamilies (void) {
  for (size_t f = 0; f < fams_.size(); f++) {
    const Family &F = fams_[f];
    const uint4 info = hFamInfo_.p[f];
    if (!info.x) continue;
    Species *S = grpSpecies_[F.g];
    S->pending = S->pending > info.x ? S->pending - info.x : 0;
    insertions_ += info.x;
    branching_ += uint32_t(hFamStats_.p[2 * f]);
    comparisons_ += uint32_t(hFamStats_.p[2 * f + 1]);
    counters[0] += info.x;  counters[1] += info.x;  counters[15] += info.x;
    if (info.y != famFounded_[f].size()) throw ApiError(APT_ERR_STATE, "family pass: founded species were not all committed");
    const uint32_t C = F.snapN + info.y;
    for (uint32_t c = 0; c < C; c++) {
      const uint32_t n1 = hChildJoin_.p[F.joinOff + 2 * c], n2 = hChildJoin_.p[F.joinOff + 2 * c + 1];
      if (!(n1 + n2)) continue;
      Species *sp = c < F.snapN ? S->children[F.snapN - 1 - c] : famFounded_[f][c - F.snapN];
      sp->count += n1 + n2;  sp->alive += n1 + n2;  sp->last = step_;
      bulkContribution(sp, S, n1, n2);
    }
  }
}

/// Device-resident births: genome of a birth the stream kernel put into an enveloppe (it came back with the
/// birth's distance row), or null when the birth is not such an entry
const uint32_t* Engine::entryRowOf (uint32_t i) const {
  if (bg_[i] < 0) return nullptr;
  const uint32_t act = hAct_.p[i];
  if (famResults_ && act == ACT_FALLTHROUGH && (hFamAct_.p[i] & FAM_CODE_MASK) >= FAM_ENTERED)
    return hRowPayload_.p + size_t(hFamRow_.p[i]) * wp_;       // entered a sub-species / founded one in the family pass
  if ((act & (ACT_CODE_MASK | ACT_UNCERTAIN)) != ACT_ENTERED) return nullptr;
  return hRowPayload_.p + size_t(act >> ACT_ROW_SHIFT) * wp_;
}

/// Payload of birth i of the current chunk as the host sees it
const uint32_t* Engine::payloadOf (uint32_t i) {
  if (curPayload_) return curPayload_ + size_t(i) * W_;
  if (const uint32_t *w = entryRowOf(i)) return w;
  if (curFtRow_ >= 0 && i == curBirth_) return hFtRows_.p + size_t(curFtRow_) * W_;
  auto it = std::lower_bound(gatherIdx_.begin(), gatherIdx_.end(), i);
  if (it != gatherIdx_.end() && *it == i) return hSlowRows_.p + size_t(it - gatherIdx_.begin()) * W_;
  // a birth handed back by a taken-over stream: fetched on demand
  lateRows_.emplace_back(W_);
  cudaCheck(cudaMemcpy(lateRows_.back().data(), curDevPayload_ + size_t(i) * W_, W_ * 4, cudaMemcpyDeviceToHost), "late row");
  return lateRows_.back().data();
}

// ---- device plumbing ----------------------------------------------------------------------------

PoolView Engine::poolView (void) const {
  return PoolView { dPool_.p, dMeta_.p, dFill_.p, K_, wp_, nSlots_ };
}

ScoreCfg Engine::scoreCfg (void) const {
  return ScoreCfg { int(p_.kind), p_.full_continuous, lutMode_ ? 1 : 0, p_.compat_thr, p_.sim_thr, p_.avg_thr,
                    dLut_.p, lutStride_, lutUploaded_ };
}

void Engine::uploadLut (void) {
  if (!lutMode_ || lutUploaded_ == classes_.size()) return;
  dLut_.reserve(hLut_.size(), true, stream_);
  const size_t off = size_t(lutUploaded_) * lutStride_;
  cudaCheck(cudaMemcpyAsync(dLut_.p + off, hLut_.data() + off, (hLut_.size() - off) * sizeof(double),
                            cudaMemcpyHostToDevice, stream_), "LUT upload");
  cudaCheck(cudaStreamSynchronize(stream_), "LUT upload sync");   // hLut_ is pageable and may be resized
  lutUploaded_ = uint32_t(classes_.size());
}

/// Brings the device pool up to date with the host state: only what changed is staged (pinned)
/// and scattered by three small kernels (rows+metadata, distance matrices, fills).
void Engine::refreshPool (void) {
  const double t0 = now();
  dPool_.reserve(size_t(nSlots_) * K_ * wp_, true, stream_);
  dMeta_.reserve(size_t(nSlots_) * K_, true, stream_);
  dPoolDist_.reserve(size_t(nSlots_) * K_ * K_, true, stream_);
  const size_t oldFillCap = dFill_.cap;
  dFill_.reserve(nSlots_, true, stream_);
  if (dFill_.cap != oldFillCap)
    cudaCheck(cudaMemsetAsync(dFill_.p + oldFillCap, 0, (dFill_.cap - oldFillCap) * sizeof(uint32_t), stream_),
              "fill clear");

  size_t rows = 0, mats = 0;
  for (Species *s: refresh_) {
    if (s->trimmed) continue;
    if (s->full_upload) s->stale_mask = s->k >= 32 ? 0xFFFFFFFFu : ((1u << s->k) - 1u);
    rows += __builtin_popcount(s->stale_mask);
    mats++;
  }
  const size_t KK = size_t(K_) * K_;
  hStageRows_.reserve(rows * wp_ + 4);  hStageMeta_.reserve(rows + 1);  hStageDist_.reserve(mats * KK + 1);
  hStageDst_.reserve(rows + mats + fillUpdates_.size() + 2 * refresh_.size() + 4);
  size_t i = 0, m = 0;
  std::vector<uint32_t> matSlots;
  for (Species *s: refresh_) {
    s->queued = false;
    if (s->trimmed) continue;
    for (uint32_t r = 0; r < s->k; r++) {
      if (!((s->stale_mask >> r) & 1u)) continue;
      memcpy(hStageRows_.p + i * wp_, repPayload(s, r), wp_ * sizeof(uint32_t));
      const RepMeta &mm = s->reps[r];
      float cls;  memcpy(&cls, &mm.cls, 4);
      hStageMeta_.p[i] = make_float4(mm.mu, mm.sin, mm.sout, cls);
      hStageDst_.p[i] = s->slot * K_ + r;
      i++;
    }
    memcpy(hStageDist_.p + m * KK, s->dist.data(), KK * sizeof(float));
    matSlots.push_back(s->slot);
    m++;
    s->stale_mask = 0;
    s->full_upload = false;
    s->dev_fill = s->k;
    setStreamable(s, s->k > 0);
    fillUpdates_.push_back(s->slot);  fillUpdates_.push_back(s->k);
  }
  refresh_.clear();
  rows = i;  mats = m;
  counters[10] += rows;
  const size_t nFill = fillUpdates_.size() / 2;
  // destinations share one pinned buffer: [row dst][matrix slots][fill slots][fill values]
  for (size_t j = 0; j < mats; j++) hStageDst_.p[rows + j] = matSlots[j];
  for (size_t j = 0; j < nFill; j++) {
    hStageDst_.p[rows + mats + j] = fillUpdates_[2 * j];
    hStageDst_.p[rows + mats + nFill + j] = fillUpdates_[2 * j + 1];
  }
  fillUpdates_.clear();
  if (rows || mats || nFill) {
    dStageRows_.reserve(rows * wp_ + 4);  dStageMeta_.reserve(rows + 1);  dStageDist_.reserve(mats * KK + 1);
    dStageDst_.reserve(rows + mats + 2 * nFill + 4);
    if (rows) {
      cudaCheck(cudaMemcpyAsync(dStageRows_.p, hStageRows_.p, rows * wp_ * sizeof(uint32_t), cudaMemcpyHostToDevice, stream_), "stage rows");
      cudaCheck(cudaMemcpyAsync(dStageMeta_.p, hStageMeta_.p, rows * sizeof(float4), cudaMemcpyHostToDevice, stream_), "stage meta");
    }
    if (mats)
      cudaCheck(cudaMemcpyAsync(dStageDist_.p, hStageDist_.p, mats * KK * sizeof(float), cudaMemcpyHostToDevice, stream_), "stage dist");
    cudaCheck(cudaMemcpyAsync(dStageDst_.p, hStageDst_.p, (rows + mats + 2 * nFill) * sizeof(uint32_t), cudaMemcpyHostToDevice, stream_), "stage dst");
    launch_scatter_rows(dStageRows_.p, dStageMeta_.p, dStageDst_.p, uint32_t(rows), wp_, dPool_.p, dMeta_.p, stream_);
    launch_scatter_blocks_f32(dStageDist_.p, dStageDst_.p + rows, uint32_t(mats), uint32_t(KK), dPoolDist_.p, stream_);
    launch_scatter_u32(dStageDst_.p + rows + mats, dStageDst_.p + rows + mats + nFill, uint32_t(nFill), dFill_.p, stream_);
    counters[4] += (rows ? 1 : 0) + (mats ? 1 : 0) + (nFill ? 1 : 0);
    // the pinned staging buffers are rewritten by the next refresh: wait for the copies
    cudaCheck(cudaStreamSynchronize(stream_), "pool refresh");
  }
  timers[3] += now() - t0;
}

/// Groups the births of a chunk into per-species streams (single parent species with a current
/// device copy) and speculative pairs (everything else).  Serial version.
void Engine::groupSerial (uint32_t n, const uint32_t *msid, const uint32_t *fsid, size_t &P, size_t &nStream,
                          uint32_t &maxStream) {
  auto touch = [&] (Species *s) {
    if (s->touch_epoch != touchEpoch_) { s->touch_epoch = touchEpoch_; counters[8] += s->dev_fill; }
  };
  auto lookup = [&] (uint32_t sid) -> Species* {
    return (sid != INV && sid < nodes_.size()) ? nodes_[sid] : nullptr;
  };
  // birthGroup_[i]: stream index | contribution class << 30 (what the birth adds to its own species,
  // :576-577: 0 primordial, 1 one parent, 2 both parents), or -1 for the host / speculative path
  for (uint32_t i = 0; i < n; i++) {
    Species *s0 = nullptr, *s1 = nullptr;
    const uint32_t m = msid[i], f = fsid[i];
    if (streamOk_ && (f == INV || m == f) && m < sidGroup_.size()) {
      // single-parent birth: the per-chunk table answers without touching the species record
      int32_t gidx = sidGroup_[m];
      if (gidx == -2) {
        Species *sp = nodes_[m];
        if (sp && sp->dev_fill) {
          sp->grp_epoch = chunkEpoch_;
          sp->grp_index = uint32_t(grpSpecies_.size());
          gidx = int32_t(sp->grp_index);
          grpSpecies_.push_back(sp);  grpCount_.push_back(0);
          touch(sp);
        } else gidx = -1;
        sidGroup_[m] = gidx;
        sidTouched_.push_back(m);
      }
      if (gidx >= 0) {
        grpCount_[size_t(gidx)]++;
        birthGroup_[i] = gidx | int32_t((m == f ? 2u : 1u) << ORDER_CLASS_SHIFT);
        nStream++;
        continue;
      }
    }
    if (m == INV && f == INV) s0 = root_;
    else if (f == INV || m == f) s0 = lookup(m);
    else { s0 = lookup(m); s1 = lookup(f); }
    if (streamOk_ && s0 && !s1 && s0->dev_fill) {
      if (s0->grp_epoch != chunkEpoch_) {
        s0->grp_epoch = chunkEpoch_;
        s0->grp_index = uint32_t(grpSpecies_.size());
        grpSpecies_.push_back(s0);  grpCount_.push_back(0);
        if (s0->sid < sidGroup_.size() && sidGroup_[s0->sid] == -2) { sidGroup_[s0->sid] = int32_t(s0->grp_index); sidTouched_.push_back(s0->sid); }
      }
      grpCount_[s0->grp_index]++;
      const uint32_t ccls = m == INV ? 0u : (m == f ? 2u : 1u);
      birthGroup_[i] = int32_t(s0->grp_index) | int32_t(ccls << ORDER_CLASS_SHIFT);
      nStream++;
      touch(s0);
      continue;
    }
    birthGroup_[i] = -1;
    if (s0) s0->pair_epoch = chunkEpoch_;                 // walked by the host: nobody else may own these species' families
    if (s1) s1->pair_epoch = chunkEpoch_;
    PairRef pr { i, -1, -1 };
    if (s0 && s0->dev_fill) { pr.p0 = int64_t(P); hPairBirth_.p[P] = i; hPairSlot_.p[P] = s0->slot; P++; touch(s0); }
    if (s1 && s1->dev_fill) { pr.p1 = int64_t(P); hPairBirth_.p[P] = i; hPairSlot_.p[P] = s1->slot; P++; touch(s1); }
    if (pr.p0 >= 0 || pr.p1 >= 0) pairOf_.push_back(pr);
  }
  const uint32_t G = uint32_t(grpSpecies_.size());
  if (G) {
    hGrpSlot_.reserve(G);  hGrpOff_.reserve(G + 1);  hOrder_.reserve(nStream);
    uint32_t off = 0;
    hGrpSid_.reserve(G);
    for (uint32_t g = 0; g < G; g++) {
      hGrpSlot_.p[g] = grpSpecies_[g]->slot;  hGrpSid_.p[g] = grpSpecies_[g]->sid;
      hGrpOff_.p[g] = off;  off += grpCount_[g];  maxStream = std::max(maxStream, grpCount_[g]);  grpCount_[g] = off - grpCount_[g];
    }
    hGrpOff_.p[G] = off;
    for (uint32_t i = 0; i < n; i++) {
      const uint32_t bg = uint32_t(birthGroup_[i]);       // class 2 sets the sign bit: compare with "none" explicitly
      if (bg != 0xFFFFFFFFu) {
        const uint32_t g = bg & ORDER_INDEX_MASK;
        hOrder_.p[grpCount_[g]++] = i | (bg & ~ORDER_INDEX_MASK);
        birthGroup_[i] = int32_t(g);
      }
    }
  }
}

/// Same result as groupSerial for large chunks: a stable parallel counting sort by species id.
/// Thread t owns a contiguous range of births; per-thread histograms give every thread its own
/// write cursor inside each stream, so arrival order inside a stream is preserved.  Every thread
/// scatters into a PRIVATE buffer first (grouped by species) and then copies its run of each stream
/// into place: writing the shared order list element by element made the threads fight over its cache
/// lines (a stream of 100 births is 7 lines shared by 12 writers; 0.10 ms of 0.19 ms, measured).
void Engine::groupParallel (uint32_t n, const uint32_t *msid, const uint32_t *fsid, size_t &P, size_t &nStream,
                            uint32_t &maxStream) {
  const unsigned T = pool_->size();
  const size_t NS = nodes_.size();
  if (sidSlot_.size() < NS) sidSlot_.resize(NS, INV);
  parCnt_.resize(T);  parRest_.resize(T);  parOrd_.resize(T);  parLoc_.resize(T);
  birthKey_.resize(n);
  sidTotal_.resize(NS);
  const uint32_t rootSid = (root_ && root_->sid < NS && nodes_[root_->sid] == root_) ? root_->sid : INV;
  const bool rootOk = rootSid != INV && sidSlot_[rootSid] != INV;
  const uint32_t per = (n + T - 1) / T;
  const uint32_t sidPer = uint32_t((NS + T - 1) / T);
  const double q0 = profile_ ? now() : 0.;
  // 1. classify every birth, count per (thread, species)
  pool_->run([&] (unsigned t) {
    auto &cnt = parCnt_[t];
    cnt.assign(NS, 0u);
    const uint32_t lo = std::min(n, t * per), hi = std::min(n, lo + per);
    for (uint32_t i = lo; i < hi; i++) {
      const uint32_t m = msid[i], f = fsid[i];
      uint32_t key = INV;
      if (f == INV || m == f) {
        if (m == INV) { if (rootOk) key = rootSid; }                                  // primordial: root, contributes nothing (:281-282)
        else if (m < NS && sidSlot_[m] != INV) key = m | ((m == f ? 2u : 1u) << ORDER_CLASS_SHIFT);
      }
      birthKey_[i] = key;
      if (key != INV) cnt[key & ORDER_INDEX_MASK]++;
    }
  });
  // 1b. per species: total over the threads, and every thread's offset inside the stream (thread t owns a range of ids)
  pool_->run([&] (unsigned t) {
    const size_t lo = std::min(NS, size_t(t) * sidPer), hi = std::min(NS, lo + sidPer);
    for (size_t sid = lo; sid < hi; sid++) {
      uint32_t run = 0;
      for (unsigned u = 0; u < T; u++) { const uint32_t c = parCnt_[u][sid]; parCnt_[u][sid] = run; run += c; }
      sidTotal_[sid] = run;
    }
  });
  const double q1 = profile_ ? now() : 0.;
  // 2. streams in ascending species id
  uint32_t off = 0;
  for (size_t sid = 0; sid < NS; sid++) {
    const uint32_t total = sidTotal_[sid];
    if (sid + 8 < NS && sidTotal_[sid + 8] && nodes_[sid + 8]) {
      __builtin_prefetch(nodes_[sid + 8], 1);  __builtin_prefetch(reinterpret_cast<const char*>(nodes_[sid + 8]) + 128, 1);
    }
    if (!total) continue;
    Species *sp = nodes_[sid];
    sp->grp_epoch = chunkEpoch_;
    sp->grp_index = uint32_t(grpSpecies_.size());
    grpSpecies_.push_back(sp);  grpCount_.push_back(total);
    if (sp->touch_epoch != touchEpoch_) { sp->touch_epoch = touchEpoch_; counters[8] += sp->dev_fill; }
    sidGroup_[sid] = int32_t(sp->grp_index);
    sidTouched_.push_back(uint32_t(sid));
    sidTotal_[sid] = off;                       // from here on: where the stream starts in the order list
    off += total;
    maxStream = std::max(maxStream, total);
  }
  nStream = off;
  const uint32_t G = uint32_t(grpSpecies_.size());
  hGrpSlot_.reserve(G);  hGrpOff_.reserve(G + 1);  hGrpSid_.reserve(G);  hOrder_.reserve(nStream);
  uint32_t o2 = 0;
  for (uint32_t g = 0; g < G; g++) {
    hGrpSlot_.p[g] = grpSpecies_[g]->slot;  hGrpSid_.p[g] = grpSpecies_[g]->sid;
    hGrpOff_.p[g] = o2;  o2 += grpCount_[g];
  }
  hGrpOff_.p[G] = o2;
  const double q2 = profile_ ? now() : 0.;
  // 3. scatter: private buffer grouped by species, then one copy per (thread, stream)
  pool_->run([&] (unsigned t) {
    auto &cur = parCnt_[t];                     // offset of this thread's run inside each stream
    auto &rest = parRest_[t];
    auto &ord = parOrd_[t];
    auto &loc = parLoc_[t];                     // cursor of each species inside the private buffer
    rest.clear();
    const uint32_t lo = std::min(n, t * per), hi = std::min(n, lo + per);
    // this thread's count per species = next thread's offset - its own (the last thread: total - its own);
    // counted again here, cheaper than keeping a second table per thread
    loc.assign(NS + 1, 0u);
    for (uint32_t i = lo; i < hi; i++) { const uint32_t key = birthKey_[i]; if (key != INV) loc[(key & ORDER_INDEX_MASK) + 1]++; }
    for (size_t sid = 0; sid < NS; sid++) loc[sid + 1] += loc[sid];
    ord.resize(loc[NS]);
    for (uint32_t i = lo; i < hi; i++) {
      const uint32_t key = birthKey_[i];
      if (key != INV) {
        const uint32_t sid = key & ORDER_INDEX_MASK;
        ord[loc[sid]++] = i | (key & ~ORDER_INDEX_MASK);
        birthGroup_[i] = sidGroup_[sid];
      } else { birthGroup_[i] = -1; rest.push_back(i); }
    }
    // loc[sid] now points behind the thread's run of species sid
    uint32_t begin = 0;
    for (size_t sid = 0; sid < NS; sid++) {
      const uint32_t end = loc[sid];
      if (end != begin) {                       // runs are a handful of entries: copied inline, no call
        uint32_t *dst = hOrder_.p + sidTotal_[sid] + cur[sid];
        const uint32_t *src = ord.data() + begin;
        for (uint32_t j = 0; j < end - begin; j++) dst[j] = src[j];
      }
      begin = end;
    }
  });
  const double q3 = profile_ ? now() : 0.;
  if (profile_) fprintf(stderr, "group: count %.3f assign %.3f scatter %.3f ms\n", 1e3*(q1-q0), 1e3*(q2-q1), 1e3*(q3-q2));
  // 4. the rest (two parent species, unknown ids, species without a device copy): speculative pairs
  for (unsigned t = 0; t < T; t++)
    for (uint32_t i: parRest_[t]) {
      const uint32_t m = msid[i], f = fsid[i];
      Species *s0 = nullptr, *s1 = nullptr;
      auto lookup = [&] (uint32_t sid) -> Species* { return (sid != INV && sid < NS) ? nodes_[sid] : nullptr; };
      if (m == INV && f == INV) s0 = root_;
      else if (f == INV || m == f) s0 = lookup(m);
      else { s0 = lookup(m); s1 = lookup(f); }
      if (s0) s0->pair_epoch = chunkEpoch_;
      if (s1) s1->pair_epoch = chunkEpoch_;
      PairRef pr { i, -1, -1 };
      auto touch = [&] (Species *s) {
        if (s->touch_epoch != touchEpoch_) { s->touch_epoch = touchEpoch_; counters[8] += s->dev_fill; }
      };
      if (s0 && s0->dev_fill) { pr.p0 = int64_t(P); hPairBirth_.p[P] = i; hPairSlot_.p[P] = s0->slot; P++; touch(s0); }
      if (s1 && s1->dev_fill) { pr.p1 = int64_t(P); hPairBirth_.p[P] = i; hPairSlot_.p[P] = s1->slot; P++; touch(s1); }
      if (pr.p0 >= 0 || pr.p1 >= 0) pairOf_.push_back(pr);
    }
}


/// The same grouping done by the device (launch_group_births): the id arrays go up, the per-species totals, the births
/// outside the streams and (later, with the results) the birth -> stream table and the order list come back; the host
/// only walks the species that have births.  Returns false when the pass has to be grouped on the host after all (a
/// stream too long for the device-side sort).
bool Engine::groupDeviceLaunch (uint32_t n, const uint32_t *msid, const uint32_t *fsid) {
  const size_t NS = nodes_.size();
  if (sidSlot_.size() < NS) { sidSlot_.resize(NS, INV);  sidSlotDirty_ = true; }
  dSidSlot_.reserve(NS);
  if (sidSlotDirty_ || dSidSlotN_ != NS) {
    cudaCheck(cudaMemcpyAsync(dSidSlot_.p, sidSlot_.data(), NS * 4, cudaMemcpyHostToDevice, streamG_), "species table");
    sidSlotDirty_ = false;  dSidSlotN_ = NS;
  }
  const double pg0 = profile_ ? now() : 0.;
  if (profile_) cudaEventRecord(evK0_, streamG_);
  dGrpKey_.reserve(n);  dOrder_.reserve(n);  dRestList_.reserve(n);
  dSidStart_.reserve(NS);  dSidCursor_.reserve(NS);  dSidGidx_.reserve(NS);  dGrpInfo_.reserve(16);
  hSidTotal_.reserve(NS);  hGrpInfo_.reserve(8);  hBirthGroup_.reserve(n);  hOrder_.reserve(n);  hRestList_.reserve(n);
  // two sets of per-species counters (and of summary words) used in turn: the scan of one pass clears the set of the next,
  // so that a steady pass needs no clearing launch; a set is cleared here when it is new or the species table grew
  const uint32_t cur = grpFlip_, nxt = cur ^ 1u;
  for (uint32_t q = 0; q < 2; q++) {
    dSidTotal_[q].reserve(NS);
    if (grpZeroedPtr_[q] != dSidTotal_[q].p) { grpZeroedPtr_[q] = dSidTotal_[q].p;  grpZeroed_[q] = 0; }
  }
  if (grpZeroed_[cur] < NS) {
    cudaCheck(cudaMemsetAsync(dSidTotal_[cur].p, 0, NS * 4, streamG_), "stream counters clear");
    grpZeroed_[cur] = NS;
  }
  if (!grpInfoReady_) { cudaCheck(cudaMemsetAsync(dGrpInfo_.p, 0, 16 * 4, streamG_), "grouping summary clear");  grpInfoReady_ = true; }
  // page-locked id arrays are read by the kernel where they are; others (tests) go through a device copy
  const uint32_t *kMsid = msid, *kFsid = fsid;
  if (!idsPinned_) {
    dMsid_.reserve(n);  dFsid_.reserve(n);
    cudaCheck(cudaMemcpyAsync(dMsid_.p, msid, size_t(n) * 4, cudaMemcpyHostToDevice, streamG_), "mother species ids");
    cudaCheck(cudaMemcpyAsync(dFsid_.p, fsid, size_t(n) * 4, cudaMemcpyHostToDevice, streamG_), "father species ids");
    kMsid = dMsid_.p;  kFsid = dFsid_.p;
  }
  const uint32_t rootSid = (root_ && root_->sid < NS && nodes_[root_->sid] == root_ && sidSlot_[root_->sid] != INV) ? root_->sid : INV;
  uint32_t *dInfo = dGrpInfo_.p + 8 * cur, *dInfoNext = dGrpInfo_.p + 8 * nxt;
  hGrpInfo_.p[3] = 1u;  hGrpInfo_.p[0] = hGrpInfo_.p[2] = hGrpInfo_.p[4] = 0u;        // (a launch that fails leaves "group on the host")
  launch_group_births(kMsid, kFsid, n, dSidSlot_.p, uint32_t(NS), rootSid, dGrpKey_.p, dSidTotal_[cur].p, dSidTotal_[nxt].p,
                      dSidStart_.p, dSidCursor_.p, dSidGidx_.p, dInfo, dInfoNext, dOrder_.p, dRestList_.p, hSidTotal_.p, hGrpInfo_.p,
                      hRestList_.p, hOrder_.p, hBirthGroup_.p, streamG_);
  if (cudaGetLastError() != cudaSuccess) {      // (a cooperative launch the device cannot place: the host groups, now and later)
    grpZeroed_[0] = grpZeroed_[1] = 0;  grpInfoReady_ = false;
    devGroupMin_ = 0xFFFFFFFFu;
    return false;
  }
  grpZeroed_[cur] = 0;  grpZeroed_[nxt] = NS;  grpFlip_ = nxt;
  counters[4] += 1;
  if (profile_) cudaEventRecord(evK1_, streamG_);
  cudaCheck(cudaEventRecord(evG_, streamG_), "event record");
  if (profile_) grpProf_[0] = now() - pg0;
  return true;
}

/// ... and its second half, once the host has done what does not depend on the grouping: waits for the kernel and books
/// the streams.  Returns false when the pass has to be grouped on the host after all.
bool Engine::groupDeviceFinish (uint32_t n, const uint32_t *msid, const uint32_t *fsid, size_t &P, size_t &nStream, uint32_t &maxStream) {
  const size_t NS = nodes_.size();
  const double pg1 = profile_ ? now() : 0.;
  cudaCheck(cudaEventSynchronize(evG_), "grouping");
  if (profile_ && n >= parMin_) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, evK0_, evK1_);
    fprintf(stderr, "group (device): enqueue %.3f wait %.3f ms; on the device %.3f ms\n", 1e3 * grpProf_[0], 1e3 * (now() - pg1), double(ms));
  }
  if (hGrpInfo_.p[3]) return false;
  // streams in ascending species id (as groupParallel)
  uint32_t off = 0;
  for (size_t sid = 0; sid < NS; sid++) {
    const uint32_t total = hSidTotal_.p[sid];
    if (sid + 8 < NS && hSidTotal_.p[sid + 8] && nodes_[sid + 8]) {
      __builtin_prefetch(nodes_[sid + 8], 1);  __builtin_prefetch(reinterpret_cast<const char*>(nodes_[sid + 8]) + 128, 1);
    }
    if (!total) continue;
    Species *sp = nodes_[sid];
    if (!sp) throw ApiError(APT_ERR_STATE, "device grouping: a stream of a species that does not exist");
    sp->grp_epoch = chunkEpoch_;
    sp->grp_index = uint32_t(grpSpecies_.size());
    grpSpecies_.push_back(sp);  grpCount_.push_back(total);
    if (sp->touch_epoch != touchEpoch_) { sp->touch_epoch = touchEpoch_; counters[8] += sp->dev_fill; }
    sidGroup_[sid] = int32_t(sp->grp_index);
    sidTouched_.push_back(uint32_t(sid));
    off += total;
    maxStream = std::max(maxStream, total);
  }
  nStream = off;
  const uint32_t G = uint32_t(grpSpecies_.size());
  if (G != hGrpInfo_.p[0] || off != hGrpInfo_.p[2]) throw ApiError(APT_ERR_STATE, "device grouping: inconsistent summary");
  hGrpSlot_.reserve(G);  hGrpOff_.reserve(G + 1);  hGrpSid_.reserve(G);
  uint32_t o2 = 0;
  for (uint32_t g = 0; g < G; g++) {
    hGrpSlot_.p[g] = grpSpecies_[g]->slot;  hGrpSid_.p[g] = grpSpecies_[g]->sid;
    hGrpOff_.p[g] = o2;  o2 += grpCount_[g];
  }
  hGrpOff_.p[G] = o2;
  // the rest (two parent species, unknown ids, species without a device copy): speculative pairs, in arrival order
  const uint32_t nRest = std::min(hGrpInfo_.p[4], n);
  std::sort(hRestList_.p, hRestList_.p + nRest);              // (appended with atomics on the device: any order)
  for (uint32_t e = 0; e < nRest; e++) {
    const uint32_t i = hRestList_.p[e];
    if (i >= n) throw ApiError(APT_ERR_STATE, "device grouping: corrupt list");
    const uint32_t m = msid[i], f = fsid[i];
    Species *s0 = nullptr, *s1 = nullptr;
    auto lookup = [&] (uint32_t sid) -> Species* { return (sid != INV && sid < NS) ? nodes_[sid] : nullptr; };
    if (m == INV && f == INV) s0 = root_;
    else if (f == INV || m == f) s0 = lookup(m);
    else { s0 = lookup(m); s1 = lookup(f); }
    if (s0) s0->pair_epoch = chunkEpoch_;
    if (s1) s1->pair_epoch = chunkEpoch_;
    PairRef pr { i, -1, -1 };
    auto touch = [&] (Species *s) {
      if (s->touch_epoch != touchEpoch_) { s->touch_epoch = touchEpoch_; counters[8] += s->dev_fill; }
    };
    if (s0 && s0->dev_fill) { pr.p0 = int64_t(P); hPairBirth_.p[P] = i; hPairSlot_.p[P] = s0->slot; P++; touch(s0); }
    if (s1 && s1->dev_fill) { pr.p1 = int64_t(P); hPairBirth_.p[P] = i; hPairSlot_.p[P] = s1->slot; P++; touch(s1); }
    if (pr.p0 >= 0 || pr.p1 >= 0) pairOf_.push_back(pr);
  }
  return true;
}

void Engine::runChunk (uint32_t n, const uint32_t *payload, const float *compat, const uint32_t *dPayload,
                       const uint32_t *gid, const uint32_t *mgid, const uint32_t *msid,
                       const uint32_t *fgid, const uint32_t *fsid, uint32_t *outSid, int32_t *outSlot,
                       uint32_t *outReplaced, uint32_t &done) {
  double t0 = now();
  chunkEpoch_++;
  curBirth_ = 0;              // (the error path settles the streams at the birth being committed)
  famResults_ = false;  famSlotsUsed_ = 0;
  // APT_PROFILE: named time marks of this chunk, printed as successive differences at its end
  std::vector<std::pair<const char*, double>> marks;
  auto mark = [&] (const char *name) { if (profile_) marks.emplace_back(name, now()); };
  mark("start");
  // the births' genomes start their way to the device before the host groups them: the copy
  // overlaps the grouping below (host buffers only; a chunk that ends up with no device work at
  // all -- an empty pool -- has paid for a copy nobody reads)
  bool birthsUploaded = false;
  auto uploadBirths = [&] (void) {
    dBirths_.reserve(size_t(n) * wp_);
    const size_t bytes = size_t(n) * wp_ * 4;
    if (wp_ == W_ && (payloadPinned_ || bytes < stageMin_ || !pool_)) {
      // pinned rows: one asynchronous DMA (small pageable batches: the driver's own staging)
      cudaCheck(cudaMemcpyAsync(dBirths_.p, payload, bytes, cudaMemcpyHostToDevice, stream_), "births");
    } else {
      // pageable rows, or rows that need padding: the workers copy slices into pinned memory and each slice
      // leaves for the device as soon as it is complete (a pageable cudaMemcpy is one thread at ~10 GB/s:
      // 4.8 ms for the 51 MB of 100 000 4096-bit genomes, measured)
      hBirths_.reserve(size_t(n) * wp_);
      const unsigned T = (pool_ && bytes >= stageMin_) ? pool_->size() : 1u;
      const uint32_t rowsPerSlice = uint32_t(std::max<size_t>(1, std::max<size_t>(STAGE_SLICE, bytes / (4 * size_t(T))) / (size_t(wp_) * 4)));
      const uint32_t nSlices = (n + rowsPerSlice - 1) / rowsPerSlice;
      std::atomic<int> err { int(cudaSuccess) };
      auto stage = [&] (unsigned t) {
        if (T > 1 && t) cudaSetDevice(p_.device);            // worker threads start on device 0
        for (uint32_t sl = t; sl < nSlices; sl += T) {
          const uint32_t r0 = sl * rowsPerSlice, r1 = std::min(n, r0 + rowsPerSlice);
          uint32_t *dst = hBirths_.p + size_t(r0) * wp_;
          if (wp_ == W_) memcpy(dst, payload + size_t(r0) * W_, size_t(r1 - r0) * W_ * 4);
          else
            for (uint32_t i = r0; i < r1; i++) {
              memcpy(hBirths_.p + size_t(i) * wp_, payload + size_t(i) * W_, W_ * 4);
              memset(hBirths_.p + size_t(i) * wp_ + W_, 0, (wp_ - W_) * 4);
            }
          const cudaError_t e = cudaMemcpyAsync(dBirths_.p + size_t(r0) * wp_, dst, size_t(r1 - r0) * wp_ * 4,
                                                cudaMemcpyHostToDevice, stream_);
          if (e != cudaSuccess) err.store(int(e));
        }
      };
      if (T > 1) pool_->run(stage); else stage(0);
      cudaCheck(cudaError_t(err.load()), "births");
    }
    birthsUploaded = true;
  };
  const bool tryDevGroup = streamOk_ && n >= devGroupMin_ && (idsPinned_ || devGroupPageable_) && nSlots_ > freeSlots_.size()
                           && nodes_.size() <= std::min<size_t>(size_t(2) * n, 32768);
  const bool earlyUpload = !dPayload && n >= 4096 && nSlots_ > freeSlots_.size();
  if (earlyUpload) uploadBirths();
  // ... and so do the clears of the result arrays (they depend on n only): the device runs them under the grouping
  bool earlyClears = false;
  auto clearResults = [&] (void) {
    dOut2_.reserve(2 * size_t(n));  dAct_.reserve(n);  dCtr_.reserve(8);
    cudaCheck(cudaMemsetAsync(dOut2_.p, 0xFF, 2 * size_t(n) * 4, stream_), "sid / slot clear");
    cudaCheck(cudaMemsetAsync(dAct_.p, 0, size_t(n) * 4, stream_), "act clear");
    earlyClears = true;
  };
  const bool earlyClear = streamOk_ && n >= 4096 && nSlots_ > freeSlots_.size();
  if (earlyClear) clearResults();
  mark("upload");
  // ---- prepare: classes; births with a single parent species join that species' stream, the
  //      others get (birth, species) pairs for the speculative scoring kernel ----
  hCls_.reserve(n);
  hBMeta_.reserve(n);
  birthGroup_.resize(n);
  pairOf_.clear();
  grpSpecies_.clear();  grpCount_.clear();
  hPairBirth_.reserve(2 * size_t(n));  hPairSlot_.reserve(2 * size_t(n));
  size_t P = 0, nStream = 0;
  touchEpoch_++;
  for (uint32_t sid: sidTouched_) if (sid < sidGroup_.size()) sidGroup_[sid] = -2;
  sidTouched_.clear();
  if (sidGroup_.size() < nodes_.size()) sidGroup_.resize(nodes_.size(), -2);
  // (the device starts grouping the births while the host looks at their compatibility classes)
  const bool devLaunched = tryDevGroup && groupDeviceLaunch(n, msid, fsid);
  // compatibility classes: one triple for the whole chunk is the common case (array == itself shifted by one triple)
  if (pool_ && n >= parMin_) {
    const unsigned T = pool_->size();
    const uint32_t per = (n - 1 + T - 1) / T;
    std::atomic<bool> same { true };
    pool_->run([&] (unsigned t) {
      const uint32_t lo = std::min(n - 1, t * per), hi = std::min(n - 1, lo + per);     // triple i against triple i + 1, i in [lo, hi)
      if (hi > lo && memcmp(compat + 3 * size_t(lo), compat + 3 * size_t(lo) + 3, size_t(hi - lo) * 3 * sizeof(float)) != 0)
        same.store(false, std::memory_order_relaxed);
    });
    chunkUniform_ = same.load();
  } else
    chunkUniform_ = n < 2 || memcmp(compat, compat + 3, size_t(n - 1) * 3 * sizeof(float)) == 0;
  if (chunkUniform_) {
    chunkCls_ = classOf(compat[0], compat[1], compat[2]);
  } else {
    for (uint32_t i = 0; i < n; i++) {
      const uint32_t cls = classOf(compat[3 * i], compat[3 * i + 1], compat[3 * i + 2]);
      hCls_.p[i] = cls;
      float fc;  memcpy(&fc, &cls, 4);
      hBMeta_.p[i] = make_float4(compat[3 * i], compat[3 * i + 1], compat[3 * i + 2], fc);
    }
  }
  mark("classes");
  uint32_t maxStream = 0;
  devGrouped_ = false;
  bg_ = nullptr;
  if (devLaunched && !streamOk_) cudaCheck(cudaEventSynchronize(evG_), "grouping");      // (the classes just switched the streams off)
  else if (devLaunched) {
    devGrouped_ = groupDeviceFinish(n, msid, fsid, P, nStream, maxStream);
    if (devGrouped_) bg_ = hBirthGroup_.p;
    else {                        // (handed back: whatever it booked is dropped)
      P = 0;  nStream = 0;  maxStream = 0;  pairOf_.clear();  grpSpecies_.clear();  grpCount_.clear();
      for (uint32_t sid: sidTouched_) if (sid < sidGroup_.size()) sidGroup_[sid] = -2;
      sidTouched_.clear();
    }
  }
  if (devGrouped_) {}
  else if (pool_ && n >= parMin_ && streamOk_ && nodes_.size() <= size_t(2) * n)
    groupParallel(n, msid, fsid, P, nStream, maxStream);
  else
    groupSerial(n, msid, fsid, P, nStream, maxStream);
  if (!bg_) bg_ = birthGroup_.data();
  mark("group");
  const uint32_t G = uint32_t(grpSpecies_.size());
  // sub-species of every stream's species, in the reference's visiting order (newest first, :506)
  size_t nChild = 0;
  const bool useChildKernel = G && childKernel_;
  if (useChildKernel) {
    size_t total = 0;
    for (uint32_t g = 0; g < G; g++) total += grpSpecies_[g]->children.size();
    hChildOff_.reserve(G + 1);  hChildSlot_.reserve(total + 1);
    if (famOk_) {
      hChildSid_.reserve(total + 1);
      uploadLut();                                         // (the table's size decides how many sub-species fit)
      famCap_ = family_capacity(wp_, K_, scoreCfg());
    }
    fams_.clear();
    grpFam_.assign(G, -1);
    // A family -- the births a stream's species rejects at top level and that species' sub-species -- is resolved on the
    // device (K3) when nothing else in this pass can touch it: the species and its parent are named by no birth outside
    // the streams, the parent has no stream (its rejected births would walk into this species), no sub-species has a
    // stream or a two-parent birth of its own, all of them are current on the device, and a species founded now becomes
    // a sub-species (strictly older parent, speciescontributors.h:177-197).
    uint32_t joinOff = 0, maxSnap = 0;
    for (uint32_t g = 0; g < G; g++) {
      hChildOff_.p[g] = uint32_t(nChild);
      Species *S = grpSpecies_[g];
      const auto &ch = S->children;
      bool ok = famOk_ && S->pair_epoch != chunkEpoch_ && S->first < step_ && ch.size() + 4 <= famCap_
                && (!S->parent || (S->parent->grp_epoch != chunkEpoch_ && S->parent->pair_epoch != chunkEpoch_));
      for (size_t idx = ch.size(); idx-- > 0; ) {
        const Species *c = ch[idx];
        const bool current = c->dev_fill && c->dev_fill == c->k && !c->stale_mask && !c->full_upload;
        if (famOk_) hChildSid_.p[nChild] = c->sid;
        hChildSlot_.p[nChild++] = current ? c->slot : INV;
        ok = ok && current && c->grp_epoch != chunkEpoch_ && c->pair_epoch != chunkEpoch_;
      }
      if (ok) {
        grpFam_[g] = int32_t(fams_.size());
        fams_.push_back(Family{ g, uint32_t(ch.size()), joinOff });
        maxSnap = std::max<uint32_t>(maxSnap, uint32_t(ch.size()));
      }
    }
    hChildOff_.p[G] = uint32_t(nChild);
    if (!fams_.empty()) {
      // species a family may found before it hands the rest of its births back; the counters of family f start at joinOff_f
      famMaxNew_ = std::min<uint32_t>(famCap_ - maxSnap, std::max<uint32_t>(16u, std::min<uint32_t>(256u, uint32_t(65536 / fams_.size()))));
      famMaxNew_ = std::max<uint32_t>(famMaxNew_, 1u);
      for (Family &F: fams_) { F.joinOff = joinOff; joinOff += 2 * (F.snapN + famMaxNew_); }
    }
  } else { fams_.clear(); grpFam_.assign(G, -1); }
  uploadLut();
  mark("children");
  timers[0] += now() - t0;
  t0 = now();

  // ---- device ----
  const uint32_t *dBirthRows = nullptr;      // [n][wp_] on the device
  if (P || G) {
    dBMeta_.reserve(n);
    if (dPayload) {
      // payloads are already resident: used in place when rows need no padding
      if (wp_ == W_) dBirthRows = dPayload;
      else {
        dBirths_.reserve(size_t(n) * wp_);
        launch_pad_rows(dPayload, n, W_, wp_, dBirths_.p, stream_);
        counters[4]++;
        dBirthRows = dBirths_.p;
      }
    } else {
      if (!birthsUploaded) uploadBirths();
      dBirthRows = dBirths_.p;
    }
    if (chunkUniform_) {
      float fc;  memcpy(&fc, &chunkCls_, 4);
      // generation after generation of the same triple: the array filled for an earlier chunk still holds it
      const float4 v = make_float4(compat[0], compat[1], compat[2], fc);
      if (!(fillPtr_ == dBMeta_.p && fillN_ >= n && memcmp(&fillValue_, &v, sizeof v) == 0)) {
        launch_fill_f4(dBMeta_.p, n, v, stream_);
        counters[4]++;
        fillPtr_ = dBMeta_.p;  fillN_ = n;  fillValue_ = v;
      }
    } else {
      cudaCheck(cudaMemcpyAsync(dBMeta_.p, hBMeta_.p, size_t(n) * sizeof(float4), cudaMemcpyHostToDevice, stream_), "birth meta");
      fillPtr_ = nullptr;
    }
  }
  if (P) {
    // speculative scores of every (birth, parent species) pair of the two-parent births
    dPairBirth_.reserve(P);  dPairSlot_.reserve(P);  dScore_.reserve(P);  dDist_.reserve(P * K_);
    hScore_.reserve(P);  hDist_.reserve(P * K_);
    cudaCheck(cudaMemcpyAsync(dPairBirth_.p, hPairBirth_.p, P * 4, cudaMemcpyHostToDevice, stream_), "pairs");
    cudaCheck(cudaMemcpyAsync(dPairSlot_.p, hPairSlot_.p, P * 4, cudaMemcpyHostToDevice, stream_), "pairs");
    cudaCheck(cudaEventRecord(evK0_, stream_), "event record");
    launch_score_pairs(poolView(), scoreCfg(), dBirthRows, dBMeta_.p, uint32_t(P), dPairBirth_.p, dPairSlot_.p,
                       dScore_.p, dDist_.p, stream_);
    cudaCheck(cudaEventRecord(evK1_, stream_), "event record");
    counters[4]++;  counters[9]++;
    cudaCheck(cudaGetLastError(), "score_pairs launch");
    cudaCheck(cudaMemcpyAsync(hScore_.p, dScore_.p, P * 4, cudaMemcpyDeviceToHost, stream_), "scores");
    cudaCheck(cudaMemcpyAsync(hDist_.p, dDist_.p, P * K_ * 4, cudaMemcpyDeviceToHost, stream_), "distances");
    counters[7] += P;
  }
  uint32_t nRows = 0;
  nFt_ = 0;  nFtRows_ = 0;
  bool ftRows = false;
  uint32_t specRows = 0, specFt = 0;        // rows / rejected births fetched ahead of knowing their number
  uint32_t specSlow = 0, nSlow = 0;
  const bool devList = streamOk_ && n >= devListMin_;      // (small passes: the host scans the result words itself)
  bool directOut = false;
  std::function<void(void)> launchFamilies;     // K3 is only launched once the pass is known to have rejected births (below)
  if (G) {
    // the ordered resolution of the single-parent births, one warp per species stream; it runs
    // AFTER the speculative kernel, which must see the pool as it was when the chunk was submitted
    // few, larger transfers: every call below costs the host ~3 us before the kernel can start
    dOrder_.reserve(nStream);  dGrpAll_.reserve(3 * size_t(G) + 1);  hGrpAll_.reserve(3 * size_t(G) + 1);  dAct_.reserve(n);
    dCtr_.reserve(8);  hCtr_.reserve(16);  dRows_.reserve((size_t(nStream) + 8 * size_t(G)) * K_);
    dOut2_.reserve(2 * size_t(n));  dGrpSum_.reserve(G);
    hAct_.reserve(n);  hOutSid_.reserve(n);  hOutSlot_.reserve(n);  hGrpSum_.reserve(G);
    uint32_t *dGrpSlot = dGrpAll_.p, *dGrpSid = dGrpAll_.p + G, *dGrpOff = dGrpAll_.p + 2 * size_t(G);
    uint32_t *dOutSid = dOut2_.p;
    int32_t *dOutSlot = reinterpret_cast<int32_t*>(dOut2_.p + n);
    memcpy(hGrpAll_.p, hGrpSlot_.p, size_t(G) * 4);
    memcpy(hGrpAll_.p + G, hGrpSid_.p, size_t(G) * 4);
    memcpy(hGrpAll_.p + 2 * size_t(G), hGrpOff_.p, (size_t(G) + 1) * 4);
    cudaCheck(cudaMemcpyAsync(dGrpAll_.p, hGrpAll_.p, (3 * size_t(G) + 1) * 4, cudaMemcpyHostToDevice, stream_), "groups");
    if (!devGrouped_) cudaCheck(cudaMemcpyAsync(dOrder_.p, hOrder_.p, nStream * 4, cudaMemcpyHostToDevice, stream_), "order");
    if (!earlyClears) clearResults();
    // (kept behind the copies on purpose: the copy-engine -> SM hand-over then sits in front of this 8-byte clear and
    // not between the two events that time the kernel)
    // [0] rows claimed, [1] births rejected at top level, [2..5] the stream kernel's time span (two 64-bit words)
    cudaCheck(cudaMemsetAsync(dCtr_.p, 0, 24, stream_), "counter clear");
    StreamArgs a;
    a.pool = poolView();
    a.pool_payload_w = dPool_.p;  a.pool_cmeta_w = dMeta_.p;  a.pool_fill_w = dFill_.p;  a.pool_dist = dPoolDist_.p;
    a.cfg = scoreCfg();  a.crit = p_.env_crit;  a.guard = GUARD;
    a.b_payload = dBirthRows;  a.b_cmeta = dBMeta_.p;
    a.n_groups = G;  a.max_stream = maxStream;  a.grp_slot = dGrpSlot;  a.grp_off = dGrpOff;  a.order = dOrder_.p;
    a.act = dAct_.p;  a.rows = dRows_.p;  a.row_counter = dCtr_.p;
    a.span = reinterpret_cast<unsigned long long*>(dCtr_.p + 2);
    a.row_payload = nullptr;
    if (!payload) { dRowPayload_.reserve(size_t(nStream) * wp_ + 4); a.row_payload = dRowPayload_.p; }
    a.grp_sid = dGrpSid;  a.out_sid = dOutSid;  a.out_slot = dOutSlot;  a.grp_sum = dGrpSum_.p;
    a.ft_list = nullptr;  a.ft_counter = nullptr;
    const uint32_t nFam = uint32_t(fams_.size());
    if (nFam) {
      dRejOrder_.reserve(nStream);  dRejFill_.reserve(nStream);  dRejCount_.reserve(G);
      a.rej_order = dRejOrder_.p;  a.rej_fill = dRejFill_.p;  a.rej_count = dRejCount_.p;
    }
    if (useChildKernel) {
      dFtList_.reserve(2 * nStream);
      // everything the sub-species pass may need is allocated with the first stream pass, not on the step in which the
      // first sub-species appears (a steady step): page-locking five small arrays there cost 1.5-3 ms on multi-GPU hosts
      dChildOff_.reserve(G + 1);  dFtFirst_.reserve(nStream);  dFtFlags_.reserve(nStream);  dFtScore_.reserve(nStream);
      dFtDist_.reserve(nStream * K_);
      hFtList_.reserve(2 * 2048);  hFtFirst_.reserve(2048);  hFtFlags_.reserve(2048);  hFtScore_.reserve(2048);
      hFtDist_.reserve(size_t(2048) * K_);
      a.ft_list = dFtList_.p;  a.ft_counter = dCtr_.p + 1;
    }
    const char *profPath = streamProf_.empty() ? nullptr : streamProf_.c_str();
    if (profPath) { dProf_.reserve(size_t(G) * 16); a.prof = dProf_.p; }
    a.ev_before = evS0_;                     // (recorded by the launcher, immediately in front of the launch)
    launch_resolve_streams(a, stream_);
    cudaCheck(cudaEventRecord(evS1_, stream_), "event record");
    counters[4]++;  counters[12]++;  counters[13] += nStream;
    cudaCheck(cudaGetLastError(), "resolve_streams launch");
    if (useChildKernel && nChild) {
      // the births the streams rejected walk the sub-species on the device too (pool as the streams left it)
      dChildOff_.reserve(G + 1);  dChildSlot_.reserve(nChild);
      dFtFirst_.reserve(nStream);  dFtFlags_.reserve(nStream);  dFtScore_.reserve(nStream);  dFtDist_.reserve(nStream * K_);
      cudaCheck(cudaMemcpyAsync(dChildOff_.p, hChildOff_.p, (G + 1) * 4, cudaMemcpyHostToDevice, stream_), "children");
      cudaCheck(cudaMemcpyAsync(dChildSlot_.p, hChildSlot_.p, nChild * 4, cudaMemcpyHostToDevice, stream_), "children");
      ChildArgs c;
      c.pool = poolView();  c.cfg = scoreCfg();  c.guard = GUARD;
      c.b_payload = dBirthRows;  c.b_cmeta = dBMeta_.p;
      c.ft_list = dFtList_.p;  c.ft_counter = dCtr_.p + 1;  c.max_ft = uint32_t(nStream);
      c.child_off = dChildOff_.p;  c.child_slot = dChildSlot_.p;
      c.out_first = dFtFirst_.p;  c.out_flags = dFtFlags_.p;  c.out_score = dFtScore_.p;  c.out_dist = dFtDist_.p;
      launch_classify_children(c, stream_);
      counters[4]++;
      cudaCheck(cudaGetLastError(), "classify_children launch");
    }
    if (nFam) launchFamilies = [&, nFam, dGrpOff, dOutSid, dOutSlot, famRowPayload = a.row_payload] (void) {
      // K3: the rejected births of every eligible family walk its sub-species on the device, in order, founding included
      // (family.cu).  Species founded there take consecutive pool slots behind the ones in use; the host makes them
      // real when it commits the founding births.
      const uint32_t reserve = famReserve(nStream);
      famSlotBase_ = nSlots_;
      const size_t slots = size_t(nSlots_) + reserve;
      dPool_.reserve(slots * K_ * wp_, true, stream_);  dMeta_.reserve(slots * K_, true, stream_);
      dPoolDist_.reserve(slots * K_ * K_, true, stream_);
      { const size_t oldCap = dFill_.cap;
        dFill_.reserve(slots, true, stream_);
        if (dFill_.cap != oldCap) cudaCheck(cudaMemsetAsync(dFill_.p + oldCap, 0, (dFill_.cap - oldCap) * 4, stream_), "fill clear"); }
      size_t joinWords = 0;
      for (const Family &F: fams_) joinWords = std::max<size_t>(joinWords, F.joinOff + 2 * (size_t(F.snapN) + famMaxNew_));
      hFamTab_.reserve(2 * size_t(nFam));  dFamTab_.reserve(2 * size_t(nFam));
      for (uint32_t f = 0; f < nFam; f++) { hFamTab_.p[f] = fams_[f].g;  hFamTab_.p[nFam + f] = fams_[f].joinOff; }
      dChildSid_.reserve(nChild + 1);  dFamAct_.reserve(n);  dFamRow_.reserve(n);  dChildJoin_.reserve(joinWords + 1);
      dNewFounder_.reserve(size_t(nFam) * famMaxNew_ * 2);  dFamInfo_.reserve(nFam);  dFamStats_.reserve(2 * size_t(nFam));
      dFamSlotCtr_.reserve(1);
      if (!(useChildKernel && nChild)) {        // (the sub-species lists usually went up for K2 just above)
        dChildOff_.reserve(G + 1);  dChildSlot_.reserve(nChild + 1);
        cudaCheck(cudaMemcpyAsync(dChildOff_.p, hChildOff_.p, (G + 1) * 4, cudaMemcpyHostToDevice, stream_), "children");
        if (nChild) cudaCheck(cudaMemcpyAsync(dChildSlot_.p, hChildSlot_.p, nChild * 4, cudaMemcpyHostToDevice, stream_), "children");
      }
      if (nChild) cudaCheck(cudaMemcpyAsync(dChildSid_.p, hChildSid_.p, nChild * 4, cudaMemcpyHostToDevice, stream_), "children");
      cudaCheck(cudaMemcpyAsync(dFamTab_.p, hFamTab_.p, 2 * size_t(nFam) * 4, cudaMemcpyHostToDevice, stream_), "families");
      cudaCheck(cudaMemsetAsync(dFamAct_.p, 0, size_t(n) * 4, stream_), "family results clear");
      cudaCheck(cudaMemsetAsync(dFamSlotCtr_.p, 0, 4, stream_), "family slot counter clear");
      FamilyArgs fa;
      fa.pool = poolView();  fa.pool.n_slots = uint32_t(slots);
      fa.pool_payload_w = dPool_.p;  fa.pool_cmeta_w = dMeta_.p;  fa.pool_fill_w = dFill_.p;  fa.pool_dist = dPoolDist_.p;
      fa.cfg = scoreCfg();  fa.crit = p_.env_crit;  fa.guard = GUARD;
      fa.b_payload = dBirthRows;  fa.b_cmeta = dBMeta_.p;
      fa.n_fam = nFam;  fa.fam_group = dFamTab_.p;  fa.fam_join_off = dFamTab_.p + nFam;
      fa.grp_off = dGrpOff;  fa.rej_order = dRejOrder_.p;  fa.rej_fill = dRejFill_.p;  fa.rej_count = dRejCount_.p;
      fa.child_off = dChildOff_.p;  fa.child_slot = dChildSlot_.p;  fa.child_sid = dChildSid_.p;
      fa.slot_counter = dFamSlotCtr_.p;  fa.slot_base = famSlotBase_;  fa.slot_limit = reserve;
      fa.max_new = famMaxNew_;
      uint32_t maxSnap = 0;
      for (const Family &F: fams_) maxSnap = std::max(maxSnap, F.snapN);
      fa.cap = std::min<uint32_t>(famCap_, (maxSnap + famMaxNew_ + 3u) & ~3u);
      fa.fam_act = dFamAct_.p;  fa.fam_row = dFamRow_.p;  fa.out_sid = dOutSid;  fa.out_slot = dOutSlot;
      fa.rows = dRows_.p;  fa.row_payload = famRowPayload;  fa.row_counter = dCtr_.p;
      fa.fam_info = dFamInfo_.p;  fa.fam_stats = dFamStats_.p;  fa.child_join = dChildJoin_.p;
      fa.new_founder = dNewFounder_.p;  fa.new_slot = dNewFounder_.p + size_t(nFam) * famMaxNew_;
      cudaCheck(cudaEventRecord(evF0_, stream_), "event record");
      launch_resolve_families(fa, stream_);
      cudaCheck(cudaEventRecord(evF1_, stream_), "event record");
      counters[4]++;
      cudaCheck(cudaGetLastError(), "resolve_families launch");
      hFamInfo_.reserve(nFam);
      cudaCheck(cudaMemcpyAsync(hFamInfo_.p, dFamInfo_.p, size_t(nFam) * sizeof(uint4), cudaMemcpyDeviceToHost, stream_), "family info");
    };
    if (!payload && useChildKernel) {
      // device-resident births: the host will want the genomes of the births the streams rejected (it walks the
      // sub-species for them); they are gathered from the device-side list right away, so that they come back
      // with the distance rows instead of costing a round trip of their own
      dFtRows_.reserve(size_t(nStream) * W_);
      // (first-use allocations of the rarely needed host-driven gather, taken here for the same reason as above)
      hFtRows_.reserve(size_t(2048) * W_);  hSlowIdx_.reserve(1024);  dSlowIdx_.reserve(1024);
      dSlowRows_.reserve(size_t(1024) * W_);  hSlowRows_.reserve(size_t(1024) * W_);
      launch_gather_listed_rows(dPayload, dFtList_.p, dCtr_.p + 1, uint32_t(nStream), W_, dFtRows_.p, stream_);
      counters[4]++;
      ftRows = true;
    }
    if (devList) {
      // the births the host will have to look at, in arrival order (it reads this list instead of every result word)
      dSlowBlk_.reserve(list_unsettled_blocks(n));  dSlowList_.reserve(2 * size_t(n));  dSlowCtr_.reserve(1);
      launch_list_unsettled(dAct_.p, nullptr, n, dSlowBlk_.p, dSlowList_.p, dSlowCtr_.p, stream_);
      counters[4] += 2;
      cudaCheck(cudaGetLastError(), "list_unsettled launch");
      cudaCheck(cudaMemcpyAsync(hCtr_.p + 9, dSlowCtr_.p, 4, cudaMemcpyDeviceToHost, stream_), "listed births");
      hSlowList_.reserve(2 * std::max<size_t>(2048, nStream / 4));
      if (estSlow_) {
        specSlow = uint32_t(std::min<size_t>(estSlow_, n));
        hSlowList_.reserve(2 * size_t(specSlow));
        cudaCheck(cudaMemcpyAsync(hSlowList_.p, dSlowList_.p, 2 * size_t(specSlow) * 4, cudaMemcpyDeviceToHost, stream_), "list");
      }
    }
    cudaCheck(cudaMemcpyAsync(hAct_.p, dAct_.p, size_t(n) * 4, cudaMemcpyDeviceToHost, stream_), "act");
    cudaCheck(cudaMemcpyAsync(hCtr_.p, dCtr_.p, 24, cudaMemcpyDeviceToHost, stream_), "counters");
    // results go straight into the caller's arrays when those are pinned (no staging copy on the host)
    directOut = outSid && outSlot && outPinned_;
    cudaCheck(cudaMemcpyAsync(directOut ? outSid : hOutSid_.p, dOutSid, size_t(n) * 4, cudaMemcpyDeviceToHost, stream_), "sids");
    cudaCheck(cudaMemcpyAsync(directOut ? reinterpret_cast<uint32_t*>(outSlot) : reinterpret_cast<uint32_t*>(hOutSlot_.p), dOutSlot,
                              size_t(n) * 4, cudaMemcpyDeviceToHost, stream_), "slots");
    cudaCheck(cudaMemcpyAsync(hGrpSum_.p, dGrpSum_.p, size_t(G) * sizeof(uint4), cudaMemcpyDeviceToHost, stream_), "sums");
    // how many distance rows (enveloppe entries) and rejected births there are is only known after the pass; about as
    // many as last time are fetched right away, so that the usual step needs no second round trip
    // (page-locked result arrays are sized by the pass -- an eighth of its births may enter an enveloppe, or be rejected,
    // before anything has to grow -- and not by the counts seen so far: those come from the shorter passes of the
    // adaptive phase, and a re-allocation would then hit the first full-size steps: 1.5-2 ms each on multi-GPU hosts)
    const size_t roomy = std::max<size_t>(1024, nStream / 8);
    hRows_.reserve(roomy * K_);
    if (!payload) hRowPayload_.reserve(roomy * wp_);
    if (ftRows) { hFtRows_.reserve(roomy * W_);  hFtList_.reserve(2 * roomy); }
    if (estRows_) {
      specRows = uint32_t(std::min<size_t>(estRows_, nStream));
      hRows_.reserve(size_t(specRows) * K_);
      cudaCheck(cudaMemcpyAsync(hRows_.p, dRows_.p, size_t(specRows) * K_ * 4, cudaMemcpyDeviceToHost, stream_), "rows");
      if (!payload) {
        hRowPayload_.reserve(size_t(specRows) * wp_);
        cudaCheck(cudaMemcpyAsync(hRowPayload_.p, dRowPayload_.p, size_t(specRows) * wp_ * 4, cudaMemcpyDeviceToHost, stream_), "row genomes");
      }
    }
    if (ftRows && estFt_) {
      specFt = uint32_t(std::min<size_t>(estFt_, nStream));
      hFtRows_.reserve(size_t(specFt) * W_);  hFtList_.reserve(2 * size_t(specFt));
      cudaCheck(cudaMemcpyAsync(hFtRows_.p, dFtRows_.p, size_t(specFt) * W_ * 4, cudaMemcpyDeviceToHost, stream_), "rejected births' rows");
      cudaCheck(cudaMemcpyAsync(hFtList_.p, dFtList_.p, 2 * size_t(specFt) * 4, cudaMemcpyDeviceToHost, stream_), "ft list");
    }
  }
  // host work that needs no device result runs under the device pass
  mark("enqueue");
  if (G && outReplaced) memset(outReplaced, 0xFF, size_t(n) * 4);
  grpState_.assign(G, GroupState{});
  for (uint32_t g = 0; g < G; g++) grpState_[g].snapChildren = uint32_t(grpSpecies_[g]->children.size());
  mark("memset");
  if (P || G) {
    cudaCheck(cudaStreamSynchronize(stream_), "device pass");
    mark("sync1");
    float ms = 0.f;
    if (P) { cudaCheck(cudaEventElapsedTime(&ms, evK0_, evK1_), "event elapsed"); timers[4] += ms; }
    if (G) {
      cudaCheck(cudaEventElapsedTime(&ms, evS0_, evS1_), "event elapsed");
      {
        // the kernel's own span -- first CTA in to last CTA out, by the device's nanosecond timer -- kept beside the event
        // time (apt_tree_kernel_span): the events also see the gap between the host's record call and its launch call
        // whenever the stream is idle at that point (it is, since the grouping moved to the device), ~5 us of no kernel
        unsigned long long sp[2];
        memcpy(sp, hCtr_.p + 2, sizeof sp);
        const unsigned long long t0 = ~sp[0], t1 = sp[1];
        if (sp[0] && t1 > t0) { streamSpanMs_ += double(t1 - t0) * 1e-6;  streamSpanN_++; }
      }
      timers[5] += ms;
      if (const char *profPath = streamProf_.empty() ? nullptr : streamProf_.c_str()) {
        // development aid: one text line per stream {launch, stream, births, kernel ms, 10 timeline words}, appended
        std::vector<unsigned long long> hp (size_t(G) * 16);
        cudaCheck(cudaMemcpy(hp.data(), dProf_.p, hp.size() * 8, cudaMemcpyDeviceToHost), "stream profile");
        if (FILE *f = fopen(profPath, "a")) {
          for (uint32_t g = 0; g < G; g++) {
            fprintf(f, "%llu %u %u %.3f", (unsigned long long) counters[12], g, hGrpOff_.p[g + 1] - hGrpOff_.p[g], double(ms));
            for (int q = 0; q < 14; q++) fprintf(f, " %llu", hp[size_t(g) * 16 + q]);
            fputc('\n', f);
          }
          fclose(f);
        }
      }
      // The family pass runs only when the streams rejected a fair share of their births (a radiation): a steady pass does
      // not pay for it.  It appends its rows behind the streams' and is awaited here.
      if (launchFamilies && hCtr_.p[1] >= std::max<uint32_t>(32u, n / 64u)) {
        launchFamilies();
        famFounded_.resize(std::max(famFounded_.size(), fams_.size()));
        for (size_t f = 0; f < fams_.size(); f++) famFounded_[f].clear();
        cudaCheck(cudaMemcpyAsync(hCtr_.p, dCtr_.p, 8, cudaMemcpyDeviceToHost, stream_), "counters");
        // (the species ids / slots of the births it decided: fetched again)
        cudaCheck(cudaMemcpyAsync(directOut ? outSid : hOutSid_.p, dOut2_.p, size_t(n) * 4, cudaMemcpyDeviceToHost, stream_), "sids");
        cudaCheck(cudaMemcpyAsync(directOut ? reinterpret_cast<uint32_t*>(outSlot) : reinterpret_cast<uint32_t*>(hOutSlot_.p), dOut2_.p + n,
                                  size_t(n) * 4, cudaMemcpyDeviceToHost, stream_), "slots");
        if (devList) {              // (births that joined a sub-species there are settled now: listed again)
          launch_list_unsettled(dAct_.p, dFamAct_.p, n, dSlowBlk_.p, dSlowList_.p, dSlowCtr_.p, stream_);
          counters[4] += 2;
          cudaCheck(cudaGetLastError(), "list_unsettled launch");
          cudaCheck(cudaMemcpyAsync(hCtr_.p + 9, dSlowCtr_.p, 4, cudaMemcpyDeviceToHost, stream_), "listed births");
          specSlow = 0;
        }
        cudaCheck(cudaStreamSynchronize(stream_), "family pass");
        float fm = 0.f;
        cudaCheck(cudaEventElapsedTime(&fm, evF0_, evF1_), "event elapsed");
        timers[6] += fm;
        specRows = 0;               // rows fetched ahead of the family pass do not hold its entries yet: fetched again below
      } else fams_.clear();
      nRows = hCtr_.p[0];
      nFt_ = (useChildKernel && nChild) ? hCtr_.p[1] : 0u;
      if (nRows > specRows) {
        hRows_.reserve(size_t(nRows) * K_);
        cudaCheck(cudaMemcpyAsync(hRows_.p, dRows_.p, size_t(nRows) * K_ * 4, cudaMemcpyDeviceToHost, stream_), "rows");
        if (!payload) {
          hRowPayload_.reserve(size_t(nRows) * wp_);
          cudaCheck(cudaMemcpyAsync(hRowPayload_.p, dRowPayload_.p, size_t(nRows) * wp_ * 4, cudaMemcpyDeviceToHost, stream_), "row genomes");
        }
      }
      if (nFt_) {
        hFtList_.reserve(2 * size_t(nFt_));  hFtFirst_.reserve(nFt_);  hFtFlags_.reserve(nFt_);
        hFtScore_.reserve(nFt_);  hFtDist_.reserve(size_t(nFt_) * K_);
        cudaCheck(cudaMemcpyAsync(hFtList_.p, dFtList_.p, 2 * size_t(nFt_) * 4, cudaMemcpyDeviceToHost, stream_), "ft list");
        cudaCheck(cudaMemcpyAsync(hFtFirst_.p, dFtFirst_.p, size_t(nFt_) * 4, cudaMemcpyDeviceToHost, stream_), "ft first");
        cudaCheck(cudaMemcpyAsync(hFtFlags_.p, dFtFlags_.p, size_t(nFt_) * 4, cudaMemcpyDeviceToHost, stream_), "ft flags");
        cudaCheck(cudaMemcpyAsync(hFtScore_.p, dFtScore_.p, size_t(nFt_) * 4, cudaMemcpyDeviceToHost, stream_), "ft score");
        cudaCheck(cudaMemcpyAsync(hFtDist_.p, dFtDist_.p, size_t(nFt_) * K_ * 4, cudaMemcpyDeviceToHost, stream_), "ft dist");
      }
      nFtRows_ = ftRows ? std::min(hCtr_.p[1], uint32_t(nStream)) : 0u;
      if (nFtRows_ > specFt) {
        hFtRows_.reserve(size_t(nFtRows_) * W_);
        cudaCheck(cudaMemcpyAsync(hFtRows_.p, dFtRows_.p, size_t(nFtRows_) * W_ * 4, cudaMemcpyDeviceToHost, stream_), "rejected births' rows");
        if (!nFt_) {
          hFtList_.reserve(2 * size_t(nFtRows_));
          cudaCheck(cudaMemcpyAsync(hFtList_.p, dFtList_.p, 2 * size_t(nFtRows_) * 4, cudaMemcpyDeviceToHost, stream_), "ft list");
        }
      }
      // what the family pass decided comes back only when it decided anything (a second round trip, radiations only)
      famResults_ = false;
      famSlotsUsed_ = 0;
      if (!fams_.empty()) {
        const uint32_t nFam = uint32_t(fams_.size());
        uint64_t decided = 0;
        for (uint32_t f = 0; f < nFam; f++) decided += hFamInfo_.p[f].x;
        if (decided) {
          size_t joinWords = 0;
          for (const Family &F: fams_) joinWords = std::max<size_t>(joinWords, F.joinOff + 2 * (size_t(F.snapN) + famMaxNew_));
          hFamAct_.reserve(n);  hFamRow_.reserve(n);  hChildJoin_.reserve(joinWords + 1);
          hNewFounder_.reserve(size_t(nFam) * famMaxNew_ * 2 + 1);  hFamStats_.reserve(2 * size_t(nFam));
          cudaCheck(cudaMemcpyAsync(hFamAct_.p, dFamAct_.p, size_t(n) * 4, cudaMemcpyDeviceToHost, stream_), "family results");
          cudaCheck(cudaMemcpyAsync(hFamRow_.p, dFamRow_.p, size_t(n) * 4, cudaMemcpyDeviceToHost, stream_), "family rows");
          cudaCheck(cudaMemcpyAsync(hChildJoin_.p, dChildJoin_.p, joinWords * 4, cudaMemcpyDeviceToHost, stream_), "family counters");
          cudaCheck(cudaMemcpyAsync(hNewFounder_.p, dNewFounder_.p, size_t(nFam) * famMaxNew_ * 2 * 4, cudaMemcpyDeviceToHost, stream_), "family founders");
          cudaCheck(cudaMemcpyAsync(hFamStats_.p, dFamStats_.p, 2 * size_t(nFam) * 8, cudaMemcpyDeviceToHost, stream_), "family stats");
          cudaCheck(cudaMemcpyAsync(hCtr_.p + 8, dFamSlotCtr_.p, 4, cudaMemcpyDeviceToHost, stream_), "family slots");
          famResults_ = true;
        }
      }
      if (devList) {
        nSlow = std::min(hCtr_.p[9], n);
        if (nSlow > specSlow) {
          hSlowList_.reserve(2 * size_t(nSlow));
          cudaCheck(cudaMemcpyAsync(hSlowList_.p, dSlowList_.p, 2 * size_t(nSlow) * 4, cudaMemcpyDeviceToHost, stream_), "list");
        }
        estSlow_ = nSlow + nSlow / 4 + 64;
      }
      if (nRows > specRows || nFt_ || nFtRows_ > specFt || famResults_ || nSlow > specSlow) cudaCheck(cudaStreamSynchronize(stream_), "rows");
      if (famResults_) {
        famSlotsUsed_ = std::min<uint32_t>(hCtr_.p[8], famReserve(nStream));
        if (famSlotsUsed_) {
          // the species founded on the device live in pool slots [famSlotBase_, +used): they become real now
          nSlots_ = famSlotBase_ + famSlotsUsed_;
          hPayload_.resize(size_t(nSlots_) * K_ * wp_, true);
        }
      }
      estRows_ = nRows + nRows / 4 + 64;
      estFt_ = ftRows ? nFtRows_ + nFtRows_ / 2 + 256 : estFt_;     // (few, and their number varies more)
    }
  }
  mark("sync2");
  timers[1] += now() - t0;
  t0 = now();

  // ---- host: ordered commit.  Births the device accepted without touching an enveloppe need no
  //      individual attention: their outputs were written by the kernel and their counters are
  //      booked per stream (settleStream).  Everything else -- enveloppe entries, births rejected at
  //      top level, two-parent births, streams handed back -- is committed here in arrival order. ----
  curPayload_ = payload;
  slow_.clear();  slowGrp_.clear();  slowAct_.clear();  requeue_.clear();  grpOwner_.clear();
  if (G && devList) {
    // the device listed the births the host has to look at (arrival order); the outputs of the others are in place
    const unsigned T = (pool_ && n >= parMin_) ? pool_->size() : 1u;
    if (!directOut) {
      const uint32_t per = (n + T - 1) / T;
      auto copyOut = [&] (unsigned t) {
        const uint32_t lo = std::min(n, t * per), hi = std::min(n, lo + per);
        if (outSid) memcpy(outSid + lo, hOutSid_.p + lo, size_t(hi - lo) * 4);
        if (outSlot) memcpy(outSlot + lo, hOutSlot_.p + lo, size_t(hi - lo) * 4);
      };
      if (T > 1) pool_->run(copyOut); else copyOut(0);
    }
    mark("commit-init");
    mark("scan");
    const size_t total = nSlow;
    slow_.resize(total);  slowGrp_.resize(total);  slowAct_.resize(total);
    const uint32_t *list = hSlowList_.p;
    for (size_t q = 0; q < total; q++) {
      if (q + 16 < total) __builtin_prefetch(&bg_[list[2 * (q + 16)]]);
      const uint32_t i = list[2 * q];
      if (i >= n) throw ApiError(APT_ERR_STATE, "device list of unsettled births is corrupt");
      const int32_t bg = bg_[i];
      slow_[q] = i;  slowGrp_[q] = bg;  slowAct_[q] = list[2 * q + 1];
      if (bg >= 0 && (list[2 * q + 1] & ACT_CODE_MASK) == ACT_ENTERED) grpState_[size_t(bg)].lastEnter = i;
    }
  } else if (G) {
    // outputs of the births the device settled + the list of births the host has to look at
    // (one pass over the result words; split over the worker pool for large chunks)
    const unsigned T = (pool_ && n >= parMin_) ? pool_->size() : 1u;
    parRest_.resize(std::max<size_t>(parRest_.size(), T));
    const uint32_t per = (n + T - 1) / T;
    auto scan = [&] (unsigned t) {
      const uint32_t lo = std::min(n, t * per), hi = std::min(n, lo + per);
      if (!directOut) {
        if (outSid) memcpy(outSid + lo, hOutSid_.p + lo, size_t(hi - lo) * 4);
        if (outSlot) memcpy(outSlot + lo, hOutSlot_.p + lo, size_t(hi - lo) * 4);
      }
      // the stream index and result word of every listed birth are kept next to it: the passes below read them
      // in list order instead of gathering them from the per-birth arrays again
      auto &rest = parRest_[t];
      rest.clear();
      for (uint32_t i = lo; i < hi; i++) {
        const int32_t bg = bg_[i];
        const uint32_t act = hAct_.p[i];
        // settled on the device: accepted with the enveloppe left alone -- at top level, or (family pass) by a sub-species
        const bool settled = bg >= 0 && (act == ACT_ACCEPTED ||
                                         (famResults_ && act == ACT_FALLTHROUGH && (hFamAct_.p[i] & FAM_CODE_MASK) == FAM_JOINED));
        if (!settled) { rest.push_back(i);  rest.push_back(uint32_t(bg));  rest.push_back(act); }
      }
    };
    mark("commit-init");
    if (T > 1) pool_->run(scan); else scan(0);
    mark("scan");
    // the workers' lists, concatenated (each worker copies its own to its offset)
    size_t total = 0;
    std::vector<size_t> at (T + 1, 0);
    for (unsigned t = 0; t < T; t++) { at[t] = total;  total += parRest_[t].size() / 3; }
    at[T] = total;
    slow_.resize(total);  slowGrp_.resize(total);  slowAct_.resize(total);
    auto merge = [&] (unsigned t) {
      const auto &rest = parRest_[t];
      size_t o = at[t];
      for (size_t q = 0; q < rest.size(); q += 3, o++) { slow_[o] = rest[q];  slowGrp_[o] = int32_t(rest[q + 1]);  slowAct_[o] = rest[q + 2]; }
    };
    if (T > 1 && total >= 1024) pool_->run(merge); else for (unsigned t = 0; t < T; t++) merge(t);
    for (size_t q = 0; q < total; q++)
      if (slowGrp_[q] >= 0 && (slowAct_[q] & ACT_CODE_MASK) == ACT_ENTERED) grpState_[size_t(slowGrp_[q])].lastEnter = slow_[q];
  } else {
    slow_.resize(n);
    slowGrp_.assign(n, -1);  slowAct_.assign(n, 0u);
    for (uint32_t i = 0; i < n; i++) slow_[i] = i;
  }
  mark("slow-list");
  // entry of the device-side list of rejected births (sub-species verdicts, fetched genomes) for slow_[q]
  ftOfSlow_.assign(slow_.size(), -1);
  ftRowOfSlow_.assign(nFtRows_ ? slow_.size() : 0, -1);
  for (uint32_t e = 0; e < std::max(nFt_, nFtRows_); e++) {
    auto it = std::lower_bound(slow_.begin(), slow_.end(), hFtList_.p[2 * e]);
    if (it != slow_.end() && *it == hFtList_.p[2 * e]) {
      if (e < nFt_) ftOfSlow_[size_t(it - slow_.begin())] = int32_t(e);
      if (e < nFtRows_) ftRowOfSlow_[size_t(it - slow_.begin())] = int32_t(e);
    }
  }
  curDevPayload_ = payload ? nullptr : dPayload;
  lateRows_.clear();
  gatherIdx_.clear();
  if (!payload && !slow_.empty()) {
    // device-resident births: the genomes of the births that entered an enveloppe came back with their distance
    // rows; the other births the host has to look at (rejected at top level, two parents, handed back) are
    // gathered on the device and fetched
    for (size_t q = 0; q < slow_.size(); q++) {
      const bool entry = slowGrp_[q] >= 0 && (slowAct_[q] & (ACT_CODE_MASK | ACT_UNCERTAIN)) == ACT_ENTERED;    // = entryRowOf(slow_[q]) != null
      if (!entry && !(nFtRows_ && ftRowOfSlow_[q] >= 0)) gatherIdx_.push_back(slow_[q]);
    }
    const size_t ns = gatherIdx_.size();
    if (ns) {
      hSlowIdx_.reserve(ns);  dSlowIdx_.reserve(ns);  dSlowRows_.reserve(ns * W_);  hSlowRows_.reserve(ns * W_);
      memcpy(hSlowIdx_.p, gatherIdx_.data(), ns * 4);
      cudaCheck(cudaMemcpyAsync(dSlowIdx_.p, hSlowIdx_.p, ns * 4, cudaMemcpyHostToDevice, stream_), "slow list");
      launch_gather_rows(dPayload, dSlowIdx_.p, uint32_t(ns), W_, dSlowRows_.p, stream_);
      counters[4]++;
      cudaCheck(cudaMemcpyAsync(hSlowRows_.p, dSlowRows_.p, ns * W_ * 4, cudaMemcpyDeviceToHost, stream_), "slow rows");
      cudaCheck(cudaStreamSynchronize(stream_), "slow rows");
    }
  }
  mark("ft+gather");
  const double w1 = profile_ ? now() : 0.;
  size_t si = 0;
  // A run of births that are plain enveloppe entries of device streams (not taken over) touches nothing
  // but each stream's own species: worker t mirrors the entries of the streams it owns (in arrival
  // order), callbacks are merged back into arrival order afterwards.  Anything else (a birth rejected at
  // top level, a two-parent birth, a stream handed back) ends the run and is committed serially; the
  // next run starts behind it.  Returns the number of births it committed from slow_[from...).
  const size_t entryParMin = std::max<size_t>(1, parMin_ / 32);
  auto parallelRun = [&] (size_t from) -> size_t {
    size_t run = 0;
    while (from + run < slow_.size()) {
      const int32_t bg = slowGrp_[from + run];
      if (bg < 0 || (slowAct_[from + run] & (ACT_CODE_MASK | ACT_UNCERTAIN)) != ACT_ENTERED) break;
      if (grpSpecies_[size_t(bg)]->div_epoch == chunkEpoch_) break;
      run++;
    }
    if (run < entryParMin) return 0;
    curBirth_ = slow_[from];
    const unsigned T = pool_->size();
    const bool wantEvents = !(p_.flags & APT_FLAG_NO_EVENTS);
    struct Local { uint64_t comparisons = 0, n = 0; std::string error; };
    std::vector<Local> loc (T);
    // callbacks of the run: at most two per entry, written to the entry's own place and read back in arrival order
    if (wantEvents) { runEv_.resize(2 * run);  runEvN_.assign(run, 0); }
    parMine_.resize(std::max<size_t>(parMine_.size(), T));
    if (grpOwner_.size() != G) {
      // stream -> worker (a table: no division per entry in the filter below).  Contiguous blocks of streams, i.e. of
      // species ids: species founded one after the other sit next to each other in memory (records, enveloppe
      // rows in the host mirror), and a round-robin deal made neighbouring workers share those cache lines
      grpOwner_.resize(G);
      for (uint32_t g = 0; g < G; g++) grpOwner_[g] = uint8_t(uint64_t(g) * T / G);
    }
    pool_->run([&] (unsigned t) {
      Local &L = loc[t];
      // this worker's entries (streams t, t + T, ...), then one pass over them with the lines each one will touch
      // requested a few entries ahead: an entry reads ~15 cache lines nobody has touched since the last generation
      auto &mine = parMine_[t];
      mine.clear();
      for (size_t q = from; q < from + run; q++) if (grpOwner_[size_t(slowGrp_[q])] == t) mine.push_back(uint32_t(q));
      auto prefetchBirth = [&] (size_t q) {
        const uint32_t j = slow_[q];
        __builtin_prefetch(&gid[j]);  __builtin_prefetch(&mgid[j]);  __builtin_prefetch(&msid[j]);
        __builtin_prefetch(&fgid[j]);  __builtin_prefetch(&fsid[j]);  __builtin_prefetch(&compat[3 * size_t(j)]);
        if (payload) __builtin_prefetch(payload + size_t(j) * W_);
        else __builtin_prefetch(hRowPayload_.p + size_t(slowAct_[q] >> ACT_ROW_SHIFT) * wp_);
        __builtin_prefetch(hRows_.p + size_t(slowAct_[q] >> ACT_ROW_SHIFT) * K_);
        const Species *sp = grpSpecies_[size_t(slowGrp_[q])];
        __builtin_prefetch(sp, 1);  __builtin_prefetch(reinterpret_cast<const char*>(sp) + 64, 1);
        __builtin_prefetch(reinterpret_cast<const char*>(sp) + 128, 1);
      };
      auto prefetchSpecies = [&] (size_t q) {
        const Species *sp = grpSpecies_[size_t(slowGrp_[q])];
        const uint32_t r = (slowAct_[q] >> ACT_SLOT_SHIFT) & 31u;
        __builtin_prefetch(sp->reps.data() + r, 1);  __builtin_prefetch(sp->contribs.data(), 1);
        const char *dm = reinterpret_cast<const char*>(sp->dist.data());
        for (size_t o = 0; o < size_t(K_) * K_ * 4; o += 64) __builtin_prefetch(dm + o, 1);
        __builtin_prefetch(&hPayload_[(size_t(sp->slot) * K_ + r) * wp_], 1);
      };
      uint64_t nDone = 0, nCmp = 0;           // (kept off the shared Local array inside the loop: neighbours share its cache lines)
      try {
        for (size_t a = 0; a < mine.size(); a++) {
          if (a + 6 < mine.size()) prefetchBirth(mine[a + 6]);
          if (a + 3 < mine.size()) prefetchSpecies(mine[a + 3]);
          const size_t q = mine[a];
          const uint32_t i = slow_[q];
          const uint32_t gi = uint32_t(slowGrp_[q]);
          Species *s0 = grpSpecies_[gi];
          if (s0->pending > 0) s0->pending--;
          touchSnap(s0);
          nCmp += s0->k;
          RepMeta g { gid[i], mgid[i], msid[i], fgid[i], fsid[i], 0, 0, chunkUniform_ ? chunkCls_ : hCls_.p[i],
                      compat[3 * i], compat[3 * i + 1], compat[3 * i + 2] };
          int32_t slot = -1;
          uint32_t replaced = INV;
          apt_event ev[2];
          const uint32_t act = slowAct_[q];
          const uint32_t *w = payload ? payload + size_t(i) * W_ : hRowPayload_.p + size_t(act >> ACT_ROW_SHIFT) * wp_;
          const int ne = mirrorEntry(w, g, s0, act, slot, replaced, ev);
          if (wantEvents) { for (int e = 0; e < ne; e++) runEv_[2 * (q - from) + e] = ev[e];  runEvN_[q - from] = uint8_t(ne); }
          s0->count++;  s0->alive++;  s0->last = step_;
          if (g.msid != INV) selfContribution(s0, 1u + (g.msid == g.fsid));
          if (outReplaced) outReplaced[i] = replaced;
          nDone++;
        }
      } catch (const std::exception &e) { L.error = e.what(); }
      L.comparisons = nCmp;  L.n = nDone;
    });
    for (const Local &L: loc) if (!L.error.empty()) throw ApiError(APT_ERR_STATE, L.error);
    uint64_t total = 0;
    for (const Local &L: loc) { comparisons_ += uint32_t(L.comparisons); total += L.n; }
    insertions_ += uint32_t(total);
    counters[0] += total;  counters[1] += total;  counters[14] += total;
    if (wantEvents)
      for (size_t q = 0; q < run; q++)
        for (uint8_t e = 0; e < runEvN_[q]; e++) events_.push_back(runEv_[2 * q + e]);
    return run;
  };
  size_t nextTry = 0;                 // first position from which a parallel run is worth looking for again
  std::vector<uint32_t> heap;       // births handed back by a taken-over stream (min-heap)
  auto heapCmp = [] (uint32_t a, uint32_t b) { return a > b; };
  uint32_t walked = 0;
  try {
    for (;;) {
      for (uint32_t b: requeue_) { heap.push_back(b); std::push_heap(heap.begin(), heap.end(), heapCmp); }
      requeue_.clear();
      if (pool_ && G && heap.empty() && si >= nextTry && slow_.size() - si >= entryParMin) {
        const size_t done = parallelRun(si);
        if (done) { si += done; continue; }
        nextTry = si + entryParMin / 4 + 1;         // a short run: do not rescan at every birth
      }
      uint32_t i;
      if (!heap.empty() && (si >= slow_.size() || heap.front() < slow_[si])) {
        std::pop_heap(heap.begin(), heap.end(), heapCmp);
        i = heap.back();  heap.pop_back();
        curFt_ = -1;  curFtRow_ = -1;
      } else if (si < slow_.size()) { curFt_ = ftOfSlow_[si];  curFtRow_ = nFtRows_ ? ftRowOfSlow_[si] : -1;  i = slow_[si++]; }
      else break;
      curBirth_ = i;
      // the births looked at are sparse: pull what the next ones touch into the cache ahead of time
      if (si + 12 < slow_.size()) {
        const uint32_t j = slow_[si + 12];
        __builtin_prefetch(&bg_[j]);  __builtin_prefetch(&hAct_.p[j]);
        __builtin_prefetch(&gid[j]);  __builtin_prefetch(&mgid[j]);  __builtin_prefetch(&msid[j]);
        __builtin_prefetch(&fgid[j]);  __builtin_prefetch(&fsid[j]);  __builtin_prefetch(&compat[3 * size_t(j)]);
        if (payload) __builtin_prefetch(payload + size_t(j) * W_);
      }
      if (si + 6 < slow_.size()) {
        const int32_t bg = bg_[slow_[si + 6]];
        if (bg >= 0) {
          const Species *sp = grpSpecies_[size_t(bg)];
          __builtin_prefetch(sp);  __builtin_prefetch(reinterpret_cast<const char*>(sp) + 64);
          __builtin_prefetch(reinterpret_cast<const char*>(sp) + 128);
        }
      }
      if (si + 3 < slow_.size()) {
        const int32_t bg = bg_[slow_[si + 3]];
        if (bg >= 0) {
          const Species *sp = grpSpecies_[size_t(bg)];
          __builtin_prefetch(sp->reps.data(), 1);  __builtin_prefetch(sp->dist.data(), 1);
          __builtin_prefetch(sp->contribs.data(), 1);
          __builtin_prefetch(&hPayload_[size_t(sp->slot) * K_ * wp_], 1);
          if (!sp->children.empty()) __builtin_prefetch(&sp->children.back());
        }
      }
      // a birth its stream rejected walks the species' sub-species, newest first: their records are cold too
      if (si + 2 < slow_.size()) {
        const uint32_t j = slow_[si + 2];
        const int32_t bg = bg_[j];
        if (bg >= 0 && (hAct_.p[j] & ACT_CODE_MASK) == ACT_FALLTHROUGH) {
          const Species *sp = grpSpecies_[size_t(bg)];
          const size_t nc = sp->children.size();
          for (size_t c = 0; c < std::min<size_t>(nc, 2); c++) {
            const char *cp = reinterpret_cast<const char*>(sp->children[nc - 1 - c]);
            __builtin_prefetch(cp, 1);  __builtin_prefetch(cp + 64, 1);  __builtin_prefetch(cp + 128, 1);
          }
        }
      }
      // a birth outside the streams (two parent species, mostly) is scored on the host against both species and their
      // sub-species: a chain of dependent cold lines (id -> record -> list of sub-species -> their records -> their
      // genomes), requested stage by stage over the births in front of it
      auto parentsOf = [&] (uint32_t j, const Species *out[2]) {
        out[0] = msid[j] < nodes_.size() ? nodes_[msid[j]] : nullptr;
        out[1] = (fsid[j] != msid[j] && fsid[j] < nodes_.size()) ? nodes_[fsid[j]] : nullptr;
      };
      if (si + 9 < slow_.size()) {
        const uint32_t j = slow_[si + 9];
        if (bg_[j] < 0) {
          if (msid[j] < nodes_.size()) __builtin_prefetch(&nodes_[msid[j]]);
          if (fsid[j] < nodes_.size()) __builtin_prefetch(&nodes_[fsid[j]]);
        }
      }
      if (si + 7 < slow_.size() && bg_[slow_[si + 7]] < 0) {
        const Species *ps[2];
        parentsOf(slow_[si + 7], ps);
        for (const Species *sp: ps)
          if (sp) {
            const char *cp = reinterpret_cast<const char*>(sp);
            __builtin_prefetch(cp);  __builtin_prefetch(cp + 64);  __builtin_prefetch(cp + 128);
          }
      }
      if (si + 5 < slow_.size() && bg_[slow_[si + 5]] < 0) {
        const Species *ps[2];
        parentsOf(slow_[si + 5], ps);
        for (const Species *sp: ps)
          if (sp) {
            __builtin_prefetch(sp->reps.data());  __builtin_prefetch(sp->contribs.data(), 1);
            __builtin_prefetch(&hPayload_[size_t(sp->slot) * K_ * wp_]);
            if (!sp->children.empty()) __builtin_prefetch(&sp->children.back());
          }
      }
      if (si + 3 < slow_.size() && bg_[slow_[si + 3]] < 0) {
        const Species *ps[2];
        parentsOf(slow_[si + 3], ps);
        for (const Species *sp: ps)
          if (sp) {
            const size_t nc = sp->children.size();
            for (size_t c = 0; c < std::min<size_t>(nc, 2); c++) {
              const char *cp = reinterpret_cast<const char*>(sp->children[nc - 1 - c]);
              __builtin_prefetch(cp, 1);  __builtin_prefetch(cp + 64, 1);  __builtin_prefetch(cp + 128, 1);
            }
          }
      }
      if (si + 1 < slow_.size() && bg_[slow_[si + 1]] < 0) {
        const Species *ps[2];
        parentsOf(slow_[si + 1], ps);
        for (const Species *sp: ps)
          if (sp && !sp->children.empty()) {
            const Species *c = sp->children.back();
            __builtin_prefetch(c->reps.data());  __builtin_prefetch(&hPayload_[size_t(c->slot) * K_ * wp_]);
          }
      }
      Species *s0 = nullptr;
      uint32_t act = 0;
      if (bg_[i] >= 0) {
        s0 = grpSpecies_[size_t(bg_[i])];
        if (s0->div_epoch == chunkEpoch_) s0 = nullptr;
        else {
          act = hAct_.p[i];
          if ((act & ACT_UNCERTAIN) || (act & ACT_CODE_MASK) == ACT_UNPROCESSED) { diverge(s0); s0 = nullptr; }
        }
      }
      RepMeta g { gid[i], mgid[i], msid[i], fgid[i], fsid[i], 0, 0, chunkUniform_ ? chunkCls_ : hCls_.p[i],
                  compat[3 * i], compat[3 * i + 1], compat[3 * i + 2] };
      uint32_t sid = INV, replaced = INV;
      int32_t slot = -1;
      const uint32_t *w = payloadOf(i);
      const double p0t = profile_ ? now() : 0.;
      int pcat = 0;                       // 0 plain entry, 1 rejected at top level (stream), 2 everything else
      if (s0 && famResults_ && act == ACT_FALLTHROUGH && hFamAct_.p[i] != FAM_UNPROCESSED) {
        pcat = 1;
        walkFamily(i, w, g, s0, hFamAct_.p[i], sid, slot, replaced);
      }
      else if (s0) { pcat = (act & ACT_CODE_MASK) == ACT_FALLTHROUGH ? 1 : 0; walkStream(w, g, s0, act, sid, slot, replaced); }
      else {
        pcat = 2;
        int64_t p0 = -1, p1 = -1;
        auto it = std::lower_bound(pairOf_.begin(), pairOf_.end(), i, [] (const PairRef &p, uint32_t b) { return p.birth < b; });
        if (it != pairOf_.end() && it->birth == i) { p0 = it->p0; p1 = it->p1; }
        walkOne(w, g, p0, p1, sid, slot, replaced);
      }
      if (profile_) { profWalk_[pcat] += now() - p0t; profWalkN_[pcat]++; }
      if (outSid) outSid[i] = sid;
      if (outSlot) outSlot[i] = slot;
      if (outReplaced) outReplaced[i] = replaced;
      counters[0]++;
      walked++;
    }
    const double w2 = profile_ ? now() : 0.;
    mark("walk");
    for (uint32_t g = 0; g < G; g++) {
      if (g + 8 < G) {
        const char *sp = reinterpret_cast<const char*>(grpSpecies_[g + 8]);
        __builtin_prefetch(sp, 1);  __builtin_prefetch(sp + 64, 1);  __builtin_prefetch(sp + 128, 1);
      }
      if (g + 4 < G) __builtin_prefetch(grpSpecies_[g + 4]->contribs.data(), 1);
      if (grpSpecies_[g]->div_epoch != chunkEpoch_) settleStream(g, false, 0, nullptr);
    }
    if (famResults_) {
      settleFamilies();
      // births placed in a species founded in this pass carry a placeholder: the ids were handed out by the walk above
      if (outSid && famSlotsUsed_) {
        const uint32_t nFam = uint32_t(fams_.size());
        for (uint32_t i = 0; i < n; i++) {
          const uint32_t v = outSid[i];
          if (v != INV && (v & FAM_NEW_SID)) {
            const uint32_t f = (v & ~FAM_NEW_SID) / famMaxNew_, j = (v & ~FAM_NEW_SID) % famMaxNew_;
            if (f >= nFam || j >= famFounded_[f].size()) throw ApiError(APT_ERR_STATE, "family pass: unknown founded species");
            outSid[i] = famFounded_[f][j]->sid;
          }
        }
      }
    }
    mark("settle");
    if (profile_ && n >= parMin_) {
      fprintf(stderr, "chunk:");
      for (size_t q = 1; q < marks.size(); q++) fprintf(stderr, " %s %.3f", marks[q].first, 1e3 * (marks[q].second - marks[q - 1].second));
      fprintf(stderr, " ms\n");
    }
    if (profile_) fprintf(stderr, "walk: scan %.3f slow(%zu) %.3f settle %.3f ms\n", 1e3*(w1-t0), slow_.size(), 1e3*(w2-w1), 1e3*(now()-w2));
  } catch (...) {
    // a birth failed (unknown species, :169-174): everything before it stays committed, as when the
    // reference throws out of addGenome; streams are settled up to that birth
    const uint32_t failed = curBirth_;
    curBirth_ = failed;
    if (famResults_) {
      // the family pass decided births beyond the failing one too: its sub-species are rewritten from the host state, and
      // the births before the failure that joined one without changing it are booked one by one (their share of
      // Stats::comparisons / branching is only known per family and is not added)
      for (uint32_t i = 0; i < failed && i < n; i++) {
        if (bg_[i] < 0 || hAct_.p[i] != ACT_FALLTHROUGH || (hFamAct_.p[i] & FAM_CODE_MASK) != FAM_JOINED) continue;
        Species *S = grpSpecies_[size_t(bg_[i])];
        const Family &F = fams_[size_t(grpFam_[S->grp_index])];
        const auto &founded = famFounded_[size_t(grpFam_[S->grp_index])];
        const uint32_t child = hFamAct_.p[i] >> FAM_CHILD_SHIFT;
        if (child >= F.snapN + founded.size()) continue;
        Species *sp = child < F.snapN ? S->children[F.snapN - 1 - child] : founded[child - F.snapN];
        if (S->pending > 0) S->pending--;
        sp->count++;  sp->alive++;  sp->last = step_;  insertions_++;
        bulkContribution(sp, S, msid[i] == fsid[i] ? 0u : 1u, msid[i] == fsid[i] ? 1u : 0u);
      }
      for (const Family &F: fams_) {
        Species *S = grpSpecies_[F.g];
        for (Species *sp: S->children) { sp->full_upload = true; queueRefresh(sp); }
      }
      // (slots of species the device founded behind the failing birth stay unused)
    }
    for (uint32_t g = 0; g < G; g++)
      if (grpSpecies_[g]->div_epoch != chunkEpoch_) {
        grpSpecies_[g]->div_epoch = chunkEpoch_;  grpSpecies_[g]->full_upload = true;  queueRefresh(grpSpecies_[g]);
        settleStream(g, true, failed, nullptr);
      }
    done += failed;
    grpState_.clear();
    timers[2] += now() - t0;
    throw;
  }
  grpState_.clear();
  done += n;
  timers[2] += now() - t0;
}

void Engine::addGenomes (uint32_t n, const uint32_t *payload, const float *compat, bool onDevice,
                         cudaStream_t userStream, const uint32_t *gid, const uint32_t *mgid, const uint32_t *msid,
                         const uint32_t *fgid, const uint32_t *fsid, uint32_t *outSid, int32_t *outSlot,
                         uint32_t *outReplaced, uint32_t *nDone) {
  uint32_t done = 0;
  if (nDone) *nDone = 0;
  if (!n) return;
  if (pool_ && n >= parMin_) pool_->wake();
  cudaCheck(cudaSetDevice(p_.device), "cudaSetDevice");
  struct AllocTime {          // timers[7]: buffer growth inside this call (process-wide clock: one host thread per tree)
    double &out, a0;
    explicit AllocTime (double &o) : out(o), a0(allocSeconds()) {}
    ~AllocTime (void) { out += allocSeconds() - a0; }
  } allocTime (timers[7]);
  payloadPinned_ = false;
  if (!onDevice && size_t(n) * W_ * 4 >= stageMin_) {
    cudaPointerAttributes a;
    payloadPinned_ = cudaPointerGetAttributes(&a, payload) == cudaSuccess && a.type == cudaMemoryTypeHost;
    cudaGetLastError();
  }
  idsPinned_ = false;
  if (n >= devGroupMin_ && streamOk_) {
    cudaPointerAttributes a0, a1;
    idsPinned_ = cudaPointerGetAttributes(&a0, msid) == cudaSuccess && a0.type == cudaMemoryTypeHost
              && cudaPointerGetAttributes(&a1, fsid) == cudaSuccess && a1.type == cudaMemoryTypeHost;
    cudaGetLastError();
  }
  outPinned_ = false;
  if (outSid && outSlot && n >= 4096) {
    cudaPointerAttributes a0, a1;
    outPinned_ = cudaPointerGetAttributes(&a0, outSid) == cudaSuccess && a0.type == cudaMemoryTypeHost
              && cudaPointerGetAttributes(&a1, outSlot) == cudaSuccess && a1.type == cudaMemoryTypeHost;
    cudaGetLastError();        // an unregistered pointer is not an error here
  }

  // device-resident payloads stay where they are: the kernels read them in place and the host only
  // fetches the rows of the births it has to look at (runChunk)
  const uint32_t *dPayload = nullptr;
  if (onDevice) {
    dPayload = payload;  payload = nullptr;
    cudaCheck(cudaEventRecord(evUser_, userStream), "event record");
    cudaCheck(cudaStreamWaitEvent(stream_, evUser_, 0), "stream wait");
  }

  try {
    uint32_t i0 = 0;
    if (!root_) {
      // the very first genome founds species 0 (:270-273); Stats::insertions is not bumped there
      root_ = makeNode({});
      RepMeta g { gid[0], mgid[0], msid[0], fgid[0], fsid[0], 0, 0, classOf(compat[0], compat[1], compat[2]),
                  compat[0], compat[1], compat[2] };
      Eval none;
      uint32_t replaced;
      std::vector<uint32_t> first (W_);
      if (dPayload) cudaCheck(cudaMemcpy(first.data(), dPayload, W_ * 4, cudaMemcpyDeviceToHost), "first row");
      else memcpy(first.data(), payload, W_ * 4);
      const int32_t slot = insertInto(first.data(), g, root_, none, false, replaced);
      if (outSid) outSid[0] = root_->sid;
      if (outSlot) outSlot[0] = slot;
      if (outReplaced) outReplaced[0] = replaced;
      i0 = 1;  done = 1;  counters[0]++;
    }
    // Adaptive device passes.  Inside one pass the device sees the tree as it was at submission: births that
    // fall through to sub-species founded or modified earlier in the same pass are scored on the host.  When that
    // happens a lot (radiations, speciation stress) shorter passes are cheaper -- the next pass sees the new
    // species on the device -- and when it does not, the whole batch is one pass.  Any cut gives the same result.
    while (i0 < n) {
      const uint32_t m = std::min(adaptChunk_, n - i0);
      const uint64_t hostScores0 = counters[3];
      refreshPool();
      runChunk(m, payload ? payload + size_t(i0) * W_ : nullptr, compat + size_t(i0) * 3,
               dPayload ? dPayload + size_t(i0) * W_ : nullptr,
               gid + i0, mgid + i0, msid + i0, fgid + i0, fsid + i0,
               outSid ? outSid + i0 : nullptr, outSlot ? outSlot + i0 : nullptr,
               outReplaced ? outReplaced + i0 : nullptr, done);
      i0 += m;
      const uint64_t hostScores = counters[3] - hostScores0;
      if (hostScores > 8 * uint64_t(m) && m > CHUNK_MIN) adaptChunk_ = std::max(CHUNK_MIN, m / 16);
      else if (hostScores > m && m > CHUNK_MIN) adaptChunk_ = std::max(CHUNK_MIN, m / 2);
      else if (hostScores < m / 8 && m == adaptChunk_) adaptChunk_ = std::min(CHUNK_MAX, adaptChunk_ * 2);
    }
  } catch (...) {
    if (nDone) *nDone = done;
    throw;
  }
  if (nDone) *nDone = done;
}

/// Copy construction of the reference (phylogenetictree.hpp:98-109 with deepcopy :123-136), into an EMPTY tree: the
/// nodes reachable from `that`'s root are copied (data, enveloppe, distances, contributors; children in their order),
/// then eligibilities and parents are recomputed tree-wide (:103).  As in the reference, the set of alive species
/// and the statistics are NOT copied, and species that are in nobody's children list (same-step orphans) are lost.
void Engine::cloneFrom (const Engine &that) {
  if (root_ || nextNodeID_ || width_) throw ApiError(APT_ERR_STATE, "cloning needs an empty tree");
  if (that.K_ != K_ || that.W_ != W_ || that.p_.kind != p_.kind) throw ApiError(APT_ERR_ARGUMENT, "cloning between trees of different shapes");
  nextNodeID_ = that.nextNodeID_;  stillborns_ = that.stillborns_;  step_ = that.step_;
  ageOrdered_ = that.ageOrdered_;
  struct Rec {
    Engine &e;  const Engine &o;
    Species* node (const Species *src) {
      Species *n = new Species;
      n->sid = src->sid;
      n->slot = e.allocSlot();
      n->first = src->first;  n->last = src->last;  n->count = src->count;  n->alive = src->alive;  n->pending = src->pending;
      n->contribs = src->contribs;
      n->k = src->k;  n->reps = src->reps;  n->dist = src->dist;
      for (uint32_t r = 0; r < n->k; r++) {
        RepMeta &m = n->reps[r];
        m.cls = e.classOf(m.mu, m.sin, m.sout);
        memcpy(&e.hPayload_[(size_t(n->slot) * e.K_ + r) * e.wp_], o.repPayload(src, r), e.wp_ * sizeof(uint32_t));
      }
      if (e.nodes_.size() <= n->sid) e.nodes_.resize(std::max<size_t>(n->sid + 1, e.nodes_.size() * 2), nullptr);
      e.nodes_[n->sid] = n;
      e.width_++;
      n->full_upload = true;
      e.queueRefresh(n);
      for (const Species *c: src->children) {
        Species *cc = node(c);
        n->children.push_back(cc);
      }
      return n;
    }
  } rec { *this, that };
  // (`that.root_` may point at a trimmed node, as in the reference; it is copied all the same, :101)
  if (that.root_) root_ = rec.node(that.root_);
  eligStale_ = true;                      // (species outside the root's subtree are lost in a copy: flags naming them go)
  updateElligibilities();
}

// ---- JSON dump (layout of phylogenetictree.hpp:887-941) ---------------------
namespace {
void num (std::string &o, double v) {
  char buf[40];
  snprintf(buf, sizeof buf, "%.17g", v);
  o += buf;
}
void pid (std::string &o, uint32_t g, uint32_t s) {
  o += "{\"g\":" + std::to_string(g) + ",\"s\":" + std::to_string(s) + "}";
}
}

std::string Engine::toJson (void) const {
  std::string o = "{\"_step\":" + std::to_string(step_) + ",\"_envSize\":" + std::to_string(K_)
                + ",\"_stillborns\":" + std::to_string(stillborns_) + ",\"alive\":[";
  bool first = true;
  for (uint32_t s: aliveSpecies_) { if (!first) o += ","; first = false; o += std::to_string(s); }
  o += "],\"nextSID\":" + std::to_string(nextNodeID_) + ",\"tree\":";
  struct Rec {
    const Engine &e;
    void node (std::string &o, const Species &n) const {
      const uint32_t K = e.K_;
      o += "{\"id\":" + std::to_string(n.sid);
      o += ",\"data\":[" + std::to_string(n.first) + "," + std::to_string(n.last) + "," + std::to_string(n.count)
         + "," + std::to_string(n.alive) + "," + std::to_string(n.pending) + "]";
      o += ",\"envlp\":[";
      for (uint32_t r = 0; r < n.k; r++) {
        const RepMeta &m = n.reps[r];
        if (r) o += ",";
        o += "[{\"c\":[";  num(o, m.mu);  o += ",";  num(o, m.sin);  o += ",";  num(o, m.sout);
        o += "],\"g\":{\"f\":";  pid(o, m.fgid, m.fsid);
        o += ",\"m\":";  pid(o, m.mgid, m.msid);
        o += ",\"s\":";  pid(o, m.gid, INV);
        o += "},\"w\":[";
        const uint32_t *w = e.repPayload(&n, r);
        for (uint32_t j = 0; j < e.W_; j++) { if (j) o += ","; o += std::to_string(w[j]); }
        o += "]}," + std::to_string(m.udata) + "]";
      }
      o += "],\"contribs\":[";
      for (size_t i = 0; i < n.contribs.size(); i++) {
        const Contributor &c = n.contribs[i];
        if (i) o += ",";
        o += "[" + std::to_string(c.sid) + "," + std::to_string(c.count) + "," + (c.elligible ? "true" : "false") + "]";
      }
      o += "],\"dists\":";
      if (n.k < 2) o += "null";
      else {
        o += "[";
        bool first = true;
        for (uint32_t i = 0; i < n.k; i++)
          for (uint32_t j = i + 1; j < n.k; j++) {
            if (!first) o += ",";
            first = false;
            o += "[" + std::to_string(i) + "," + std::to_string(j) + ",";  num(o, n.dist[i * K + j]);  o += "]";
          }
        o += "]";
      }
      o += ",\"children\":";
      if (n.children.empty()) o += "null";
      else {
        o += "[";
        for (size_t i = 0; i < n.children.size(); i++) { if (i) o += ","; node(o, *n.children[i]); }
        o += "]";
      }
      o += "}";
    }
  } rec { *this };
  if (root_) rec.node(o, *root_); else o += "null";
  o += "}";
  return o;
}

} // namespace apt

// ---- JSON load (phylogenetictree.hpp:908-927, :946-967) ------------------------------------------------
namespace apt {
namespace {

/// Minimal JSON value + recursive-descent parser (objects, arrays, numbers, true/false/null, strings)
struct JVal {
  enum Type { NUL, BOOL, NUM, STR, ARR, OBJ } type = NUL;
  double num = 0;
  bool b = false;
  std::string str;
  std::vector<JVal> arr;
  std::vector<std::pair<std::string, JVal>> obj;
  const JVal& at (const char *key) const {
    for (const auto &kv: obj) if (kv.first == key) return kv.second;
    throw ApiError(APT_ERR_ARGUMENT, std::string("JSON: missing key '") + key + "'");
  }
  const JVal& idx (size_t i) const {
    if (type != ARR || i >= arr.size()) throw ApiError(APT_ERR_ARGUMENT, "JSON: bad array access");
    return arr[i];
  }
  uint32_t u32 (void) const {
    if (type != NUM) throw ApiError(APT_ERR_ARGUMENT, "JSON: number expected");
    return uint32_t(num);
  }
  size_t size (void) const { return type == ARR ? arr.size() : 0; }
};

struct JParser {
  const char *p, *end;
  [[noreturn]] void fail (const char *what) const { throw ApiError(APT_ERR_ARGUMENT, std::string("JSON: ") + what); }
  void ws (void) { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
  bool lit (const char *s) { const size_t n = strlen(s); if (size_t(end - p) >= n && !memcmp(p, s, n)) { p += n; return true; } return false; }
  std::string string (void) {
    if (p >= end || *p != '"') fail("string expected");
    p++;
    std::string s;
    while (p < end && *p != '"') { if (*p == '\\' && p + 1 < end) p++; s += *p++; }
    if (p >= end) fail("unterminated string");
    p++;
    return s;
  }
  JVal value (void) {
    ws();
    if (p >= end) fail("unexpected end");
    JVal v;
    if (*p == '{') {
      v.type = JVal::OBJ;  p++;  ws();
      if (p < end && *p == '}') { p++; return v; }
      for (;;) {
        ws();
        std::string k = string();
        ws();
        if (p >= end || *p != ':') fail("':' expected");
        p++;
        v.obj.emplace_back(std::move(k), value());
        ws();
        if (p < end && *p == ',') { p++; continue; }
        if (p < end && *p == '}') { p++; return v; }
        fail("',' or '}' expected");
      }
    }
    if (*p == '[') {
      v.type = JVal::ARR;  p++;  ws();
      if (p < end && *p == ']') { p++; return v; }
      for (;;) {
        v.arr.push_back(value());
        ws();
        if (p < end && *p == ',') { p++; continue; }
        if (p < end && *p == ']') { p++; return v; }
        fail("',' or ']' expected");
      }
    }
    if (*p == '"') { v.type = JVal::STR; v.str = string(); return v; }
    if (lit("null")) return v;
    if (lit("true")) { v.type = JVal::BOOL; v.b = true; return v; }
    if (lit("false")) { v.type = JVal::BOOL; v.b = false; return v; }
    char *e = nullptr;
    v.num = strtod(p, &e);
    if (e == p) fail("value expected");
    v.type = JVal::NUM;  p = e;
    return v;
  }
};

} // namespace

/// fromJson (:946-967): rebuilds the hierarchy (:908-927), then re-derives every node's parent from its
/// stored contributors in ascending species id (updateContributions(n, {}, true), :961-962) -- which,
/// as in the reference, re-attaches sub-species in THAT order.  The device pool is rebuilt from the
/// loaded enveloppes at the next insertion.  The tree must be empty.
void Engine::fromJson (const std::string &text) {
  if (root_ || nextNodeID_ || width_) throw ApiError(APT_ERR_STATE, "fromJson needs an empty tree");
  JParser ps { text.data(), text.data() + text.size() };
  const JVal j = ps.value();
  const uint32_t envSize = j.at("_envSize").u32();
  if (envSize != K_)
    throw ApiError(APT_ERR_ARGUMENT, "Current configuration file specifies an enveloppe size of " + std::to_string(K_)
                   + " whereas the provided PTree was built with " + std::to_string(envSize));     // :949-954
  // nothing of the input is trusted: ids are bounded by the dump's own nextSID (a later makeNode must not land on
  // an existing node) and by what a tree can hold, and a dump that fails half-way leaves the tree empty again
  const uint32_t nextSID = j.at("nextSID").u32();
  if (nextSID > (1u << 28)) throw ApiError(APT_ERR_ARGUMENT, "JSON: nextSID out of range");
  struct Rollback {
    Engine &e;  bool armed = true;
    ~Rollback (void) {
      if (!armed) return;
      for (Species *n: e.nodes_) delete n;
      e.nodes_.clear();  e.root_ = nullptr;  e.width_ = 0;  e.nextNodeID_ = 0;  e.step_ = 0;  e.stillborns_ = 0;
      e.ageOrdered_ = true;  e.eligStale_ = false;
      e.aliveSpecies_.clear();  e.refresh_.clear();  e.fillUpdates_.clear();  e.freeSlots_.clear();
      for (uint32_t s = e.nSlots_; s-- > 0; ) e.freeSlots_.push_back(s);       // every slot is free again
      std::fill(e.sidSlot_.begin(), e.sidSlot_.end(), INV);
      e.sidSlotDirty_ = true;
    }
  } rollback { *this };
  step_ = j.at("_step").u32();
  stillborns_ = j.at("_stillborns").u32();

  struct Rec {
    Engine &e;  const uint32_t nextSID;
    Species* node (const JVal &jn) {                                        // rebuildHierarchy, :908-927
      const uint32_t id = jn.at("id").u32();
      if (id >= nextSID) throw ApiError(APT_ERR_ARGUMENT, "JSON: species id not below nextSID");
      if (id < e.nodes_.size() && e.nodes_[id]) throw ApiError(APT_ERR_ARGUMENT, "JSON: duplicate species id");
      std::unique_ptr<Species> holder (new Species);
      Species *n = holder.get();
      n->sid = id;
      n->slot = e.allocSlot();
      n->reps.resize(e.K_);
      n->dist.assign(size_t(e.K_) * e.K_, 0.f);
      for (const JVal &c: jn.at("contribs").arr)
        n->contribs.push_back({ c.idx(0).u32(), c.idx(1).u32(), c.idx(2).type == JVal::BOOL ? c.idx(2).b : c.idx(2).num != 0 });
      const JVal &d = jn.at("data");
      n->first = d.idx(0).u32();  n->last = d.idx(1).u32();  n->count = d.idx(2).u32();
      n->alive = d.idx(3).u32();  n->pending = d.idx(4).u32();
      const JVal &env = jn.at("envlp");
      if (env.size() > e.K_) throw ApiError(APT_ERR_ARGUMENT, "JSON: enveloppe larger than rsetSize");
      for (size_t r = 0; r < env.size(); r++) {
        const JVal &g = env.idx(r).idx(0);
        RepMeta &m = n->reps[r];
        m.mu = float(g.at("c").idx(0).num);  m.sin = float(g.at("c").idx(1).num);  m.sout = float(g.at("c").idx(2).num);
        m.cls = e.classOf(m.mu, m.sin, m.sout);
        const JVal &ge = g.at("g");
        m.fgid = ge.at("f").at("g").u32();  m.fsid = ge.at("f").at("s").u32();
        m.mgid = ge.at("m").at("g").u32();  m.msid = ge.at("m").at("s").u32();
        m.gid = ge.at("s").at("g").u32();
        m.timestamp = 0;                      // not serialised by the reference either (node.hpp:54-63)
        m.udata = env.idx(r).idx(1).type == JVal::NUM ? env.idx(r).idx(1).u32() : m.gid;
        const JVal &w = g.at("w");
        if (w.size() != e.W_) throw ApiError(APT_ERR_ARGUMENT, "JSON: payload width does not match the tree");
        uint32_t *dst = &e.hPayload_[(size_t(n->slot) * e.K_ + r) * e.wp_];
        for (uint32_t i = 0; i < e.W_; i++) dst[i] = w.idx(i).u32();
      }
      n->k = uint32_t(env.size());
      for (const JVal &dd: jn.at("dists").arr) {
        const uint32_t a = dd.idx(0).u32(), b = dd.idx(1).u32();
        if (a >= e.K_ || b >= e.K_) throw ApiError(APT_ERR_ARGUMENT, "JSON: distance index out of range");
        n->dist[a * e.K_ + b] = n->dist[b * e.K_ + a] = float(dd.idx(2).num);
      }
      if (e.nodes_.size() <= n->sid) e.nodes_.resize(std::max<size_t>(n->sid + 1, e.nodes_.size() * 2), nullptr);
      e.nodes_[n->sid] = holder.release();
      e.width_++;
      n->full_upload = true;
      e.queueRefresh(n);
      for (const JVal &c: jn.at("children").arr) node(c);
      return n;
    }
  } rec { *this, nextSID };
  const JVal &jt = j.at("tree");
  if (jt.type == JVal::OBJ) root_ = rec.node(jt);
  aliveSpecies_.clear();
  for (const JVal &a: j.at("alive").arr) aliveSpecies_.insert(a.u32());
  nextNodeID_ = nextSID;

  // Ensure correct parenting (:960-962): contributors.update({}) = stable_sort + currentMain on stored flags
  for (Species *n: nodes_) {
    if (!n) continue;
    std::stable_sort(n->contribs.begin(), n->contribs.end(),
                     [] (const Contributor &l, const Contributor &r) { return l.count > r.count; });
    Species *newMC = updateParent(n, currentMain(n));
    if (newMC) newMC->children.push_back(n);
  }
  // a dump written by hand may flag a younger (or missing) contributor as eligible: elligibile() then keeps its subtree walk
  for (const Species *n: nodes_) {
    if (!n) continue;
    for (const Contributor &c: n->contribs)
      if (c.elligible && (c.sid >= nodes_.size() || !nodes_[c.sid] || nodes_[c.sid]->first >= n->first)) ageOrdered_ = false;
  }
  eligStale_ = true;                      // flags as stored: the next eligibility pass recomputes them all
  rollback.armed = false;
}

} // namespace apt
