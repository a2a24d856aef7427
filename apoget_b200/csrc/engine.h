// Tree engine of libapt_b200.so: PhylogeneticTree<GENOME,UDATA> for device-encoded genomes.
//
// Split of labour ("Ordered resolution"):
//   device  K1 scores every birth of a batch against the enveloppes of its parents' species, and
//           K2 against those species' sub-species, from the pool as it stood when the batch was
//           submitted (speculation);
//   host    one ordered walk commits births in arrival order.  A speculative (birth, species)
//           result is used as is when that enveloppe has not changed since the pool was refreshed,
//           patched (only the changed representatives are re-measured) when it has, and computed
//           from scratch for species founded inside the batch -- so every birth sees exactly the
//           state the reference's sequential addGenome would have shown it.
// The hierarchy bookkeeping (contributors, parents, trimming) follows the reference line by line;
// citations are relative to /root/reference/src/core/tree.
#ifndef APT_B200_ENGINE_H
#define APT_B200_ENGINE_H

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "apt_b200.h"
#include "kernels.cuh"

namespace apt {

struct ApiError : std::runtime_error {
  int status;
  ApiError (int status, const std::string &what) : std::runtime_error(what), status(status) {}
};

void cudaCheck (cudaError_t e, const char *what);
/// APT_PROFILE: every (re)allocation of a grow-only buffer is reported on stderr with the time it took
void noteAlloc (const char *kind, size_t bytes, double t0);
double allocClock (void);
/// seconds this process spent growing those buffers (cudaMalloc / cudaMallocHost and the frees that go with them)
double allocSeconds (void);

/// Grow-only device buffer
template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;
  ~DevBuf (void) { if (p) cudaFree(p); }
  void reserve (size_t n, bool keep = false, cudaStream_t st = nullptr) {
    if (n <= cap) return;
    // few, generous steps: every growth is a cudaMalloc + cudaFree (a device-wide synchronisation each; measured at 5-30 ms
    // apiece inside a speciating run, 300 of them over seven generations when buffers doubled from 1 024 elements)
    size_t ncap = cap ? cap : (size_t(1) << 16);
    while (ncap < n) ncap *= 4;
    T *q = nullptr;
    const double t0 = allocClock();
    cudaCheck(cudaMalloc(&q, ncap * sizeof(T)), "cudaMalloc");
    if (p) {
      if (keep && cap) {
        cudaCheck(cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st), "pool growth copy");
        cudaCheck(cudaStreamSynchronize(st), "pool growth sync");
      }
      cudaFree(p);
    }
    p = q;  cap = ncap;
    noteAlloc("device", ncap * sizeof(T), t0);
  }
};

/// Grow-only pinned host buffer
template <typename T>
struct PinBuf {
  T *p = nullptr;
  size_t cap = 0;
  ~PinBuf (void) { if (p) cudaFreeHost(p); }
  void reserve (size_t n) {
    if (n <= cap) return;
    // at least 25 % of headroom: page-locking is slow (1.5 ms per call in a process that sees several GPUs) and
    // result sizes creep up from step to step
    size_t ncap = cap ? cap : (size_t(1) << 14);
    while (ncap < n + n / 4) ncap *= 4;
    T *q = nullptr;
    const double t0 = allocClock();
    cudaCheck(cudaMallocHost(&q, ncap * sizeof(T)), "cudaMallocHost");
    if (p) cudaFreeHost(p);
    p = q;  cap = ncap;
    noteAlloc("pinned", ncap * sizeof(T), t0);
  }
};

/// Host mirror of the pool's genome rows.  Grows by realloc: the C library remaps large blocks instead of copying them (a
/// std::vector moves the old rows into a new block every time 10 000 species are founded in one pass).  New rows are
/// zeroed in one sweep -- leaving them untouched was measured slower: the first touches then happen scattered in the walk.
struct HostRows {
  uint32_t *p = nullptr;
  size_t n = 0, cap = 0;
  HostRows (void) = default;
  HostRows (const HostRows&) = delete;
  HostRows& operator= (const HostRows&) = delete;
  ~HostRows (void) { free(p); }
  void resize (size_t m, bool zeroNew) {
    if (m > cap) {
      const size_t ncap = std::max<size_t>(std::max<size_t>(m, cap + cap / 2), size_t(1) << 16);
      void *q = realloc(p, ncap * sizeof(uint32_t));
      if (!q) throw std::bad_alloc();
      p = static_cast<uint32_t*>(q);
      if (zeroNew) memset(p + cap, 0, (ncap - cap) * sizeof(uint32_t));
      cap = ncap;
    }
    n = m;
  }
  size_t size (void) const { return n; }
  uint32_t& operator[] (size_t i) { return p[i]; }
  const uint32_t& operator[] (size_t i) const { return p[i]; }
};

/// Fixed set of worker threads for the O(n) host loops of large chunks (grouping births into
/// streams).  run(fn) executes fn(t) for t in [0, size) -- t = 0 on the caller -- and returns when
/// all are done.  No OpenMP on purpose: the library is loaded next to other runtimes.
class WorkerPool {
public:
  explicit WorkerPool (unsigned n);
  ~WorkerPool (void);
  unsigned size (void) const { return n_; }
  void run (const std::function<void(unsigned)> &fn);
  void wake (void);
private:
  void loop (unsigned t);
  unsigned n_;
  std::vector<std::thread> threads_;
  std::mutex m_;
  std::condition_variable wake_, done_;
  const std::function<void(unsigned)> *fn_ = nullptr;
  std::atomic<uint64_t> gen_ { 0 };
  std::atomic<unsigned> pending_ { 0 };
  std::atomic<bool> workerFailed_ { false };
  bool stop_ = false;
  uint64_t hint_ = 0;
};

// speciescontributors.h:16-86 / :89-100
struct Contributor { uint32_t sid, count; bool elligible; };
struct Contribution { uint32_t species, count; };
/// What one birth contributes: its mother's and (when different) its father's species -- never more than two entries,
/// kept inline (the reference's std::vector costs a heap round trip per birth on the commit path)
struct Contributions {
  Contribution v[2];
  uint32_t n = 0;
  Contributions (void) = default;
  Contributions (std::initializer_list<Contribution> l) { for (const Contribution &c: l) push_back(c); }
  void push_back (const Contribution &c) {
    if (n >= 2) throw ApiError(APT_ERR_STATE, "more than two contributions for one birth");
    v[n++] = c;
  }
  size_t size (void) const { return n; }
  Contribution& operator[] (size_t i) { return v[i]; }
  const Contribution& operator[] (size_t i) const { return v[i]; }
  const Contribution* begin (void) const { return v; }
  const Contribution* end (void) const { return v + n; }
};

/// Everything the tree keeps about a representative besides its payload (node.hpp:25-85)
struct RepMeta {
  uint32_t gid, mgid, msid, fgid, fsid;   // genealogy of the genome in this slot
  uint32_t timestamp;                     // insertion step, node.hpp:26
  uint32_t udata;                         // gid the slot's user-data object was last built from (:708-710)
  uint32_t cls;                           // compatibility class (index of its (mu, s_in, s_out) triple)
  float mu, sin, sout;
};

// node.hpp:16-243 + speciesdata.hpp:15-31
struct Species {
  uint32_t sid = 0, slot = 0;
  Species *parent = nullptr;
  std::vector<Species*> children;
  uint32_t first = 0, last = 0, count = 0, alive = 0, pending = 0;
  std::vector<Contributor> contribs;
  uint32_t k = 0;                         // rset.size()
  std::vector<RepMeta> reps;              // K entries, k used
  std::vector<float> dist;                // K x K, symmetric, valid for i != j < k (DistanceMap)
  // relation between this host state and the copy in the device pool
  uint32_t dev_fill = 0;                  // representatives the device copy holds
  uint32_t stale_mask = 0;                // slots whose device copy is out of date
  bool queued = false;                    // already in the refresh list
  bool trimmed = false;                   // removed by stillborn trimming
  uint32_t cand_epoch = 0, cand_rank = 0; // position in the current group's candidate list (K2)
  // per-chunk state (valid when the epoch field equals the engine's chunk epoch)
  uint32_t grp_epoch = 0, grp_index = 0;  // this species has a device-resolved stream in the current chunk
  uint32_t div_epoch = 0;                 // ... and the host took that stream over (device results discarded)
  uint32_t snap_epoch = 0, snap_mask = 0, snap_fill = 0;  // slots changed since the chunk's device snapshot
  uint32_t self_idx = 0;                  // where this species' own entry sits in contribs (hint)
  uint32_t child_change_epoch = 0;        // chunk in which a sub-species was moved into / out of this node
  bool full_upload = false;               // the device copy (rows, matrix, fill) must be rewritten entirely
  uint32_t touch_epoch = 0;               // last scoring launch that read this enveloppe
  uint32_t pair_epoch = 0;                // chunk in which a birth outside the streams names this species as a parent
  uint32_t fam_epoch = 0;                 // chunk in which the family pass (K3) changed this enveloppe
  bool extinct (void) const { return alive == 0 && pending == 0; }   // node.hpp:157-159
};

struct Eval {
  float score;
  bool exact;                             // computed with reference-identical arithmetic
  float d[KMAX];
};

class Engine {
public:
  explicit Engine (const apt_params &p);
  ~Engine (void);

  void addGenomes (uint32_t n, const uint32_t *payload, const float *compat, bool onDevice, cudaStream_t userStream,
                   const uint32_t *gid, const uint32_t *mgid, const uint32_t *msid,
                   const uint32_t *fgid, const uint32_t *fsid,
                   uint32_t *outSid, int32_t *outSlot, uint32_t *outReplaced, uint32_t *nDone);
  /// fn(ctx, lo, hi) over disjoint slices of [0, n) on the engine's worker threads (the caller's thread included);
  /// returns when all slices are done.  For the binding's own O(n) loops (encoding genomes, filling results).
  void parallelFor (uint32_t n, void (*fn)(void *ctx, uint32_t lo, uint32_t hi), void *ctx);
  void delGenome (uint32_t sid);
  void candidacy (uint32_t msid, uint32_t fsid, int dir);
  void setStep (uint32_t s) { step_ = s; }
  void doStep (uint32_t s, const uint32_t *sids, uint32_t n);

  void stats (uint32_t out[8]) const;
  void resetStats (void) { insertions_ = deletions_ = comparisons_ = branching_ = 0; }
  std::vector<apt_event>& events (void) { return events_; }
  std::string toJson (void) const;
  void fromJson (const std::string &text);
  void cloneFrom (const Engine &that);
  std::vector<uint32_t>& trimmedSpecies (void) { return trimmedSids_; }
  uint32_t stillborns (void) const { return stillborns_; }
  Species* nodeAt (uint32_t sid) const;
  Species* root (void) const { return root_; }
  const std::set<uint32_t>& aliveSpecies (void) const { return aliveSpecies_; }
  const apt_params& params (void) const { return p_; }
  uint64_t counters[16] = {};
  double timers[8] = {};
  double streamSpanMs_ = 0.;  uint64_t streamSpanN_ = 0;     // stream kernel: first CTA in to last CTA out, summed over its launches

private:
  apt_params p_;
  uint32_t K_, W_, wp_;
  bool lutMode_;
  uint32_t lutStride_;
  uint32_t devListMin_ = 4096;               // passes of at least this many births get their list of unsettled births from the device (APT_DEVLIST_MIN)
  size_t lutMaxBytes_ = size_t(256) << 20;   // largest compatibility table kept (APT_LUT_MAX_BYTES); beyond it: generic mode

  // ---- tree state (phylogenetictree.hpp:370-389) ----
  uint32_t nextNodeID_ = 0, stillborns_ = 0, step_ = 0;
  Species *root_ = nullptr;
  std::vector<Species*> nodes_;           // indexed by SID; null = never existed or trimmed
  uint32_t width_ = 0;
  std::vector<Species*> graveyard_;
  std::set<uint32_t> aliveSpecies_;
  uint32_t insertions_ = 0, deletions_ = 0, comparisons_ = 0, branching_ = 0;
  std::vector<apt_event> events_;
  std::vector<uint32_t> trimmedSids_;     // species removed by stillborn trimming since the binding last asked

  // ---- host mirror of the pool ----
  HostRows hPayload_;                     // [slot][K][wp]
  std::vector<uint32_t> freeSlots_;
  uint32_t nSlots_ = 0;
  std::vector<Species*> refresh_;         // species with stale device rows
  std::vector<uint32_t> fillUpdates_;     // (slot, fill) pairs for the device `fill` array

  // ---- compatibility classes ----
  struct Triple { float mu, sin, sout; bool operator< (const Triple &o) const {
    return mu != o.mu ? mu < o.mu : sin != o.sin ? sin < o.sin : sout < o.sout; } };
  std::map<Triple, uint32_t> classIds_;
  std::vector<Triple> classes_;
  std::vector<double> hLut_;              // [class][lutStride_], host libm values
  uint32_t lutUploaded_ = 0;              // classes already in the device table
  Triple lastTriple_ { 0, 0, 0 };  uint32_t lastClass_ = 0xFFFFFFFFu;

  // ---- device state ----
  cudaStream_t stream_ = nullptr;
  cudaEvent_t evUser_ = nullptr, evK0_ = nullptr, evK1_ = nullptr;
  uint32_t touchEpoch_ = 0;
  DevBuf<uint32_t> dPool_;  DevBuf<float4> dMeta_;  DevBuf<uint32_t> dFill_;  DevBuf<double> dLut_;
  DevBuf<uint32_t> dBirths_, dRaw_;  DevBuf<float4> dBMeta_;  DevBuf<float> dRawCompat_;  DevBuf<uint32_t> dCls_;
  DevBuf<uint32_t> dPairBirth_, dPairSlot_;  DevBuf<float> dScore_, dDist_;
  DevBuf<uint32_t> dStageRows_, dStageDst_;  DevBuf<float4> dStageMeta_;
  DevBuf<float> dPoolDist_, dStageDist_, dRows_;  DevBuf<uint32_t> dOrder_, dGrpAll_, dAct_, dCtr_, dOut2_;
  PinBuf<float> hStageDist_, hRows_;  PinBuf<uint32_t> hOrder_, hGrpSlot_, hGrpOff_, hAct_, hGrpAll_, hCtr_;
  std::vector<int32_t> birthGroup_;       // stream index of each birth of the chunk, -1 = host/K1 path
  std::vector<Species*> grpSpecies_;
  std::vector<uint32_t> grpCount_;
  /// Host view of one device-resolved stream while the chunk is being committed
  struct GroupState {
    uint32_t cursor = 0;                  // position in the stream's order list reached by catch-up scans
    uint32_t scanned[3] = {0, 0, 0};      // accepted-unchanged births seen by those scans, per contribution class
    uint32_t selfApplied = 0;             // self-contribution already credited to the species' own entry
    uint32_t snapChildren = 0;            // sub-species the stream's species had when the chunk was submitted
    int64_t lastEnter = -1;               // last birth of the chunk the device let into this enveloppe
  };
  std::vector<GroupState> grpState_;
  // parallel grouping
  std::unique_ptr<WorkerPool> pool_;
  uint32_t parMin_ = 32768;               // chunks at least this large are grouped by the worker pool
  std::vector<uint32_t> sidSlot_;         // sid -> pool slot when the species can be streamed (device copy current), else APT_INVALID
  std::vector<std::vector<uint32_t>> parRest_, parCnt_, parOrd_, parLoc_;   // per worker: births left to the host path, offsets per species, private order buffer and its cursors
  std::vector<uint32_t> sidTotal_;        // per species: births of the chunk, then the start of its stream in the order list

  std::vector<uint32_t> birthKey_;
  void groupSerial (uint32_t n, const uint32_t *msid, const uint32_t *fsid, size_t &P, size_t &nStream, uint32_t &maxStream);
  void groupParallel (uint32_t n, const uint32_t *msid, const uint32_t *fsid, size_t &P, size_t &nStream, uint32_t &maxStream);
  void setStreamable (const Species *s, bool on);
  std::vector<int32_t> sidGroup_;         // per-chunk cache: sid -> stream index, -1 not streamable, -2 not looked up yet
  std::vector<uint32_t> sidTouched_;
  DevBuf<uint4> dGrpSum_;
  PinBuf<uint32_t> hGrpSid_, hOutSid_;  PinBuf<int32_t> hOutSlot_;  PinBuf<uint4> hGrpSum_;
  std::vector<uint32_t> slow_, requeue_;  // births the host has to look at (ascending) / handed back by a taken-over stream
  std::vector<int32_t> slowGrp_;          // stream index of slow_[q] (-1: none) and ...
  std::vector<uint32_t> slowAct_;         // ... its result word, kept next to it (sequential passes instead of gathers)
  std::vector<apt_event> runEv_;  std::vector<uint8_t> runEvN_;   // callbacks of a parallel run, two places per entry
  std::vector<uint8_t> grpOwner_;         // worker that mirrors the entries of stream g (pool sizes stay far below 256)
  std::vector<std::vector<uint32_t>> parMine_;   // per worker: positions in slow_ of the entries it mirrors
  uint32_t curBirth_ = 0;                 // index (in the chunk) of the birth being committed
  const uint32_t *curDevPayload_ = nullptr;
  DevBuf<uint32_t> dSlowIdx_, dSlowRows_, dRowPayload_;  PinBuf<uint32_t> hSlowIdx_, hSlowRows_, hRowPayload_;
  std::vector<uint32_t> gatherIdx_;     // births of the chunk whose genomes were gathered into hSlowRows_ (ascending)
  std::vector<std::vector<uint32_t>> lateRows_;
  // sub-species kernel (K2)
  bool childKernel_ = true;
  DevBuf<uint32_t> dChildOff_, dChildSlot_, dFtList_, dFtFirst_, dFtFlags_;  DevBuf<float> dFtScore_, dFtDist_;
  PinBuf<uint32_t> hChildOff_, hChildSlot_, hFtList_, hFtFirst_, hFtFlags_;  PinBuf<float> hFtScore_, hFtDist_;
  std::vector<int32_t> ftOfSlow_;         // entry of the sub-species results for slow_[i], -1 = none
  uint32_t nFt_ = 0;
  uint32_t nFtRows_ = 0;                 // device-resident births: genomes of the rejected births fetched with the rows
  std::vector<int32_t> ftRowOfSlow_;
  int64_t curFtRow_ = -1;
  // grouping of a pass's births on the device (launch_group_births): for callers whose id arrays are page-locked
  DevBuf<uint32_t> dMsid_, dFsid_, dSidSlot_, dGrpKey_, dSidTotal_[2], dSidStart_, dSidCursor_, dSidGidx_, dGrpInfo_, dRestList_;
  uint32_t grpFlip_ = 0;                     // which of the two counter sets the next grouping uses (its scan clears the other one)
  size_t grpZeroed_[2] = { 0, 0 };           // leading entries of each set known to be zero
  const uint32_t *grpZeroedPtr_[2] = { nullptr, nullptr };
  bool grpInfoReady_ = false;
  PinBuf<uint32_t> hSidTotal_, hGrpInfo_, hRestList_;
  PinBuf<int32_t> hBirthGroup_;
  cudaEvent_t evG_ = nullptr;
  cudaStream_t streamG_ = nullptr;          // the grouping kernel runs beside the genomes' upload
  bool sidSlotDirty_ = true;                 // sidSlot_ changed since it was last uploaded
  size_t dSidSlotN_ = 0;
  uint32_t devGroupMin_ = 4096;              // passes of at least this many births are grouped on the device (APT_DEVGROUP_MIN)
  bool devGroupPageable_ = false;            // (tests) also when the id arrays are not page-locked (APT_DEVGROUP_PAGEABLE)
  bool idsPinned_ = false, devGrouped_ = false;
  uint32_t estRest_ = 0;
  const int32_t *bg_ = nullptr;              // the pass's birth -> stream table: birthGroup_ (host grouping) or hBirthGroup_ (device)
  bool groupDeviceLaunch (uint32_t n, const uint32_t *msid, const uint32_t *fsid);
  bool groupDeviceFinish (uint32_t n, const uint32_t *msid, const uint32_t *fsid, size_t &P, size_t &nStream, uint32_t &maxStream);
  double grpProf_[2] = { 0., 0. };
  // the births of a pass the host has to look at, listed in arrival order by the device (launch_list_unsettled)
  DevBuf<uint32_t> dSlowBlk_, dSlowList_, dSlowCtr_;  PinBuf<uint32_t> hSlowList_;
  uint32_t estSlow_ = 0;
  uint32_t estRows_ = 0, estFt_ = 0;     // enveloppe entries / rejected births of the last device pass (+25 %): fetched speculatively with the results
  DevBuf<uint32_t> dFtRows_;  PinBuf<uint32_t> hFtRows_;
  int64_t curFt_ = -1;                    // sub-species result entry of the birth being committed
  // sub-species resolution on the device (K3, family.cu): the rejected births of a stream against its species'
  // sub-species, in order, founding included
  bool famOk_ = false;                    // the family kernel can serve this tree's configuration
  uint32_t famCap_ = 0;                   // sub-species (at submission + founded) one family may hold
  struct Family { uint32_t g, snapN, joinOff; };    // stream index, sub-species at submission, offset of its counters
  std::vector<Family> fams_;
  std::vector<std::vector<Species*>> famFounded_;   // per family: the species it founded in this pass, in founding order
  std::vector<int32_t> grpFam_;           // stream -> family, -1 = none
  uint32_t famMaxNew_ = 0, famSlotBase_ = 0, famSlotsUsed_ = 0;
  bool famResults_ = false;               // the pass produced family results (fetched)
  uint32_t forcedSlot_ = 0xFFFFFFFFu;     // pool slot the next makeNode takes (a species the device founded lives there already)
  DevBuf<uint32_t> dRejOrder_, dRejFill_, dRejCount_, dFamTab_, dChildSid_, dFamAct_, dFamRow_, dChildJoin_, dNewFounder_, dFamSlotCtr_;
  DevBuf<uint4> dFamInfo_;  DevBuf<unsigned long long> dFamStats_;
  PinBuf<uint32_t> hFamTab_, hChildSid_, hFamAct_, hFamRow_, hChildJoin_, hNewFounder_;
  PinBuf<uint4> hFamInfo_;  PinBuf<unsigned long long> hFamStats_;
  void walkFamily (uint32_t i, const uint32_t *w, RepMeta &g, Species *s0, uint32_t fact, uint32_t &sid, int32_t &slot, uint32_t &replaced);
  void settleFamilies (void);
  /// pool slots kept ready behind the ones in use for species the family pass founds in one device pass
  static uint32_t famReserve (size_t nStream) { return uint32_t(std::min<size_t>(nStream / 4 + 1024, size_t(1) << 16)); }
  void bulkContribution (Species *c, Species *from, uint32_t n1, uint32_t n2);
  uint32_t chunkEpoch_ = 0;
  bool chunkUniform_ = false;             // every birth of the chunk carries the same compatibility triple
  const float4 *fillPtr_ = nullptr;  uint32_t fillN_ = 0;  float4 fillValue_ {};   // what the births' device meta array was last filled with (uniform chunks)
  uint32_t chunkCls_ = 0;
  bool streamOk_ = false;                 // the stream kernel can serve this tree's configuration
  cudaEvent_t evS0_ = nullptr, evS1_ = nullptr, evF0_ = nullptr, evF1_ = nullptr;
  static constexpr uint32_t CHUNK_MIN = 1024, CHUNK_MAX = 1u << 20;
  uint32_t adaptChunk_ = 4096;          // births per device pass, adapted to how much of the last pass fell back to the host
  bool ageOrdered_ = true;                // every contributor flagged eligible is older than its species (elligibile)
  bool eligStale_ = false;                // a species was removed, or flags came from a dump, since the last tree-wide eligibility pass
  const Species *freshNode_ = nullptr;    // the node makeNode is linking into the tree (elligibile: it has no subtree)
  double profWalk_[3] = {0, 0, 0};  uint64_t profWalkN_[3] = {0, 0, 0};   // APT_PROFILE: serial commit by category
  bool profile_ = false;                // APT_PROFILE: per-phase host timings on stderr
  std::string streamProf_;              // APT_STREAM_PROF: file the per-stream kernel timelines are appended to
  bool outPinned_ = false;              // the caller's result arrays of the current call are pinned host memory
  bool payloadPinned_ = false;          // ... and so are its payload rows (else large batches are staged by the workers)
  size_t stageMin_ = size_t(1) << 20;   // payload bytes from which pageable rows are staged by the worker pool (APT_STAGE_MIN)
  static constexpr size_t STAGE_SLICE = size_t(256) << 10;
  DevBuf<unsigned long long> dProf_;    // per-stream timelines, only with APT_STREAM_PROF=<file>
  PinBuf<uint32_t> hBirths_, hPairBirth_, hPairSlot_, hStageRows_, hStageDst_, hCls_;
  PinBuf<float4> hBMeta_, hStageMeta_;
  PinBuf<float> hScore_, hDist_;

  // ---- per-chunk scratch ----
  struct PairRef { uint32_t birth; int64_t p0, p1; };   // speculative pair indices of a two-parent birth
  std::vector<PairRef> pairOf_;           // ascending by birth
  const uint32_t *curPayload_ = nullptr;  // host payload of the chunk being walked (W_ words per row)

  // hierarchy (line-by-line restatements, see engine.cpp)
  bool elligibile (uint32_t lhs, uint32_t rhs) const;
  uint32_t currentMain (Species *n) const;
  uint32_t contributorsUpdate (Species *n, const Contributions &ctbs);
  Species* updateParent (Species *n, uint32_t sid);
  Species* nodeUpdate (Species *n, const Contributions &c);
  void updateElligibilities (void);
  Species* makeNode (const Contributions &c);
  void updateContributions (Species *s, const Contributions &c);
  void performStillbornTrimming (void);
  void emit (uint32_t t, uint32_t a, uint32_t b, uint32_t c) {
    if (!(p_.flags & APT_FLAG_NO_EVENTS)) events_.push_back({t, a, b, c});
  }

  // scoring on the host (exact) and resolution
  uint32_t classOf (float mu, float sin, float sout);
  double compatOf (uint32_t cls, float mu, float sin, float sout, double d) const;
  double distanceExact (const uint32_t *a, const uint32_t *b) const;
  const uint32_t* repPayload (const Species *s, uint32_t r) const { return &hPayload_[(size_t(s->slot) * K_ + r) * wp_]; }
  float scoreFrom (const double *c, uint32_t k) const;
  void evalExact (const uint32_t *w, const RepMeta &g, Species *s, Eval &e);
  void evalSpeculated (const uint32_t *w, const RepMeta &g, Species *s, int64_t pair, Eval &e);
  bool trusted (float score) const;
  int32_t insertInto (const uint32_t *w, const RepMeta &g, Species *s, const Eval &e, bool haveCache, uint32_t &replaced);
  void walkOne (const uint32_t *w, RepMeta &g, int64_t p0, int64_t p1, uint32_t &sid, int32_t &slot, uint32_t &replaced);
  void markStale (Species *s, uint32_t r);
  void touchSnap (Species *s);
  void diverge (Species *s);
  void queueRefresh (Species *s) { if (!s->queued) { s->queued = true; refresh_.push_back(s); } }
  void walkStream (const uint32_t *w, RepMeta &g, Species *s0, uint32_t act, uint32_t &sid, int32_t &slot, uint32_t &replaced);
  void resolveBelow (const uint32_t *w, const RepMeta &g, Species *s0, Species *s1, Species *&best, float &bestScore, Eval &bestEval);
  void resolveBelowStream (const uint32_t *w, const RepMeta &g, Species *s0, int64_t ft, Species *&best, float &bestScore, Eval &bestEval);
  void finishBirth (const uint32_t *w, RepMeta &g, Species *best, float bestScore, const Eval &bestEval,
                    const Contributions &contrib, uint32_t &sid, int32_t &slot, uint32_t &replaced);
  void selfContribution (Species *s, uint32_t count);
  int mirrorEntry (const uint32_t *w, const RepMeta &g, Species *s0, uint32_t act, int32_t &slot, uint32_t &replaced, apt_event ev[2]);
  void catchUpSelf (Species *s, uint32_t before);
  void settleStream (uint32_t g, bool diverging, uint32_t before, std::vector<uint32_t> *requeue);

  // device plumbing
  uint32_t allocSlot (void);
  void refreshPool (void);
  void uploadLut (void);
  PoolView poolView (void) const;
  ScoreCfg scoreCfg (void) const;
  const uint32_t* payloadOf (uint32_t i);
  const uint32_t* entryRowOf (uint32_t i) const;
  void runChunk (uint32_t n, const uint32_t *payload, const float *compat, const uint32_t *dPayload,
                 const uint32_t *gid, const uint32_t *mgid, const uint32_t *msid,
                 const uint32_t *fgid, const uint32_t *fsid, uint32_t *outSid, int32_t *outSlot,
                 uint32_t *outReplaced, uint32_t &done);
};

} // namespace apt

#endif
