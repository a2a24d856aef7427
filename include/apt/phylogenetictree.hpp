// apt/phylogenetictree.hpp -- drop-in for APOGeT's phylogeny::PhylogeneticTree<GENOME,UDATA>
// (reference: src/core/tree/phylogenetictree.hpp) whose species-assignment hot path runs on a
// B200 through libapt_b200.so (C ABI: include/apt_b200.h).
//
// Same class name, member names, argument meaning and callback order as the reference; the types
// below mirror src/core/tree/treetypes.h (GID, SID, PID, Genealogy, InsertionResult),
// callbacks.hpp (Callbacks_t) and ptreeconfig.h (config::PTree).  A host program switches by
// including this header, using namespace `apt` instead of `phylogeny`, and describing how its
// genome is laid out on the device:
//
//   template <> struct apt::DeviceEncoding<MyGenome> {
//     static constexpr apt_kind kind = APT_BITS;      // or APT_F32
//     static constexpr uint32_t words = 8;            // 32-bit payload words
//     static void encode (const MyGenome &g, uint32_t *payload, float compat[3]);
//   };
//
// where payload/compat must be consistent with the genome's own distance() / compatibility()
// (Hamming or L2 over the payload; the two-sided gaussian of src/core/crossover.h:76-92 with
// compat = {optimalDistance, inbreedTolerance, outbreedTolerance}).  New with respect to the
// reference: addGenomes(first, last, out), whose result is defined as that of calling addGenome on
// every element in order.  Not provided ("out of scope"): copy construction /
// assignment (:98-118; trees are movable) and logTo (:874-879).
// toJson / saveTo / fromJson / readFrom need nlohmann::json: define APT_WITH_JSON before including.
#ifndef APT_PHYLOGENETICTREE_HPP
#define APT_PHYLOGENETICTREE_HPP

#include <cstdint>
#include <deque>
#ifdef APT_DROPIN_PROFILE
#include <chrono>
#include <cstdio>
#endif
#include <limits>
#include <map>
#include <memory>
#include <new>
#include <ostream>
#include <set>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <iterator>
#include <vector>

#include "../apt_b200.h"

#ifdef APT_WITH_JSON
#include <fstream>
#include <nlohmann/json.hpp>
#endif

namespace apt {

using uint = unsigned int;

// ---- treetypes.h ----------------------------------------------------------------------------
enum class GID : uint { INVALID = uint(-1) };                        // treetypes.h:24-26
enum class SID : uint { INVALID = uint(-1) };                        // treetypes.h:74-76
inline std::ostream& operator<< (std::ostream &os, GID g) { return os << uint(g); }
inline std::ostream& operator<< (std::ostream &os, SID s) { return os << uint(s); }
using LivingSet = std::set<SID>;                                     // treetypes.h:82

/// treetypes.h:32-71
class GIDManager {
  uint next;
public:
  GIDManager (void) { setNext(GID(0)); }
  void setNext (GID n) { next = uint(n) + 1; }
  GID operator() (void) {
    if (next == std::numeric_limits<uint>::max())
      throw std::out_of_range("Exhausted all possible identifers for the underlying type");
    return GID(next++);
  }
  explicit operator GID (void) const { return GID(next); }
};

/// treetypes.h:85-119
struct PID {
  GID gid;
  SID sid;
  explicit PID (void) : PID(GID::INVALID) {}
  explicit PID (GID gid) : gid(gid), sid(SID::INVALID) {}
  bool valid (void) const { return gid != GID::INVALID; }
  friend bool operator== (const PID &l, const PID &r) { return l.gid == r.gid && l.sid == r.sid; }
};

/// treetypes.h:122-188
struct Genealogy {
  PID mother, father, self;
  uint generation = 0;
  void setSID (SID sid) { self.sid = sid; }
  void updateAfterCloning (GIDManager &m) { mother = self; father = PID(); self = PID(m()); generation++; }
  void updateAfterCrossing (const Genealogy &mo, const Genealogy &fa, GIDManager &m) {
    mother = mo.self;  father = fa.self;  self = PID(m());
    generation = std::max(mo.generation, fa.generation) + 1;
  }
  void setAsPrimordial (GIDManager &m) { self = PID(m()); mother = PID(); father = PID(); generation = 0; }
};

/// treetypes.h:191-195
template <typename UserData>
struct InsertionResult {
  SID sid;
  UserData *udata;
};

/// phylogenetictree.hpp:28-34
struct NoUserData {
  NoUserData (void) {}
  NoUserData (GID) {}
  void removedFromEnveloppe (void) const {}
};

/// callbacks.hpp:16-45 -- specialise for your tree type to receive the events
template <typename PT>
struct Callbacks_t {
  void onStepped (uint, const LivingSet &) {}
  void onNewSpecies (SID, SID) {}
  void onGenomeEntersEnveloppe (SID, GID) {}
  void onGenomeLeavesEnveloppe (SID, GID) {}
  void onMajorContributorChanged (SID, SID, SID) {}
};

namespace config {
/// ptreeconfig.h:17-69 with the defaults of ptreeconfig.cpp:7-19.  Values are read when a tree is
/// constructed (the reference latches rsetSize there too, phylogenetictree.hpp:90).
struct PTree {
#define APT_PARAM(TYPE, NAME, VALUE) static TYPE& NAME (void) { static TYPE v = VALUE; return v; }
  APT_PARAM(double, compatibilityThreshold, .1)
  APT_PARAM(double, similarityThreshold, .5)
  APT_PARAM(double, avgCompatibilityThreshold, .25)
  APT_PARAM(uint, rsetSize, 5)
  APT_PARAM(bool, simpleNewSpecies, true)
  APT_PARAM(uint, stillbornTrimmingPeriod, 100)
  APT_PARAM(float, stillbornTrimmingThreshold, .25f)
  APT_PARAM(float, stillbornTrimmingDelay, 4.f)
  APT_PARAM(uint, stillbornTrimmingMinDelay, 200)
  APT_PARAM(bool, DEBUG_FULL_CONTINUOUS, true)
  APT_PARAM(int, DEBUG_ENV_CRIT, 1)
#undef APT_PARAM
};
} // namespace config

/// How a genome is laid out for the device; specialise per genome type (see the file header).
/// Allocator of the staging arrays handed to the engine (encoded genomes, results): page-locked memory
/// (apt_host_alloc), so the births' copy to the device is one asynchronous DMA and results land in place.
template <typename T>
struct PinnedAllocator {
  using value_type = T;
  PinnedAllocator (void) = default;
  template <typename U> PinnedAllocator (const PinnedAllocator<U>&) {}
  T* allocate (std::size_t n) {
    void *p = nullptr;
    if (apt_host_alloc(uint64_t(n) * sizeof(T), &p) != APT_OK || !p) throw std::bad_alloc();
    return static_cast<T*>(p);
  }
  void deallocate (T *p, std::size_t) { apt_host_free(p); }
  /// resize() default-initialises (no zero fill): every element is written before it is read
  template <typename U, typename... A> void construct (U *p, A&&... a) {
    if constexpr (sizeof...(A) == 0) ::new (static_cast<void*>(p)) U;
    else ::new (static_cast<void*>(p)) U(std::forward<A>(a)...);
  }
  template <typename U> bool operator== (const PinnedAllocator<U>&) const { return true; }
  template <typename U> bool operator!= (const PinnedAllocator<U>&) const { return false; }
};
template <typename T> using PinnedVector = std::vector<T, PinnedAllocator<T>>;

template <typename GENOME> struct DeviceEncoding;

// ---- phylogenetictree.hpp ------------------------------------------------------------------------
template <typename GENOME, typename UDATA = NoUserData>
class PhylogeneticTree {
public:
  using Genome = GENOME;
  using UserData = UDATA;
  using Callbacks = Callbacks_t<PhylogeneticTree<Genome, UserData>>;
  using Config = config::PTree;
  using InsertionResult = apt::InsertionResult<UserData>;
  using Encoding = DeviceEncoding<GENOME>;

  /// node.hpp:25-85
  struct Representative {
    uint timestamp;
    Genome genome;
    std::unique_ptr<UserData> userData;
  };

  /// What the reference exposes through nodeAt(sid)-> : id, parent, children, data, rset
  struct Node {
    SID id, parent;                                  ///< parent == SID::INVALID for parentless nodes
    struct { uint firstAppearance, lastAppearance, count, currentlyAlive, pendingCandidates; } data;
    std::vector<SID> children;
    const std::vector<Representative> *rset;
  };

  /// phylogenetictree.hpp:343-355
  struct Stats {
    uint insertions = 0, deletions = 0, comparisons = 0, branching = 0;
    friend std::ostream& operator<< (std::ostream &os, const Stats &s) {
      return os << " " << s.insertions << " " << s.deletions << " " << s.comparisons << " " << s.branching;
    }
  };
  struct StatsHeader {
    friend std::ostream& operator<< (std::ostream &os, const StatsHeader&) {
      return os << " PTInsertions PTDeletions PTComparisons PTBranching";
    }
  };

  /// Create an empty PTree (:88-95) on CUDA device `device`
  explicit PhylogeneticTree (int device = 0) : _callbacks(nullptr) {
    apt_params p;
    apt_params_default(&p);
    p.kind = Encoding::kind;  p.words = Encoding::words;
    p.rset_size = Config::rsetSize();
    p.full_continuous = Config::DEBUG_FULL_CONTINUOUS();  p.env_crit = Config::DEBUG_ENV_CRIT();
    p.compat_thr = Config::compatibilityThreshold();  p.sim_thr = Config::similarityThreshold();
    p.avg_thr = Config::avgCompatibilityThreshold();
    p.trim_period = Config::stillbornTrimmingPeriod();  p.trim_thr = Config::stillbornTrimmingThreshold();
    p.trim_delay = Config::stillbornTrimmingDelay();  p.trim_min_delay = Config::stillbornTrimmingMinDelay();
    p.device = device;
    _rsetSize = p.rset_size;
    check(apt_tree_create(&p, &_tree));
  }

  ~PhylogeneticTree (void) { if (_tree) apt_tree_destroy(_tree); }
  PhylogeneticTree (const PhylogeneticTree&) = delete;
  PhylogeneticTree& operator= (const PhylogeneticTree&) = delete;
  PhylogeneticTree (PhylogeneticTree &&that) noexcept
    : _tree(that._tree), _rsetSize(that._rsetSize), _aliveSpecies(std::move(that._aliveSpecies)),
      _callbacks(nullptr), _rsets(std::move(that._rsets)) { that._tree = nullptr; }

  // == Accessors (:155-202) ==
  Callbacks* callbacks (void) { return _callbacks; }
  uint width (void) const { return st(5); }
  uint step (void) const { return st(6); }
  SID nextNodeID (void) const { return SID(st(4)); }
  const LivingSet& aliveSpecies (void) const { return _aliveSpecies; }
  SID root (void) const { uint32_t r; check(apt_tree_root(_tree, &r)); return SID(r); }

  /// \return a view of the node with \p sid (throws like :168-177 when it does not exist)
  Node nodeAt (SID sid) const {
    apt_node_info i;
    check(apt_tree_node_info(_tree, uint32_t(sid), &i));
    Node n;
    n.id = SID(i.sid);  n.parent = SID(i.parent_sid);
    n.data = { i.data[0], i.data[1], i.data[2], i.data[3], i.data[4] };
    std::vector<uint32_t> c (i.n_children);
    if (!c.empty()) check(apt_tree_node_children(_tree, uint32_t(sid), c.data(), uint32_t(c.size())));
    for (uint32_t s: c) n.children.push_back(SID(s));
    static const std::vector<Representative> empty;
    const auto *rs = rsetOf(sid);
    n.rset = rs ? rs : &empty;
    return n;
  }

  /// :181-187
  UserData* getUserData (const PID &pid) const {
    nodeAt(pid.sid);                                           // throws for unknown species
    if (const auto *rs = rsetOf(pid.sid))
      for (const Representative &ep: *rs)
        if (ep.genome.genealogy().self.gid == pid.gid) return ep.userData.get();
    return nullptr;
  }

  // == Modifiers (:216-221) ==
  void setCallbacks (Callbacks *c) const { _callbacks = c; }
  void setStep (uint step) { check(apt_tree_set_step(_tree, step)); }

  // == Core API ==
  /// :245-262
  template <typename IT, typename F>
  void step (uint step, IT begin, IT end, F sidExtractor) {
    _aliveSpecies.clear();
    std::vector<uint32_t> sids;
    for (IT it = begin; it != end; ++it) {
      SID s = sidExtractor(*it);
      _aliveSpecies.insert(s);
      sids.push_back(uint32_t(s));
    }
    check(apt_tree_step(_tree, step, sids.data(), uint32_t(sids.size())));
    dropTrimmed();
    replayEvents(nullptr, nullptr, 0);
  }

  /// :268-302
  InsertionResult addGenome (const Genome &g) {
    InsertionResult r;
    addGenomes(&g, &g + 1, &r);
    return r;
  }

  /// Batched insertion: identical, by definition, to addGenome on each element of [first, last)
  /// in order (results, callbacks and their order included).  `out` may be null.
  template <typename IT>
  void addGenomes (IT first, IT last, InsertionResult *out) {
#ifdef APT_DROPIN_PROFILE
    const auto p0 = std::chrono::steady_clock::now();
#endif
    _pl.clear();  _cp.clear();  _ids.clear();
    std::vector<const Genome*> gs;
    gs.reserve(size_t(std::distance(first, last)));
    for (IT it = first; it != last; ++it) gs.push_back(&*it);
    const size_t n = gs.size();
    if (!n) return;
    _pl.resize(n * Encoding::words);  _cp.resize(n * 3);  _ids.resize(n * 5);
    _sid.resize(n);  _slot.resize(n);  _rep.resize(n);
    // genome objects -> staging arrays, on the engine's worker threads for large batches: DeviceEncoding::encode is
    // called concurrently for DIFFERENT genomes (define APT_DROPIN_SERIAL_ENCODE if yours is not a pure function of its genome)
    struct EncodeCtx { PhylogeneticTree *self; const Genome *const *gs; size_t n; } ectx { this, gs.data(), n };
    auto encodeSlice = [] (void *c, uint32_t lo, uint32_t hi) {
      EncodeCtx &x = *static_cast<EncodeCtx*>(c);
      PhylogeneticTree &t = *x.self;
      const size_t n = x.n;
      for (size_t i = lo; i < hi; i++) {
        Encoding::encode(*x.gs[i], &t._pl[i * Encoding::words], &t._cp[i * 3]);
        const Genealogy &ge = x.gs[i]->genealogy();
        t._ids[i] = uint32_t(ge.self.gid);
        t._ids[n + i] = uint32_t(ge.mother.gid);      t._ids[2 * n + i] = uint32_t(ge.mother.sid);
        t._ids[3 * n + i] = uint32_t(ge.father.gid);  t._ids[4 * n + i] = uint32_t(ge.father.sid);
        t._sid[i] = APT_INVALID;  t._slot[i] = -1;  t._rep[i] = APT_INVALID;
      }
    };
#ifdef APT_DROPIN_SERIAL_ENCODE
    encodeSlice(&ectx, 0, uint32_t(n));
#else
    check(apt_tree_parallel_for(_tree, uint32_t(n), +encodeSlice, &ectx));
#endif
#ifdef APT_DROPIN_PROFILE
    const auto p1 = std::chrono::steady_clock::now();
#endif
    uint32_t done = 0;
    const int rc = apt_tree_add_genomes(_tree, uint32_t(n), _pl.data(), _cp.data(), &_ids[0], &_ids[n], &_ids[2 * n],
                                        &_ids[3 * n], &_ids[4 * n], _sid.data(), _slot.data(), _rep.data(), &done);
#ifdef APT_DROPIN_PROFILE
    const auto p2 = std::chrono::steady_clock::now();
#endif
    // births committed before an error stay committed, exactly as when the reference throws
    replayEvents(gs.data(), out, done);
#ifdef APT_DROPIN_PROFILE
    const auto p3 = std::chrono::steady_clock::now();
    auto ms = [] (auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    if (n >= 4096) fprintf(stderr, "dropin: encode %.3f engine %.3f replay %.3f ms\n", ms(p0, p1), ms(p1, p2), ms(p2, p3));
#endif
    check(rc);
  }

  /// :305-317
  void delGenome (const Genome &g) { check(apt_tree_del_genome(_tree, uint32_t(g.genealogy().self.sid))); }

  /// :320-329
  void registerCandidate (const Genealogy &g) { check(apt_tree_register_candidate(_tree, uint32_t(g.mother.sid), uint32_t(g.father.sid), +1)); }
  void unregisterCandidate (const Genealogy &g) { check(apt_tree_register_candidate(_tree, uint32_t(g.mother.sid), uint32_t(g.father.sid), -1)); }

  // == Stats (:335-365) ==
  void resetStats (void) { check(apt_tree_reset_stats(_tree)); }
  const Stats& stats (void) const {
    uint32_t s[8];
    check(apt_tree_stats(_tree, s));
    _stats.insertions = s[0]; _stats.deletions = s[1]; _stats.comparisons = s[2]; _stats.branching = s[3];
    return _stats;
  }

  /// The engine behind this tree (counters, timers, node accessors of apt_b200.h)
  apt_tree* handle (void) const { return _tree; }

#ifdef APT_WITH_JSON
  using json = nlohmann::json;

  /// :934-941 / :887-904; genomes and user data are serialised by the user's own to_json
  static void toJson (json &j, const PhylogeneticTree &pt) {
    uint32_t s[8];
    check(apt_tree_stats(pt._tree, s));
    j["_step"] = s[6];
    j["_envSize"] = pt._rsetSize;
    j["_stillborns"] = s[7];
    json alive = json::array();
    for (SID a: pt._aliveSpecies) alive.push_back(uint(a));
    j["alive"] = alive;
    const SID r = pt.root();
    j["tree"] = r == SID::INVALID ? json() : pt.nodeJson(r);
    j["nextSID"] = s[4];
  }

  /// :946-967 -- into an EMPTY tree.  Genomes and user data are rebuilt by the user's own from_json and
  /// re-encoded for the device; parents are re-derived from the stored contributors as in the reference.
  static void fromJson (const json &j, PhylogeneticTree &pt) {
    json e;                                   // the same dump with device-encoded genomes, for the engine
    e["_step"] = j["_step"];  e["_envSize"] = j["_envSize"];  e["_stillborns"] = j["_stillborns"];
    e["alive"] = j["alive"].is_null() ? json::array() : j["alive"];
    e["nextSID"] = j["nextSID"];
    e["tree"] = j["tree"].is_null() ? json() : pt.loadNode(j["tree"]);
    const std::string text = e.dump();
    check(apt_tree_from_json(pt._tree, text.data(), text.size()));
    pt._aliveSpecies.clear();
    for (const auto &a: e["alive"]) pt._aliveSpecies.insert(SID(a.get<uint>()));
  }

  /// :1004-1016
  static PhylogeneticTree readFrom (const std::string &filename, int device = 0) {
    std::ifstream ifs (filename);
    if (!ifs) throw std::invalid_argument("Unable to open '" + filename + "' for reading");
    PhylogeneticTree pt (device);
    fromJson(json::parse(ifs), pt);
    return pt;
  }

  /// :984-1001
  bool saveTo (const std::string &filename) const {
    std::ofstream ofs (filename);
    if (!ofs) return false;
    saveTo(ofs, -1);
    return true;
  }
  void saveTo (std::ostream &os, int indent = 2) const {
    json j;
    toJson(j, *this);
    os << j.dump(indent);
  }
#endif

private:
  apt_tree *_tree = nullptr;
  uint _rsetSize;
  LivingSet _aliveSpecies;
  mutable Callbacks *_callbacks;
  mutable Stats _stats;
  /// host-side genome objects per enveloppe slot, indexed by species id (ids are handed out densely, :195); a deque:
  /// element addresses survive growth (Node views point at them) and a look-up is two hops instead of a hash chain
  std::deque<std::vector<Representative>> _rsets;
  const std::vector<Representative>* rsetOf (SID sid) const {
    return (uint32_t(sid) < _rsets.size() && !_rsets[uint32_t(sid)].empty()) ? &_rsets[uint32_t(sid)] : nullptr;
  }
  std::vector<Representative>& rsetFor (SID sid) {
    if (uint32_t(sid) >= _rsets.size()) _rsets.resize(size_t(uint32_t(sid)) + 1);
    return _rsets[uint32_t(sid)];
  }
  PinnedVector<uint32_t> _pl, _sid, _rep, _ids;
  PinnedVector<float> _cp;
  PinnedVector<int32_t> _slot;

  static void check (int rc) {
    if (rc == APT_OK) return;
    const std::string msg = apt_last_error();
    if (rc == APT_ERR_ARGUMENT) throw std::invalid_argument(msg);   // utils::Thrower's default (:169-174)
    throw std::runtime_error(msg);
  }

  uint st (int i) const { uint32_t s[8]; check(apt_tree_stats(_tree, s)); return s[i]; }

  /// Forget the genome objects of species removed by stillborn trimming (:816-860)
  void dropTrimmed (void) {
    for (size_t sid = 0; sid < _rsets.size(); sid++) {
      if (_rsets[sid].empty()) continue;
      apt_node_info i;
      if (apt_tree_node_info(_tree, uint32_t(sid), &i) != APT_OK) std::vector<Representative>().swap(_rsets[sid]);
    }
  }

  /// Fires the recorded callbacks in the reference's order and mirrors enveloppe changes into the
  /// host-side genome / user-data objects (:657-726): for each committed birth, its events first,
  /// then the slot update, so handlers observe the state the reference would show them.
  void replayEvents (const Genome *const *gs, InsertionResult *out, uint32_t n) {
    const uint32_t ne = apt_tree_event_count(_tree);
    std::vector<apt_event> ev (ne);
    if (ne) apt_tree_events(_tree, ev.data(), ne, 1);
    size_t e = 0;
    const uint stepNow = step();           // (one call: step() goes through the C ABI)
    auto fire = [&] (const apt_event &x) {
      if (!_callbacks) return;
      switch (x.type) {
      case APT_EV_STEPPED: _callbacks->onStepped(x.a, _aliveSpecies); break;
      case APT_EV_NEW_SPECIES: _callbacks->onNewSpecies(SID(x.a), SID(x.b)); break;
      case APT_EV_ENTERS_ENVELOPPE: _callbacks->onGenomeEntersEnveloppe(SID(x.a), GID(x.b)); break;
      case APT_EV_LEAVES_ENVELOPPE: _callbacks->onGenomeLeavesEnveloppe(SID(x.a), GID(x.b)); break;
      case APT_EV_MAJOR_CONTRIBUTOR: _callbacks->onMajorContributorChanged(SID(x.a), SID(x.b), SID(x.c)); break;
      }
    };
    // most births neither enter an enveloppe nor fire a callback: their results are filled in bulk (worker threads),
    // the ordered pass below only stops at the others
    if (out) {
      struct FillCtx { const uint32_t *sid; InsertionResult *out; } fctx { _sid.data(), out };
      auto fillSlice = [] (void *c, uint32_t lo, uint32_t hi) {
        FillCtx &x = *static_cast<FillCtx*>(c);
        for (uint32_t i = lo; i < hi; i++) x.out[i] = InsertionResult{ SID(x.sid[i]), nullptr };
      };
      check(apt_tree_parallel_for(_tree, n, +fillSlice, &fctx));
    }
    const int32_t *slotp = _slot.data();
    const uint32_t *sidp = _sid.data();
    for (uint32_t i = 0; i < n; i++) {
      {
        // nothing happens at a birth that neither enters an enveloppe nor is the subject of the next pending callback:
        // runs of those (95 % of a steady generation) are skipped in a loop over two arrays and two constants
        const bool pending = e < ne;
        const uint32_t ea = pending ? ev[e].a : 0u, eb = pending ? ev[e].b : 0u;
        while (i < n && slotp[i] < 0 && !(pending && (sidp[i] == ea || sidp[i] == eb))) i++;
        if (i == n) break;
      }
      const SID sid = SID(_sid[i]);
      const uint32_t gid = _ids[i];
      if (e < ne && ev[e].type == APT_EV_NEW_SPECIES && ev[e].b == _sid[i]) fire(ev[e++]);
      UserData *ud = nullptr;
      if (_slot[i] >= 0) {
        auto &rset = rsetFor(sid);
        if (_rep[i] != APT_INVALID) {
          if (e < ne && ev[e].type == APT_EV_LEAVES_ENVELOPPE) fire(ev[e++]);
          if (e < ne && ev[e].type == APT_EV_ENTERS_ENVELOPPE && ev[e].b == gid) fire(ev[e++]);
          Representative &ep = rset[size_t(_slot[i])];
          ep.userData->removedFromEnveloppe();               // :708
          ud = ep.userData.get();
          *ep.userData = UserData(GID(_rep[i]));             // rebuilt from the OUTGOING genome's id, :710
          ep.genome = *gs[i];
          ep.timestamp = stepNow;
        } else {
          if (e < ne && ev[e].type == APT_EV_ENTERS_ENVELOPPE && ev[e].b == gid) fire(ev[e++]);
          Representative ep;
          ep.timestamp = stepNow;
          ep.genome = *gs[i];
          ep.userData = std::make_unique<UserData>(GID(gid));
          rset.push_back(std::move(ep));
          ud = rset.back().userData.get();
        }
      }
      if (e < ne && ev[e].type == APT_EV_MAJOR_CONTRIBUTOR && ev[e].a == _sid[i]) fire(ev[e++]);
      if (out && ud) out[i].udata = ud;
    }
    while (e < ne) fire(ev[e++]);
  }

#ifdef APT_WITH_JSON
  /// One node of a dump -> the engine's layout (genome as {c, g, w}); fills the host-side enveloppe objects
  json loadNode (const json &jn) {
    json out;
    const SID sid = SID(jn["id"].get<uint>());
    out["id"] = jn["id"];  out["data"] = jn["data"];  out["contribs"] = jn["contribs"].is_null() ? json::array() : jn["contribs"];
    out["dists"] = jn["dists"].is_null() ? json::array() : jn["dists"];
    json env = json::array(), children = json::array();
    auto &rset = rsetFor(sid);
    if (!jn["envlp"].is_null())
      for (const auto &jr: jn["envlp"]) {
        Representative ep;
        ep.timestamp = 0;
        ep.genome = jr[0].get<Genome>();
        ep.userData = std::make_unique<UserData>(GID::INVALID);        // node.hpp:59-63
        *ep.userData = jr[1].get<UserData>();
        uint32_t payload[Encoding::words];
        float compat[3];
        Encoding::encode(ep.genome, payload, compat);
        const Genealogy &ge = ep.genome.genealogy();
        json g;
        g["c"] = { compat[0], compat[1], compat[2] };
        g["g"] = { { "f", { { "g", uint(ge.father.gid) }, { "s", uint(ge.father.sid) } } },
                   { "m", { { "g", uint(ge.mother.gid) }, { "s", uint(ge.mother.sid) } } },
                   { "s", { { "g", uint(ge.self.gid) }, { "s", uint(ge.self.sid) } } } };
        g["w"] = std::vector<uint32_t>(payload, payload + Encoding::words);
        env.push_back({ g, uint(ge.self.gid) });
        rset.push_back(std::move(ep));
      }
    if (!jn["children"].is_null())
      for (const auto &jc: jn["children"]) children.push_back(loadNode(jc));
    out["envlp"] = env;  out["children"] = children;
    return out;
  }

  json nodeJson (SID sid) const {
    apt_node_info i;
    check(apt_tree_node_info(_tree, uint32_t(sid), &i));
    json j, jd, jc, je = json::array(), jct = json::array();
    if (const auto *rs = rsetOf(sid))
      for (const Representative &r: *rs) je.push_back({ r.genome, *r.userData });
    std::vector<uint32_t> ct (size_t(i.n_contributors) * 3);
    if (!ct.empty()) check(apt_tree_node_contributors(_tree, uint32_t(sid), ct.data(), i.n_contributors));
    for (uint32_t c = 0; c < i.n_contributors; c++) jct.push_back({ ct[3 * c], ct[3 * c + 1], bool(ct[3 * c + 2]) });
    std::vector<float> d (size_t(_rsetSize) * _rsetSize);
    check(apt_tree_node_distances(_tree, uint32_t(sid), d.data()));
    for (uint32_t a = 0; a < i.fill; a++)
      for (uint32_t b = a + 1; b < i.fill; b++) jd.push_back({ a, b, d[a * _rsetSize + b] });
    std::vector<uint32_t> ch (i.n_children);
    if (!ch.empty()) check(apt_tree_node_children(_tree, uint32_t(sid), ch.data(), i.n_children));
    for (uint32_t c: ch) jc.push_back(nodeJson(SID(c)));
    j["id"] = i.sid;
    j["data"] = { i.data[0], i.data[1], i.data[2], i.data[3], i.data[4] };
    j["envlp"] = je;
    j["contribs"] = jct;
    j["dists"] = jd;
    j["children"] = jc;
    return j;
  }
#endif
};

} // namespace apt

#endif // APT_PHYLOGENETICTREE_HPP
