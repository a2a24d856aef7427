// apt/phylogenetictree.hpp -- drop-in for APOGeT's phylogeny::PhylogeneticTree<GENOME,UDATA>
// (reference: src/core/tree/phylogenetictree.hpp) whose species-assignment hot path runs on a
// B200 through libapt_b200.so (C ABI: include/apt_b200.h).
//
// Same namespaces, class names, member names, argument meaning and callback order as the reference: the types
// below mirror src/core/tree/treetypes.h (GID, SID, PID, Genealogy, InsertionResult), node.hpp (Node, Node::Ptr,
// Representative), speciesdata.hpp, speciescontributors.h (the read-only part), callbacks.hpp (Callbacks_t) in
// namespace `phylogeny`, and ptreeconfig.h (::config::PTree).  A host program switches by including this header
// instead of "kgd/apt/core/tree/phylogenetictree.hpp" and describing how its genome is laid out on the device
// (tests/cpp/test_dropin.cpp compiles unchanged against either header); `apt` is an alias of `phylogeny`:
//
//   template <> struct phylogeny::DeviceEncoding<MyGenome> {
//     static constexpr apt_kind kind = APT_BITS;      // or APT_F32
//     static constexpr uint32_t words = 8;            // 32-bit payload words
//     static void encode (const MyGenome &g, uint32_t *payload, float compat[3]);
//   };
//
// where payload/compat must be consistent with the genome's own distance() / compatibility()
// (Hamming or L2 over the payload; the two-sided gaussian of src/core/crossover.h:76-92 with
// compat = {optimalDistance, inbreedTolerance, outbreedTolerance}).  New with respect to the
// reference: addGenomes(first, last, out), whose result is defined as that of calling addGenome on
// every element in order; trees are movable as well as copyable.
// nodeAt() / root() / Node::children() / Node::parent() hand out Node::Ptr handles like the reference's; a handle's
// public fields (data, contributors, distances; rset is always live) are brought up to date whenever it is obtained
// through one of those accessors -- fetch it again after mutating the tree instead of keeping it across mutations.
// toJson / saveTo / fromJson / readFrom need nlohmann::json: define APT_WITH_JSON before including.
#ifndef APT_PHYLOGENETICTREE_HPP
#define APT_PHYLOGENETICTREE_HPP

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <deque>
#include <fstream>
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
#include <nlohmann/json.hpp>
#endif

namespace config {
/// ptreeconfig.h:17-69 with the defaults of ptreeconfig.cpp:7-19, in the reference's own (global) namespace `config`.
/// Values are read when a tree is constructed (the reference latches rsetSize there too, phylogenetictree.hpp:90);
/// every accessor returns a reference that can be assigned to (the reference reads them from a config file).
struct PTree {
  using uint = unsigned int;
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

namespace phylogeny {

namespace config = ::config;
using uint = unsigned int;
#ifdef APT_WITH_JSON
using json = nlohmann::json;
#endif

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
  friend bool operator== (const Genealogy &l, const Genealogy &r) {
    return l.mother == r.mother && l.father == r.father && l.self == r.self && l.generation == r.generation;
  }
};

#ifdef APT_WITH_JSON
// treetypes.h:104-107, 158-169 (the generation is not serialised by the reference either)
inline void to_json (json &j, const PID &p) { j["g"] = uint(p.gid); j["s"] = uint(p.sid); }
inline void from_json (const json &j, PID &p) { p.gid = GID(j["g"].get<uint>()); p.sid = SID(j["s"].get<uint>()); }
inline void to_json (json &j, const Genealogy &g) { j["m"] = g.mother; j["f"] = g.father; j["s"] = g.self; }
inline void from_json (const json &j, Genealogy &g) { from_json(j["m"], g.mother); from_json(j["f"], g.father); from_json(j["s"], g.self); }
#endif

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
#ifdef APT_WITH_JSON
  friend void to_json (json &j, const NoUserData&) { j = nullptr; }
  friend void from_json (const json&, NoUserData&) {}
#endif
  friend void assertEqual (const NoUserData&, const NoUserData&, bool) {}
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

/// speciesdata.hpp:15-31
struct SpeciesData {
  uint firstAppearance, lastAppearance, count, currentlyAlive, pendingCandidates;
};

/// speciescontributors.h:16-86, the read-only part (the engine maintains the collection)
class Contributor {
  SID _speciesID;  uint _count;  bool _elligible;
public:
  Contributor (void) : Contributor(SID::INVALID, uint(-1), false) {}
  Contributor (SID sid, uint initialCount, bool elligible) : _speciesID(sid), _count(initialCount), _elligible(elligible) {}
  SID speciesID (void) const { return _speciesID; }
  uint count (void) const { return _count; }
  bool elligible (void) const { return _elligible; }
};

/// speciescontributors.h:118-206, the read-only part
class Contributors {
  SID nodeID;
  std::vector<Contributor> vec;
public:
  Contributors (void) : Contributors(SID::INVALID) {}
  Contributors (SID id) : nodeID(id) {}
  Contributors (SID id, std::vector<Contributor> &&v) : nodeID(id), vec(std::move(v)) {}
  SID getNodeID (void) const { return nodeID; }
  const std::vector<Contributor>& data (void) const { return vec; }
  auto begin (void) const { return vec.begin(); }
  auto end (void) const { return vec.end(); }
};

namespace _details {
/// treetypes.h:233-259
template <typename T>
struct ordered_pair {
  T first, second;
  ordered_pair (T a, T b) : first(std::min(a, b)), second(std::max(a, b)) {}
  friend bool operator< (const ordered_pair &l, const ordered_pair &r) { return l.first != r.first ? l.first < r.first : l.second < r.second; }
};
using DistanceMap = std::map<ordered_pair<uint>, float>;

/// What the reference's utils::assertEqual does on a mismatch is Tools' business (not vendored); here: std::logic_error
[[noreturn]] inline void unequal (const std::string &what) { throw std::logic_error("assertEqual: " + what + " differ"); }
template <typename T> void same (const T &l, const T &r, const char *what) { if (!(l == r)) unequal(what); }
} // namespace _details

template <typename GENOME, typename UDATA> class PhylogeneticTree;

/// node.hpp:16-243.  Handles to the species of a tree: obtained from PhylogeneticTree::nodeAt / root and from
/// children() / parent(); `rset` holds the host-side genome and user-data objects of the enveloppe (always live),
/// the other public fields mirror the engine's state as of the moment the handle was obtained.
template <typename GENOME, typename UDATA>
struct Node {
  using Ptr = std::shared_ptr<Node>;
  using Collection = std::map<SID, Ptr>;

  /// node.hpp:25-85
  struct Representative {
    uint timestamp = 0;
    GENOME genome;
    std::unique_ptr<UDATA> userData;
    Representative (void) = default;
    Representative (const Representative &that) : genome(that.genome), userData(std::make_unique<UDATA>(*that.userData)) {}   // :34-36
    Representative (Representative &&) = default;
    Representative& operator= (Representative that) { std::swap(genome, that.genome); std::swap(userData, that.userData); return *this; }
    static Representative make (const GENOME &g) {
      Representative r;
      r.genome = g;
      r.userData = std::make_unique<UDATA>(g.genealogy().self.gid);
      return r;
    }
#ifdef APT_WITH_JSON
    friend void to_json (json &j, const Representative &p) { j = { p.genome, *p.userData }; }
    friend void from_json (const json &j, Representative &p) {
      p.genome = j[0].get<GENOME>();
      p.userData = std::make_unique<UDATA>(GID::INVALID);
      *p.userData = j[1].get<UDATA>();
    }
#endif
    friend void assertEqual (const Representative &l, const Representative &r, bool deepcopy) {
      assertEqual(l.genome, r.genome, deepcopy);                     // the user's, found by ADL as in the reference
      if (bool(l.userData) != bool(r.userData)) _details::unequal("user data");
      if (l.userData) {
        if (deepcopy && l.userData.get() == r.userData.get()) _details::unequal("user data addresses (shallow copy)");
        assertEqual(*l.userData, *r.userData, deepcopy);
      }
    }
  };

  SpeciesData data {};
  Contributors contributors;
  std::vector<Representative> rset;
  _details::DistanceMap distances;

  explicit Node (SID id) : contributors(id) {}
  SID id (void) const { return contributors.getNodeID(); }
  Node* parent (void) { return linked(), _parent; }
  const Node* parent (void) const { return linked(), _parent; }
  const std::vector<Ptr>& children (void) const { return linked(), _children; }
  Ptr& child (size_t i) { return linked(), _children[i]; }
  auto representativeGenome (uint i) const { return rset[i].genome; }
  auto representativeId (uint i) const { return representativeGenome(i).genealogy().self.gid; }
  bool extinct (void) const { return data.currentlyAlive == 0 && data.pendingCandidates == 0; }

  /// node.hpp:212-219: this node and its sub-species, in dot format
  void logTo (std::ostream &os) const {
    os << "\t" << id() << ";\n";
    for (const Ptr &n: children()) {
      os << "\t" << id() << " -> " << n->id() << ";\n";
      n->logTo(os);
    }
  }

  /// node.hpp:222-235
  friend void assertEqual (const Node &l, const Node &r, bool deepcopy) {
    using _details::same;
    const Node *lp = l.parent(), *rp = r.parent();
    same(bool(lp), bool(rp), "parents");
    if (lp && rp) same(lp->id(), rp->id(), "parent ids");
    same(l.id(), r.id(), "species ids");
    same(l.data.firstAppearance, r.data.firstAppearance, "first appearances");
    same(l.data.lastAppearance, r.data.lastAppearance, "last appearances");
    same(l.data.count, r.data.count, "counts");
    same(l.data.currentlyAlive, r.data.currentlyAlive, "live counts");
    same(l.data.pendingCandidates, r.data.pendingCandidates, "pending candidates");
    same(l.contributors.data().size(), r.contributors.data().size(), "contributor counts");
    for (size_t i = 0; i < l.contributors.data().size(); i++) {
      const Contributor &a = l.contributors.data()[i], &b = r.contributors.data()[i];
      if (a.speciesID() != b.speciesID() || a.count() != b.count() || a.elligible() != b.elligible()) _details::unequal("contributors");
    }
    same(l.rset.size(), r.rset.size(), "enveloppe sizes");
    for (size_t i = 0; i < l.rset.size(); i++) assertEqual(l.rset[i], r.rset[i], deepcopy);
    same(l.distances.size(), r.distances.size(), "distance map sizes");
    for (auto a = l.distances.begin(), b = r.distances.begin(); a != l.distances.end(); ++a, ++b)
      if (a->first.first != b->first.first || a->first.second != b->first.second || a->second != b->second) _details::unequal("distances");
    const auto &lc = l.children(), &rc = r.children();
    same(lc.size(), rc.size(), "sub-species counts");
    for (size_t i = 0; i < lc.size(); i++) {
      if (deepcopy && lc[i].get() == rc[i].get()) _details::unequal("node addresses (shallow copy)");
      assertEqual(*lc[i], *rc[i], deepcopy);
    }
  }

private:
  friend class PhylogeneticTree<GENOME, UDATA>;
  const PhylogeneticTree<GENOME, UDATA> *_owner = nullptr;
  mutable Node *_parent = nullptr;
  mutable std::vector<Ptr> _children;
  mutable uint64_t _epoch = 0;          ///< tree modification count the scalar fields were read at
  mutable uint64_t _linkEpoch = 0;      ///< ... and the one parent / children were resolved (and refreshed) at
  void linked (void) const;             // defined below the tree
};

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
  using InsertionResult = phylogeny::InsertionResult<UserData>;
  using Encoding = DeviceEncoding<GENOME>;

  using Node = phylogeny::Node<Genome, UserData>;
  using Node_ptr = typename Node::Ptr;
  using Nodes = typename Node::Collection;
  using Representative = typename Node::Representative;

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

  /// Constructs a deep copy of that PTree (:98-109): the nodes reachable from its root, eligibilities recomputed;
  /// as in the reference the copy has no callbacks, no alive-species set and zeroed statistics
  PhylogeneticTree (const PhylogeneticTree &that) : _rsetSize(that._rsetSize), _callbacks(nullptr) {
    check(apt_tree_clone(that._tree, &_tree));
    for (const Node_ptr &tn: that._nodes) {
      if (!tn) continue;
      apt_node_info i;
      if (apt_tree_node_info(_tree, uint32_t(tn->id()), &i) != APT_OK) continue;     // not reachable from the root: not copied (:123-136)
      nodeFor(tn->id())->rset = tn->rset;
    }
  }
  /// Assigns that PTree to this one (:112-115)
  PhylogeneticTree& operator= (PhylogeneticTree that) { swap(*this, that); return *this; }
  PhylogeneticTree (PhylogeneticTree &&that) noexcept
    : _tree(that._tree), _rsetSize(that._rsetSize), _aliveSpecies(std::move(that._aliveSpecies)),
      _callbacks(that._callbacks), _nodes(std::move(that._nodes)), _root(std::move(that._root)), _epoch(that._epoch + 1) {
    that._tree = nullptr;  that._callbacks = nullptr;
    for (const Node_ptr &n: _nodes) if (n) n->_owner = this;
  }
  /// :138-147 (alive species and statistics travel with the engine here; the reference forgets to swap them)
  friend void swap (PhylogeneticTree &l, PhylogeneticTree &r) {
    using std::swap;
    swap(l._tree, r._tree);  swap(l._rsetSize, r._rsetSize);  swap(l._aliveSpecies, r._aliveSpecies);
    swap(l._callbacks, r._callbacks);  swap(l._nodes, r._nodes);  swap(l._root, r._root);
    l._epoch = r._epoch = std::max(l._epoch, r._epoch) + 1;
    for (const Node_ptr &n: l._nodes) if (n) n->_owner = &l;
    for (const Node_ptr &n: r._nodes) if (n) n->_owner = &r;
  }

  // == Accessors (:155-202) ==
  Callbacks* callbacks (void) { return _callbacks; }
  uint width (void) const { return st(5); }
  uint step (void) const { return st(6); }
  SID nextNodeID (void) const { return SID(st(4)); }
  const LivingSet& aliveSpecies (void) const { return _aliveSpecies; }

  /// \return a smart pointer to the root (can be null), :158-160
  const Node_ptr& root (void) const {
    uint32_t r;
    check(apt_tree_root(_tree, &r));
    _root = (r == APT_INVALID) ? Node_ptr() : synced(SID(r), true);
    return _root;
  }

  /// \return the node with \p sid (throws like :168-177 when it does not exist)
  const Node_ptr& nodeAt (SID sid) const { return synced(sid, true); }

  /// :181-187
  UserData* getUserData (const PID &pid) const {
    apt_node_info i;
    check(apt_tree_node_info(_tree, uint32_t(pid.sid), &i));   // throws for unknown species
    if (const Node *n = nodeOf(pid.sid))
      for (const Representative &ep: n->rset)
        if (ep.genome.genealogy().self.gid == pid.gid) return ep.userData.get();
    return nullptr;
  }

  // == Modifiers (:216-221) ==
  void setCallbacks (Callbacks *c) const { _callbacks = c; }
  void setStep (uint step) { _epoch++; check(apt_tree_set_step(_tree, step)); }

  // == Core API ==
  /// :245-262
  template <typename IT, typename F>
  void step (uint step, IT begin, IT end, F sidExtractor) {
    _epoch++;
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
    std::vector<const Genome*> gs;
    gs.reserve(size_t(std::distance(first, last)));
    for (IT it = first; it != last; ++it) gs.push_back(&*it);
    const size_t n = gs.size();
    if (!n) return;
    _epoch++;
    // (no clear() in front of these: a vector re-initialises every element it grows by, 7 MB per 100 000 births on one
    // thread; every element is written below -- payload rows are zeroed by the slice that encodes them)
    _pl.resize(n * Encoding::words);  _cp.resize(n * 3);  _ids.resize(n * 5);
    _sid.resize(n);  _slot.resize(n);  _rep.resize(n);
    // genome objects -> staging arrays, on the engine's worker threads for large batches: DeviceEncoding::encode is
    // called concurrently for DIFFERENT genomes (define APT_DROPIN_SERIAL_ENCODE if yours is not a pure function of its genome)
    struct EncodeCtx { PhylogeneticTree *self; const Genome *const *gs; size_t n; } ectx { this, gs.data(), n };
    auto encodeSlice = [] (void *c, uint32_t lo, uint32_t hi) {
      EncodeCtx &x = *static_cast<EncodeCtx*>(c);
      PhylogeneticTree &t = *x.self;
      const size_t n = x.n;
      std::memset(&t._pl[size_t(lo) * Encoding::words], 0, size_t(hi - lo) * Encoding::words * sizeof(uint32_t));
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
  void delGenome (const Genome &g) { _epoch++; check(apt_tree_del_genome(_tree, uint32_t(g.genealogy().self.sid))); }

  /// :320-329
  void registerCandidate (const Genealogy &g) { _epoch++; check(apt_tree_register_candidate(_tree, uint32_t(g.mother.sid), uint32_t(g.father.sid), +1)); }
  void unregisterCandidate (const Genealogy &g) { _epoch++; check(apt_tree_register_candidate(_tree, uint32_t(g.mother.sid), uint32_t(g.father.sid), -1)); }

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

protected:
  /// Dump this PTree into dot file \p filename (:874-879; protected there too)
  void logTo (const std::string &filename) const {
    std::ofstream ofs (filename);
    ofs << "digraph {\n";
    if (const Node_ptr &r = root()) r->logTo(ofs);      // (the reference dereferences a null root here)
    ofs << "}\n";
  }
public:

  /// Asserts that two phylogenetic trees are equal (:970-981): root and nodes (recursively: parents, data, contributors,
  /// enveloppes through the genome's and user data's own assertEqual, distances, sub-species), alive species, next
  /// species id, enveloppe size, stillborns, step.  \p deepcopy additionally requires distinct objects.
  /// Throws std::logic_error on the first difference.
  friend void assertEqual (const PhylogeneticTree &lhs, const PhylogeneticTree &rhs, bool deepcopy) {
    using _details::same;
    const Node_ptr lr = lhs.root(), rr = rhs.root();
    same(bool(lr), bool(rr), "roots");
    if (lr) {
      if (deepcopy && lr.get() == rr.get()) _details::unequal("root addresses (shallow copy)");
      assertEqual(*lr, *rr, deepcopy);
    }
    same(lhs.width(), rhs.width(), "node counts");
    uint32_t ls[8], rs[8];
    check(apt_tree_stats(lhs._tree, ls));  check(apt_tree_stats(rhs._tree, rs));
    for (uint32_t sid = 0, seen = 0; sid < ls[4] && seen < ls[5]; sid++) {       // _nodes: also the species no list holds
      apt_node_info li, ri;
      const bool lin = apt_tree_node_info(lhs._tree, sid, &li) == APT_OK, rin = apt_tree_node_info(rhs._tree, sid, &ri) == APT_OK;
      same(lin, rin, "node sets");
      if (!lin) continue;
      seen++;
      const Node_ptr ln = lhs.nodeAt(SID(sid)), rn = rhs.nodeAt(SID(sid));
      if (deepcopy && ln.get() == rn.get()) _details::unequal("node addresses (shallow copy)");
      assertEqual(*ln, *rn, deepcopy);
    }
    same(lhs._aliveSpecies, rhs._aliveSpecies, "alive species");
    same(ls[4], rs[4], "next species ids");
    same(lhs._rsetSize, rhs._rsetSize, "enveloppe sizes");
    same(ls[7], rs[7], "stillborn counts");
    same(ls[6], rs[6], "steps");
  }

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
    uint32_t r;
    check(apt_tree_root(pt._tree, &r));
    j["tree"] = r == APT_INVALID ? json() : pt.nodeJson(SID(r));
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
    // the reference's tree leaves readFrom through its COPY constructor (no move constructor exists there), which
    // drops the set of alive species (:98-109): programs written against it see an empty set after loading
    return PhylogeneticTree(static_cast<const PhylogeneticTree&>(pt));
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
  /// handles = host-side objects of the species, indexed by species id (ids are handed out densely, :195): the
  /// enveloppe's genome / user-data objects live in their `rset`
  mutable std::deque<Node_ptr> _nodes;
  mutable Node_ptr _root;
  mutable uint64_t _epoch = 1;            ///< bumped by every mutating call: handles refresh their fields when it moved
  const Node* nodeOf (SID sid) const { return uint32_t(sid) < _nodes.size() ? _nodes[uint32_t(sid)].get() : nullptr; }
  const Node_ptr& nodeFor (SID sid) const {
    if (uint32_t(sid) >= _nodes.size()) _nodes.resize(size_t(uint32_t(sid)) + 1);
    Node_ptr &n = _nodes[uint32_t(sid)];
    if (!n) { n = std::make_shared<Node>(sid);  n->_owner = this; }
    return n;
  }
  std::vector<Representative>& rsetFor (SID sid) { return nodeFor(sid)->rset; }
  const std::vector<Representative>* rsetOf (SID sid) const {
    const Node *n = nodeOf(sid);
    return (n && !n->rset.empty()) ? &n->rset : nullptr;
  }

  /// The handle of species \p sid with its scalar fields (data, contributors, distances) as the engine holds them now;
  /// \p links: parent and sub-species handles resolved (and refreshed the same way) too.  Throws for unknown species.
  const Node_ptr& synced (SID sid, bool links) const {
    apt_node_info i;
    check(apt_tree_node_info(_tree, uint32_t(sid), &i));
    const Node_ptr &n = nodeFor(sid);
    if (n->_epoch != _epoch) {
      n->data = { i.data[0], i.data[1], i.data[2], i.data[3], i.data[4] };
      std::vector<uint32_t> ct (size_t(i.n_contributors) * 3);
      if (!ct.empty()) check(apt_tree_node_contributors(_tree, uint32_t(sid), ct.data(), i.n_contributors));
      std::vector<Contributor> cv;
      for (uint32_t c = 0; c < i.n_contributors; c++) cv.emplace_back(SID(ct[3 * c]), ct[3 * c + 1], ct[3 * c + 2] != 0);
      n->contributors = Contributors(sid, std::move(cv));
      std::vector<float> d (size_t(_rsetSize) * _rsetSize);
      check(apt_tree_node_distances(_tree, uint32_t(sid), d.data()));
      n->distances.clear();
      for (uint32_t a = 0; a < i.fill; a++)
        for (uint32_t b = a + 1; b < i.fill; b++) n->distances.emplace(_details::ordered_pair<uint>(a, b), d[a * _rsetSize + b]);
      n->_epoch = _epoch;
    }
    if (links && n->_linkEpoch != _epoch) {
      n->_linkEpoch = _epoch;
      n->_parent = i.parent_sid == APT_INVALID ? nullptr : synced(SID(i.parent_sid), false).get();
      std::vector<uint32_t> ch (i.n_children);
      if (!ch.empty()) check(apt_tree_node_children(_tree, uint32_t(sid), ch.data(), i.n_children));
      n->_children.clear();
      for (uint32_t c: ch) n->_children.push_back(synced(SID(c), false));
    }
    return n;
  }
  friend struct phylogeny::Node<Genome, UserData>;

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

  /// Forget the objects of species removed by stillborn trimming (:816-860): the engine lists them
  void dropTrimmed (void) {
    uint32_t n = apt_tree_trimmed(_tree, nullptr, 0, 0);
    if (!n) return;
    std::vector<uint32_t> sids (n);
    n = apt_tree_trimmed(_tree, sids.data(), n, 1);
    for (uint32_t k = 0; k < n; k++)
      if (sids[k] < _nodes.size()) _nodes[sids[k]].reset();
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

/// parent / sub-species handles of a node, resolved against the tree as it is now
template <typename GENOME, typename UDATA>
void Node<GENOME, UDATA>::linked (void) const {
  if (_owner && _linkEpoch != _owner->_epoch) _owner->synced(id(), true);
}

} // namespace phylogeny

namespace apt = phylogeny;

#endif // APT_PHYLOGENETICTREE_HPP
