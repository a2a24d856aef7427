// TEST INFRASTRUCTURE -- CPU restatement of APOGeT's species-assignment path.
//
// This file is the parity ORACLE, not the product: only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it. It restates, in plain sequential
// C++ (no CUDA, no batching), what the reference does per birth; every function cites the
// reference lines it follows (paths relative to /root/reference/src/core). Parity is PINNED:
// tests/test_oracle.py replays the committed traces through this file and compares per-birth
// species ids, the callback stream, Stats and the full JSON dump with tests/golden/*, which
// were produced by the unmodified reference (oracle/_ref/apt_ref_driver, see
// tests/golden/make_golden.py); in the build container the same test also runs the reference
// binary live.
//
// Genome model (user code in the reference's contract, phylogenetictree.hpp:441-442,467-468):
// fixed-length payload of `words` 32-bit words; kind BITS -> Hamming distance, kind F32 -> L2
// distance accumulated in double in index order; compatibility(d) is the two-sided gaussian of
// crossover.h:76-92 with per-genome (mu, sigma_in, sigma_out).

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "trace.h"

namespace {

constexpr uint32_t INVALID = 0xFFFFFFFFu;   // SID::INVALID / GID::INVALID, treetypes.h:24-26,74-76

struct Params {
  uint32_t kind, words, K;
  int fullContinuous;       // DEBUG_FULL_CONTINUOUS, ptreeconfig.cpp:18
  int envCrit;              // DEBUG_ENV_CRIT, ptreeconfig.cpp:19
  double compatThr, simThr, avgThr;   // ptreeconfig.cpp:7-9
  uint32_t trimPeriod;      // ptreeconfig.cpp:13
  float trimThr, trimDelay; // ptreeconfig.cpp:14-15
  uint32_t trimMinDelay;    // ptreeconfig.cpp:16
};

struct Genome {
  const uint32_t *w;        // payload (not owned)
  float mu, sin, sout;
  uint32_t gid, mgid, msid, fgid, fsid;
};

// crossover.h:89-92 (formula documented at :76-86); utils::gauss is from the un-vendored Tools
double compatibility (float mu, float sin, float sout, double d) {
  double sigma = d < mu ? sin : sout;
  return std::exp(-((d - mu) * (d - mu)) / (2. * sigma * sigma));
}

double distanceOf (uint32_t kind, uint32_t words, const uint32_t *a, const uint32_t *b) {
  if (kind == apt_trace::F32) {
    double s = 0;
    for (uint32_t i = 0; i < words; i++) {
      float fa, fb;
      memcpy(&fa, a + i, 4);
      memcpy(&fb, b + i, 4);
      double t = double(fa) - double(fb);
      s += t * t;
    }
    return std::sqrt(s);
  }
  uint32_t s = 0;
  for (uint32_t i = 0; i < words; i++) s += __builtin_popcount(a[i] ^ b[i]);
  return double(s);
}

// treetypes.h:200-227
struct DCCache {
  std::vector<float> distances, compatibilities;
  void clear (void) { distances.clear(); compatibilities.clear(); }
};

// treetypes.h:262-266
struct EnveloppeContribution { bool better; uint32_t than; float value; };

// Intra-enveloppe distances (treetypes.h:233-259): symmetric K x K matrix standing for the map
// keyed by ordered_pair; `has` mirrors which keys exist.
struct DistMat {
  uint32_t K;
  std::vector<float> v;
  std::vector<uint8_t> has;
  explicit DistMat (uint32_t K = 0) : K(K), v(size_t(K) * K, 0.f), has(size_t(K) * K, 0) {}
  float at (uint32_t i, uint32_t j) const {
    if (!has[i * K + j]) throw std::out_of_range("DistanceMap::at");
    return v[i * K + j];
  }
  void set (uint32_t i, uint32_t j, float d) {
    v[i * K + j] = v[j * K + i] = d;
    has[i * K + j] = has[j * K + i] = 1;
  }
  // values in map iteration order: keys (i,j), i<j, lexicographic
  template <typename F> void forEach (F f) const {
    for (uint32_t i = 0; i < K; i++)
      for (uint32_t j = i + 1; j < K; j++)
        if (has[i * K + j]) f(i, j, v[i * K + j]);
  }
  size_t size (void) const { size_t n = 0; forEach([&] (uint32_t, uint32_t, float) { n++; }); return n; }
};

// ---- enveloppe criteria (tree/enveloppecriteria.cpp) -----------------------------------------

float initValue (void) {
  // `ec.value = -std::numeric_limits<double>::max()` stored into a float (:46, :93, :143, :211)
  return -std::numeric_limits<float>::infinity();
}

// enveloppecriteria.cpp:40-84
EnveloppeContribution maxAverage (const DistMat &e, const std::vector<float> &g, uint32_t k) {
  EnveloppeContribution ec { false, uint32_t(-1), initValue() };
  for (uint32_t i = 0; i < k; i++) {
    double c = 0;
    for (uint32_t j = 0; j < k; j++) {
      if (i == j) continue;
      c += - e.at(i, j) + g.at(j);      // float expression accumulated into a double (:62)
    }
    if (ec.value < c) { ec.value = float(c); ec.than = i; }
  }
  ec.better = (ec.value > 0);
  return ec;
}

// enveloppecriteria.cpp:87-133 (default criterion)
EnveloppeContribution maxMinDist (const DistMat &e, const std::vector<float> &g, uint32_t k) {
  EnveloppeContribution ec { false, uint32_t(-1), initValue() };
  for (uint32_t i = 0; i < k; i++) {
    float minBase = std::numeric_limits<float>::max(), minNew = std::numeric_limits<float>::max();
    for (uint32_t j = 0; j < k; j++) {
      if (i == j) continue;
      minBase = std::min(minBase, e.at(i, j));
      minNew = std::min(minNew, g.at(j));
    }
    double c = - minBase + minNew;      // float arithmetic, widened (:117)
    if (ec.value < c) { ec.value = float(c); ec.than = i; }
  }
  ec.better = (ec.value > 0);
  return ec;
}

// enveloppecriteria.cpp:29-37
void avgAndStdDev (const DistMat &m, double &avg, double &stdDev) {
  const double n = double(m.size());
  avg = 0;
  m.forEach([&] (uint32_t, uint32_t, float v) { avg += v; });
  avg /= n;
  stdDev = 0;
  m.forEach([&] (uint32_t, uint32_t, float v) { stdDev += std::pow(avg - v, 2); });
  stdDev = std::sqrt(stdDev / n);
}

// enveloppecriteria.cpp:136-202
EnveloppeContribution maxAvgMinStdDev (const DistMat &e, const std::vector<float> &g, uint32_t k) {
  EnveloppeContribution ec { false, uint32_t(-1), initValue() };
  double baseAVG, baseStdDev;
  avgAndStdDev(e, baseAVG, baseStdDev);
  for (uint32_t i = 0; i < k; i++) {
    DistMat n = e;
    for (uint32_t j = 0; j < k; j++)
      if (i != j) { (void)n.at(i, j); n.set(i, j, g[j]); }
    double newAVG, newStdDev;
    avgAndStdDev(n, newAVG, newStdDev);
    double c = - baseAVG + newAVG + baseStdDev - newStdDev;
    if (ec.value < c) { ec.value = float(c); ec.than = i; }
  }
  ec.better = (ec.value > 0);
  return ec;
}

// enveloppecriteria.cpp:205-272, including its sizing quirk: d_i and d_g are constructed with k
// zeros and then k-1 values are appended (:231, :238-239), so the zeros take part in the ranking.
EnveloppeContribution maxWeightedDist2Avg (const DistMat &e, const std::vector<float> &g, uint32_t k) {
  EnveloppeContribution ec { false, uint32_t(-1), initValue() };
  double A = 0;
  e.forEach([&] (uint32_t, uint32_t, float v) { A += v; });
  A /= double(e.size());
  auto weight = [A] (double d) { return 1 - exp(- (d-A)*(d-A) / (2. * A * A / 16.)); };
  auto sortedDesc = [] (std::vector<double> v) {
    std::sort(v.begin(), v.end(), [] (double a, double b) { return a > b; });
    return v;
  };
  for (uint32_t i = 0; i < k; i++) {
    std::vector<double> d_i (k), d_g (k);
    for (uint32_t j = 0; j < k; j++) {
      if (i == j) continue;
      d_i.push_back(e.at(i, j));
      d_g.push_back(g[j]);
    }
    const auto s_i = sortedDesc(d_i), s_g = sortedDesc(d_g);
    double c = 0;
    for (uint32_t j = 0; j + 1 < k; j++) {
      double nc = - s_i[j], pc = + s_g[j];
      c += weight(pc) * (nc + pc);
    }
    if (ec.value < c) { ec.value = float(c); ec.than = i; }
  }
  ec.better = (ec.value > 0);
  return ec;
}

// enveloppecriteria.cpp:275-289
EnveloppeContribution computeContribution (int crit, const DistMat &e, const std::vector<float> &g,
                                           uint32_t k) {
  switch (crit) {
  case 0: return maxAverage(e, g, k);
  case 1: return maxMinDist(e, g, k);
  case 2: return maxAvgMinStdDev(e, g, k);
  case 3: return maxWeightedDist2Avg(e, g, k);
  default: throw std::logic_error("No function for this use case");
  }
}

// ---- species bookkeeping ----------------------------------------------------------------------

// speciescontributors.h:16-86
struct Contributor { uint32_t sid, count; bool elligible; };
// speciescontributors.h:89-100
struct Contribution { uint32_t species, count; };
using Contributions = std::vector<Contribution>;

// node.hpp:25-85 (+ the genome fields our toy genome serialises)
struct Rep {
  uint32_t timestamp;
  std::vector<uint32_t> w;
  float mu, sin, sout;
  uint32_t gid, mgid, msid, fgid, fsid;
  uint32_t udata;           // gid the slot's user-data object was last built from
};

// node.hpp:16-243, speciesdata.hpp:15-31
struct Node {
  uint32_t id;
  Node *parent = nullptr;
  std::vector<Node*> children;
  uint32_t firstAppearance = 0, lastAppearance = 0, count = 0, currentlyAlive = 0, pendingCandidates = 0;
  std::vector<Contributor> contributors;
  std::vector<Rep> rset;
  DistMat distances;
  bool extinct (void) const { return currentlyAlive == 0 && pendingCandidates == 0; }  // node.hpp:157-159
};

struct Event { uint32_t type, a, b, c; };

struct Tree {
  Params p;
  uint32_t nextNodeID = 0, stillborns = 0, step = 0;
  Node *root = nullptr;
  std::map<uint32_t, Node*> nodes;
  std::vector<Node*> graveyard;         // trimmed nodes (the reference frees them)
  std::set<uint32_t> aliveSpecies;
  uint32_t insertions = 0, deletions = 0, comparisons = 0, branching = 0;  // phylogenetictree.hpp:343-355
  std::vector<Event> events;
  bool logEvents = true;
  std::string error;

  ~Tree (void) {
    // the root may have been trimmed out of `nodes` (it then sits in the graveyard)
    for (auto &kv: nodes) delete kv.second;
    for (Node *n: graveyard) delete n;
  }

  void emit (uint32_t t, uint32_t a, uint32_t b, uint32_t c) { if (logEvents) events.push_back({t, a, b, c}); }

  // phylogenetictree.hpp:168-177
  Node* nodeAt (uint32_t sid) {
    if (sid == INVALID) throw std::invalid_argument("SID::INVALID is (by definition) invalid");
    auto it = nodes.find(sid);
    if (it == nodes.end()) throw std::invalid_argument("No node found for species " + std::to_string(sid));
    return it->second;
  }

  // speciescontributors.h:177-197
  bool elligibile (uint32_t lhs, uint32_t rhs) {
    Node *n = nodes.at(lhs);
    auto pit = nodes.find(rhs);
    if (pit == nodes.end()) return false;
    Node *q = pit->second;
    if (n->firstAppearance <= q->firstAppearance) return false;
    while (q && q != n) q = q->parent;
    return q != n;
  }

  // speciescontributors.cpp:67-83
  uint32_t currentMain (Node *n) {
    auto &vec = n->contributors;
    if (vec.empty()) return INVALID;
    size_t i = 1;
    Contributor mc = vec[0];
    while (!mc.elligible && i < vec.size()) mc = vec[i++];
    return mc.elligible ? mc.sid : INVALID;
  }

  // speciescontributors.cpp:15-65
  uint32_t contributorsUpdate (Node *n, Contributions ctbs) {
    auto &vec = n->contributors;
    size_t i = 0, unprocessed = ctbs.size();
    ctbs.erase(std::remove_if(ctbs.begin(), ctbs.end(),
                              [] (const Contribution &c) { return c.species == INVALID; }), ctbs.end());
    while (i < vec.size() && unprocessed > 0) {
      auto it = std::find_if(ctbs.begin(), ctbs.end(),
                             [&] (const Contribution &c) { return c.species == vec[i].sid; });
      if (it != ctbs.end()) {
        vec[i].count += it->count;
        unprocessed--;
        ctbs.erase(it);
      }
      i++;
    }
    for (const Contribution &c: ctbs)
      vec.push_back({c.species, c.count, elligibile(n->id, c.species)});
    std::stable_sort(vec.begin(), vec.end(),
                     [] (const Contributor &l, const Contributor &r) { return l.count > r.count; });
    return currentMain(n);
  }

  // node.hpp:239-241
  Node* updateParent (Node *n, uint32_t sid) {
    return n->parent = (sid == INVALID) ? nullptr : nodes.at(sid);
  }

  // node.hpp:182-185
  Node* nodeUpdate (Node *n, const Contributions &c) { return updateParent(n, contributorsUpdate(n, c)); }

  // node.hpp:191-194, speciescontributors.cpp:85-90
  Node* nodeUpdateElligibilities (Node *n) {
    for (Contributor &c: n->contributors) c.elligible = elligibile(n->id, c.sid);
    return updateParent(n, currentMain(n));
  }

  // phylogenetictree.hpp:773-783
  void updateElligibilities (void) {
    for (auto &kv: nodes) nodeUpdateElligibilities(kv.second);
  }

  // phylogenetictree.hpp:398-426
  Node* makeNode (const Contributions &contrib) {
    Node *n = new Node;
    n->id = nextNodeID++;
    n->distances = DistMat(p.K);
    n->firstAppearance = n->lastAppearance = step;
    nodes[n->id] = n;
    nodeUpdate(n, contrib);
    if (n->parent) n->parent->children.push_back(n);
    emit(1, n->parent ? n->parent->id : INVALID, n->id, 0);
    return n;
  }

  // phylogenetictree.hpp:430-452 and :456-478, selected as in :483-490
  float speciesMatchingScore (const Genome &g, Node *s, DCCache &cache) {
    const uint32_t k = uint32_t(s->rset.size());
    cache.clear();
    if (p.fullContinuous) {
      float avgCompat = 0;
      for (const Rep &r: s->rset) {
        double d = distanceOf(p.kind, p.words, g.w, r.w.data());
        double c = std::min(compatibility(g.mu, g.sin, g.sout, d), compatibility(r.mu, r.sin, r.sout, d));
        comparisons++;
        avgCompat += c;                                       // float += double, :472
        cache.distances.push_back(float(d));  cache.compatibilities.push_back(float(c));
      }
      return avgCompat / float(k) - p.avgThr;                 // :477
    } else {
      uint32_t matable = 0;
      for (const Rep &r: s->rset) {
        double d = distanceOf(p.kind, p.words, g.w, r.w.data());
        double c = std::min(compatibility(g.mu, g.sin, g.sout, d), compatibility(r.mu, r.sin, r.sout, d));
        comparisons++;
        if (c >= p.compatThr) matable++;                      // :446
        cache.distances.push_back(float(d));  cache.compatibilities.push_back(float(c));
      }
      return matable - p.simThr * k;                          // :451
    }
  }

  // phylogenetictree.hpp:493-548
  void findBestDerived (const Genome &g, const std::vector<Node*> &species, Node *&bestSpecies,
                        float &bestScore, DCCache &bestCache) {
    DCCache cache;
    const size_t S = species.size();
    std::vector<size_t> left (S);     // children not yet visited; visiting newest first (:506)
    for (size_t i = 0; i < S; i++) left[i] = species[i]->children.size();
    uint32_t done = 0, allDone = (1u << S) - 1;
    size_t k = 0;
    while (done != allDone) {
      if (left[k] == 0) done |= (1u << k);
      else {
        branching++;
        Node *sub = species[k]->children[left[k] - 1];
        float score = speciesMatchingScore(g, sub, cache);
        if (bestScore < score) { bestSpecies = sub; bestScore = score; bestCache = cache; }
        if (bestScore > 0) return;
        left[k]--;
      }
      k = (k + 1) % S;
    }
  }

  // phylogenetictree.hpp:657-726
  bool insertInto (const Genome &g, Node *s, const DCCache &cache) {
    const uint32_t k = uint32_t(s->rset.size());
    bool entered = false;
    auto makeRep = [&] (void) {
      Rep r;
      r.timestamp = step;
      r.w.assign(g.w, g.w + p.words);
      r.mu = g.mu; r.sin = g.sin; r.sout = g.sout;
      r.gid = g.gid; r.mgid = g.mgid; r.msid = g.msid; r.fgid = g.fgid; r.fsid = g.fsid;
      r.udata = g.gid;
      return r;
    };
    if (k < p.K) {
      s->rset.push_back(makeRep());
      entered = true;
      emit(2, s->id, g.gid, 0);
      for (uint32_t i = 0; i < k; i++) s->distances.set(i, k, cache.distances[i]);
    } else {
      EnveloppeContribution ec = computeContribution(p.envCrit, s->distances, cache.distances, k);
      if (ec.better) {
        Rep &ep = s->rset[ec.than];
        const uint32_t epId = ep.gid;
        emit(3, s->id, epId, 0);
        emit(2, s->id, g.gid, 0);
        entered = true;
        Rep r = makeRep();
        r.udata = epId;                 // *ep.userData = UserData(ep_id), :710
        ep = r;
        for (uint32_t i = 0; i < k; i++)
          if (i != ec.than) s->distances.set(i, ec.than, cache.distances[i]);
      }
    }
    s->count++;
    s->currentlyAlive++;
    s->lastAppearance = step;
    return entered;
  }

  // phylogenetictree.hpp:741-769
  void updateContributions (Node *s, const Contributions &contrib, bool fromFile = false) {
    Node *oldMC = s->parent, *newMC = nodeUpdate(s, contrib);
    if (oldMC != newMC) {
      if (!newMC) throw std::logic_error("species lost its major contributor");
      if (oldMC) {
        auto &c = oldMC->children;
        c.erase(std::remove(c.begin(), c.end(), s), c.end());
      }
      newMC->children.push_back(s);
      if (!fromFile) {
        updateElligibilities();
        // the reference dereferences a null oldMC here (:765-766); we report INVALID instead
        emit(4, s->id, oldMC ? oldMC->id : INVALID, newMC->id);
      }
    }
  }

  // phylogenetictree.hpp:730-738
  void updateSpeciesContents (const Genome &g, Node *s, const DCCache &cache, const Contributions &ctb,
                              uint32_t &sid, bool &entered) {
    entered = insertInto(g, s, cache);
    if (!ctb.empty()) updateContributions(s, ctb);
    sid = s->id;
  }

  // phylogenetictree.hpp:268-302 and :553-650
  void addGenome (const Genome &g, uint32_t &sid, bool &entered) {
    if (!root) {
      root = makeNode({});
      return updateSpeciesContents(g, root, DCCache{}, {}, sid, entered);
    }
    const uint32_t mSID = g.msid, fSID = g.fsid;
    Node *s0 = nullptr, *s1 = nullptr;
    if (mSID == INVALID && fSID == INVALID) s0 = root;
    else if (fSID == INVALID || mSID == fSID) s0 = nodeAt(mSID);
    else { s0 = nodeAt(mSID); s1 = nodeAt(fSID); }
    if (s0->pendingCandidates > 0) s0->pendingCandidates--;
    if (s1 && s1->pendingCandidates > 0) s1->pendingCandidates--;

    DCCache cache, bestCache;
    Node *best = nullptr;
    float bestScore = -std::numeric_limits<float>::max();
    std::vector<Node*> species { s0 };
    Contributions contrib { { mSID, 1u + (mSID == fSID) } };
    if (s1) { species.push_back(s1); contrib.push_back({ fSID, 1 }); }

    std::map<uint32_t, float> scores;
    for (Node *s: species) {
      float score = speciesMatchingScore(g, s, cache);
      if (bestScore < score) { best = s; bestScore = score; bestCache = cache; }
      scores[s->id] = score;
    }
    // std::sort with a `>=` comparator on two elements (:599-602): libstdc++'s insertion sort
    // moves the second element first iff comp(second, first), i.e. score1 >= score0.
    if (contrib.size() == 2 && scores.at(contrib[1].species) >= scores.at(contrib[0].species))
      std::swap(contrib[0], contrib[1]);

    if (!(bestScore > 0)) findBestDerived(g, species, best, bestScore, bestCache);
    if (bestScore > 0) updateSpeciesContents(g, best, bestCache, contrib, sid, entered);
    else {
      Node *sub = makeNode(contrib);
      updateSpeciesContents(g, sub, DCCache{}, {}, sid, entered);
    }
    insertions++;
  }

  // phylogenetictree.hpp:305-317
  void delGenome (uint32_t sid) {
    Node *n = nodeAt(sid);
    n->lastAppearance = step;
    n->currentlyAlive--;
    deletions++;
  }

  // phylogenetictree.hpp:786-792
  void candidacy (uint32_t mSID, uint32_t fSID, int dir) {
    if (mSID != INVALID) nodeAt(mSID)->pendingCandidates += dir;
    if (mSID != fSID && fSID != INVALID) nodeAt(fSID)->pendingCandidates += dir;
  }

  // phylogenetictree.hpp:816-860
  void performStillbornTrimming (void) {
    const float T = p.trimThr, D = p.trimDelay, MD = float(p.trimMinDelay);
    for (auto it = nodes.begin(); it != nodes.end(); ) {
      Node &s = *it->second;
      bool remove = false;
      if (s.children.empty() && s.extinct()) {
        bool underfilled = (s.rset.size() < T * p.K);
        uint32_t liveTime = s.lastAppearance - s.firstAppearance;
        uint32_t deadTime = step - s.lastAppearance;
        if (underfilled && std::max(MD, liveTime * D) < deadTime) {
          if (s.parent) {
            auto &c = s.parent->children;
            c.erase(std::remove(c.begin(), c.end(), &s), c.end());
          }
          stillborns++;
          remove = true;
        }
      }
      if (remove) { graveyard.push_back(it->second); it = nodes.erase(it); }
      else ++it;
    }
  }

  // phylogenetictree.hpp:245-262
  void doStep (uint32_t s, const uint32_t *sids, uint32_t n) {
    aliveSpecies.clear();
    for (uint32_t i = 0; i < n; i++) aliveSpecies.insert(sids[i]);
    for (uint32_t sid: aliveSpecies) nodeAt(sid)->lastAppearance = s;
    step = s;
    if (p.trimPeriod > 0 && (step % p.trimPeriod) == 0) performStillbornTrimming();
    emit(0, s, uint32_t(aliveSpecies.size()), 0);
  }

  // ---- JSON dump, layout of phylogenetictree.hpp:887-941 ---------------
  static void num (std::string &o, double v) {
    char buf[40];
    snprintf(buf, sizeof buf, "%.17g", v);
    o += buf;
  }
  static void pid (std::string &o, uint32_t g, uint32_t s) {
    o += "{\"g\":" + std::to_string(g) + ",\"s\":" + std::to_string(s) + "}";
  }
  void nodeJson (std::string &o, const Node &n) const {
    o += "{\"id\":" + std::to_string(n.id);
    o += ",\"data\":[" + std::to_string(n.firstAppearance) + "," + std::to_string(n.lastAppearance) + ","
       + std::to_string(n.count) + "," + std::to_string(n.currentlyAlive) + ","
       + std::to_string(n.pendingCandidates) + "]";
    o += ",\"envlp\":";
    if (n.rset.empty()) o += "[]";     // a std::vector converts to an array even when empty
    else {
      o += "[";
      for (size_t i = 0; i < n.rset.size(); i++) {
        const Rep &r = n.rset[i];
        if (i) o += ",";
        o += "[{\"c\":[";  num(o, r.mu);  o += ",";  num(o, r.sin);  o += ",";  num(o, r.sout);
        o += "],\"g\":{\"f\":";  pid(o, r.fgid, r.fsid);
        o += ",\"m\":";  pid(o, r.mgid, r.msid);
        o += ",\"s\":";  pid(o, r.gid, INVALID);
        o += "},\"w\":[";
        for (uint32_t j = 0; j < p.words; j++) { if (j) o += ","; o += std::to_string(r.w[j]); }
        o += "]}," + std::to_string(r.udata) + "]";
      }
      o += "]";
    }
    o += ",\"contribs\":";
    if (n.contributors.empty()) o += "[]";
    else {
      o += "[";
      for (size_t i = 0; i < n.contributors.size(); i++) {
        const Contributor &c = n.contributors[i];
        if (i) o += ",";
        o += "[" + std::to_string(c.sid) + "," + std::to_string(c.count) + ","
           + (c.elligible ? "true" : "false") + "]";
      }
      o += "]";
    }
    o += ",\"dists\":";
    if (n.distances.size() == 0) o += "null";
    else {
      o += "[";
      bool first = true;
      n.distances.forEach([&] (uint32_t i, uint32_t j, float v) {
        if (!first) o += ",";
        first = false;
        o += "[" + std::to_string(i) + "," + std::to_string(j) + ",";  num(o, v);  o += "]";
      });
      o += "]";
    }
    o += ",\"children\":";
    if (n.children.empty()) o += "null";
    else {
      o += "[";
      for (size_t i = 0; i < n.children.size(); i++) { if (i) o += ","; nodeJson(o, *n.children[i]); }
      o += "]";
    }
    o += "}";
  }
  std::string toJson (void) const {
    std::string o = "{\"_step\":" + std::to_string(step) + ",\"_envSize\":" + std::to_string(p.K)
                  + ",\"_stillborns\":" + std::to_string(stillborns) + ",\"alive\":[";
    bool first = true;
    for (uint32_t s: aliveSpecies) { if (!first) o += ","; first = false; o += std::to_string(s); }
    o += "],\"nextSID\":" + std::to_string(nextNodeID) + ",\"tree\":";
    if (root) nodeJson(o, *root); else o += "null";
    o += "}";
    return o;
  }
};

template <typename F>
int guarded (Tree *t, F f) {
  try { f(); return 0; }
  catch (const std::exception &e) { t->error = e.what(); return 1; }
}

} // namespace

// ---- C interface (loaded with ctypes by oracle/__init__.py) -------------------------------------
extern "C" {

void* apto_create (uint32_t kind, uint32_t words, uint32_t K, int fullContinuous, int envCrit,
                   double compatThr, double simThr, double avgThr,
                   uint32_t trimPeriod, float trimThr, float trimDelay, uint32_t trimMinDelay) {
  Tree *t = new Tree;
  t->p = Params{ kind, words, K, fullContinuous, envCrit, compatThr, simThr, avgThr,
                 trimPeriod, trimThr, trimDelay, trimMinDelay };
  return t;
}

void apto_destroy (void *h) { delete static_cast<Tree*>(h); }

const char* apto_error (void *h) { return static_cast<Tree*>(h)->error.c_str(); }

void apto_log_events (void *h, int on) { static_cast<Tree*>(h)->logEvents = on != 0; }

int apto_add_genome (void *h, const uint32_t *payload, const float *compat, uint32_t gid,
                     uint32_t mgid, uint32_t msid, uint32_t fgid, uint32_t fsid,
                     uint32_t *outSid, uint8_t *outEntered) {
  Tree *t = static_cast<Tree*>(h);
  return guarded(t, [&] {
    Genome g { payload, compat[0], compat[1], compat[2], gid, mgid, msid, fgid, fsid };
    uint32_t sid;  bool entered;
    t->addGenome(g, sid, entered);
    *outSid = sid;  *outEntered = entered;
  });
}

int apto_del_genome (void *h, uint32_t sid) {
  Tree *t = static_cast<Tree*>(h);
  return guarded(t, [&] { t->delGenome(sid); });
}

int apto_candidacy (void *h, uint32_t msid, uint32_t fsid, int dir) {
  Tree *t = static_cast<Tree*>(h);
  return guarded(t, [&] { t->candidacy(msid, fsid, dir); });
}

void apto_set_step (void *h, uint32_t step) { static_cast<Tree*>(h)->step = step; }

int apto_step (void *h, uint32_t step, const uint32_t *sids, uint32_t n) {
  Tree *t = static_cast<Tree*>(h);
  return guarded(t, [&] { t->doStep(step, sids, n); });
}

/// out[8] = insertions, deletions, comparisons, branching, nextSID, width, step, stillborns
void apto_stats (void *h, uint32_t *out) {
  Tree *t = static_cast<Tree*>(h);
  out[0] = t->insertions; out[1] = t->deletions; out[2] = t->comparisons; out[3] = t->branching;
  out[4] = t->nextNodeID; out[5] = uint32_t(t->nodes.size()); out[6] = t->step; out[7] = t->stillborns;
}

uint32_t apto_event_count (void *h) { return uint32_t(static_cast<Tree*>(h)->events.size()); }
void apto_events (void *h, uint32_t *out) {
  Tree *t = static_cast<Tree*>(h);
  memcpy(out, t->events.data(), t->events.size() * sizeof(Event));
}

/// Returns the dump length; copies at most `cap` bytes into `out` (call with cap=0 to size it)
uint64_t apto_to_json (void *h, char *out, uint64_t cap) {
  std::string s = static_cast<Tree*>(h)->toJson();
  if (out && cap) memcpy(out, s.data(), std::min<uint64_t>(cap, s.size()));
  return s.size();
}

/// Replays a whole trace exactly as oracle/ref_driver.cpp does; returns the seconds spent in the
/// addGenome loop (negative on error). maxBirths == 0 replays everything, otherwise stops after
/// the generation that crosses it. nDone receives the number of births replayed.
double apto_replay (void *h, const char *path, uint32_t maxBirths, uint32_t *outSid,
                    uint8_t *outEntered, uint32_t *nDone, double *genSeconds) {
  Tree *t = static_cast<Tree*>(h);
  double seconds = 0;
  int rc = guarded(t, [&] {
    apt_trace::Trace tr = apt_trace::load(path);
    if (tr.h.kind != t->p.kind || tr.h.words != t->p.words) throw std::invalid_argument("trace/tree genome mismatch");
    const uint32_t W = tr.h.words;
    std::vector<uint32_t> sidOf (tr.h.n_births, INVALID);
    uint32_t done = 0;
    for (uint32_t gen = 0; gen < tr.h.n_gens; gen++) {
      const uint32_t b0 = tr.gen_birth_off[gen], b1 = tr.gen_birth_off[gen+1];
      auto start = std::chrono::steady_clock::now();
      for (uint32_t b = b0; b < b1; b++) {
        const uint32_t m = tr.mother[b], f = tr.father[b];
        Genome g { &tr.payload[size_t(b) * W], tr.compat[3*b], tr.compat[3*b+1], tr.compat[3*b+2], b + 1,
                   m == apt_trace::NONE ? INVALID : m + 1, m == apt_trace::NONE ? INVALID : sidOf[m],
                   f == apt_trace::NONE ? INVALID : f + 1, f == apt_trace::NONE ? INVALID : sidOf[f] };
        bool entered;
        t->addGenome(g, sidOf[b], entered);
        if (outSid) outSid[b] = sidOf[b];
        if (outEntered) outEntered[b] = entered;
      }
      const double gs = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
      seconds += gs;
      if (genSeconds) genSeconds[gen] = gs;
      for (uint32_t i = tr.gen_death_off[gen]; i < tr.gen_death_off[gen+1]; i++)
        t->delGenome(sidOf[tr.deaths[i]]);
      std::vector<uint32_t> sids;
      for (uint32_t i = tr.gen_alive_off[gen]; i < tr.gen_alive_off[gen+1]; i++)
        sids.push_back(sidOf[tr.alive[i]]);
      t->doStep(gen + 1, sids.data(), uint32_t(sids.size()));
      done = b1;
      if (maxBirths && done >= maxBirths) break;
    }
    if (nDone) *nDone = done;
  });
  return rc ? -1. : seconds;
}

// ---- stateless pieces, for kernel-level parity tests -------------------------------------------

double apto_distance (uint32_t kind, uint32_t words, const uint32_t *a, const uint32_t *b) {
  return distanceOf(kind, words, a, b);
}

double apto_compatibility (float mu, float sin, float sout, double d) { return compatibility(mu, sin, sout, d); }

/// One species score (phylogenetictree.hpp:430-478): birth vs k representatives.
/// outD[k] receives the cached float distances. Returns the score.
float apto_score (uint32_t kind, uint32_t words, uint32_t k, int fullContinuous,
                  double compatThr, double simThr, double avgThr,
                  const uint32_t *birth, const float *birthCompat,
                  const uint32_t *reps, const float *repCompat, float *outD) {
  if (fullContinuous) {
    float avg = 0;
    for (uint32_t i = 0; i < k; i++) {
      double d = distanceOf(kind, words, birth, reps + size_t(i) * words);
      double c = std::min(compatibility(birthCompat[0], birthCompat[1], birthCompat[2], d),
                          compatibility(repCompat[3*i], repCompat[3*i+1], repCompat[3*i+2], d));
      avg += c;
      if (outD) outD[i] = float(d);
    }
    return avg / float(k) - avgThr;
  }
  uint32_t matable = 0;
  for (uint32_t i = 0; i < k; i++) {
    double d = distanceOf(kind, words, birth, reps + size_t(i) * words);
    double c = std::min(compatibility(birthCompat[0], birthCompat[1], birthCompat[2], d),
                        compatibility(repCompat[3*i], repCompat[3*i+1], repCompat[3*i+2], d));
    if (c >= compatThr) matable++;
    if (outD) outD[i] = float(d);
  }
  return matable - simThr * k;
}

/// Dense sweep checker: every birth against every species (apto_score per pair, i.e. phylogenetictree.hpp:430-478),
/// folded into the per-birth winners the CUDA sweep reports -- species with the strictly greatest score (lowest
/// index on ties), lowest species with score > 0 (or INVALID), number of species with score > 0.  Species with an
/// empty enveloppe are skipped.  Split over `threads` host threads by birth (every pair is independent).  For
/// bit-strings the two exp() per pair are taken from per-triple tables filled by the same compatibility() call
/// for every integer distance: identical doubles, 100x fewer libm calls (1e9 pairs stay a matter of seconds).
void apto_score_dense (uint32_t kind, uint32_t words, uint32_t K, int fullContinuous,
                       double compatThr, double simThr, double avgThr,
                       uint32_t nBirths, const uint32_t *births, const float *birthCompat,
                       uint32_t nSpecies, const uint32_t *reps, const float *repCompat, const uint32_t *fill,
                       uint32_t threads, uint32_t *outBest, float *outBestScore, uint32_t *outFirst, uint32_t *outCount) {
  struct Triple { float a, b, c; bool operator< (const Triple &o) const { return a != o.a ? a < o.a : b != o.b ? b < o.b : c < o.c; } };
  std::map<Triple, uint32_t> ids;
  std::vector<std::vector<double>> tables;
  std::vector<uint32_t> bcls (nBirths), rcls (size_t(nSpecies) * K, 0);
  const uint32_t maxD = 32 * words;
  auto classOf = [&] (const float *t) {
    const Triple key { t[0], t[1], t[2] };
    auto it = ids.find(key);
    if (it != ids.end()) return it->second;
    const uint32_t id = uint32_t(tables.size());
    ids.emplace(key, id);
    tables.emplace_back();
    if (kind == 0) {
      tables.back().resize(maxD + 1);
      for (uint32_t d = 0; d <= maxD; d++) tables.back()[d] = compatibility(t[0], t[1], t[2], double(d));
    }
    return id;
  };
  for (uint32_t b = 0; b < nBirths; b++) bcls[b] = classOf(birthCompat + 3 * size_t(b));
  for (uint32_t s = 0; s < nSpecies; s++)
    for (uint32_t r = 0; r < std::min(fill[s], K); r++) rcls[size_t(s) * K + r] = classOf(repCompat + 3 * (size_t(s) * K + r));
  auto work = [&] (uint32_t lo, uint32_t hi) {
    for (uint32_t b = lo; b < hi; b++) {
      uint32_t best = INVALID, first = INVALID, count = 0;
      float bestScore = -std::numeric_limits<float>::infinity();
      const uint32_t *g = births + size_t(b) * words;
      const float *gc = birthCompat + 3 * size_t(b);
      for (uint32_t s = 0; s < nSpecies; s++) {
        const uint32_t k = std::min(fill[s], K);
        if (!k) continue;
        float score;
        if (kind == 0) {
          float avg = 0;
          uint32_t matable = 0;
          for (uint32_t i = 0; i < k; i++) {
            const double d = distanceOf(kind, words, g, reps + (size_t(s) * K + i) * words);
            const double c = std::min(tables[bcls[b]][uint32_t(d)], tables[rcls[size_t(s) * K + i]][uint32_t(d)]);
            if (fullContinuous) avg += c; else if (c >= compatThr) matable++;
          }
          score = fullContinuous ? float(avg / float(k) - avgThr) : float(matable - simThr * k);
        } else
          score = apto_score(kind, words, k, fullContinuous, compatThr, simThr, avgThr, g, gc,
                             reps + size_t(s) * K * words, repCompat + 3 * size_t(s) * K, nullptr);
        if (best == INVALID || bestScore < score) { best = s; bestScore = score; }
        if (score > 0) { count++; if (first == INVALID) first = s; }
      }
      outBest[b] = best;  outBestScore[b] = bestScore;  outFirst[b] = first;  outCount[b] = count;
    }
  };
  const uint32_t T = std::max(1u, std::min(threads, nBirths));
  std::vector<std::thread> pool;
  const uint32_t per = (nBirths + T - 1) / T;
  for (uint32_t t = 1; t < T; t++) pool.emplace_back(work, std::min(nBirths, t * per), std::min(nBirths, (t + 1) * per));
  work(0, std::min(nBirths, per));
  for (auto &th: pool) th.join();
}

/// Enveloppe criterion (enveloppecriteria.cpp:275-289) on a full k x k symmetric matrix
/// (row-major, diagonal ignored). Returns better; than/value through the out pointers.
int apto_env_contribution (int crit, uint32_t k, const float *edist, const float *gdist,
                           uint32_t *than, float *value) {
  DistMat e (k);
  for (uint32_t i = 0; i < k; i++)
    for (uint32_t j = i + 1; j < k; j++) e.set(i, j, edist[i * k + j]);
  std::vector<float> g (gdist, gdist + k);
  EnveloppeContribution ec = computeContribution(crit, e, g, k);
  *than = ec.than;  *value = ec.value;
  return ec.better;
}

} // extern "C"
