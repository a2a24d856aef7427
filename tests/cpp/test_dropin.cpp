// Test program for the C++ drop-in (include/apt/phylogenetictree.hpp): replays a birth trace the way
// a host simulation written against the reference's API would -- user genome type, user data,
// specialised Callbacks_t -- and writes the same .res / .json artefacts as oracle/ref_driver.cpp so
// the Python tests can diff them against the golden files produced by the unmodified reference.
//
// The SAME source compiles against the reference's own header: with -DAPT_TEST_REFERENCE it includes
// src/core/tree/phylogenetictree.hpp of /root/reference (over the Tools stand-in of oracle/shim) instead of ours.
// What differs between the two builds is confined to the three APT_TEST_REFERENCE blocks below: the include, how a
// configuration value is set (the reference reads a config file through Tools), and the batch entry point + device
// encoding, which the reference does not have.
#ifdef APT_TEST_REFERENCE
#include "core/tree/phylogenetictree.hpp"
#define SET_CONFIG(NAME, VALUE) config::PTree::NAME##_mut() = VALUE
#else
#define APT_WITH_JSON
#include "apt/phylogenetictree.hpp"
#define SET_CONFIG(NAME, VALUE) config::PTree::NAME() = VALUE
#endif

#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <utility>

#include "../../oracle/trace.h"     // trace reader only (test infrastructure)

using namespace phylogeny;
using json = nlohmann::json;

/// The host program's genome: payload + the three BOC compatibility genes, with the operators the reference's
/// template contract asks for (distance, compatibility, genealogy; src/core/tree/phylogenetictree.hpp:441-442, :276)
template <uint32_t W, int KIND>
struct Tagged {
  Genealogy gen;
  std::array<uint32_t, W> w;
  float mu = 0, sin = 1, sout = 1;
  Tagged (void) : gen(), w() {}
  const Genealogy& genealogy (void) const { return gen; }
  double compatibility (double d) const {                              // crossover.h:89-92
    const double s = d < mu ? sin : sout;
    return std::exp(-((d - mu) * (d - mu)) / (2. * s * s));
  }
  friend double distance (const Tagged &a, const Tagged &b) {
    if (KIND == 1) {
      double s = 0;
      for (uint32_t i = 0; i < W; i++) {
        float fa, fb;
        memcpy(&fa, &a.w[i], 4);  memcpy(&fb, &b.w[i], 4);
        const double t = double(fa) - double(fb);
        s += t * t;
      }
      return std::sqrt(s);
    }
    uint32_t s = 0;
    for (uint32_t i = 0; i < W; i++) s += __builtin_popcount(a.w[i] ^ b.w[i]);
    return double(s);
  }
  friend void to_json (json &j, const Tagged &g) { j["g"] = g.gen; j["c"] = {g.mu, g.sin, g.sout}; j["w"] = g.w; }
  friend void from_json (const json &j, Tagged &g) {
    g.gen = j["g"].get<Genealogy>();
    g.mu = j["c"][0];  g.sin = j["c"][1];  g.sout = j["c"][2];
    g.w = j["w"].get<std::array<uint32_t, W>>();
  }
  friend void assertEqual (const Tagged &a, const Tagged &b, bool) {
    if (!(a.gen.self == b.gen.self) || !(a.gen.mother == b.gen.mother) || !(a.gen.father == b.gen.father) || a.w != b.w
        || a.mu != b.mu || a.sin != b.sin || a.sout != b.sout) throw std::logic_error("genomes differ");
  }
};

#ifndef APT_TEST_REFERENCE
namespace phylogeny {
template <uint32_t W, int KIND>
struct DeviceEncoding<Tagged<W, KIND>> {
  static constexpr apt_kind kind = apt_kind(KIND);
  static constexpr uint32_t words = W;
  static void encode (const Tagged<W, KIND> &g, uint32_t *payload, float compat[3]) {
    memcpy(payload, g.w.data(), W * 4);
    compat[0] = g.mu;  compat[1] = g.sin;  compat[2] = g.sout;
  }
};
}
#endif

struct UD {
  uint32_t gid = uint32_t(-1);
  int removed = 0;
  UD (void) {}
  UD (GID g) : gid(uint32_t(g)) {}
  void removedFromEnveloppe (void) { removed++; }
  friend void to_json (json &j, const UD &u) { j = u.gid; }
  friend void from_json (const json &j, UD &u) { u.gid = j.get<uint32_t>(); }
  friend void assertEqual (const UD &a, const UD &b, bool) { if (a.gid != b.gid) throw std::logic_error("user data differ"); }
};

struct Event { uint32_t type, a, b, c; };
static std::vector<Event> g_events;

namespace phylogeny {
template <uint32_t W, int KIND>
struct Callbacks_t<PhylogeneticTree<Tagged<W, KIND>, UD>> {
  void onStepped (uint step, const LivingSet &living) { g_events.push_back({0, step, uint32_t(living.size()), 0}); }
  void onNewSpecies (SID pid, SID sid) { g_events.push_back({1, uint32_t(pid), uint32_t(sid), 0}); }
  void onGenomeEntersEnveloppe (SID sid, GID gid) { g_events.push_back({2, uint32_t(sid), uint32_t(gid), 0}); }
  void onGenomeLeavesEnveloppe (SID sid, GID gid) { g_events.push_back({3, uint32_t(sid), uint32_t(gid), 0}); }
  void onMajorContributorChanged (SID sid, SID o, SID n) { g_events.push_back({4, uint32_t(sid), uint32_t(o), uint32_t(n)}); }
};
}

template <uint32_t W, int KIND>
int run (const apt_trace::Trace &t, const std::string &prefix, bool oneByOne) {
  using G = Tagged<W, KIND>;
  using PT = PhylogeneticTree<G, UD>;
  struct Exposed : PT { using PT::logTo; };        // logTo is protected in the reference (:872-879)
  Exposed tree;
  typename PT::Callbacks cb;
  tree.setCallbacks(&cb);

  const uint32_t B = t.h.n_births;
  std::vector<uint32_t> sidOf (B, uint32_t(-1));
  std::vector<uint8_t> entered (B, 0);
  auto make = [&] (uint32_t b) {
    G g;
    memcpy(g.w.data(), &t.payload[size_t(b) * W], W * 4);
    g.mu = t.compat[3*b];  g.sin = t.compat[3*b+1];  g.sout = t.compat[3*b+2];
    g.gen.self = PID(GID(b + 1));  g.gen.self.sid = SID(sidOf[b]);
    const uint32_t m = t.mother[b], f = t.father[b];
    if (m != apt_trace::NONE) { g.gen.mother = PID(GID(m + 1)); g.gen.mother.sid = SID(sidOf[m]); }
    if (f != apt_trace::NONE) { g.gen.father = PID(GID(f + 1)); g.gen.father.sid = SID(sidOf[f]); }
    return g;
  };
  int bad = 0;
  std::vector<double> addMs;          // time inside addGenome(s) per generation: encode + engine + callback replay
  for (uint32_t gen = 0; gen < t.h.n_gens; gen++) {
    const uint32_t b0 = t.gen_birth_off[gen], b1 = t.gen_birth_off[gen+1];
    std::vector<G> batch;
    for (uint32_t b = b0; b < b1; b++) batch.push_back(make(b));
    std::vector<typename PT::InsertionResult> res (b1 - b0);
    const auto c0 = std::chrono::steady_clock::now();
#ifndef APT_TEST_REFERENCE
    if (!oneByOne) tree.addGenomes(batch.begin(), batch.end(), res.data());
    else
#endif
    for (uint32_t b = b0; b < b1; b++) res[b - b0] = tree.addGenome(batch[b - b0]);
    addMs.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - c0).count());
    for (uint32_t b = b0; b < b1; b++) {
      sidOf[b] = uint32_t(res[b - b0].sid);
      entered[b] = res[b - b0].udata != nullptr;
      // InsertionResult.udata must be the enveloppe slot's own object (:665-725)
      PID pid (GID(b + 1));  pid.sid = res[b - b0].sid;
      UD *ud = tree.getUserData(pid);
      if (b + 1 == b1 && entered[b] && ud != res[b - b0].udata) bad++;
    }
    for (uint32_t i = t.gen_death_off[gen]; i < t.gen_death_off[gen+1]; i++) tree.delGenome(make(t.deaths[i]));
    const uint32_t *a0 = t.alive.data() + t.gen_alive_off[gen], *a1 = t.alive.data() + t.gen_alive_off[gen+1];
    tree.step(gen + 1, a0, a1, [&] (uint32_t b) { return SID(sidOf[b]); });
  }
  const auto &st = tree.stats();
  const PT &ctree = tree;       // (the reference's non-const nodeAt / nextNodeID overloads are protected, :206-208, :225-230)
  { std::ofstream ofs (prefix + ".json");  tree.saveTo(ofs, -1); }
  {
    // saveTo -> readFrom -> saveTo: must equal what the reference's own round trip produces
    PT copy = PT::readFrom(prefix + ".json");
    std::ofstream ofs (prefix + ".rt.json");
    copy.saveTo(ofs, -1);
    if (copy.width() != tree.width() || std::as_const(copy).nextNodeID() != ctree.nextNodeID()) bad++;
  }
  {
    // node handles (node.hpp:92-164): root()->..., nodeAt(sid)->..., children() / parent() walks
    uint32_t reached = 0, entries = 0;
    std::vector<const typename PT::Node*> todo { ctree.root().get() };
    while (!todo.empty()) {
      const typename PT::Node *n = todo.back();  todo.pop_back();
      reached++;
      entries += uint32_t(n->rset.size());
      if (ctree.nodeAt(n->id())->data.count != n->data.count) bad++;
      if (n->distances.size() != n->rset.size() * (n->rset.size() - 1) / 2) bad++;
      for (const auto &c: n->children()) {
        if (c->parent() != n) bad++;
        todo.push_back(c.get());
      }
    }
    // copy construction (:98-109) keeps what the root reaches and nothing else; assignment (:112-115) goes through it
    PT copy (tree);
    PT assigned;
    assigned = tree;
    if (copy.width() != reached || assigned.width() != reached) bad++;
    if (!copy.aliveSpecies().empty()) bad++;
    int deepChecked = 0;
    if (reached == tree.width()) {
      deepChecked = 1;
      // every species hangs below the root: the copy must compare equal, as a DEEP copy -- up to the alive set, which
      // the copy constructor drops (:98-109)
      PT twin (tree);
      try {
        const auto &stp = [&] (uint32_t b) { return SID(sidOf[b]); };
        const uint32_t last = t.h.n_gens - 1;
        twin.step(tree.step(), t.alive.data() + t.gen_alive_off[last], t.alive.data() + t.gen_alive_off[last + 1], stp);
        assertEqual(tree, twin, true);
      } catch (const std::logic_error &e) { fprintf(stderr, "copy differs: %s\n", e.what()); bad++; }
    }
    tree.logTo(prefix + ".dot");
    FILE *fs = fopen((prefix + ".walk").c_str(), "w");
    if (fs) { fprintf(fs, "%u %u %u %d\n", reached, entries, copy.width(), deepChecked); fclose(fs); }
  }
  FILE *f = fopen((prefix + ".res").c_str(), "wb");
  if (!f) return 2;
  uint32_t n = B, ne = uint32_t(g_events.size());
  uint32_t tail[8] = { st.insertions, st.deletions, st.comparisons, st.branching, uint32_t(ctree.nextNodeID()),
                       tree.width(), tree.step(), 0 };
  fwrite(&n, 4, 1, f);  fwrite(sidOf.data(), 4, n, f);  fwrite(entered.data(), 1, n, f);
  fwrite(&ne, 4, 1, f);  fwrite(g_events.data(), sizeof(Event), ne, f);  fwrite(tail, 4, 8, f);
  fclose(f);
  // error contract: unknown species -> std::invalid_argument, like utils::Thrower (:169-174)
  try { ctree.nodeAt(SID(0x7FFFFFFF)); bad++; } catch (const std::invalid_argument &) {}
  printf("{\"births\": %u, \"species\": %u, \"bad\": %d, \"add_ms_per_generation\": [", B, uint32_t(ctree.nextNodeID()), bad);
  for (size_t i = 0; i < addMs.size(); i++) printf("%s%.3f", i ? ", " : "", addMs[i]);
  printf("]}\n");
  return bad ? 3 : 0;
}

int main (int argc, char **argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s <trace> <out-prefix> <rset> [--one-by-one] [--crit N] [--semi]\n", argv[0]); return 2; }
  bool one = false;
  SET_CONFIG(rsetSize, std::stoul(argv[3]));
  for (int i = 4; i < argc; i++) {
    std::string a = argv[i];
    if (a == "--one-by-one") one = true;
    else if (a == "--semi") SET_CONFIG(DEBUG_FULL_CONTINUOUS, false);
    else if (a == "--crit" && i + 1 < argc) SET_CONFIG(DEBUG_ENV_CRIT, std::stoi(argv[++i]));
  }
  try {
    apt_trace::Trace t = apt_trace::load(argv[1]);
    if (t.h.kind == apt_trace::BITS && t.h.words == 8) return run<8, 0>(t, argv[2], one);
    if (t.h.kind == apt_trace::BITS && t.h.words == 128) return run<128, 0>(t, argv[2], one);
    if (t.h.kind == apt_trace::F32 && t.h.words == 64) return run<64, 1>(t, argv[2], one);
    fprintf(stderr, "unsupported trace shape\n");
    return 2;
  } catch (const std::exception &e) {
    fprintf(stderr, "drop-in test failed: %s\n", e.what());
    return 1;
  }
}
