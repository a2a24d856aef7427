// TEST INFRASTRUCTURE -- replays a birth trace through the UNMODIFIED reference tree.
//
// This translation unit includes the reference's own headers where they lie under
// /root/reference (phylogenetictree.hpp and friends) and is linked with the reference's own
// ptreeconfig.cpp, treetypes.cpp, enveloppecriteria.cpp and speciescontributors.cpp by
// oracle/Makefile; the only non-reference code is the Tools stand-in under oracle/shim/ and the
// toy genome below (the reference ships no genome of its own: src/tests/basicviewer.cpp:11-20 is
// an empty decoy). Output goes to oracle/_ref/ only. It is the parity oracle ("reference" kind)
// and the CPU baseline; nothing on the product path may execute it.
//
// With -DAPT_REF_CREF_PATCH the eligibility binder of node.hpp:174-178 captures the node map by
// reference instead of by value ("biggest surprise"): same results, no
// O(#species) copy per insertion. Used for timing only so speed-ups are not inflated.

#include <array>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <functional>

#ifdef APT_REF_CREF_PATCH
namespace std {
template <typename F, typename A, typename B, typename C>
auto apt_bind_cref_last (F f, A a, B b, const C &c) { return std::bind(f, a, b, std::cref(c)); }
}
#define bind(...) apt_bind_cref_last(__VA_ARGS__)
#endif

#include "core/tree/phylogenetictree.hpp"

#ifdef APT_REF_CREF_PATCH
#undef bind
#endif

#include "trace.h"

using phylogeny::GID;
using phylogeny::SID;
using phylogeny::PID;
using phylogeny::Genealogy;
using json = nlohmann::json;

// ---------------------------------------------------------------------------------------------
// Toy genome: fixed-length payload (bit-string with Hamming distance, or float vector with L2)
// plus the three BOC compatibility genes of crossover.h:43-49; compatibility(d) is the two-sided
// gaussian of crossover.h:89-92.
template <uint32_t W, bool IS_F32>
struct Toy {
  Genealogy gen;
  std::array<uint32_t, W> w;
  float mu, sin, sout;

  Toy (void) : gen(), w(), mu(0), sin(1), sout(1) {}

  const Genealogy& genealogy (void) const { return gen; }
  Genealogy& genealogy (void) { return gen; }

  double compatibility (double d) const {
    return utils::gauss(d, mu, d < mu ? sin : sout);
  }

  friend double distance (const Toy &a, const Toy &b) {
    if constexpr (IS_F32) {
      double s = 0;
      for (uint32_t i = 0; i < W; i++) {
        float fa, fb;
        memcpy(&fa, &a.w[i], 4);
        memcpy(&fb, &b.w[i], 4);
        double t = double(fa) - double(fb);
        s += t * t;
      }
      return std::sqrt(s);
    } else {
      uint32_t s = 0;
      for (uint32_t i = 0; i < W; i++) s += __builtin_popcount(a.w[i] ^ b.w[i]);
      return double(s);
    }
  }

  friend void to_json (json &j, const Toy &g) {
    j["g"] = g.gen;
    j["c"] = {g.mu, g.sin, g.sout};
    j["w"] = g.w;
  }

  friend void from_json (const json &j, Toy &g) {
    g.gen = j["g"];
    g.mu = j["c"][0];  g.sin = j["c"][1];  g.sout = j["c"][2];
    g.w = j["w"];
  }

  friend void assertEqual (const Toy &a, const Toy &b, bool) {
    if (!(a.gen == b.gen) || a.w != b.w || a.mu != b.mu || a.sin != b.sin || a.sout != b.sout)
      throw std::logic_error("genomes differ");
  }
};

/// User data recording which genome id the slot object was (re)built from
/// (phylogenetictree.hpp:708-710 rebuilds it from the OUTGOING genome's id).
struct UD {
  uint32_t gid = uint32_t(-1);
  UD (void) {}
  UD (GID g) : gid(uint32_t(g)) {}
  void removedFromEnveloppe (void) {}
  friend void to_json (json &j, const UD &u) { j = u.gid; }
  friend void from_json (const json &j, UD &u) { u.gid = j; }
  friend void assertEqual (const UD &a, const UD &b, bool) {
    if (a.gid != b.gid) throw std::logic_error("user data differ");
  }
};

struct Event { uint32_t type, a, b, c; };
static std::vector<Event> g_events;
static bool g_logEvents = true;

namespace phylogeny {
template <uint32_t W, bool F>
struct Callbacks_t<PhylogeneticTree<Toy<W,F>, UD>> {
  void onStepped (uint step, const LivingSet &living) {
    if (g_logEvents) g_events.push_back({0, step, uint32_t(living.size()), 0});
  }
  void onNewSpecies (SID pid, SID sid) {
    if (g_logEvents) g_events.push_back({1, uint32_t(pid), uint32_t(sid), 0});
  }
  void onGenomeEntersEnveloppe (SID sid, GID gid) {
    if (g_logEvents) g_events.push_back({2, uint32_t(sid), uint32_t(gid), 0});
  }
  void onGenomeLeavesEnveloppe (SID sid, GID gid) {
    if (g_logEvents) g_events.push_back({3, uint32_t(sid), uint32_t(gid), 0});
  }
  void onMajorContributorChanged (SID sid, SID oldMC, SID newMC) {
    if (g_logEvents) g_events.push_back({4, uint32_t(sid), uint32_t(oldMC), uint32_t(newMC)});
  }
};
}

struct Options {
  std::string trace, prefix;
  bool writeJson = true, writeRes = true, roundtrip = false;
  uint32_t maxBirths = 0;   // 0 = all; otherwise stop after the generation that crosses it
};

template <uint32_t W, bool IS_F32>
int run (const apt_trace::Trace &t, const Options &o) {
  using G = Toy<W, IS_F32>;
  using PT = phylogeny::PhylogeneticTree<G, UD>;
  using CB = typename PT::Callbacks;

  const uint32_t B = t.h.n_births;
  std::vector<uint32_t> sidOf (B, uint32_t(-1));
  std::vector<uint8_t> entered (B, 0);

  PT tree;
  CB callbacks;
  tree.setCallbacks(&callbacks);

  auto makeGenome = [&] (uint32_t b) {
    G g;
    memcpy(g.w.data(), &t.payload[size_t(b) * W], W * 4);
    g.mu = t.compat[3*b];  g.sin = t.compat[3*b+1];  g.sout = t.compat[3*b+2];
    g.gen.self = PID(GID(b + 1));
    g.gen.self.sid = SID(sidOf[b]);
    g.gen.generation = 0;
    uint32_t m = t.mother[b], f = t.father[b];
    if (m != apt_trace::NONE) { g.gen.mother = PID(GID(m + 1));  g.gen.mother.sid = SID(sidOf[m]); }
    if (f != apt_trace::NONE) { g.gen.father = PID(GID(f + 1));  g.gen.father.sid = SID(sidOf[f]); }
    return g;
  };

  double seconds = 0;
  std::vector<double> genSeconds;
  uint32_t done = 0, gens = 0;
  std::vector<G> batch;
  for (uint32_t gen = 0; gen < t.h.n_gens; gen++) {
    const uint32_t b0 = t.gen_birth_off[gen], b1 = t.gen_birth_off[gen+1];
    batch.clear();
    for (uint32_t b = b0; b < b1; b++) batch.push_back(makeGenome(b));

    auto start = std::chrono::steady_clock::now();
    for (uint32_t b = b0; b < b1; b++) {
      auto r = tree.addGenome(batch[b - b0]);
      sidOf[b] = uint32_t(r.sid);
      entered[b] = (r.udata != nullptr);
    }
    genSeconds.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count());
    seconds += genSeconds.back();

    for (uint32_t i = t.gen_death_off[gen]; i < t.gen_death_off[gen+1]; i++)
      tree.delGenome(makeGenome(t.deaths[i]));

    const uint32_t *a0 = t.alive.data() + t.gen_alive_off[gen],
                   *a1 = t.alive.data() + t.gen_alive_off[gen+1];
    tree.step(gen + 1, a0, a1, [&] (uint32_t b) { return SID(sidOf[b]); });

    done = b1;  gens = gen + 1;
    if (o.maxBirths && done >= o.maxBirths) break;
  }

  const auto &st = tree.stats();

  if (o.writeJson) {
    std::ofstream ofs (o.prefix + ".json");
    tree.saveTo(ofs, -1);
  }

  bool roundtripOk = true;
  if (o.roundtrip) {
    // toJson -> fromJson -> toJson must be a fixed point (phylogenetictree.hpp:934-967)
    json j0, j1;
    PT::toJson(j0, tree);
    PT copy;
    PT::fromJson(j0, copy);
    PT::toJson(j1, copy);
    roundtripOk = (j0 == j1);
    std::ofstream ofs (o.prefix + ".rt.json");      // what the reference's own fromJson makes of its dump
    ofs << j1.dump(-1);
  }

  if (o.writeRes) {
    FILE *f = fopen((o.prefix + ".res").c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot write %s.res\n", o.prefix.c_str()); return 2; }
    uint32_t n = done, ne = uint32_t(g_events.size());
    uint32_t tail[8] = { st.insertions, st.deletions, st.comparisons, st.branching,
                         uint32_t(std::as_const(tree).nextNodeID()), tree.width(), tree.step(), 0 };
    fwrite(&n, 4, 1, f);
    fwrite(sidOf.data(), 4, n, f);
    fwrite(entered.data(), 1, n, f);
    fwrite(&ne, 4, 1, f);
    fwrite(g_events.data(), sizeof(Event), ne, f);
    fwrite(tail, 4, 8, f);
    fclose(f);
  }

  printf("{\"impl\": \"%s\", \"births\": %u, \"gens\": %u, \"seconds\": %.6f, "
         "\"births_per_s\": %.1f, \"evals_per_s\": %.1f, \"insertions\": %u, \"deletions\": %u, "
         "\"comparisons\": %u, \"branching\": %u, \"next_sid\": %u, \"width\": %u, "
         "\"roundtrip_ok\": %s, \"gen_seconds\": [",
#ifdef APT_REF_CREF_PATCH
         "reference+cref",
#else
         "reference",
#endif
         done, gens, seconds, done / seconds, st.comparisons / seconds,
         st.insertions, st.deletions, st.comparisons, st.branching,
         uint32_t(std::as_const(tree).nextNodeID()), tree.width(), roundtripOk ? "true" : "false");
  for (size_t i = 0; i < genSeconds.size(); i++) printf("%s%.6f", i ? ", " : "", genSeconds[i]);
  printf("]}\n");
  return 0;   // roundtrip_ok is reported, not enforced: re-parented children may be re-ordered by fromJson
}

template <bool IS_F32>
int dispatch (const apt_trace::Trace &t, const Options &o) {
  switch (t.h.words) {
  case 1:   return run<1, IS_F32>(t, o);
  case 2:   return run<2, IS_F32>(t, o);
  case 3:   return run<3, IS_F32>(t, o);
  case 4:   return run<4, IS_F32>(t, o);
  case 5:   return run<5, IS_F32>(t, o);
  case 8:   return run<8, IS_F32>(t, o);
  case 16:  return run<16, IS_F32>(t, o);
  case 32:  return run<32, IS_F32>(t, o);
  case 64:  return run<64, IS_F32>(t, o);
  case 128: return run<128, IS_F32>(t, o);
  default:
    fprintf(stderr, "unsupported payload width %u words\n", t.h.words);
    return 2;
  }
}

int main (int argc, char **argv) {
  using PTC = config::PTree;
  Options o;
  std::vector<std::string> pos;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto next = [&] { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); }
                      return std::string(argv[++i]); };
    if (a == "--rset") PTC::rsetSize_mut() = std::stoul(next());
    else if (a == "--compat-thr") PTC::compatibilityThreshold_mut() = std::stod(next());
    else if (a == "--sim-thr") PTC::similarityThreshold_mut() = std::stod(next());
    else if (a == "--avg-thr") PTC::avgCompatibilityThreshold_mut() = std::stod(next());
    else if (a == "--semi") PTC::DEBUG_FULL_CONTINUOUS_mut() = false;
    else if (a == "--crit") PTC::DEBUG_ENV_CRIT_mut() = std::stoi(next());
    else if (a == "--trim-period") PTC::stillbornTrimmingPeriod_mut() = std::stoul(next());
    else if (a == "--trim-thr") PTC::stillbornTrimmingThreshold_mut() = std::stof(next());
    else if (a == "--trim-delay") PTC::stillbornTrimmingDelay_mut() = std::stof(next());
    else if (a == "--trim-min-delay") PTC::stillbornTrimmingMinDelay_mut() = std::stoul(next());
    else if (a == "--no-json") o.writeJson = false;
    else if (a == "--no-res") o.writeRes = false;
    else if (a == "--no-events") g_logEvents = false;
    else if (a == "--roundtrip") o.roundtrip = true;
    else if (a == "--max-births") o.maxBirths = std::stoul(next());
    else pos.push_back(a);
  }
  if (pos.size() != 2) {
    fprintf(stderr, "usage: %s [options] <trace> <out-prefix>\n", argv[0]);
    return 2;
  }
  o.trace = pos[0];  o.prefix = pos[1];

  try {
    apt_trace::Trace t = apt_trace::load(o.trace);
    return t.h.kind == apt_trace::F32 ? dispatch<true>(t, o) : dispatch<false>(t, o);
  } catch (const std::exception &e) {
    fprintf(stderr, "reference driver failed: %s\n", e.what());
    return 1;
  }
}
