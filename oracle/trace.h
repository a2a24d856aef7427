// TEST INFRASTRUCTURE -- birth-trace file reader shared by the oracle programs.
//
// A trace is the recorded input of a host simulation: every birth (payload, compatibility
// parameters, parents), every death and the per-generation alive list. The simulation itself is
// not part of APOGeT ("host simulation (NOT in the reference)"); traces are
// produced by apoget_b200/synth.py and replayed identically through (a) the unmodified
// reference (ref_driver.cpp), (b) the CPU restatement (apt_oracle.cpp) and (c) the CUDA path.
//
// Layout (little endian): Header, then u32 gen_birth_off[G+1], gen_death_off[G+1],
// gen_alive_off[G+1], mother[B], father[B] (birth index or 0xFFFFFFFF), f32 compat[B*3]
// (mu, sigma_in, sigma_out), u32 payload[B*words], u32 deaths[D], u32 alive[A].
// GID of birth i is i+1 (the reference's GIDManager issues 1,2,3..., treetypes.h:41-48).
#ifndef APT_ORACLE_TRACE_H
#define APT_ORACLE_TRACE_H

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace apt_trace {

enum Kind : uint32_t { BITS = 0, F32 = 1 };
static constexpr uint32_t NONE = 0xFFFFFFFFu;

struct Header {
  char magic[8];        // "APTTRACE"
  uint32_t version;     // 1
  uint32_t kind;        // Kind
  uint32_t words;       // 32-bit words per payload
  uint32_t n_births;
  uint32_t n_gens;
  uint32_t n_deaths;
  uint32_t n_alive;
  uint32_t reserved;
};

struct Trace {
  Header h;
  std::vector<uint32_t> gen_birth_off, gen_death_off, gen_alive_off, mother, father;
  std::vector<float> compat;
  std::vector<uint32_t> payload, deaths, alive;
};

template <typename T>
inline void readVec (FILE *f, std::vector<T> &v, size_t n) {
  v.resize(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n) throw std::runtime_error("truncated trace");
}

inline Trace load (const std::string &path) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot open trace " + path);
  Trace t;
  if (fread(&t.h, sizeof(Header), 1, f) != 1 || memcmp(t.h.magic, "APTTRACE", 8) != 0
      || t.h.version != 1) {
    fclose(f);
    throw std::runtime_error("bad trace header in " + path);
  }
  const size_t B = t.h.n_births, G = t.h.n_gens;
  readVec(f, t.gen_birth_off, G + 1);
  readVec(f, t.gen_death_off, G + 1);
  readVec(f, t.gen_alive_off, G + 1);
  readVec(f, t.mother, B);
  readVec(f, t.father, B);
  readVec(f, t.compat, B * 3);
  readVec(f, t.payload, B * size_t(t.h.words));
  readVec(f, t.deaths, t.h.n_deaths);
  readVec(f, t.alive, t.h.n_alive);
  fclose(f);
  return t;
}

} // namespace apt_trace

#endif
