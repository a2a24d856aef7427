// resolve_streams_kernel -- the ordered resolution pass on the device.
//
// A "stream" is the ordered list of the births of one batch whose only top-level candidate is one
// given species (no second parent species).  For those births the reference's sequential
// addGenome (phylogenetictree.hpp:268-302, 553-650) reads and writes nothing but that species'
// enveloppe as long as the birth is accepted at top level (:612-613) -- so streams of different
// species are independent, and one CTA (NW warps) owns one stream:
//
//   * the enveloppe (representatives, compat metadata, K x K distance matrix, per-row minima)
//     lives in shared memory for the whole stream;
//   * NW*32 births are scored at once, one thread per birth (:456-478 / :430-452), each thread
//     keeping its k distances and compatibilities in shared memory;
//   * in-chunk dependencies are then resolved exactly: every thread decides under the current
//     state (accept :612; append :668-677 or criterion enveloppecriteria.cpp:87-133); the FIRST
//     thread (= earliest birth) whose decision changes the enveloppe commits -- all earlier ones
//     are final -- and the later ones re-measure only the slot that changed and re-sum their score
//     in the reference's order.  Repeat until the chunk is final.
//
// Births rejected at top level (score <= 0) do not touch the enveloppe; they are flagged and the
// host walks the sub-species for them (:623).  When a decision cannot be taken with
// reference-identical arithmetic (device exp() inside the guard band) the warp stops and hands the
// rest of the stream to the host.  At the end the warp writes the enveloppe back to the pool.
#include "kernels.cuh"
#include "criteria.h"

#include <cooperative_groups.h>
#include <math_constants.h>
#include <algorithm>
#include <cstdlib>

namespace apt {

namespace {

constexpr float FMAX = 3.402823466e+38f;  // std::numeric_limits<float>::max()

__device__ __forceinline__ uint32_t popc4 (uint4 a, uint4 b) {
  return __popc(a.x ^ b.x) + __popc(a.y ^ b.y) + __popc(a.z ^ b.z) + __popc(a.w ^ b.w);
}

__device__ __forceinline__ double sqdiff4 (uint4 a, uint4 b) {
  double t0 = double(__uint_as_float(a.x)) - double(__uint_as_float(b.x));
  double t1 = double(__uint_as_float(a.y)) - double(__uint_as_float(b.y));
  double t2 = double(__uint_as_float(a.z)) - double(__uint_as_float(b.z));
  double t3 = double(__uint_as_float(a.w)) - double(__uint_as_float(b.w));
  return t0 * t0 + t1 * t1 + t2 * t2 + t3 * t3;
}

__device__ __forceinline__ double compat_generic (float4 m, double d) {     // crossover.h:89-92
  const double mu = double(m.x);
  const double s = d < mu ? double(m.y) : double(m.z);
  const double t = d - mu;
  return exp(-(t * t) / (2. * s * s));
}

template <bool LUT, class WS>
__device__ __forceinline__ double pair_compat (const ScoreCfg &cfg, const WS &ws, float4 bm, float4 rm, double d) {
  double a, b;
  if (LUT) {
    const uint32_t di = uint32_t(d);
    a = ws.lut(size_t(__float_as_uint(bm.w)) * cfg.lut_stride + di);
    b = ws.lut(size_t(__float_as_uint(rm.w)) * cfg.lut_stride + di);
  } else {
    a = compat_generic(bm, d);
    b = compat_generic(rm, d);
  }
  return b < a ? b : a;
}

/// Dynamic shared memory of the CTA.  Every access goes through this symbol (plus a byte offset), so the
/// compiler knows the address space: generic pointers kept in a struct made it re-derive the shared
/// window from SR_CgaCtaId in front of every access (57 S2R in the event loop, ~1.3 us per event).
extern __shared__ __align__(16) unsigned char apt_smem[];

/// Shared-window address of apt_smem, laundered through a shuffle so that it is a plain register value:
/// otherwise ptxas re-derives it (S2R SR_CgaCtaId + LEA) in front of accesses instead of keeping it live.
__device__ __forceinline__ uint32_t smem_window (void) {
  return __shfl_sync(0xffffffffu, uint32_t(__cvta_generic_to_shared(apt_smem)), 0);
}
/// Unconditional shared-memory accesses by window address (a load inside a C++ conditional is turned into a branch)
__device__ __forceinline__ float lds_f32 (uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

/// The stream's state in shared memory (T = threads of the CTA = births per step): byte offsets
struct WarpState {
  uint32_t oReps, oRmeta, oD, oMinBase, od, oc, oPre, oEv, oCtl, oLut;
  const double *lutG;   // compatibility table in global memory (used when it does not fit in shared memory)
  bool lutS;            // table staged in shared memory at oLut
  uint32_t T;
  uint32_t sb;          // shared-window address of apt_smem (smem_window)
  __device__ __forceinline__ uint4& reps (uint32_t i) const { return reinterpret_cast<uint4*>(apt_smem + oReps)[i]; }        // [K][WQ]
  __device__ __forceinline__ float4& rmeta (uint32_t i) const { return reinterpret_cast<float4*>(apt_smem + oRmeta)[i]; }    // [K]
  __device__ __forceinline__ float& D (uint32_t i) const { return reinterpret_cast<float*>(apt_smem + oD)[i]; }              // [K][K]
  __device__ __forceinline__ float& minBase (uint32_t i) const { return reinterpret_cast<float*>(apt_smem + oMinBase)[i]; }  // [K]
  __device__ __forceinline__ float& d (uint32_t i) const { return reinterpret_cast<float*>(apt_smem + od)[i]; }              // [K][T] distance of thread t to representative r at r*T + t
  __device__ __forceinline__ double& c (uint32_t i) const { return reinterpret_cast<double*>(apt_smem + oc)[i]; }            // [K][T] compatibilities
  __device__ __forceinline__ float& pre (uint32_t i) const { return reinterpret_cast<float*>(apt_smem + oPre)[i]; }          // [K+1][T] running average BEFORE representative r is added (pre[k] = the total)
  __device__ __forceinline__ uint32_t& ev (uint32_t i) const { return reinterpret_cast<uint32_t*>(apt_smem + oEv)[i]; }      // [T] per event of a step: thread << 8 | distances
  __device__ __forceinline__ uint32_t& ctl (uint32_t i) const { return reinterpret_cast<uint32_t*>(apt_smem + oCtl)[i]; }    // [64] exchange words
  __device__ __forceinline__ double lut (size_t i) const {
    return lutS ? reinterpret_cast<const double*>(apt_smem + oLut)[i] : __ldg(lutG + i);
  }
};

/// A thread's distances to the k representatives.  With the enveloppe size known at compile time (KT > 0)
/// they live in registers and every loop over them is unrolled -- the criterion on the per-event critical
/// path then has no shared-memory round trips in it; they are mirrored to shared memory because the OTHER
/// threads need the entering birth's row.  The compatibilities stay in shared memory: the decisions only
/// need their sum (decision_score), the exact chain reads them when a score is too close to call.
template <int KT>
struct Lane {
  float d[KT ? KT : 1];
};

#define APT_FOR_K(r, k) _Pragma("unroll") for (uint32_t r = 0; r < (KT ? uint32_t(KT) : (k)); r++) if (KT == 0 || r < (k))

/// distance + compatibility of this lane's birth to representative r
template <int WQ, int KIND, bool LUT>
__device__ __forceinline__ void measure_one (const ScoreCfg &cfg, const uint4 (&bw)[WQ], float4 bm,
                                             const WarpState &ws, uint32_t r, float &df, double &c) {
  if (KIND == 0 && LUT) {
    // Hamming distance stays an integer: it indexes the table directly (no double round trip)
    uint32_t acc = 0;
#pragma unroll
    for (int v = 0; v < WQ; v++) acc += popc4(bw[v], ws.reps(r * WQ + v));
    const float4 rm = ws.rmeta(r);
    const double ca = ws.lut(size_t(__float_as_uint(bm.w)) * cfg.lut_stride + acc);
    const double cb = ws.lut(size_t(__float_as_uint(rm.w)) * cfg.lut_stride + acc);
    df = float(acc);
    c = cb < ca ? cb : ca;
    return;
  }
  double d;
  if (KIND == 0) {
    uint32_t acc = 0;
#pragma unroll
    for (int v = 0; v < WQ; v++) acc += popc4(bw[v], ws.reps(r * WQ + v));
    d = double(acc);
  } else {
    double acc = 0.;
#pragma unroll
    for (int v = 0; v < WQ; v++) acc += sqdiff4(bw[v], ws.reps(r * WQ + v));
    d = sqrt(acc);
  }
  df = float(d);
  c = pair_compat<LUT>(cfg, ws, bm, ws.rmeta(r), d);
}

/// (re-)measures representative r (read from row src) and stores the result in the lane's column; returns by how much the lane's
/// compatibility to slot r changed (`replaces`: the slot held a representative before; else it is new)
template <int WQ, int KIND, bool LUT, int KT>
__device__ __forceinline__ double measure (const ScoreCfg &cfg, const uint4 (&bw)[WQ], float4 bm,
                                           const WarpState &ws, Lane<KT> &ln, uint32_t src, uint32_t r, int lane, bool replaces) {
  float df;  double c, old;
  measure_one<WQ, KIND, LUT>(cfg, bw, bm, ws, src, df, c);        // src: the row that holds slot r's new genome (staging row)
  ws.d(r * ws.T + lane) = df;
  if (KT > 0) {
#pragma unroll
    for (int q = 0; q < (KT ? KT : 1); q++) if (uint32_t(q) == r) ln.d[q] = df;
  }
  old = ws.c(r * ws.T + lane);
  ws.c(r * ws.T + lane) = c;
  return replaces ? c - old : c;
}


/// phylogenetictree.hpp:465-477 / :439-451 over this lane's cached compatibilities.  The float
/// accumulation is order sensitive, so it stays sequential -- but representatives before `from` are
/// unchanged since the last call and their partial sum is kept (pre[from]).
template <int KT>
__device__ __forceinline__ float resum (const ScoreCfg &cfg, const WarpState &ws, const Lane<KT> &ln, uint32_t k,
                                        int lane, uint32_t from) {
  if (cfg.full_continuous) {
    float avg = from ? ws.pre(from * ws.T + lane) : 0.f;
    APT_FOR_K(r, k) {
      if (r >= from) {
        ws.pre(r * ws.T + lane) = avg;
        avg = float(__dadd_rn(double(avg), ws.c(r * ws.T + lane)));
      }
    }
    ws.pre(k * ws.T + lane) = avg;
    return float(__dsub_rn(double(__fdiv_rn(avg, float(k))), cfg.avg_thr));
  }
  uint32_t matable = 0;
  APT_FOR_K(r, k) matable += (ws.c(r * ws.T + lane) >= cfg.compat_thr);
  return float(__dsub_rn(double(matable), __dmul_rn(cfg.sim_thr, double(k))));
}

/// The score as far as the decisions need it (its sign; in the guarded mode also whether it is inside the guard band).
/// Continuous mode: the reference's float accumulation (resum) is a chain of k dependent F2F / DADD / F2F triples on the
/// critical path of every event.  A plain double sum of the same compatibilities, sumC / k - threshold, differs from it
/// by less than 2^-24 * 1.001 * ((k + 1) / 2 + 1) -- every partial sum is <= k and picks up one float rounding, the
/// quotient one more -- so outside a band eight times that wide (plus the guard) the two agree in sign and in the guard
/// test, and the exact chain only runs for the rare birth inside it (from the first representative whose compatibility
/// changed since the chain last ran: dirtyFrom).  No score leaves this kernel, only the decisions.
template <int KT>
__device__ __forceinline__ float decision_score (const ScoreCfg &cfg, const WarpState &ws, const Lane<KT> &ln, uint32_t k,
                                                 int lane, double sumC, float guard, uint32_t &dirtyFrom) {
  if (cfg.full_continuous) {
    // (k is uniform; a full enveloppe of compile-time size needs no division)
    const double invK = (KT > 0 && k == uint32_t(KT)) ? 1. / double(KT ? KT : 1) : 1. / double(k);
    const double approx = __dsub_rn(__dmul_rn(sumC, invK), cfg.avg_thr);
    const float band = guard + float(k + 8u) * 2.4e-7f;            // 4 * 2^-24 * (k + 8) >= 8 x the error bound
    if (fabs(approx) > double(band)) return approx > 0. ? 1.f : -1.f;
    const float exact = resum<KT>(cfg, ws, ln, k, lane, dirtyFrom);
    dirtyFrom = k;
    return exact;
  }
  return resum<KT>(cfg, ws, ln, k, lane, 0);
}

/// maxMinDist (enveloppecriteria.cpp:87-133) for this lane's distances against a FULL enveloppe.
/// :117 widens a float difference to double and :125 compares it with the float maximum: the
/// comparison of the two floats is the same predicate, so it is taken in float.
template <int KT>
__device__ __forceinline__ bool criterion_maxMinDist (const WarpState &ws, const Lane<KT> &ln, uint32_t K, int lane,
                                                      uint32_t &than) {
  float m1 = FMAX, m2 = FMAX;
  uint32_t i1 = 0xFFFFFFFFu;
  APT_FOR_K(r, K) {
    const float x = KT ? ln.d[KT ? r : 0] : ws.d(r * ws.T + lane);
    if (x < m1) { m2 = m1; m1 = x; i1 = r; }
    else if (x < m2) m2 = x;
  }
  float value = -CUDART_INF_F;
  than = 0xFFFFFFFFu;
  APT_FOR_K(i, K) {
    const float minNew = (i == i1) ? m2 : m1;                    // min over j != i of gdist[j]
    const float c = __fadd_rn(-ws.minBase(i), minNew);           // :117
    if (value < c) { value = c; than = i; }                      // :125-128, first index wins ties
  }
  return value > 0.f;                                            // :131
}

/// The same criterion for bit-string genomes, where every distance is a small integer held exactly by a float: the
/// differences of :117 are exact, so max_i (minNew_i - minBase_i) is reached either at i1 (the representative nearest
/// to the birth, whose minNew is the second smallest distance m2) or at the row with the smallest minBase among the
/// others (minNew = m1 for all of them; smallest minBase, first index on ties, = first index of the largest
/// difference).  The two smallest (minBase, index) pairs of the enveloppe are found once per event by the warp that
/// updates the matrix (keys below) instead of by every thread: key = minBase << 5 | row, FLT_MAX as 0xFFFFFFE0 | row.
__device__ __forceinline__ uint32_t minbase_key (float m, uint32_t row) {
  return (m >= 16777216.f ? 0xFFFFFFE0u : (uint32_t(m) << 5)) | row;
}
template <int KT>
__device__ __forceinline__ bool criterion_maxMinDist_int (const WarpState &ws, const Lane<KT> &ln, uint32_t K, int lane,
                                                          uint32_t key1, uint32_t key2, uint32_t &than) {
  float m1 = FMAX, m2 = FMAX;
  uint32_t i1 = 0xFFFFFFFFu;
  APT_FOR_K(r, K) {
    const float x = KT ? ln.d[KT ? r : 0] : ws.d(r * ws.T + lane);
    if (x < m1) { m2 = m1; m1 = x; i1 = r; }
    else if (x < m2) m2 = x;
  }
  const uint32_t keyA = (key1 & 31u) == i1 ? key2 : key1;            // smallest minBase over the rows other than i1
  const uint32_t iA = keyA & 31u;
  const float cA = __fadd_rn(-float(keyA >> 5), m1);                 // :117 for every i != i1 collapses to this one
  const float cB = __fadd_rn(-ws.minBase(i1), m2);                   // ... and i1 itself
  // :125-128: the first index reaching the maximum wins
  const bool takeB = cB > cA || (cB == cA && i1 < iA);
  than = takeB ? i1 : iA;
  return (takeB ? cB : cA) > 0.f;                                    // :131
}

/// maxAverage (enveloppecriteria.cpp:40-84)
template <int KT>
__device__ __forceinline__ bool criterion_maxAverage (const WarpState &ws, const Lane<KT> &ln, uint32_t K, int lane,
                                                      uint32_t &than) {
  float value = -CUDART_INF_F;
  than = 0xFFFFFFFFu;
  for (uint32_t i = 0; i < K; i++) {
    double c = 0.;
    APT_FOR_K(j, K) {
      if (j != i)
        c = __dadd_rn(c, double(__fadd_rn(-ws.D(i * K + j), KT ? ln.d[KT ? j : 0] : ws.d(j * ws.T + lane))));   // :62
    }
    if (double(value) < c) { value = float(c); than = i; }
  }
  return value > 0.f;
}

/// Criteria 2 and 3 (enveloppecriteria.cpp:136-272, criteria.h) against the full enveloppe in shared memory.  Kept out
/// of line: they are debug options of the reference (DEBUG_ENV_CRIT) and must not cost the default criteria registers.
/// Returns than | better << 30 | uncertain << 31.
__device__ __noinline__ uint32_t criterion_other (int crit, uint32_t oD, uint32_t K, const float *gd) {
  const float *Dm = reinterpret_cast<const float*>(apt_smem + oD);
  auto D = [&] (uint32_t i, uint32_t j) { return Dm[i * K + j]; };
  auto g = [&] (uint32_t j) { return gd[j]; };
  const CritResult r = crit == 2 ? crit_maxAvgMinStdDev(D, g, K)
                                 : crit_maxWeightedDist2Avg<KMAX>(D, g, K, [] (double x) { return exp(x); });
  return (r.than & 0x3FFFFFFFu) | (r.better ? 0x40000000u : 0u) | (r.uncertain ? 0x80000000u : 0u);
}

/// Shared-memory bytes of one stream (without the table)
__host__ __device__ inline size_t stream_smem (uint32_t K, uint32_t wq, uint32_t T, bool ktConst = false) {
  (void) ktConst;
  return size_t(K) * T * 8 + size_t(K + T / 32) * wq * 16 + size_t(K + T / 32) * 16 + size_t(K) * K * 4 + size_t(K) * 4
       + size_t(K) * T * 4 + size_t(K + 1) * T * 4 + size_t(T) * 4 + 160 * 4;
}

__device__ __forceinline__ void prefetch_l2 (const void *p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

// Event word exchanged between the warps of a stream: earliest thread first (high bits), then what it does
constexpr uint32_t EV_TID_SHIFT = 8, EV_UNCERTAIN = 0x80u, EV_SLOT_MASK = 0x3Fu;

// 7 CTAs of 4 warps per SM keep 1 000 streams in one wave on 148 SMs (register-resident variant: <= 72 registers)
//
// CLUSTER: one very long stream (a species that holds all the births of a pass) spread over the CTAs of a thread-block
// cluster, as resolve_families_kernel does (family.cu): every CTA keeps its own copy of the enveloppe, a step decides
// T births per CTA (position in the step = CTA rank * T + thread), the earliest event is elected across the cluster
// through distributed shared memory, and its CTA pushes the new representative and the entering birth's distances
// into the other copies.
template <int WQ, int KIND, bool LUT, int NW, int KT, bool CRITX = false, bool CLUSTER = false>
__global__ void __launch_bounds__(NW * 32, (NW == 4 && KT > 0 && WQ <= 2) ? 7 : 1)
resolve_streams_kernel (StreamArgs a) {
  namespace cg = cooperative_groups;
  constexpr uint32_t T = NW * 32;
  constexpr uint32_t NONE = 0xFFFFFFFFu;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t crank = 0, csize = 1;
  if (CLUSTER) { cg::cluster_group cl = cg::this_cluster();  crank = cl.block_rank();  csize = cl.num_blocks(); }
  const uint32_t g = blockIdx.x / csize;
  const uint32_t gpos = crank * T + uint32_t(tid);            // this thread's position among the births of a step
  const uint32_t TS = T * csize;                               // births per step
  const uint32_t K = KT ? uint32_t(KT) : a.pool.K;
  uint32_t base = 0;
  WarpState ws;
  ws.T = T;
  ws.oc = base;        base += K * T * 8;
  ws.oReps = base;     base += (K + NW) * WQ * 16;        // rows K ...: one staging row per warp (the genome its candidate
  ws.oRmeta = base;    base += (K + NW) * 16;             // would put into the enveloppe, visible to all at the event barrier)
  ws.oD = base;        base += K * K * 4;
  ws.oMinBase = base;  base += K * 4;
  ws.od = base;        base += K * T * 4;
  ws.oPre = base;      base += (K + 1) * T * 4;
  ws.oEv = base;       base += T * 4;
  ws.oCtl = base;      base += 160 * 4;
  // ws.ctl(0..NW): event word per warp; ws.ctl(33): first output row of the step; ws.ctl(36..39): counters;
  // cluster: ws.ctl(64..95) the CTAs' events (two sets), ws.ctl(96..111) their rejected births, ws.ctl(112..143) the entering birth's distances
  ws.oLut = (base + 15u) & ~15u;
  ws.lutG = a.cfg.lut;
  ws.lutS = LUT && a.lut_in_smem;
  ws.sb = smem_window();

  // ---- everything the stream needs first is requested before anything is waited for: the
  //      stream's extent, then (one round trip later) the enveloppe, the first births' indices and
  //      the table, then (another one) those births' genomes ----
  // optional timeline of this stream (development aid, APT_STREAM_PROF): {globaltimer ns at start, SM clock at
  // start, after the enveloppe is loaded, after the first births are scored, at the end, events, steps, SM id,
  // cycles between barriers on the score/decide side, on the enveloppe-update side}
  unsigned long long *prof = (a.prof && crank == 0) ? a.prof + size_t(g) * 16 : nullptr;
  uint32_t profEvents = 0, profSteps = 0;
  long long profA = 0, profB = 0, profT = 0;     // cycles between barriers: score/decide side, enveloppe update side
  if (prof && tid == 0) {
    unsigned long long gt;  uint32_t smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    prof[0] = gt;  prof[1] = clock64();  prof[7] = smid;
  }
  if (a.span && tid == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    atomicMax(a.span, ~gt);
  }
  const uint32_t slot = a.grp_slot[g], sid = a.grp_sid[g];
  const uint32_t g0 = a.grp_off[g], g1 = a.grp_off[g + 1];
  const uint4 *poolRows = reinterpret_cast<const uint4*>(a.pool.payload) + size_t(slot) * K * WQ;
  const uint4 *birthRows = reinterpret_cast<const uint4*>(a.b_payload);
  uint32_t ordNext = (g0 + gpos < g1) ? __ldg(a.order + g0 + gpos) : 0u;
  uint32_t k = min(__ldg(a.pool.fill + slot), K);
  for (uint32_t i = tid; i < K * WQ; i += T) ws.reps(i) = poolRows[i];          // rows past the fill are never read
  for (uint32_t i = tid; i < K; i += T) ws.rmeta(i) = a.pool.cmeta[size_t(slot) * K + i];
  // the matrix is kept with FLT_MAX on the diagonal and in unused rows / columns: a row's minimum over its
  // other representatives (minBase, enveloppecriteria.cpp:105-113) is then a plain minimum over K entries
  for (uint32_t i = tid; i < K * K; i += T) {
    const float v = a.pool_dist[size_t(slot) * K * K + i];
    const uint32_t row = i / K, col = i - row * K;
    ws.D(i) = (row < k && col < k && row != col) ? v : FMAX;
  }
  if (ws.lutS) {
    double *lut_s = reinterpret_cast<double*>(apt_smem + ws.oLut);
    const uint32_t nd = a.cfg.lut_classes * a.cfg.lut_stride;
    for (uint32_t i = tid; i < nd; i += T) lut_s[i] = __ldg(a.cfg.lut + i);
  }
  if (tid < 4) ws.ctl(36 + tid) = 0;
  uint4 bw[WQ];
  float4 bm = make_float4(0.f, 1.f, 1.f, 0.f);
  if (g0 + gpos < g1) {
    const uint32_t b = ordNext & ORDER_INDEX_MASK;
#pragma unroll
    for (int v = 0; v < WQ; v++) bw[v] = __ldg(birthRows + size_t(b) * WQ + v);
    bm = __ldg(a.b_cmeta + b);
  } else {
#pragma unroll
    for (int v = 0; v < WQ; v++) bw[v] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  for (uint32_t r = tid; r < k; r += T) {
    float m = FMAX;
#pragma unroll (KT ? KT : 1)
    for (uint32_t j = 0; j < K; j++) m = fminf(m, ws.D(r * K + j));
    ws.minBase(r) = m;
  }
  __syncthreads();
  if (KIND == 0 && warp == 0) {                 // criterion_maxMinDist_int's keys (rows past k: FLT_MAX)
    const uint32_t key = (uint32_t(lane) < K) ? minbase_key(uint32_t(lane) < k ? ws.minBase(lane) : FMAX, lane) : 0xFFFFFFFFu;
    const uint32_t k1 = __reduce_min_sync(0xffffffffu, key);
    const uint32_t k2 = __reduce_min_sync(0xffffffffu, key == k1 ? 0xFFFFFFFFu : key);
    if (lane == 0) { ws.ctl(34) = k1;  ws.ctl(35) = k2; }
  }
  __syncthreads();
  if (prof && tid == 0) prof[2] = clock64();
  Lane<KT> ln;
  const float bandGuard = LUT ? 0.f : a.guard;      // decision_score: the exact table needs no guard band

  uint32_t changed = 0;         // slots to write back (uniform)
  uint32_t nRej = 0;            // births rejected at top level so far (uniform)
  bool aborted = false;
  uint32_t nAcc0 = 0, nAcc1 = 0, nAcc2 = 0;     // per thread: accepted, enveloppe unchanged, by contribution class

  uint32_t elect = 0;           // which set of election words the next event uses (cluster)
  for (uint32_t cb = g0; cb < g1 && !aborted; cb += TS) {
    const uint32_t nb = min(TS, g1 - cb);
    const bool active = gpos < nb;
    const uint32_t ord = ordNext;
    const uint32_t b = ord & ORDER_INDEX_MASK, ccls = ord >> ORDER_CLASS_SHIFT;
    if (cb != g0) {
      if (active) {
#pragma unroll
        for (int v = 0; v < WQ; v++) bw[v] = __ldg(birthRows + size_t(b) * WQ + v);
        bm = __ldg(a.b_cmeta + b);
      } else {
#pragma unroll
        for (int v = 0; v < WQ; v++) bw[v] = make_uint4(0, 0, 0, 0);
      }
    }
    // the next step's births start their way up from HBM while this step is resolved
    ordNext = 0u;
    if (cb + TS + gpos < g1) {
      ordNext = __ldg(a.order + cb + TS + gpos);
      const char *nx = reinterpret_cast<const char*>(birthRows + size_t(ordNext & ORDER_INDEX_MASK) * WQ);
#pragma unroll
      for (int o = 0; o < WQ * 16; o += 128) prefetch_l2(nx + o);
      prefetch_l2(a.b_cmeta + (ordNext & ORDER_INDEX_MASK));
    }
    float score = 0.f;
    double sumC = 0.;               // sum of this birth's compatibilities (decision_score)
    uint32_t dirtyFrom = 0;         // the exact chain's partial sums are valid up to here
    if (active) {
      APT_FOR_K(r, k) {
        float df;  double c;
        measure_one<WQ, KIND, LUT>(a.cfg, bw, bm, ws, r, df, c);
        ws.d(r * T + tid) = df;
        if (KT > 0) ln.d[KT ? r : 0] = df;
        ws.c(r * T + tid) = c;
        sumC += c;
      }
      score = decision_score<KT>(a.cfg, ws, ln, k, tid, sumC, bandGuard, dirtyFrom);
    }
    if (prof && tid == 0 && cb == g0) prof[3] = clock64();
    profSteps++;
    bool mine = active;           // not final yet
    if (prof) profT = clock64();
    // Nothing leaves the chip inside the event loop: scattered global stores ahead of the shared-memory
    // traffic of the next event cost ~1.3 us per event (measured with APT_STREAM_PROF; the LDS queue up
    // behind them).  Each thread keeps its verdict in registers and the step's results are written once.
    uint32_t myAct = ACT_UNPROCESSED, myEv = 0, nEv = 0, myK = 0;
    bool rejected = false;

    for (;;) {
      bool uncertain = mine && !LUT && fabsf(score) <= a.guard;
      const bool accept = mine && score > 0.f;
      bool modifies = false;
      uint32_t than = 0;
      if (accept && !uncertain) {
        if (k < K) { modifies = true; than = k; }                       // :668-677
        else if (K > 1) {
          if (a.crit == 1) {
            if (KIND == 0) modifies = criterion_maxMinDist_int<KT>(ws, ln, K, tid, ws.ctl(34), ws.ctl(35), than);
            else modifies = criterion_maxMinDist<KT>(ws, ln, K, tid, than);
          }
          else if (a.crit == 0) modifies = criterion_maxAverage<KT>(ws, ln, K, tid, than);
          else if (CRITX) {           // criteria 2 / 3: instantiated for the generic variant only (launchNW)
            float gd[KMAX];
            APT_FOR_K(r, K) gd[r] = KT ? ln.d[KT ? r : 0] : ws.d(r * ws.T + tid);
            const uint32_t res = criterion_other(a.crit, ws.oD, K, gd);
            than = res & 0x3FFFFFFFu;
            uncertain = (res >> 31) != 0u;                               // criterion 3 inside its error band: the host decides
            modifies = !uncertain && ((res >> 30) & 1u);
          }
        }
      }
      // earliest thread of the CTA whose decision is an event (changes the enveloppe / needs the host),
      // published together with what it does: one barrier tells every thread who and where
      const uint32_t em = __ballot_sync(0xffffffffu, uncertain || modifies);
      if (em) {
        const int first = __ffs(em) - 1;
        const uint32_t word = (gpos << EV_TID_SHIFT) | (uncertain ? EV_UNCERTAIN : (than & EV_SLOT_MASK));
        const uint32_t w0 = __shfl_sync(0xffffffffu, word, first);
        if (lane == 0) ws.ctl(warp) = w0;
        // the warp's candidate shows its genome before the barrier: should it be the event, the later births measure
        // it at once, while warp 0 updates the matrix (nobody waits for the owner to publish it after the barrier)
        if (!CLUSTER && lane == first && !uncertain) {
#pragma unroll
          for (int v = 0; v < WQ; v++) ws.reps((K + warp) * WQ + v) = bw[v];
          ws.rmeta(K + warp) = bm;
        }
      } else if (lane == 0) ws.ctl(warp) = NONE;
      __syncthreads();
      if (prof) { const long long t = clock64(); profA += t - profT; profT = t; }
      uint32_t event = NONE;
#pragma unroll
      for (int w = 0; w < NW; w++) event = min(event, ws.ctl(w));
      if (CLUSTER) {
        // the CTAs' earliest events meet in everybody's control words (two sets used in turn: a CTA that runs ahead
        // must not overwrite what a slower one still reads)
        if (uint32_t(tid) < csize) *cg::this_cluster().map_shared_rank(&ws.ctl(64 + 16 * elect + crank), tid) = event;
        cg::this_cluster().sync();
        event = NONE;
        for (uint32_t r = 0; r < csize; r++) event = min(event, ws.ctl(64 + 16 * elect + r));
        elect ^= 1u;
      }
      const uint32_t stop = event == NONE ? NONE : event >> EV_TID_SHIFT;
      if (mine && gpos < stop) {                    // final: every earlier decision stands
        myAct = accept ? ACT_ACCEPTED : ACT_FALLTHROUGH;
        rejected = !accept;
        myK = k;                                    // the enveloppe fill this birth's top-level score saw
        mine = false;
      }
      if (event == NONE) break;
      if (event & EV_UNCERTAIN) {                     // the host takes the stream over from here
        if (gpos == stop) myAct = ACT_UNCERTAIN;
        aborted = true;
        break;
      }
      profEvents++;
      const uint32_t sl = event & EV_SLOT_MASK;
      const bool append = (k < K);
      const uint32_t kOld = k;
      if (append) k++;
      const bool ownEvent = !CLUSTER || stop / T == crank;       // (uniform in the CTA)
      if (gpos == stop) {
#pragma unroll
        for (int v = 0; v < WQ; v++) ws.reps(sl * WQ + v) = bw[v];
        ws.rmeta(sl) = bm;
        myAct = ACT_ENTERED | (append ? 0u : ACT_REPLACED) | (sl << ACT_SLOT_SHIFT);
        myEv = nEv;
        ws.ev(nEv) = (uint32_t(tid) << 8) | kOld;    // whose distances are the row, and how many
        mine = false;
      }
      if (ownEvent) nEv++;
      if (CLUSTER) {
        // the event's CTA pushes the new representative into the other copies of the enveloppe, and the entering birth's
        // distances (the new row / column of the matrix) into everybody's, its own included
        if (ownEvent) {
          __syncthreads();
          const uint32_t ls = stop - crank * T;
          for (uint32_t r = 0; r < csize; r++) {
            unsigned char *peer = cg::this_cluster().map_shared_rank(apt_smem, r);
            if (r != crank) {
              for (uint32_t v = tid; v < uint32_t(WQ); v += T) reinterpret_cast<uint4*>(peer + ws.oReps)[sl * WQ + v] = ws.reps(sl * WQ + v);
              if (tid == 0) reinterpret_cast<float4*>(peer + ws.oRmeta)[sl] = ws.rmeta(sl);
            }
            for (uint32_t j = tid; j < K; j += T) reinterpret_cast<float*>(peer + ws.oCtl)[112 + j] = ws.d(j * T + ls);
          }
        }
        cg::this_cluster().sync();
      }
      // the entering birth's distances are the new row / column sl of the matrix (:676-677 / :712-715).
      // Lane r of warp 0 owns row r (only its column sl changes), one lane of warp 1 owns row sl (the birth's
      // own distances): no thread needs another's writes for its new minimum.  Straight-line code, loads
      // unconditional -- this sits on the critical path of every event.
      if (KIND == 0) {
        // bit strings: warp 0 does all rows in one instruction stream (lane sl: the birth's own row), because the
        // rows' minima meet in a warp reduction right after (criterion_maxMinDist_int's keys)
        if (warp == 0) {
          // (cluster: the entering birth's distances were pushed into the control words; else they are its thread's column)
          const uint32_t aEnt = CLUSTER ? ws.sb + ws.oCtl + 112 * 4 : ws.sb + ws.od + stop * 4, sEnt = CLUSTER ? 4u : T * 4u;
          const uint32_t r = min(uint32_t(tid), K - 1u), aRow = ws.sb + ws.oD + r * K * 4;
          const bool row = uint32_t(tid) < k, own = uint32_t(tid) == sl;
          const float dr = lds_f32(aEnt + r * sEnt);
          float m = FMAX;
#pragma unroll (KT ? KT : 1)
          for (uint32_t j = 0; j < K; j++) {
            const float e = lds_f32(aEnt + j * sEnt);           // columns past kOld hold stale values: masked
            const float w = lds_f32(aRow + j * 4);
            const float v = own ? ((j == sl || j >= kOld) ? FMAX : e) : (j == sl ? dr : w);
            if (own) ws.D(sl * K + j) = v;
            m = fminf(m, v);
          }
          if (row && !own) ws.D(r * K + sl) = dr;
          if (row) ws.minBase(r) = m;
          const uint32_t key = (uint32_t(tid) < K) ? minbase_key(row ? m : FMAX, tid) : 0xFFFFFFFFu;
          const uint32_t k1 = __reduce_min_sync(0xffffffffu, key);
          const uint32_t k2 = __reduce_min_sync(0xffffffffu, key == k1 ? 0xFFFFFFFFu : key);
          if (lane == 0) { ws.ctl(34) = k1;  ws.ctl(35) = k2; }
        }
      } else {
        const uint32_t aEnt = CLUSTER ? ws.sb + ws.oCtl + 112 * 4 : ws.sb + ws.od + stop * 4, sEnt = CLUSTER ? 4u : T * 4u;
        if (uint32_t(tid) < k && uint32_t(tid) != sl) {
          const uint32_t r = tid, aRow = ws.sb + ws.oD + r * K * 4;
          const float dr = lds_f32(aEnt + r * sEnt);
          float m = FMAX;
#pragma unroll (KT ? KT : 1)
          for (uint32_t j = 0; j < K; j++) {
            const float w = lds_f32(aRow + j * 4);
            m = fminf(m, j == sl ? dr : w);
          }
          ws.D(r * K + sl) = dr;
          ws.minBase(r) = m;
        } else if (tid == 32) {
          float m = FMAX;
#pragma unroll (KT ? KT : 1)
          for (uint32_t j = 0; j < K; j++) {
            const float e = lds_f32(aEnt + j * sEnt);           // columns past kOld hold stale values: masked
            const float v = (j == sl || j >= kOld) ? FMAX : e;
            ws.D(sl * K + j) = v;
            m = fminf(m, v);
          }
          ws.minBase(sl) = m;
        }
      }
      // later births see the new representative: re-measure that slot only (from the winner's staging row: the matrix
      // update above and this run side by side), the score follows from the change of that one compatibility
      if (mine) {
        sumC += measure<WQ, KIND, LUT, KT>(a.cfg, bw, bm, ws, ln, CLUSTER ? sl : K + (stop >> 5), sl, tid, !append);
        dirtyFrom = min(dirtyFrom, sl);
        score = decision_score<KT>(a.cfg, ws, ln, k, tid, sumC, bandGuard, dirtyFrom);
      }
      __syncthreads();            // the criterion at the top of the loop reads the matrix' row minima
      if (prof) { const long long t = clock64(); profB += t - profT; profT = t; }
      changed |= 1u << sl;
    }

    // ---- the step's results leave the chip: one row claim for all its enveloppe entries, one result
    //      word per birth, the cached distance row of every entering birth (what the host mirrors),
    //      and the births rejected at top level queue up for the sub-species kernel ----
    if (nEv && tid == 0) ws.ctl(33) = atomicAdd(a.row_counter, nEv);
    const uint32_t ftm = __ballot_sync(0xffffffffu, rejected);
    // ... and, for the sub-species resolution on the device (family.cu), line up per stream IN ARRIVAL ORDER:
    // the warps' counts meet in shared memory across the barrier below
    if (a.rej_order && lane == 0) ws.ctl(40 + warp) = uint32_t(__popc(ftm));
    if (ftm && a.ft_list) {
      uint32_t fbase = 0;
      if (lane == 0) fbase = atomicAdd(a.ft_counter, uint32_t(__popc(ftm)));
      fbase = __shfl_sync(0xffffffffu, fbase, 0);
      if (rejected) {
        const uint32_t at = fbase + __popc(ftm & ((1u << lane) - 1u));
        a.ft_list[2 * at] = b;  a.ft_list[2 * at + 1] = g;
      }
    }
    __syncthreads();
    if (a.rej_order) {
      uint32_t before = 0, all = 0;
#pragma unroll
      for (int w = 0; w < NW; w++) { const uint32_t c = ws.ctl(40 + w); all += c; before += (w < warp) ? c : 0u; }
      if (CLUSTER) {
        // arrival order across the cluster: the CTAs of lower rank come first
        if (uint32_t(tid) < csize) *cg::this_cluster().map_shared_rank(&ws.ctl(96 + crank), tid) = all;
        cg::this_cluster().sync();
        all = 0;
        for (uint32_t r = 0; r < csize; r++) { const uint32_t c = ws.ctl(96 + r); all += c; before += (r < crank) ? c : 0u; }
      }
      if (rejected) {
        const uint32_t at = g0 + nRej + before + __popc(ftm & ((1u << lane) - 1u));
        a.rej_order[at] = ord;
        a.rej_fill[at] = myK;                     // the enveloppe fill its top-level score saw (Stats::comparisons, :444 / :470)
      }
      nRej += all;
    }
    const uint32_t rowBase = nEv ? ws.ctl(33) : 0u;
    if (active) {
      const uint32_t code = myAct & ACT_CODE_MASK;
      if (code == ACT_ENTERED) {
        myAct |= (rowBase + myEv) << ACT_ROW_SHIFT;
        a.out_slot[b] = int32_t((myAct >> ACT_SLOT_SHIFT) & 31u);
        if (a.row_payload) {          // device-resident births: the host mirrors the new representative from here
          uint4 *dst = reinterpret_cast<uint4*>(a.row_payload) + size_t(rowBase + myEv) * WQ;
#pragma unroll
          for (int v = 0; v < WQ; v++) dst[v] = bw[v];
        }
      }
      if (code == ACT_ENTERED || code == ACT_ACCEPTED) a.out_sid[b] = sid;
      if (code == ACT_ACCEPTED) { nAcc0 += (ccls == 0); nAcc1 += (ccls == 1); nAcc2 += (ccls == 2); }
      a.act[b] = myAct;
    }
    for (uint32_t idx = tid; idx < nEv * K; idx += T) {
      const uint32_t e = idx / K, r = idx - e * K, info = ws.ev(e);
      if (r < (info & 0xFFu)) a.rows[size_t(rowBase + e) * K + r] = ws.d(r * T + (info >> 8));
    }
    __syncthreads();
  }

  // ---- per-stream totals ----
  atomicAdd(&ws.ctl(36), nAcc0);  atomicAdd(&ws.ctl(37), nAcc1);  atomicAdd(&ws.ctl(38), nAcc2);
  __syncthreads();
  if (CLUSTER) {
    // rank 0 sums the CTAs' counters and writes the enveloppe back (the copies are identical); the others stay until it has
    cg::this_cluster().sync();
    if (crank == 0 && tid == 0) {
      uint32_t s0 = 0, s1 = 0, s2 = 0;
      for (uint32_t r = 0; r < csize; r++) {
        const uint32_t *pc = cg::this_cluster().map_shared_rank(&ws.ctl(0), r);
        s0 += pc[36];  s1 += pc[37];  s2 += pc[38];
      }
      a.grp_sum[g] = make_uint4(s0, s1, s2, aborted ? 1u : 0u);
      if (a.rej_order) a.rej_count[g] = nRej;
    }
    cg::this_cluster().sync();
    if (crank != 0) return;
  } else if (tid == 0) {
    a.grp_sum[g] = make_uint4(ws.ctl(36), ws.ctl(37), ws.ctl(38), aborted ? 1u : 0u);
    if (a.rej_order) a.rej_count[g] = nRej;
  }

  // ---- write the enveloppe back (the host mirrors the same changes from act/rows) ----
  if (changed && !aborted) {
    uint4 *outRows = reinterpret_cast<uint4*>(a.pool_payload_w) + size_t(slot) * K * WQ;
    for (uint32_t i = tid; i < k * WQ; i += T)
      if ((changed >> (i / WQ)) & 1u) outRows[i] = ws.reps(i);
    for (uint32_t r = tid; r < k; r += T)
      if ((changed >> r) & 1u) a.pool_cmeta_w[size_t(slot) * K + r] = ws.rmeta(r);
    for (uint32_t i = tid; i < K * K; i += T) {
      const uint32_t row = i / K, col = i - row * K;
      if (row < k && col < k && row != col) a.pool_dist[size_t(slot) * K * K + i] = ws.D(i);
    }
    if (tid == 0) a.pool_fill_w[slot] = k;
  }
  if (prof && tid == 0) { prof[4] = clock64();  prof[5] = profEvents;  prof[6] = profSteps;  prof[8] = profA;  prof[9] = profB; }
  if (a.span && tid == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    atomicMax(a.span + 1, gt);
  }
}

/// `cluster` > 1: CTAs per stream (the CLUSTER instantiation); returns false when the device cannot place such clusters
template <int WQ, int NW, int KT, bool CRITX = false, bool CLUSTER = false>
bool launchWQ (const StreamArgs &a, cudaStream_t st, uint32_t cluster = 1) {
  const uint32_t K = a.pool.K;
  size_t smem = (stream_smem(K, WQ, NW * 32, KT > 0) + 15) & ~size_t(15);
  StreamArgs b = a;
  const size_t lutBytes = size_t(a.cfg.lut_classes) * a.cfg.lut_stride * sizeof(double);
  b.lut_in_smem = a.cfg.use_lut && lutBytes > 0 && lutBytes <= (size_t(48) << 10) && smem + lutBytes <= (size_t(200) << 10);
  if (b.lut_in_smem) smem += lutBytes;
  bool ok = true;
  auto go = [&] (auto kernel) {
    // (the attribute call costs microseconds: made when the size changes, not per launch.  Keyed by the kernel: every
    // variant has the same pointer TYPE, so this lambda -- and a plain static in it -- is shared by all of them)
    struct Set { const void *fn; size_t smem; int dev; };
    static Set sets[4] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    Set *e = nullptr;
    for (Set &q: sets) if (q.fn == reinterpret_cast<const void*>(kernel) || (!q.fn && !e)) { e = &q; if (q.fn) break; }
    if (!e || e->fn != reinterpret_cast<const void*>(kernel) || e->smem != smem || e->dev != dev) {
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
      if (e) { e->fn = reinterpret_cast<const void*>(kernel);  e->smem = smem;  e->dev = dev; }
    }
    // (preload: the runtime loads a kernel on its first use -- milliseconds that would otherwise land in the first pass)
    if (a.preload) { cudaFuncAttributes fa;  cudaFuncGetAttributes(&fa, kernel);  return; }
    if (a.ev_before) cudaEventRecord(a.ev_before, st);
    if (!CLUSTER) { kernel<<<a.n_groups, NW * 32, smem, st>>>(b);  return; }
    cudaLaunchConfig_t lc = {};
    cudaLaunchAttribute at[1];
    lc.blockDim = dim3(NW * 32);  lc.dynamicSmemBytes = smem;  lc.stream = st;  lc.attrs = at;  lc.numAttrs = 1;
    at[0].id = cudaLaunchAttributeClusterDimension;
    if (cluster > 8) cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    uint32_t cl = cluster;
    for (; cl > 1; cl /= 2) {             // the largest cluster the device can schedule
      lc.gridDim = dim3(a.n_groups * cl);
      at[0].val.clusterDim.x = cl;  at[0].val.clusterDim.y = 1;  at[0].val.clusterDim.z = 1;
      int nClusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nClusters, kernel, &lc) == cudaSuccess && nClusters >= 1) break;
      cudaGetLastError();
    }
    if (cl > 1) cudaLaunchKernelEx(&lc, kernel, b); else ok = false;
  };
  if (a.cfg.kind == 0) {
    if (a.cfg.use_lut) go(resolve_streams_kernel<WQ, 0, true, NW, KT, CRITX, CLUSTER>); else go(resolve_streams_kernel<WQ, 0, false, NW, KT, CRITX, CLUSTER>);
  } else go(resolve_streams_kernel<WQ, 1, false, NW, KT, CRITX, CLUSTER>);
  return ok;
}

/// Warps per stream: wide CTAs pay off when some stream is long; genomes that fill the register
/// file (>= 2048 bits per thread) stay at 4 warps.  Enveloppe sizes 5 (the reference's default,
/// ptreeconfig.cpp:10) and 10 (BASELINE.json) get the register-resident variant for short genomes.
template <int WQ>
void launchNW (const StreamArgs &a, cudaStream_t st) {
  // criteria 2 and 3 (debug options of the reference): one variant, any enveloppe size, 4 warps
  if (a.crit >= 2) { launchWQ<WQ, 4, 0, true>(a, st); return; }
  // a pass of few streams one of which is very long (one species holds the births: the first generations of a
  // radiation): that many CTAs of 4 warps per stream as keep all streams on the chip at once, 16 at most
  const char *env = getenv("APT_STREAM_CLUSTER_MIN");          // (read per launch: tests switch it)
  const uint32_t clusterMin = env ? uint32_t(std::max(1, atoi(env))) : 2048u;
  if (a.preload) launchWQ<WQ, 4, 0, false, true>(a, st);
  else if (a.max_stream >= clusterMin && a.n_groups * 2 <= 148) {
    uint32_t cl = 2;
    while (cl < 16 && a.n_groups * cl * 2 <= 148 && cl * 128 < a.max_stream) cl *= 2;
    if (launchWQ<WQ, 4, 0, false, true>(a, st, cl)) return;
  }
  const bool wide = a.max_stream > 1024 && WQ <= 8
                    && ((stream_smem(a.pool.K, WQ, 512) + 15) & ~size_t(15)) <= (size_t(150) << 10);
  if constexpr (WQ <= 4) {
    if (a.pool.K == 10) { if (wide) launchWQ<WQ, 16, 10>(a, st); else launchWQ<WQ, 4, 10>(a, st); return; }
    if (a.pool.K == 5) { if (wide) launchWQ<WQ, 16, 5>(a, st); else launchWQ<WQ, 4, 5>(a, st); return; }
  }
  if constexpr (WQ <= 8) { if (wide) { launchWQ<WQ, 16, 0>(a, st); return; } }
  launchWQ<WQ, 4, 0>(a, st);
}

} // namespace

bool stream_kernel_supports (uint32_t wp, uint32_t K, int crit) {
  const uint32_t wq = wp >> 2;
  if (!(wq == 1 || wq == 2 || wq == 4 || wq == 8 || wq == 16 || wq == 32)) return false;
  if (crit < 0 || crit > 3) return false;
  return ((stream_smem(K, wq, 128) + 15) & ~size_t(15)) <= size_t(150) << 10;
}

void launch_resolve_streams (const StreamArgs &a, cudaStream_t stream) {
  if (!a.n_groups && !a.preload) return;
  switch (a.pool.wp >> 2) {
  case 1: launchNW<1>(a, stream); break;
  case 2: launchNW<2>(a, stream); break;
  case 4: launchNW<4>(a, stream); break;
  case 8: launchNW<8>(a, stream); break;
  case 16: launchNW<16>(a, stream); break;
  case 32: launchNW<32>(a, stream); break;
  default: break;     // callers check stream_kernel_supports()
  }
}

} // namespace apt
