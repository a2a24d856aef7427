// resolve_families_kernel -- the sub-species walk of rejected births on the device, in order.
//
// A birth that its parent species S rejects at top level (score <= 0, phylogenetictree.hpp:612) visits the direct
// sub-species of S newest first and joins the first one whose score is > 0 (findBestDerived, :493-548); there it is
// appended while the enveloppe is not full (:668-677), replaces a representative when the criterion says so
// (:684-718, enveloppecriteria.cpp) or changes nothing; if no sub-species accepts it, it founds a new one (:637-643),
// which the NEXT rejected birth of S visits first.  While a species radiates this is what every birth does, and each
// one depends on what the previous ones did.
//
// One CTA owns one "family" -- the rejected births of one stream, in arrival order (the stream kernel lines them up),
// and the sub-species of its species: those the pool held at submission plus the ones founded in this pass, all in
// shared memory.  128 births are decided per step, one thread each, against the state at the start of the step; the
// earliest birth whose decision changes anything (a new species, an append, a replacement) commits, every earlier one
// is final, and the later ones re-score only the sub-species that changed -- a thread keeps one "accepts me" bit per
// sub-species, so "the first accepting sub-species in visiting order" is a bit scan.  Births that join a full
// enveloppe without changing it (the bulk, once the niches are populated) cost no event at all.
//
// A family with very many births (the first generations of a radiation: ONE family of 100 000) is spread over a thread
// block CLUSTER: every CTA keeps its own copy of the family's state, a step decides 256 births per CTA -- position in
// the step = CTA rank * 256 + thread -- the earliest event is elected across the cluster through distributed shared
// memory, its CTA applies it and pushes the changed sub-species into the other copies.  Scoring a step's births against
// all sub-species, the bulk of the work when events are rare, then runs on up to 16 SMs instead of one.
//
// The host only commits what changed -- new species (it hands out the ids in birth order, :225-230), enveloppe
// entries -- and books the rest per sub-species from the kernel's counters (engine.cpp).  A family whose state stops
// fitting (sub-species capacity, reserved pool slots), or whose next decision needs exact arithmetic (scores inside
// the guard band without the compatibility table), hands the REST of its births back to the host's walk.
#include "kernels.cuh"

#include <cooperative_groups.h>
#include <math_constants.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace apt {

namespace {

constexpr float FMAX = 3.402823466e+38f;
constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr int FT = 256;                 // most threads (= births per step) a family CTA runs with; 128 when there are many families

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

extern __shared__ __align__(16) unsigned char fam_smem[];

/// Byte offsets of the family's state in shared memory
struct FamState {
  uint32_t oReps, oMeta, oD, oMinB, oFill, oSid, oChanged, oJoin, oAcc, oUnc, oDc, oEv, oCtl, oLut;
  uint32_t K, cap, MW, T;
  const double *lutG;
  bool lutS;
  __device__ __forceinline__ uint4& reps (uint32_t c, uint32_t r, uint32_t v, uint32_t WQ) const { return reinterpret_cast<uint4*>(fam_smem + oReps)[(c * K + r) * WQ + v]; }
  __device__ __forceinline__ float4& meta (uint32_t c, uint32_t r) const { return reinterpret_cast<float4*>(fam_smem + oMeta)[c * K + r]; }
  __device__ __forceinline__ float& D (uint32_t c, uint32_t i) const { return reinterpret_cast<float*>(fam_smem + oD)[c * K * K + i]; }
  __device__ __forceinline__ float& minB (uint32_t c, uint32_t r) const { return reinterpret_cast<float*>(fam_smem + oMinB)[c * K + r]; }
  __device__ __forceinline__ uint32_t& fill (uint32_t c) const { return reinterpret_cast<uint32_t*>(fam_smem + oFill)[c]; }
  __device__ __forceinline__ uint32_t& sid (uint32_t c) const { return reinterpret_cast<uint32_t*>(fam_smem + oSid)[c]; }
  __device__ __forceinline__ uint32_t& changed (uint32_t c) const { return reinterpret_cast<uint32_t*>(fam_smem + oChanged)[c]; }
  __device__ __forceinline__ uint32_t& join (uint32_t i) const { return reinterpret_cast<uint32_t*>(fam_smem + oJoin)[i]; }
  __device__ __forceinline__ uint32_t& acc (uint32_t w, uint32_t t) const { return reinterpret_cast<uint32_t*>(fam_smem + oAcc)[w * T + t]; }
  __device__ __forceinline__ uint32_t& unc (uint32_t w, uint32_t t) const { return reinterpret_cast<uint32_t*>(fam_smem + oUnc)[w * T + t]; }
  __device__ __forceinline__ float& dc (uint32_t r, uint32_t t) const { return reinterpret_cast<float*>(fam_smem + oDc)[r * T + t]; }
  __device__ __forceinline__ uint32_t& ev (uint32_t i) const { return reinterpret_cast<uint32_t*>(fam_smem + oEv)[i]; }
  __device__ __forceinline__ uint32_t& ctl (uint32_t i) const { return reinterpret_cast<uint32_t*>(fam_smem + oCtl)[i]; }
  __device__ __forceinline__ double lut (size_t i) const { return lutS ? reinterpret_cast<const double*>(fam_smem + oLut)[i] : __ldg(lutG + i); }
};

__host__ __device__ inline size_t fam_child_bytes (uint32_t K, uint32_t wq) {
  return size_t(K) * wq * 16 + size_t(K) * 16 + size_t(K) * K * 4 + size_t(K) * 4 + 4 + 4 + 4 + 8;
}
__host__ __device__ inline size_t fam_fixed_bytes (uint32_t K, uint32_t cap) {
  const size_t MW = (cap + 31) / 32;
  return 2 * MW * FT * 4 + size_t(K) * FT * 4 + size_t(FT) * 4 + 96 * 4 + 64;
}

/// speciesMatchingScore (:456-478 / :430-452) of this thread's birth against sub-species c, representatives in order;
/// `store`: the k distances go to the thread's cache row (they become the matrix row if the birth enters)
template <int WQ, int KIND, bool LUT>
__device__ __forceinline__ float fam_score (const ScoreCfg &cfg, const FamState &fs, const uint4 (&bw)[WQ], float4 bm,
                                            uint32_t c, uint32_t k, bool store, int tid) {
  float avg = 0.f;
  uint32_t matable = 0;
  for (uint32_t r = 0; r < k; r++) {
    double d;
    if (KIND == 0) {
      uint32_t a = 0;
#pragma unroll
      for (int v = 0; v < WQ; v++) a += popc4(bw[v], fs.reps(c, r, v, WQ));
      d = double(a);
    } else {
      double a = 0.;
#pragma unroll
      for (int v = 0; v < WQ; v++) a += sqdiff4(bw[v], fs.reps(c, r, v, WQ));
      d = sqrt(a);
    }
    const float4 rm = fs.meta(c, r);
    double ca, cb;
    if (LUT) {
      const uint32_t di = uint32_t(d);
      ca = fs.lut(size_t(__float_as_uint(bm.w)) * cfg.lut_stride + di);
      cb = fs.lut(size_t(__float_as_uint(rm.w)) * cfg.lut_stride + di);
    } else { ca = compat_generic(bm, d); cb = compat_generic(rm, d); }
    const double cc = cb < ca ? cb : ca;
    avg = float(__dadd_rn(double(avg), cc));
    matable += (cc >= cfg.compat_thr);
    if (store) fs.dc(r, tid) = float(d);
  }
  if (cfg.full_continuous) return float(__dsub_rn(double(__fdiv_rn(avg, float(k))), cfg.avg_thr));
  return float(__dsub_rn(double(matable), __dmul_rn(cfg.sim_thr, double(k))));
}

// event kinds (the word every warp publishes: thread << 8 | kind; smaller = earlier)
constexpr uint32_t EVK_FOUND = 1, EVK_ENTER = 2, EVK_STOP = 3;

template <int WQ, int KIND, bool LUT, int NT, bool CLUSTER>
__global__ void __launch_bounds__(NT, 1)
resolve_families_kernel (FamilyArgs a) {
  namespace cg = cooperative_groups;
  constexpr int FT = NT;                  // (shadows the sizing constant: everything below is per launch width)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t crank = 0, csize = 1;          // this CTA's rank in the family's cluster, CTAs per family
  if (CLUSTER) { cg::cluster_group cl = cg::this_cluster();  crank = cl.block_rank();  csize = cl.num_blocks(); }
  auto cluster_sync = [&] (void) { if (CLUSTER) cg::this_cluster().sync(); else __syncthreads(); };
  const uint32_t f = blockIdx.x / csize, g = a.fam_group[f];
  const uint32_t gpos = crank * FT + uint32_t(tid);           // this thread's position among the births of a step
  const uint32_t K = a.pool.K, cap = a.cap, MW = (cap + 31) / 32;
  FamState fs;
  fs.K = K;  fs.cap = cap;  fs.MW = MW;  fs.T = FT;
  uint32_t base = 0;
  fs.oReps = base;     base += cap * K * WQ * 16;
  fs.oMeta = base;     base += cap * K * 16;
  fs.oD = base;        base += cap * K * K * 4;
  fs.oMinB = base;     base += cap * K * 4;
  fs.oFill = base;     base += cap * 4;
  fs.oSid = base;      base += cap * 4;
  fs.oChanged = base;  base += cap * 4;
  fs.oJoin = base;     base += cap * 8;
  fs.oAcc = base;      base += MW * FT * 4;
  fs.oUnc = base;      base += MW * FT * 4;
  fs.oDc = base;       base += K * FT * 4;
  fs.oEv = base;       base += FT * 4;
  base = (base + 7u) & ~7u;        // (64-bit totals live in the control words)
  fs.oCtl = base;      base += 96 * 4;      // 0..7 warps' events, 8..11 the event, 16 new slot, 20 births done, 32..35 totals, 40..71 the cluster's events (two sets)
  fs.oLut = (base + 15u) & ~15u;
  fs.lutG = a.cfg.lut;
  fs.lutS = LUT && a.lut_in_smem;

  const uint32_t g0 = a.grp_off[g];
  const uint32_t nRej = a.rej_count[g];
  if (nRej == 0) {                        // nothing was rejected: the usual case outside radiations
    if (threadIdx.x == 0 && crank == 0) { a.fam_info[f] = make_uint4(0u, 0u, 0u, 0u); a.fam_stats[2 * size_t(f)] = 0ull; a.fam_stats[2 * size_t(f) + 1] = 0ull; }
    return;
  }
  const uint32_t c0 = a.child_off[g], Cs = a.child_off[g + 1] - c0;     // sub-species at submission (host: Cs < cap)
  const uint4 *birthRows = reinterpret_cast<const uint4*>(a.b_payload);

  // ---- the snapshot sub-species come in from the pool ----
  for (uint32_t i = tid; i < cap; i += FT) {
    const bool live = i < Cs;
    const uint32_t slot = live ? __ldg(a.child_slot + c0 + i) : 0u;
    fs.fill(i) = live ? min(__ldg(a.pool.fill + slot), K) : 0u;
    fs.sid(i) = live ? __ldg(a.child_sid + c0 + i) : NONE;
    fs.changed(i) = 0u;
    fs.join(2 * i) = 0u;  fs.join(2 * i + 1) = 0u;
  }
  for (uint32_t i = tid; i < Cs * K * WQ; i += FT) {
    const uint32_t c = i / (K * WQ), rest = i - c * K * WQ;
    const uint32_t slot = __ldg(a.child_slot + c0 + c);
    reinterpret_cast<uint4*>(fam_smem + fs.oReps)[i] = reinterpret_cast<const uint4*>(a.pool.payload)[size_t(slot) * K * WQ + rest];
  }
  for (uint32_t i = tid; i < Cs * K; i += FT) {
    const uint32_t c = i / K, r = i - c * K;
    fs.meta(c, r) = a.pool.cmeta[size_t(__ldg(a.child_slot + c0 + c)) * K + r];
  }
  if (fs.lutS) {
    double *lut_s = reinterpret_cast<double*>(fam_smem + fs.oLut);
    const uint32_t nd = a.cfg.lut_classes * a.cfg.lut_stride;
    for (uint32_t i = tid; i < nd; i += FT) lut_s[i] = __ldg(a.cfg.lut + i);
  }
  if (tid < 8) fs.ctl(32 + tid) = 0u;            // 32..35: 64-bit totals (branching, comparisons)
  __syncthreads();
  // matrices with FLT_MAX on the diagonal and in unused cells (a row minimum needs no masks), then the row minima
  for (uint32_t i = tid; i < Cs * K * K; i += FT) {
    const uint32_t c = i / (K * K), rest = i - c * K * K, row = rest / K, col = rest - row * K;
    const uint32_t k = fs.fill(c);
    const float v = a.pool_dist[size_t(__ldg(a.child_slot + c0 + c)) * K * K + rest];
    fs.D(c, rest) = (row < k && col < k && row != col) ? v : FMAX;
  }
  __syncthreads();
  for (uint32_t i = tid; i < Cs * K; i += FT) {
    const uint32_t c = i / K, r = i - c * K;
    float m = FMAX;
    for (uint32_t j = 0; j < K; j++) m = fminf(m, fs.D(c, r * K + j));
    fs.minB(c, r) = m;
  }
  __syncthreads();

  uint32_t C = Cs, nNew = 0, nDone = 0;           // uniform: sub-species now, founded here, births decided
  uint32_t elect = 0;                             // which set of election words the next event uses (cluster)
  bool handedBack = false;
  unsigned long long branching = 0, comparisons = 0;        // this thread's births
  const uint32_t newSidBase = FAM_NEW_SID | (f * a.max_new);

  // sub-species visited up to and including j, newest first (:506): how many, and the sum of their fills
  auto visited = [&] (uint32_t j, uint32_t &nb, uint32_t &nc) {
    nb = 0;  nc = 0;
    if (j == NONE) { for (uint32_t c = 0; c < C; c++) { nb++; nc += fs.fill(c); } return; }
    for (uint32_t c = C; c-- > Cs; ) { nb++; nc += fs.fill(c); if (c == j) return; }
    for (uint32_t c = 0; c < Cs; c++) { nb++; nc += fs.fill(c); if (c == j) return; }
  };

  for (uint32_t cb = 0; cb < nRej && !handedBack; cb += FT * csize) {
    const uint32_t nb = min(uint32_t(FT) * csize, nRej - cb);
    const bool active = gpos < nb;
    const uint32_t ord = active ? __ldg(a.rej_order + g0 + cb + gpos) : 0u;
    const uint32_t topFill = active ? __ldg(a.rej_fill + g0 + cb + gpos) : 0u;
    const uint32_t b = ord & ORDER_INDEX_MASK, ccls = ord >> ORDER_CLASS_SHIFT;
    uint4 bw[WQ];
    float4 bm = make_float4(0.f, 1.f, 1.f, 0.f);
    if (active) {
#pragma unroll
      for (int v = 0; v < WQ; v++) bw[v] = __ldg(birthRows + size_t(b) * WQ + v);
      bm = __ldg(a.b_cmeta + b);
    } else {
#pragma unroll
      for (int v = 0; v < WQ; v++) bw[v] = make_uint4(0, 0, 0, 0);
    }
    // ---- every sub-species as it stands now ----
    for (uint32_t w = 0; w < MW; w++) { fs.acc(w, tid) = 0u; fs.unc(w, tid) = 0u; }
    if (active)
      for (uint32_t c = 0; c < C; c++) {
        const uint32_t k = fs.fill(c);
        if (!k) continue;
        const float sc = fam_score<WQ, KIND, LUT>(a.cfg, fs, bw, bm, c, k, false, tid);
        if (!LUT && fabsf(sc) <= a.guard) fs.unc(c >> 5, tid) |= 1u << (c & 31);
        else if (sc > 0.f) fs.acc(c >> 5, tid) |= 1u << (c & 31);
      }
    bool mine = active;
    uint32_t cached = NONE;                        // sub-species the thread's distance cache describes
    uint32_t myAct = FAM_UNPROCESSED, myEv = 0, nEv = 0;

    for (;;) {
      // ---- this birth under the current state: first sub-species, newest first, that accepts it (:540-541) ----
      uint32_t j = NONE, kind = 0, than = 0;
      bool uncertain = false;
      if (mine) {
        for (uint32_t w = MW; w-- > 0 && j == NONE; ) {          // founded here: highest id first
          uint32_t m = fs.acc(w, tid) | fs.unc(w, tid);
          if (w * 32 < Cs) m &= (Cs - w * 32 >= 32) ? 0u : ~((1u << (Cs - w * 32)) - 1u);
          if (m) j = w * 32 + 31 - __clz(m);
        }
        for (uint32_t w = 0; w < MW && j == NONE && w * 32 < Cs; w++) {     // at submission: rank order
          uint32_t m = fs.acc(w, tid) | fs.unc(w, tid);
          if (Cs - w * 32 < 32) m &= (1u << (Cs - w * 32)) - 1u;
          if (m) j = w * 32 + __ffs(m) - 1;
        }
        if (j == NONE) kind = EVK_FOUND;                                        // :637-643
        else if ((fs.unc(j >> 5, tid) >> (j & 31)) & 1u) { uncertain = true; kind = EVK_STOP; }
        else {
          const uint32_t k = fs.fill(j);
          if (k < K) { kind = EVK_ENTER; than = k; }                            // :668-677
          else if (K > 1) {
            if (cached != j) { fam_score<WQ, KIND, LUT>(a.cfg, fs, bw, bm, j, k, true, tid); cached = j; }
            float value = -CUDART_INF_F;
            uint32_t best = NONE;
            if (a.crit == 1) {                                                  // maxMinDist, enveloppecriteria.cpp:87-133
              float m1 = FMAX, m2 = FMAX;
              uint32_t i1 = NONE;
              for (uint32_t r = 0; r < K; r++) {
                const float x = fs.dc(r, tid);
                if (x < m1) { m2 = m1; m1 = x; i1 = r; } else if (x < m2) m2 = x;
              }
              for (uint32_t i = 0; i < K; i++) {
                const float cval = __fadd_rn(-fs.minB(j, i), i == i1 ? m2 : m1);
                if (value < cval) { value = cval; best = i; }
              }
            } else {                                                            // maxAverage, :40-84
              for (uint32_t i = 0; i < K; i++) {
                double cval = 0.;
                for (uint32_t q = 0; q < K; q++)
                  if (q != i) cval = __dadd_rn(cval, double(__fadd_rn(-fs.D(j, i * K + q), fs.dc(q, tid))));
                if (double(value) < cval) { value = float(cval); best = i; }
              }
            }
            if (value > 0.f) { kind = EVK_ENTER; than = best; }
          }
        }
      }
      // earliest thread whose decision is an event; one barrier tells everybody who
      const uint32_t em = __ballot_sync(0xffffffffu, kind != 0u);
      const int first = em ? __ffs(em) - 1 : 0;
      const uint32_t firstKind = __shfl_sync(0xffffffffu, kind, first);
      if (lane == 0) fs.ctl(warp) = em ? (((crank * FT + uint32_t(warp) * 32u + uint32_t(first)) << 8) | firstKind) : NONE;
      __syncthreads();
      uint32_t event = NONE;
#pragma unroll
      for (int w = 0; w < FT / 32; w++) event = min(event, fs.ctl(w));
      if (CLUSTER) {
        // the CTAs' earliest events meet in everybody's control words (40 + rank), one remote store per peer
        // (two sets of words, used in turn: a CTA that runs ahead must not overwrite what a slower one still reads)
        if (uint32_t(tid) < csize) *cg::this_cluster().map_shared_rank(&fs.ctl(40 + 16 * elect + crank), tid) = event;
        cg::this_cluster().sync();
        event = NONE;
        for (uint32_t r = 0; r < csize; r++) event = min(event, fs.ctl(40 + 16 * elect + r));
        elect ^= 1u;
      }
      const uint32_t stop = event == NONE ? NONE : event >> 8;
      uint32_t evKind = event == NONE ? 0u : (event & 0xFFu);
      if (mine && gpos < stop) {
        // final: joins sub-species j and leaves its enveloppe alone (:687-691)
        myAct = FAM_JOINED | (j << FAM_CHILD_SHIFT);
        uint32_t vb, vc;
        visited(j, vb, vc);
        branching += vb;  comparisons += vc + topFill;
        atomicAdd(&fs.join(2 * j + (ccls == 2 ? 1 : 0)), 1u);
        a.out_sid[b] = fs.sid(j);
        mine = false;
      }
      if (event == NONE) break;
      cluster_sync();                  // the final births above read the state the event below is about to change
      // ---- the event: applied by its own thread, published through shared memory ----
      if (gpos == stop) {
        uint32_t jj = j, slot = than;
        if (evKind == EVK_FOUND && (nNew == a.max_new || C == cap)) evKind = EVK_STOP;     // no room: hand the rest back
        if (evKind == EVK_FOUND) {
          const uint32_t ps = atomicAdd(a.slot_counter, 1u);
          if (ps >= a.slot_limit) evKind = EVK_STOP;                                       // (the slots of such a claim are simply never used)
          else fs.ctl(16 + 0) = a.slot_base + ps;
        }
        if (evKind == EVK_FOUND) {
          jj = C;  slot = 0;
          a.new_slot[size_t(f) * a.max_new + nNew] = fs.ctl(16);
          for (uint32_t i = 0; i < K * K; i++) fs.D(jj, i) = FMAX;
          for (uint32_t r = 0; r < K; r++) fs.minB(jj, r) = FMAX;
          fs.fill(jj) = 0u;
          fs.sid(jj) = newSidBase + nNew;
          a.new_founder[size_t(f) * a.max_new + nNew] = b;
        }
        if (evKind != EVK_STOP) {
          const uint32_t kOld = fs.fill(jj);
          uint32_t vb, vc;
          visited(evKind == EVK_FOUND ? NONE : jj, vb, vc);
          branching += vb;  comparisons += vc + topFill;
          if (evKind == EVK_ENTER && cached != jj) { fam_score<WQ, KIND, LUT>(a.cfg, fs, bw, bm, jj, kOld, true, tid); cached = jj; }
          // the birth's distances are row / column `slot` of the matrix (:676-677 / :712-715)
          for (uint32_t r = 0; r < K; r++) {
            const float v = (r < kOld && r != slot) ? fs.dc(r, tid) : FMAX;
            fs.D(jj, slot * K + r) = v;
            if (r != slot) fs.D(jj, r * K + slot) = v;
          }
          const uint32_t kNew = (slot == kOld) ? kOld + 1 : kOld;
          for (uint32_t r = 0; r < kNew; r++) {
            float m = FMAX;
            for (uint32_t q = 0; q < K; q++) m = fminf(m, fs.D(jj, r * K + q));
            fs.minB(jj, r) = m;
          }
#pragma unroll
          for (int v = 0; v < WQ; v++) fs.reps(jj, slot, v, WQ) = bw[v];
          fs.meta(jj, slot) = bm;
          fs.fill(jj) = kNew;
          fs.changed(jj) |= 1u << slot;
          myAct = (evKind == EVK_FOUND ? FAM_FOUNDED : (FAM_ENTERED | (slot == kOld ? 0u : FAM_REPLACED)))
                | (slot << FAM_SLOT_SHIFT) | (jj << FAM_CHILD_SHIFT);
          myEv = nEv;
          fs.ev(nEv) = (uint32_t(tid) << 8) | (evKind == EVK_FOUND ? 0u : kOld);
          a.out_sid[b] = fs.sid(jj);
          a.out_slot[b] = int32_t(slot);
          mine = false;
        }
        fs.ctl(8) = evKind;  fs.ctl(9) = jj;  fs.ctl(11) = slot;
      }
      const bool ownEvent = stop / FT == crank;               // (uniform in the CTA)
      if (CLUSTER) {
        // the event's CTA pushes what changed -- the sub-species' matrix, row minima, the new representative, fill, id --
        // into the other CTAs' copies of the family
        if (ownEvent) {
          __syncthreads();
          const uint32_t ek = fs.ctl(8), ej = fs.ctl(9), es = fs.ctl(11);
          for (uint32_t r = 0; r < csize; r++) {
            if (r == crank) continue;
            unsigned char *peer = cg::this_cluster().map_shared_rank(fam_smem, r);
            if (ek != EVK_STOP) {
              for (uint32_t i = tid; i < K * K; i += FT) reinterpret_cast<float*>(peer + fs.oD)[ej * K * K + i] = fs.D(ej, i);
              for (uint32_t i = tid; i < K; i += FT) reinterpret_cast<float*>(peer + fs.oMinB)[ej * K + i] = fs.minB(ej, i);
              for (uint32_t v = tid; v < uint32_t(WQ); v += FT) reinterpret_cast<uint4*>(peer + fs.oReps)[(ej * K + es) * WQ + v] = fs.reps(ej, es, v, WQ);
              if (tid == 0) {
                reinterpret_cast<float4*>(peer + fs.oMeta)[ej * K + es] = fs.meta(ej, es);
                reinterpret_cast<uint32_t*>(peer + fs.oFill)[ej] = fs.fill(ej);
                reinterpret_cast<uint32_t*>(peer + fs.oSid)[ej] = fs.sid(ej);
                reinterpret_cast<uint32_t*>(peer + fs.oChanged)[ej] = fs.changed(ej);
              }
            }
            if (tid == 0) { reinterpret_cast<uint32_t*>(peer + fs.oCtl)[8] = ek;  reinterpret_cast<uint32_t*>(peer + fs.oCtl)[9] = ej; }
          }
        }
        cg::this_cluster().sync();
      } else __syncthreads();
      evKind = fs.ctl(8);
      const uint32_t jj = fs.ctl(9);
      if (evKind == EVK_STOP) { handedBack = true; break; }
      if (ownEvent) nEv++;
      if (evKind == EVK_FOUND) { C++; nNew++; }
      // later births see the change: re-score that sub-species only
      if (mine) {
        const float sc = fam_score<WQ, KIND, LUT>(a.cfg, fs, bw, bm, jj, fs.fill(jj), false, tid);
        const uint32_t bit = 1u << (jj & 31);
        uint32_t am = fs.acc(jj >> 5, tid) & ~bit, um = fs.unc(jj >> 5, tid) & ~bit;
        if (!LUT && fabsf(sc) <= a.guard) um |= bit; else if (sc > 0.f) am |= bit;
        fs.acc(jj >> 5, tid) = am;  fs.unc(jj >> 5, tid) = um;
        if (cached == jj) cached = NONE;
      }
    }

    // ---- the step's results leave the chip ----
    if (nEv && tid == 0) fs.ctl(10) = atomicAdd(a.row_counter, nEv);
    __syncthreads();
    const uint32_t rowBase = nEv ? fs.ctl(10) : 0u;
    if (active && myAct != FAM_UNPROCESSED) {
      const uint32_t code = myAct & FAM_CODE_MASK;
      if (code != FAM_JOINED) {
        a.fam_row[b] = rowBase + myEv;
        if (a.row_payload) {
          uint4 *dst = reinterpret_cast<uint4*>(a.row_payload) + size_t(rowBase + myEv) * WQ;
#pragma unroll
          for (int v = 0; v < WQ; v++) dst[v] = bw[v];
        }
      }
      a.fam_act[b] = myAct;
    }
    for (uint32_t idx = tid; idx < nEv * K; idx += FT) {
      const uint32_t e = idx / K, r = idx - e * K, info = fs.ev(e);
      if (r < (info & 0xFFu)) a.rows[size_t(rowBase + e) * K + r] = fs.dc(r, info >> 8);
    }
    nDone += __syncthreads_count(active && myAct != FAM_UNPROCESSED);
  }

  // ---- totals, per-sub-species counters, and the enveloppes go back to the pool ----
  atomicAdd(reinterpret_cast<unsigned long long*>(&fs.ctl(32)), branching);
  atomicAdd(reinterpret_cast<unsigned long long*>(&fs.ctl(34)), comparisons);
  if (tid == 0) fs.ctl(20) = nDone;
  cluster_sync();
  const uint32_t jo = a.fam_join_off[f];
  if (CLUSTER) {
    // the CTAs' totals and per-sub-species counters are summed by rank 0, which also writes the enveloppes back (the copies
    // are identical); the others stay until it has read them
    if (crank == 0) {
      cg::cluster_group cl = cg::this_cluster();
      if (tid == 0) {
        unsigned long long br = 0, cmp = 0;
        uint32_t done = 0;
        for (uint32_t r = 0; r < csize; r++) {
          const uint32_t *pc = cl.map_shared_rank(&fs.ctl(0), r);
          done += pc[20];
          br += *reinterpret_cast<const unsigned long long*>(pc + 32);
          cmp += *reinterpret_cast<const unsigned long long*>(pc + 34);
        }
        a.fam_info[f] = make_uint4(done, nNew, handedBack ? 1u : 0u, C);
        a.fam_stats[2 * size_t(f)] = br;
        a.fam_stats[2 * size_t(f) + 1] = cmp;
      }
      for (uint32_t i = tid; i < 2 * C; i += FT) {
        uint32_t sum = 0;
        for (uint32_t r = 0; r < csize; r++) sum += cl.map_shared_rank(&fs.join(0), r)[i];
        a.child_join[jo + i] = sum;
      }
    }
    cg::this_cluster().sync();
    if (crank != 0) return;
  } else {
    if (tid == 0) {
      a.fam_info[f] = make_uint4(nDone, nNew, handedBack ? 1u : 0u, C);
      a.fam_stats[2 * size_t(f)] = *reinterpret_cast<unsigned long long*>(&fs.ctl(32));
      a.fam_stats[2 * size_t(f) + 1] = *reinterpret_cast<unsigned long long*>(&fs.ctl(34));
    }
    for (uint32_t i = tid; i < 2 * C; i += FT) a.child_join[jo + i] = fs.join(i);
  }
  for (uint32_t c = 0; c < C; c++) {
    const uint32_t ch = fs.changed(c);
    if (!ch) continue;
    const uint32_t slot = c < Cs ? __ldg(a.child_slot + c0 + c) : a.new_slot[size_t(f) * a.max_new + (c - Cs)];
    const uint32_t k = fs.fill(c);
    uint4 *outRows = reinterpret_cast<uint4*>(a.pool_payload_w) + size_t(slot) * K * WQ;
    for (uint32_t i = tid; i < k * WQ; i += FT)
      if ((ch >> (i / WQ)) & 1u) outRows[i] = reinterpret_cast<const uint4*>(fam_smem + fs.oReps)[size_t(c) * K * WQ + i];
    for (uint32_t r = tid; r < k; r += FT)
      if ((ch >> r) & 1u) a.pool_cmeta_w[size_t(slot) * K + r] = fs.meta(c, r);
    for (uint32_t i = tid; i < K * K; i += FT) {
      const uint32_t row = i / K, col = i - row * K;
      if (row < k && col < k && row != col) a.pool_dist[size_t(slot) * K * K + i] = fs.D(c, i);
    }
    if (tid == 0) a.pool_fill_w[slot] = k;
  }
}

size_t fam_lut_bytes (const ScoreCfg &cfg) {
  const size_t b = size_t(cfg.lut_classes) * cfg.lut_stride * sizeof(double);
  return (cfg.use_lut && b > 0 && b <= (size_t(48) << 10)) ? ((b + 15) & ~size_t(15)) : 0;
}

} // namespace

bool family_kernel_supports (uint32_t wp, uint32_t K, int crit) {
  const uint32_t wq = wp >> 2;
  if (!(wq == 1 || wq == 2 || wq == 4 || wq == 8 || wq == 16 || wq == 32)) return false;
  return (crit == 0 || crit == 1) && K >= 1 && K <= KMAX;
}

uint32_t family_capacity (uint32_t wp, uint32_t K, const ScoreCfg &cfg) {
  const uint32_t wq = wp >> 2;
  const size_t budget = (size_t(200) << 10) - fam_lut_bytes(cfg);
  uint32_t cap = 1024;
  while (cap >= 12 && fam_child_bytes(K, wq) * cap + fam_fixed_bytes(K, cap) > budget) cap -= 4;
  return cap >= 12 ? cap : 0u;           // (4096-bit genomes, K = 10: 28 sub-species)
}

void launch_resolve_families (const FamilyArgs &a, cudaStream_t st) {
  if (!a.n_fam) return;
  const uint32_t wq = a.pool.wp >> 2;
  FamilyArgs b = a;
  const size_t lutBytes = fam_lut_bytes(a.cfg);
  size_t smem = ((fam_child_bytes(a.pool.K, wq) * a.cap + fam_fixed_bytes(a.pool.K, a.cap)) + 15) & ~size_t(15);
  b.lut_in_smem = lutBytes > 0;
  smem += lutBytes;
  // few families (the first generations of a radiation: ONE family of 100 000 births) get 256 threads per CTA and a
  // cluster of CTAs each: 16 (more than the portable 8: opted into, and checked against what the device can place), 8,
  // 4 or 2 CTAs, as many as keep the clusters of all families on the chip at once
  const bool wide = a.n_fam <= 74;
  const char *env = getenv("APT_FAMILY_CLUSTER");                // (read per launch: tests switch it)
  const int maxCluster = env ? atoi(env) : 16;
  uint32_t want = 1;
  if (wide) for (uint32_t c = 2; c <= uint32_t(std::max(1, maxCluster)) && c <= 16 && a.n_fam * c <= 148; c *= 2) want = c;
  auto go = [&] (auto kernel, auto clustered, int threads) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaFuncSetAttribute(clustered, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (a.preload) { cudaFuncAttributes fa;  cudaFuncGetAttributes(&fa, kernel);  cudaFuncGetAttributes(&fa, clustered);  return; }
    uint32_t cl = want;
    cudaLaunchConfig_t lc = {};
    cudaLaunchAttribute at[1];
    lc.blockDim = dim3(threads);  lc.dynamicSmemBytes = smem;  lc.stream = st;  lc.attrs = at;  lc.numAttrs = 1;
    at[0].id = cudaLaunchAttributeClusterDimension;
    if (cl > 8) cudaFuncSetAttribute(clustered, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (; cl > 1; cl /= 2) {             // the largest cluster the device can schedule with this much shared memory
      lc.gridDim = dim3(a.n_fam * cl);
      at[0].val.clusterDim.x = cl;  at[0].val.clusterDim.y = 1;  at[0].val.clusterDim.z = 1;
      int nClusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nClusters, clustered, &lc) == cudaSuccess && nClusters >= 1) break;
      cudaGetLastError();
    }
    if (cl > 1) cudaLaunchKernelEx(&lc, clustered, b);
    else kernel<<<a.n_fam, threads, smem, st>>>(b);
  };
  auto pick = [&] (auto wqc) {
    constexpr int WQ = decltype(wqc)::value;
    if (wide) {
      if (a.cfg.kind == 0) {
        if (a.cfg.use_lut) go(resolve_families_kernel<WQ, 0, true, 256, false>, resolve_families_kernel<WQ, 0, true, 256, true>, 256);
        else go(resolve_families_kernel<WQ, 0, false, 256, false>, resolve_families_kernel<WQ, 0, false, 256, true>, 256);
      } else go(resolve_families_kernel<WQ, 1, false, 256, false>, resolve_families_kernel<WQ, 1, false, 256, true>, 256);
    } else {
      want = 1;
      if (a.cfg.kind == 0) {
        if (a.cfg.use_lut) go(resolve_families_kernel<WQ, 0, true, 128, false>, resolve_families_kernel<WQ, 0, true, 128, false>, 128);
        else go(resolve_families_kernel<WQ, 0, false, 128, false>, resolve_families_kernel<WQ, 0, false, 128, false>, 128);
      } else go(resolve_families_kernel<WQ, 1, false, 128, false>, resolve_families_kernel<WQ, 1, false, 128, false>, 128);
    }
  };
  switch (wq) {
  case 1: pick(std::integral_constant<int, 1>{}); break;
  case 2: pick(std::integral_constant<int, 2>{}); break;
  case 4: pick(std::integral_constant<int, 4>{}); break;
  case 8: pick(std::integral_constant<int, 8>{}); break;
  case 16: pick(std::integral_constant<int, 16>{}); break;
  case 32: pick(std::integral_constant<int, 32>{}); break;
  default: break;
  }
}

} // namespace apt
