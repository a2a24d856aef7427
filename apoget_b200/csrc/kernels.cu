// CUDA kernels of the species-assignment hot path (sm_100a).
//
// What they compute follows the reference's arithmetic exactly (file:line relative to
// /root/reference/src/core):
//   score        tree/phylogenetictree.hpp:456-478 (continuous) and :430-452 (semi-continuous)
//   compat       crossover.h:76-92
//   criterion    tree/enveloppecriteria.cpp:40-84 (maxAverage), :87-133 (maxMinDist)
// How they compute it is ours: SoA enveloppe pool with uint4-aligned rows, lanes striding the
// contiguous enveloppe block (coalesced 16-byte loads), __popc on the integer pipe with
// shuffle reductions, and the order-sensitive float accumulation (`float avg += double c`,
// :465-472) kept strictly sequential -- shuffles / shared memory only GATHER the k values.
#include "kernels.cuh"

#include <cooperative_groups.h>
#include "criteria.h"

#include <math_constants.h>

#include <algorithm>

namespace apt {

namespace {

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

// crossover.h:89-92; same operation order as the host (no a*b+c pattern, so nothing to contract)
__device__ __forceinline__ double compat_generic (float4 m, double d) {
  const double mu = double(m.x);
  const double s = d < mu ? double(m.y) : double(m.z);
  const double t = d - mu;
  return exp(-(t * t) / (2. * s * s));
}

template <bool LUT>
__device__ __forceinline__ double pair_compat (const ScoreCfg &cfg, float4 bm, float4 rm, double d) {
  if (LUT) {
    const uint32_t di = uint32_t(d);
    const double a = __ldg(cfg.lut + size_t(__float_as_uint(bm.w)) * cfg.lut_stride + di);
    const double b = __ldg(cfg.lut + size_t(__float_as_uint(rm.w)) * cfg.lut_stride + di);
    return b < a ? b : a;                       // std::min(a, b), :468
  } else {
    const double a = compat_generic(bm, d), b = compat_generic(rm, d);
    return b < a ? b : a;
  }
}

/// Sequential tail of the score (phylogenetictree.hpp:465-477 / :439-451) over the k gathered
/// compatibilities, evaluated redundantly by every lane (broadcast shared-memory reads).
__device__ __forceinline__ float score_tail (const ScoreCfg &cfg, const double *c, uint32_t k) {
  if (cfg.full_continuous) {
    float avg = 0.f;
    for (uint32_t r = 0; r < k; r++) avg = float(__dadd_rn(double(avg), c[r]));
    return float(__dsub_rn(double(__fdiv_rn(avg, float(k))), cfg.avg_thr));
  }
  uint32_t matable = 0;
  for (uint32_t r = 0; r < k; r++) matable += (c[r] >= cfg.compat_thr);
  return float(__dsub_rn(double(matable), __dmul_rn(cfg.sim_thr, double(k))));
}

/// One (birth, species) score by one warp.  LPR lanes share a representative (each takes every
/// LPR-th uint4 of the row), 32/LPR representatives are processed per pass.
template <int LPR, int KIND, bool LUT>
__device__ __forceinline__ float warp_score (const PoolView &pool, const ScoreCfg &cfg,
                                             const uint4 *__restrict__ bq, float4 bm, uint32_t slot,
                                             double *c_smem, float *out_d, int lane) {
  constexpr int RPI = 32 / LPR;
  const uint32_t k = min(__ldg(pool.fill + slot), pool.K);
  const uint32_t wp4 = pool.wp >> 2;
  const uint4 *rq = reinterpret_cast<const uint4*>(pool.payload) + size_t(slot) * pool.K * wp4;
  const float4 *rmeta = pool.cmeta + size_t(slot) * pool.K;
  const int g = lane / LPR, l = lane % LPR;

  for (uint32_t r0 = 0; r0 < k; r0 += RPI) {
    const uint32_t r = r0 + g;
    double d;
    if (KIND == 0) {
      uint32_t acc = 0;
      if (r < k)
        for (uint32_t v = l; v < wp4; v += LPR) acc += popc4(__ldg(rq + size_t(r) * wp4 + v), bq[v]);
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      d = double(acc);
    } else {
      double acc = 0.;
      if (r < k)
        for (uint32_t v = l; v < wp4; v += LPR) acc += sqdiff4(__ldg(rq + size_t(r) * wp4 + v), bq[v]);
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      d = sqrt(acc);
    }
    if (l == 0 && r < k) {
      c_smem[r] = pair_compat<LUT>(cfg, bm, __ldg(rmeta + r), d);
      if (out_d) out_d[r] = float(d);           // DCCache::push_back(float d, ..), treetypes.h:218-220
    }
  }
  __syncwarp();
  const float score = score_tail(cfg, c_smem, k);
  __syncwarp();
  return score;
}

constexpr int WARPS_PER_CTA = 8;

template <int LPR, int KIND, bool LUT>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
score_pairs_kernel (PoolView pool, ScoreCfg cfg, const uint32_t *__restrict__ b_payload,
                    const float4 *__restrict__ b_cmeta, uint32_t n_pairs,
                    const uint32_t *__restrict__ pair_birth, const uint32_t *__restrict__ pair_slot,
                    float *__restrict__ out_score, float *__restrict__ out_dist) {
  __shared__ double c_smem[WARPS_PER_CTA][KMAX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t p = blockIdx.x * WARPS_PER_CTA + warp;
  if (p >= n_pairs) return;
  const uint32_t b = __ldg(pair_birth + p), slot = __ldg(pair_slot + p);
  const uint4 *bq = reinterpret_cast<const uint4*>(b_payload) + size_t(b) * (pool.wp >> 2);
  const float s = warp_score<LPR, KIND, LUT>(pool, cfg, bq, __ldg(b_cmeta + b), slot, c_smem[warp],
                                             out_dist + size_t(p) * pool.K, lane);
  if (lane == 0) out_score[p] = s;
}

/// K2: sub-species walk of one rejected birth per warp (see ChildArgs)
template <int LPR, int KIND, bool LUT>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
classify_children_kernel (ChildArgs a) {
  __shared__ double c_smem[WARPS_PER_CTA][KMAX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // the number of rejected births is only known on the device: a bounded grid strides over the list
  const uint32_t nft = min(__ldg(a.ft_counter), a.max_ft);
  for (uint32_t e = blockIdx.x * WARPS_PER_CTA + warp; e < nft; e += gridDim.x * WARPS_PER_CTA) {
    const uint32_t b = __ldg(a.ft_list + 2 * e), g = __ldg(a.ft_list + 2 * e + 1);
    const uint32_t c0 = __ldg(a.child_off + g), c1 = __ldg(a.child_off + g + 1);
    const uint4 *bq = reinterpret_cast<const uint4*>(a.b_payload) + size_t(b) * (a.pool.wp >> 2);
    const float4 bm = __ldg(a.b_cmeta + b);
    float *od = a.out_dist + size_t(e) * a.pool.K;
    uint32_t first = 0xFFFFFFFFu, flags = 0;
    float score = 0.f;
    for (uint32_t c = c0; c < c1; c++) {
      const uint32_t slot = __ldg(a.child_slot + c);
      if (slot == 0xFFFFFFFFu) { first = c - c0; flags = 1u; break; }          // the host knows this one only
      score = warp_score<LPR, KIND, LUT>(a.pool, a.cfg, bq, bm, slot, c_smem[warp], od, lane);
      const bool unsure = !LUT && (fabsf(score) <= a.guard || !a.cfg.full_continuous);
      if (unsure) { first = c - c0; flags = 1u; break; }
      if (score > 0.f) { first = c - c0; break; }                              // :540-541
    }
    if (lane == 0) { a.out_first[e] = first;  a.out_flags[e] = flags;  a.out_score[e] = score; }
    __syncwarp();
  }
}

/// Dense sweep, first version: one warp per birth walking every slot of the pool.
template <int LPR, int KIND, bool LUT>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
classify_dense_kernel (PoolView pool, ScoreCfg cfg, const uint32_t *__restrict__ b_payload,
                       const float4 *__restrict__ b_cmeta, uint32_t n_births,
                       uint32_t *__restrict__ out_best, float *__restrict__ out_best_score,
                       uint32_t *__restrict__ out_first, uint32_t *__restrict__ out_count) {
  __shared__ double c_smem[WARPS_PER_CTA][KMAX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t b = blockIdx.x * WARPS_PER_CTA + warp;
  if (b >= n_births) return;
  const uint4 *bq = reinterpret_cast<const uint4*>(b_payload) + size_t(b) * (pool.wp >> 2);
  const float4 bm = __ldg(b_cmeta + b);
  uint32_t best = 0xFFFFFFFFu, first = 0xFFFFFFFFu, count = 0;
  float bestScore = -CUDART_INF_F;
  for (uint32_t s = 0; s < pool.n_slots; s++) {
    if (__ldg(pool.fill + s) == 0) continue;
    const float sc = warp_score<LPR, KIND, LUT>(pool, cfg, bq, bm, s, c_smem[warp], nullptr, lane);
    if (best == 0xFFFFFFFFu || bestScore < sc) { best = s; bestScore = sc; }
    if (sc > 0.f) { count++; if (first == 0xFFFFFFFFu) first = s; }
  }
  if (lane == 0) {
    out_best[b] = best;  out_best_score[b] = bestScore;  out_first[b] = first;  out_count[b] = count;
  }
}

template <int KIND>
__global__ void pairwise_kernel (uint32_t wp4, const uint4 *__restrict__ payload, uint32_t n_pairs,
                                 const uint32_t *__restrict__ ia, const uint32_t *__restrict__ ib,
                                 float *__restrict__ out) {
  const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (p >= n_pairs) return;
  const uint4 *a = payload + size_t(__ldg(ia + p)) * wp4, *b = payload + size_t(__ldg(ib + p)) * wp4;
  if (KIND == 0) {
    uint32_t acc = 0;
    for (uint32_t v = lane; v < wp4; v += 32) acc += popc4(__ldg(a + v), __ldg(b + v));
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[p] = float(acc);
  } else {
    double acc = 0.;
    for (uint32_t v = lane; v < wp4; v += 32) acc += sqdiff4(__ldg(a + v), __ldg(b + v));
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[p] = float(sqrt(acc));
  }
}

// crossover.h:234-240: dist = distance(mother, father); compat = mother.compatibility(dist)
template <int KIND>
__global__ void mate_compat_kernel (uint32_t wp4, const uint4 *__restrict__ payload, const float4 *__restrict__ cmeta,
                                    uint32_t n_pairs, const uint32_t *__restrict__ im, const uint32_t *__restrict__ ifa,
                                    float *__restrict__ out_dist, double *__restrict__ out_compat) {
  const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (p >= n_pairs) return;
  const uint32_t m = __ldg(im + p);
  const uint4 *a = payload + size_t(m) * wp4, *b = payload + size_t(__ldg(ifa + p)) * wp4;
  double d;
  if (KIND == 0) {
    uint32_t acc = 0;
    for (uint32_t v = lane; v < wp4; v += 32) acc += popc4(__ldg(a + v), __ldg(b + v));
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    d = double(acc);
  } else {
    double acc = 0.;
    for (uint32_t v = lane; v < wp4; v += 32) acc += sqdiff4(__ldg(a + v), __ldg(b + v));
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    d = sqrt(acc);
  }
  if (lane == 0) { out_dist[p] = float(d);  out_compat[p] = compat_generic(__ldg(cmeta + m), d); }
}

// enveloppecriteria.cpp:40-84 (crit 0), :87-133 (crit 1), :136-202 (crit 2), :205-272 (crit 3); one thread per candidate genome
__global__ void env_contribution_kernel (int crit, uint32_t K, uint32_t n, const float *__restrict__ edist,
                                         const float *__restrict__ gdist, uint8_t *__restrict__ out_better,
                                         uint32_t *__restrict__ out_than, float *__restrict__ out_value) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *e = edist + size_t(i) * K * K, *g = gdist + size_t(i) * K;
  if (crit >= 2) {
    // :136-202 (exact) / :205-272 (device exp(): decisions within 1e-13 of a tie may differ from libm's)
    auto D = [&] (uint32_t a, uint32_t b) { return e[a * K + b]; };
    auto G = [&] (uint32_t j) { return g[j]; };
    const CritResult r = crit == 2 ? crit_maxAvgMinStdDev(D, G, K)
                                   : crit_maxWeightedDist2Avg<KMAX>(D, G, K, [] (double x) { return exp(x); });
    out_better[i] = r.better;  out_than[i] = r.than;  out_value[i] = r.value;
    return;
  }
  float value = -CUDART_INF_F;      // -DBL_MAX stored into a float, :46 / :93
  uint32_t than = 0xFFFFFFFFu;
  for (uint32_t a = 0; a < K; a++) {
    double c;
    if (crit == 1) {
      float minBase = 3.402823466e+38f, minNew = 3.402823466e+38f;
      for (uint32_t j = 0; j < K; j++) {
        if (j == a) continue;
        minBase = fminf(minBase, e[a * K + j]);
        minNew = fminf(minNew, g[j]);
      }
      c = double(__fadd_rn(-minBase, minNew));                       // :117
    } else {
      c = 0.;
      for (uint32_t j = 0; j < K; j++) {
        if (j == a) continue;
        c = __dadd_rn(c, double(__fadd_rn(-e[a * K + j], g[j])));   // :62
      }
    }
    if (double(value) < c) { value = float(c); than = a; }           // :77-80 / :125-128
  }
  out_better[i] = value > 0.f;                                        // :83 / :131
  out_than[i] = than;
  out_value[i] = value;
}

__global__ void pack_births_kernel (const uint32_t *__restrict__ src, const float *__restrict__ compat,
                                    const uint32_t *__restrict__ cls, uint32_t n, uint32_t words, uint32_t wp,
                                    uint32_t *__restrict__ dst, float4 *__restrict__ meta) {
  const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = size_t(n) * wp;
  if (t < total) {
    const uint32_t row = uint32_t(t / wp), w = uint32_t(t % wp);
    dst[t] = w < words ? src[size_t(row) * words + w] : 0u;
  }
  if (t < n)
    meta[t] = make_float4(compat[3 * t], compat[3 * t + 1], compat[3 * t + 2],
                          __uint_as_float(cls ? cls[t] : 0u));
}

__global__ void fill_f4_kernel (float4 *__restrict__ dst, uint32_t n, float4 value) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = value;
}

__global__ void pad_rows_kernel (const uint32_t *__restrict__ src, uint32_t n, uint32_t words, uint32_t wp,
                                 uint32_t *__restrict__ dst) {
  const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t < size_t(n) * wp) {
    const uint32_t row = uint32_t(t / wp), w = uint32_t(t % wp);
    dst[t] = w < words ? src[size_t(row) * words + w] : 0u;
  }
}

__global__ void gather_rows_kernel (const uint32_t *__restrict__ src, const uint32_t *__restrict__ idx, uint32_t n,
                                    uint32_t words, uint32_t *__restrict__ dst) {
  const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t < size_t(n) * words) dst[t] = src[size_t(idx[t / words]) * words + t % words];
}

__global__ void scatter_rows_kernel (const uint4 *__restrict__ sp, const float4 *__restrict__ sm,
                                     const uint32_t *__restrict__ dst_row, uint32_t n_rows, uint32_t wp4,
                                     uint4 *__restrict__ pp, float4 *__restrict__ pm) {
  const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = size_t(n_rows) * wp4;
  if (t < total) {
    const uint32_t row = uint32_t(t / wp4), v = uint32_t(t % wp4);
    pp[size_t(dst_row[row]) * wp4 + v] = sp[t];
  }
  if (t < n_rows) pm[dst_row[t]] = sm[t];
}

__global__ void scatter_u32_kernel (const uint32_t *__restrict__ idx, const uint32_t *__restrict__ val, uint32_t n,
                                    uint32_t *__restrict__ dst) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] = val[i];
}

__global__ void scatter_blocks_f32_kernel (const float *__restrict__ src, const uint32_t *__restrict__ idx, uint32_t n,
                                           uint32_t len, float *__restrict__ dst) {
  const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t < size_t(n) * len) dst[size_t(idx[t / len]) * len + t % len] = src[t];
}

int lanes_per_rep (uint32_t wp) {
  uint32_t q = wp >> 2, l = 1;
  while (l < q && l < 32) l <<= 1;
  return int(l);
}

template <template <int, int, bool> class Launcher, typename... Args>
void dispatch (int lpr, int kind, bool lut, Args&&... args) {
#define APT_CASE(L)                                                                   \
  case L:                                                                             \
    if (kind == 0) { if (lut) Launcher<L, 0, true>::run(args...); else Launcher<L, 0, false>::run(args...); } \
    else Launcher<L, 1, false>::run(args...);                                        \
    break;
  switch (lpr) { APT_CASE(1) APT_CASE(2) APT_CASE(4) APT_CASE(8) APT_CASE(16) APT_CASE(32) }
#undef APT_CASE
}

template <int LPR, int KIND, bool LUT>
struct ScorePairsLauncher {
  static void run (const PoolView &pool, const ScoreCfg &cfg, const uint32_t *bp, const float4 *bm, uint32_t n,
                   const uint32_t *pb, const uint32_t *ps, float *os, float *od, cudaStream_t st) {
    const uint32_t grid = (n + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    score_pairs_kernel<LPR, KIND, LUT><<<grid, WARPS_PER_CTA * 32, 0, st>>>(pool, cfg, bp, bm, n, pb, ps, os, od);
  }
};

template <int LPR, int KIND, bool LUT>
struct ChildrenLauncher {
  static void run (const ChildArgs &a, cudaStream_t st) {
    // at most 8 CTAs per SM of a B200: an empty list (the common case) costs one small launch, not max_ft warps
    const uint32_t grid = std::min<uint32_t>((a.max_ft + WARPS_PER_CTA - 1) / WARPS_PER_CTA, 148u * 8u);
    classify_children_kernel<LPR, KIND, LUT><<<grid, WARPS_PER_CTA * 32, 0, st>>>(a);
  }
};

template <int LPR, int KIND, bool LUT>
struct DenseLauncher {
  static void run (const PoolView &pool, const ScoreCfg &cfg, const uint32_t *bp, const float4 *bm, uint32_t n,
                   uint32_t *ob, float *obs, uint32_t *of, uint32_t *oc, cudaStream_t st) {
    const uint32_t grid = (n + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    classify_dense_kernel<LPR, KIND, LUT><<<grid, WARPS_PER_CTA * 32, 0, st>>>(pool, cfg, bp, bm, n, ob, obs, of, oc);
  }
};

} // namespace

void launch_score_pairs (const PoolView &pool, const ScoreCfg &cfg, const uint32_t *b_payload,
                         const float4 *b_cmeta, uint32_t n_pairs, const uint32_t *pair_birth,
                         const uint32_t *pair_slot, float *out_score, float *out_dist, cudaStream_t stream) {
  if (!n_pairs) return;
  dispatch<ScorePairsLauncher>(lanes_per_rep(pool.wp), cfg.kind, cfg.use_lut != 0, pool, cfg, b_payload, b_cmeta,
                               n_pairs, pair_birth, pair_slot, out_score, out_dist, stream);
}

void launch_classify_children (const ChildArgs &a, cudaStream_t stream) {
  if (!a.max_ft) return;
  dispatch<ChildrenLauncher>(lanes_per_rep(a.pool.wp), a.cfg.kind, a.cfg.use_lut != 0, a, stream);
}

void launch_classify_dense (const PoolView &pool, const ScoreCfg &cfg, const uint32_t *b_payload,
                            const float4 *b_cmeta, uint32_t n_births, uint32_t *out_best_species,
                            float *out_best_score, uint32_t *out_first_accept, uint32_t *out_n_accept,
                            cudaStream_t stream) {
  if (!n_births) return;
  if (dense_tiled_supports(pool.wp, pool.K)) {
    launch_classify_dense_tiled(pool, cfg, b_payload, b_cmeta, n_births, out_best_species, out_best_score,
                                out_first_accept, out_n_accept, stream);
    return;
  }
  dispatch<DenseLauncher>(lanes_per_rep(pool.wp), cfg.kind, cfg.use_lut != 0, pool, cfg, b_payload, b_cmeta,
                          n_births, out_best_species, out_best_score, out_first_accept, out_n_accept, stream);
}

namespace {
__global__ void publish_records_kernel (const uint32_t *__restrict__ best, const float *__restrict__ score,
                                        const uint32_t *__restrict__ first, const uint32_t *__restrict__ count,
                                        uint32_t n, ShardSink sink) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  const uint32_t off = sink.species_offset, be = best[b], fi = first[b];
  const uint4 rec = make_uint4(be == 0xFFFFFFFFu ? be : be + off, __float_as_uint(score[b]),
                               fi == 0xFFFFFFFFu ? fi : fi + off, count[b]);
  #pragma unroll
  for (uint32_t p = 0; p < MAX_PEERS; p++)
    if (p < sink.world) sink.peer[p][size_t(sink.rank) * n + b] = rec;
}

__global__ void combine_shards_kernel (const uint4 *__restrict__ rec, uint32_t n, uint32_t world,
                                       uint32_t *__restrict__ out_best, float *__restrict__ out_score,
                                       uint32_t *__restrict__ out_first, uint32_t *__restrict__ out_count) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  uint32_t best = 0xFFFFFFFFu, first = 0xFFFFFFFFu, count = 0;
  float score = -CUDART_INF_F;
  for (uint32_t r = 0; r < world; r++) {
    const uint4 q = rec[size_t(r) * n + b];
    if (q.x != 0xFFFFFFFFu) {
      const float s = __uint_as_float(q.y);
      if (best == 0xFFFFFFFFu || score < s) { best = q.x; score = s; }     // shards hold increasing species ranges
    }
    first = min(first, q.z);
    count += q.w;
  }
  out_best[b] = best;  out_score[b] = score;  out_first[b] = first;  out_count[b] = count;
}
} // namespace

void launch_publish_records (const uint32_t *best, const float *best_score, const uint32_t *first, const uint32_t *count,
                             uint32_t n_births, const ShardSink &sink, cudaStream_t stream) {
  if (!n_births || !sink.world) return;
  publish_records_kernel<<<(n_births + 255) / 256, 256, 0, stream>>>(best, best_score, first, count, n_births, sink);
}

void launch_combine_shards (const uint4 *rec, uint32_t n_births, uint32_t world, uint32_t *out_best_species,
                            float *out_best_score, uint32_t *out_first_accept, uint32_t *out_n_accept,
                            cudaStream_t stream) {
  if (!n_births) return;
  combine_shards_kernel<<<(n_births + 255) / 256, 256, 0, stream>>>(rec, n_births, world, out_best_species, out_best_score,
                                                                    out_first_accept, out_n_accept);
}

void launch_pack_births (const uint32_t *src_payload, const float *src_compat, const uint32_t *cls, uint32_t n,
                         uint32_t words, uint32_t wp, uint32_t *dst_payload, float4 *dst_cmeta,
                         cudaStream_t stream) {
  if (!n) return;
  const size_t total = size_t(n) * wp;
  pack_births_kernel<<<uint32_t((total + 255) / 256), 256, 0, stream>>>(src_payload, src_compat, cls, n, words, wp,
                                                                       dst_payload, dst_cmeta);
}

void launch_fill_f4 (float4 *dst, uint32_t n, float4 value, cudaStream_t stream) {
  if (!n) return;
  fill_f4_kernel<<<(n + 255) / 256, 256, 0, stream>>>(dst, n, value);
}

void launch_pad_rows (const uint32_t *src, uint32_t n, uint32_t words, uint32_t wp, uint32_t *dst, cudaStream_t stream) {
  if (!n) return;
  const size_t total = size_t(n) * wp;
  pad_rows_kernel<<<uint32_t((total + 255) / 256), 256, 0, stream>>>(src, n, words, wp, dst);
}

/// Rows of the births a device-side list names: entry e of `list` is the pair (birth, stream), *count entries
/// (a device counter: the host does not know it yet); grid-stride over (entry, word).
__global__ void gather_listed_rows_kernel (const uint32_t *__restrict__ src, const uint32_t *__restrict__ list,
                                           const uint32_t *__restrict__ count, uint32_t max_entries, uint32_t words,
                                           uint32_t *__restrict__ dst) {
  const size_t total = size_t(min(*count, max_entries)) * words;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const size_t e = i / words, w = i - e * words;
    dst[i] = src[size_t(list[2 * e]) * words + w];
  }
}

// ---- the births of a pass the host has to look at, in arrival order ------------------------------------------------
// A birth is settled on the device when it was accepted with the enveloppe left alone -- at top level (stream kernel) or
// by a sub-species (family pass).  Everything else (enveloppe entries, rejected births, births outside the streams whose
// result word is still clear, streams handed back) is listed as {birth, result word}: the host reads this list instead of
// scanning one result word per birth.  Two small launches: per-block counts, then the ordered write.
constexpr uint32_t LIST_THREADS = 256, LIST_PER_THREAD = 8, LIST_PER_BLOCK = LIST_THREADS * LIST_PER_THREAD;

template <int MODE>
__device__ __forceinline__ bool birth_unsettled (uint32_t act, const uint32_t *fam_act, uint32_t i) {
  if (MODE == 1) return act == 0xFFFFFFFFu;           // (grouping: `act` is the birth's stream key; none = outside the streams)
  if (act == ACT_ACCEPTED) return false;
  return !(fam_act && act == ACT_FALLTHROUGH && (fam_act[i] & FAM_CODE_MASK) == FAM_JOINED);
}

__device__ __forceinline__ uint32_t block_sum (uint32_t v, uint32_t *warpSums) {      // every thread gets the total
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) warpSums[warp] = v;
  __syncthreads();
  uint32_t t = 0;
  for (uint32_t w = 0; w < LIST_THREADS / 32; w++) t += warpSums[w];
  __syncthreads();
  return t;
}

template <int MODE>
__global__ void __launch_bounds__(LIST_THREADS)
unsettled_count_kernel (const uint32_t *act, const uint32_t *fam_act, uint32_t n, uint32_t *block_count) {
  __shared__ uint32_t warpSums[LIST_THREADS / 32];
  const uint32_t i0 = blockIdx.x * LIST_PER_BLOCK + threadIdx.x * LIST_PER_THREAD;
  uint32_t c = 0;
#pragma unroll
  for (uint32_t j = 0; j < LIST_PER_THREAD; j++)
    if (i0 + j < n) c += birth_unsettled<MODE>(act[i0 + j], fam_act, i0 + j);
  const uint32_t t = block_sum(c, warpSums);
  if (threadIdx.x == 0) block_count[blockIdx.x] = t;
}

template <int MODE>
__global__ void __launch_bounds__(LIST_THREADS)
unsettled_list_kernel (const uint32_t *act, const uint32_t *fam_act, uint32_t n, const uint32_t *block_count,
                       uint32_t *list, uint32_t *total) {
  __shared__ uint32_t warpSums[LIST_THREADS / 32];
  __shared__ uint32_t warpOff[LIST_THREADS / 32];
  // entries of the blocks in front of this one
  uint32_t before = 0;
  for (uint32_t b = threadIdx.x; b < blockIdx.x; b += LIST_THREADS) before += block_count[b];
  before = block_sum(before, warpSums);
  const uint32_t i0 = blockIdx.x * LIST_PER_BLOCK + threadIdx.x * LIST_PER_THREAD;
  uint32_t a[LIST_PER_THREAD], flags = 0, c = 0;
#pragma unroll
  for (uint32_t j = 0; j < LIST_PER_THREAD; j++) {
    a[j] = 0;
    if (i0 + j < n) {
      a[j] = act[i0 + j];
      if (birth_unsettled<MODE>(a[j], fam_act, i0 + j)) { flags |= 1u << j; c++; }
    }
  }
  // exclusive scan of the threads' counts: inside the warp, then across the warps
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= uint32_t(o)) incl += v; }
  if (lane == 31) warpSums[warp] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (uint32_t w = 0; w < LIST_THREADS / 32; w++) { warpOff[w] = run; run += warpSums[w]; }
    if (blockIdx.x == gridDim.x - 1) *total = before + run;
  }
  __syncthreads();
  uint32_t at = before + warpOff[warp] + incl - c;
#pragma unroll
  for (uint32_t j = 0; j < LIST_PER_THREAD; j++)
    if ((flags >> j) & 1u) { list[2 * size_t(at)] = i0 + j;  list[2 * size_t(at) + 1] = a[j];  at++; }
}

uint32_t list_unsettled_blocks (uint32_t n) { return (n + LIST_PER_BLOCK - 1) / LIST_PER_BLOCK; }

void launch_list_unsettled (const uint32_t *act, const uint32_t *fam_act, uint32_t n, uint32_t *block_count,
                            uint32_t *list, uint32_t *total, cudaStream_t stream) {
  if (!n) return;
  const uint32_t blocks = list_unsettled_blocks(n);
  unsettled_count_kernel<0><<<blocks, LIST_THREADS, 0, stream>>>(act, fam_act, n, block_count);
  unsettled_list_kernel<0><<<blocks, LIST_THREADS, 0, stream>>>(act, fam_act, n, block_count, list, total);
}

// ---- grouping the births of a pass into per-species streams on the device -------------------------------------------
// What Engine::groupParallel does on the host (a stable counting sort of the single-parent births by species id), for
// callers whose id arrays sit in page-locked memory.  ONE cooperative launch, phases separated by grid barriers (separate
// launches and copies cost more in gaps than the work itself: 90 us for 30 us of work):
//   0. keys + per-species totals (global atomics); births outside the streams are appended to a list; the id arrays are
//      read where they are (page-locked host memory, or a device copy)
//   1. block 0 scans the species: where each stream starts, its rank among the streams, the summary -- written to the
//      host directly -- and clears the counters of the NEXT call (two sets, used in turn)
//   2. scatter through per-species cursors (order inside a stream arbitrary); totals and the rest list go to the host
//   3. every stream is put back into arrival order (a block sorts one by birth index in shared memory; a stream too long
//      for that makes the host group the pass instead) and written out for the stream kernel and for the host
constexpr uint32_t GROUP_SORT_MAX = 2048, GROUP_THREADS = 256;

struct GroupArgs {
  const uint32_t *msid, *fsid;      ///< [n] (device-accessible)
  uint32_t n;
  const uint32_t *sid_slot;         ///< [ns]
  uint32_t ns, root_sid;
  uint32_t *key;                    ///< [n] scratch
  uint32_t *total, *total_next;     ///< [ns] counters of this call (zero on entry) / of the next one (cleared here)
  uint32_t *start, *cursor, *gidx;  ///< [ns]
  uint32_t *info, *info_next;       ///< [8]
  uint32_t *order;                  ///< [n] device: what the stream kernel reads
  uint32_t *rest;                   ///< [n] device scratch: births outside the streams
  // results for the host, written in place (page-locked memory)
  uint32_t *h_total, *h_info, *h_rest, *h_order;
  int32_t *h_birth_group;
};

__global__ void __launch_bounds__(GROUP_THREADS)
group_births_kernel (GroupArgs a) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ uint32_t sm[GROUP_SORT_MAX];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t gsize = gridDim.x * GROUP_THREADS, gtid = blockIdx.x * GROUP_THREADS + tid;

  // ---- 0: keys, totals, rest list ----
  for (uint32_t i = gtid; i < a.n; i += gsize) {
    const uint32_t m = a.msid[i], f = a.fsid[i];
    uint32_t k = 0xFFFFFFFFu;
    if (f == 0xFFFFFFFFu || m == f) {
      if (m == 0xFFFFFFFFu) k = a.root_sid;                               // primordial: the root's stream, contributes nothing
      else if (m < a.ns && __ldg(a.sid_slot + m) != 0xFFFFFFFFu) k = m | ((m == f ? 2u : 1u) << ORDER_CLASS_SHIFT);
    }
    a.key[i] = k;
    if (k != 0xFFFFFFFFu) atomicAdd(a.total + (k & ORDER_INDEX_MASK), 1u);
    else a.rest[atomicAdd(a.info + 4, 1u)] = i;                           // (few: the host puts them back into arrival order)
  }
  grid.sync();

  // ---- 1: block 0 scans the species ----
  if (blockIdx.x == 0) {
    uint32_t *wsum = sm, *wcnt = sm + 8, *wmax = sm + 16, *carry = sm + 24;      // carry: sum, count, max
    if (tid < 3) carry[tid] = 0u;
    __syncthreads();
    for (uint32_t base = 0; base < a.ns; base += GROUP_THREADS) {
      const uint32_t sid = base + tid;
      const uint32_t t = sid < a.ns ? a.total[sid] : 0u, nz = t ? 1u : 0u;
      uint32_t is = t, ic = nz, mx = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, is, o), y = __shfl_up_sync(0xffffffffu, ic, o);
        if (lane >= uint32_t(o)) { is += x; ic += y; }
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      }
      if (lane == 31) { wsum[warp] = is; wcnt[warp] = ic; }
      if (lane == 0) wmax[warp] = mx;
      __syncthreads();
      uint32_t os = carry[0], oc = carry[1];
      for (uint32_t w = 0; w < warp; w++) { os += wsum[w]; oc += wcnt[w]; }
      if (sid < a.ns) {
        const uint32_t st = os + is - t;
        a.start[sid] = st;  a.cursor[sid] = st;  a.gidx[sid] = oc + ic - nz;
        a.total_next[sid] = 0u;                 // the NEXT call's counters: no clearing launch
      }
      __syncthreads();
      if (tid == GROUP_THREADS - 1) { carry[0] = os + is; carry[1] = oc + ic; }
      if (tid == 0) { uint32_t m = carry[2]; for (uint32_t w = 0; w < GROUP_THREADS / 32; w++) m = max(m, wmax[w]); carry[2] = m; }
      __syncthreads();
    }
    if (tid == 0) {
      const uint32_t over = carry[2] > GROUP_SORT_MAX ? 1u : 0u, nrest = a.info[4];
      a.info[0] = carry[1];  a.info[1] = carry[2];  a.info[2] = carry[0];  a.info[3] = over;
      a.h_info[0] = carry[1];  a.h_info[1] = carry[2];  a.h_info[2] = carry[0];  a.h_info[3] = over;  a.h_info[4] = nrest;
    }
    if (tid < 8) a.info_next[tid] = 0u;
  }
  grid.sync();

  // ---- 2: scatter; totals and the rest list to the host ----
  for (uint32_t i = gtid; i < a.n; i += gsize) {
    const uint32_t k = a.key[i];
    if (k == 0xFFFFFFFFu) { a.h_birth_group[i] = -1; continue; }
    const uint32_t sid = k & ORDER_INDEX_MASK;
    a.order[atomicAdd(a.cursor + sid, 1u)] = i | (k & ~ORDER_INDEX_MASK);
    a.h_birth_group[i] = int32_t(a.gidx[sid]);
  }
  for (uint32_t sid = gtid; sid < a.ns; sid += gsize) a.h_total[sid] = a.total[sid];
  { const uint32_t nrest = a.info[4];  for (uint32_t e = gtid; e < nrest; e += gsize) a.h_rest[e] = a.rest[e]; }
  if (a.info[3]) return;                        // (uniform: the host groups this pass itself)
  grid.sync();

  // ---- 3: every stream back into arrival order ----
  for (uint32_t sid = blockIdx.x; sid < a.ns; sid += gridDim.x) {
    const uint32_t t = a.total[sid];
    if (!t) continue;
    uint32_t *o = a.order + a.start[sid], *ho = a.h_order + a.start[sid];
    if (t == 1) { if (tid == 0) ho[0] = o[0]; continue; }
    uint32_t p2 = 2;
    while (p2 < t) p2 <<= 1;
    for (uint32_t i = tid; i < p2; i += GROUP_THREADS) sm[i] = i < t ? o[i] : 0xFFFFFFFFu;
    __syncthreads();
    for (uint32_t k = 2; k <= p2; k <<= 1)
      for (uint32_t j = k >> 1; j > 0; j >>= 1) {
        for (uint32_t i = tid; i < p2; i += GROUP_THREADS) {
          const uint32_t x = i ^ j;
          if (x > i) {
            const uint32_t u = sm[i], v = sm[x];
            // (padding entries are all ones: larger than any index; real entries compare by birth index)
            const uint32_t ku = u == 0xFFFFFFFFu ? u : (u & ORDER_INDEX_MASK), kv = v == 0xFFFFFFFFu ? v : (v & ORDER_INDEX_MASK);
            if ((ku > kv) == ((i & k) == 0)) { sm[i] = v; sm[x] = u; }
          }
        }
        __syncthreads();
      }
    for (uint32_t i = tid; i < t; i += GROUP_THREADS) { const uint32_t v = sm[i];  o[i] = v;  ho[i] = v; }
    __syncthreads();
  }
}

void launch_group_births (const uint32_t *msid, const uint32_t *fsid, uint32_t n, const uint32_t *sid_slot, uint32_t ns,
                          uint32_t root_sid, uint32_t *key, uint32_t *total, uint32_t *total_next, uint32_t *start, uint32_t *cursor,
                          uint32_t *gidx, uint32_t *info, uint32_t *info_next, uint32_t *order, uint32_t *rest, uint32_t *h_total,
                          uint32_t *h_info, uint32_t *h_rest, uint32_t *h_order, int32_t *h_birth_group, cudaStream_t stream) {
  if (!n || !ns) return;
  static int perSm = 0, sms = 0;
  if (!perSm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, group_births_kernel, int(GROUP_THREADS), 0);
    if (perSm < 1) perSm = 1;
  }
  GroupArgs a { msid, fsid, n, sid_slot, ns, root_sid, key, total, total_next, start, cursor, gidx, info, info_next, order, rest,
                h_total, h_info, h_rest, h_order, h_birth_group };
  // every block must be resident at once (grid barriers): at most what the device holds, at least enough for the births
  const uint32_t want = std::max<uint32_t>((n + GROUP_THREADS * 4 - 1) / (GROUP_THREADS * 4), std::min<uint32_t>(ns, 1024u));
  const uint32_t grid = std::max(1u, std::min<uint32_t>(want, uint32_t(sms) * uint32_t(std::min(perSm, 4))));
  void *args[] = { &a };
  cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(group_births_kernel), dim3(grid), dim3(GROUP_THREADS), args, 0, stream);
}

void launch_gather_listed_rows (const uint32_t *src, const uint32_t *list, const uint32_t *count, uint32_t max_entries,
                                uint32_t words, uint32_t *dst, cudaStream_t stream) {
  if (!max_entries) return;
  const size_t total = size_t(max_entries) * words;
  const uint32_t blocks = uint32_t(std::min<size_t>((total + 255) / 256, 148 * 8));
  gather_listed_rows_kernel<<<blocks, 256, 0, stream>>>(src, list, count, max_entries, words, dst);
}

void launch_gather_rows (const uint32_t *src, const uint32_t *idx, uint32_t n, uint32_t words, uint32_t *dst,
                         cudaStream_t stream) {
  if (!n) return;
  const size_t total = size_t(n) * words;
  gather_rows_kernel<<<uint32_t((total + 255) / 256), 256, 0, stream>>>(src, idx, n, words, dst);
}

void launch_scatter_rows (const uint32_t *stage_payload, const float4 *stage_cmeta, const uint32_t *dst_row,
                          uint32_t n_rows, uint32_t wp, uint32_t *pool_payload, float4 *pool_cmeta,
                          cudaStream_t stream) {
  if (!n_rows) return;
  const size_t total = size_t(n_rows) * (wp >> 2);
  scatter_rows_kernel<<<uint32_t((total + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(stage_payload), stage_cmeta, dst_row, n_rows, wp >> 2,
      reinterpret_cast<uint4*>(pool_payload), pool_cmeta);
}

void launch_scatter_u32 (const uint32_t *idx, const uint32_t *val, uint32_t n, uint32_t *dst, cudaStream_t stream) {
  if (!n) return;
  scatter_u32_kernel<<<(n + 255) / 256, 256, 0, stream>>>(idx, val, n, dst);
}

void launch_scatter_blocks_f32 (const float *src, const uint32_t *idx, uint32_t n, uint32_t len, float *dst,
                                cudaStream_t stream) {
  if (!n) return;
  const size_t total = size_t(n) * len;
  scatter_blocks_f32_kernel<<<uint32_t((total + 255) / 256), 256, 0, stream>>>(src, idx, n, len, dst);
}

void launch_env_contribution (int crit, uint32_t K, uint32_t n, const float *edist, const float *gdist,
                              uint8_t *out_better, uint32_t *out_than, float *out_value, cudaStream_t stream) {
  if (!n) return;
  env_contribution_kernel<<<(n + 127) / 128, 128, 0, stream>>>(crit, K, n, edist, gdist, out_better, out_than, out_value);
}

void launch_mate_compat (int kind, uint32_t wp, const uint32_t *payload, const float4 *cmeta, uint32_t n_pairs,
                         const uint32_t *idx_mother, const uint32_t *idx_father, float *out_dist, double *out_compat,
                         cudaStream_t stream) {
  if (!n_pairs) return;
  const uint32_t grid = uint32_t((size_t(n_pairs) * 32 + 255) / 256);
  if (kind == 0)
    mate_compat_kernel<0><<<grid, 256, 0, stream>>>(wp >> 2, reinterpret_cast<const uint4*>(payload), cmeta, n_pairs,
                                                    idx_mother, idx_father, out_dist, out_compat);
  else
    mate_compat_kernel<1><<<grid, 256, 0, stream>>>(wp >> 2, reinterpret_cast<const uint4*>(payload), cmeta, n_pairs,
                                                    idx_mother, idx_father, out_dist, out_compat);
}

void launch_pairwise (int kind, uint32_t wp, const uint32_t *payload, uint32_t n_pairs, const uint32_t *idx_a,
                      const uint32_t *idx_b, float *out, cudaStream_t stream) {
  if (!n_pairs) return;
  const uint32_t grid = uint32_t((size_t(n_pairs) * 32 + 255) / 256);
  if (kind == 0)
    pairwise_kernel<0><<<grid, 256, 0, stream>>>(wp >> 2, reinterpret_cast<const uint4*>(payload), n_pairs, idx_a, idx_b, out);
  else
    pairwise_kernel<1><<<grid, 256, 0, stream>>>(wp >> 2, reinterpret_cast<const uint4*>(payload), n_pairs, idx_a, idx_b, out);
}

} // namespace apt
