// classify_dense_tiled_kernel -- every birth against every species of the pool, with the score,
// the acceptance test and the per-birth winners fused so that nothing but 16 bytes per birth
// leaves the chip (a 1 M x 100 k sweep is 1e11 distances: they cannot be materialised).
//
// Arithmetic per (birth, species) is the reference's (phylogenetictree.hpp:456-478 / :430-452);
// the schedule is a register-tiled "outer product" over the integer pipe:
//   * a CTA owns 64 births (staged once in shared memory) and streams the pool through shared
//     memory in chunks of whole species (<= 64 representative rows, contiguous in HBM);
//   * each of the 256 threads accumulates a 4 x 4 tile of (birth, representative) distances in
//     registers: per 16-byte step it reads 4 + 4 uint4 from shared memory and issues 64
//     XOR+POPC (bit-strings) or 64 x 2 FP32 ops (float vectors) -- 8 operand bytes per popc32
//     instead of 8 per popc32 *per pair* -- rows are padded to an odd number of uint4 so the
//     16 row addresses of a warp fall in different bank groups;
//   * distances go to a padded shared tile; one thread per (birth, species) -- two species at a
//     time, so that two dependency chains are in flight -- then walks the k representatives IN
//     ORDER (the float accumulation is order sensitive), applies the compatibility table and the
//     threshold, and folds the score into per-birth running winners.
//
// The reference's `float avg += double c` (:465-472) is float(double(avg) + c) per representative.
// Here the running average lives in a double that always holds a float value, and the rounding to
// float precision is done on the FP64 pipe (round_to_float, round24.h: exact, tested on the host
// against the hardware conversion) -- the two conversions per representative would otherwise share
// the quarter-rate pipe with POPC, which is the one this kernel is bound by.
#include "kernels.cuh"
#include "round24.h"

#include <math_constants.h>

#include <type_traits>

namespace apt {

namespace {

constexpr int TB = 64;          // births per CTA
constexpr int TR = 64;          // representative rows per chunk
constexpr int BT = 4, RT = 4;   // register tile
constexpr int DPAD = TB + 4;    // distance tile row pitch (words): a warp's 8 x 4 tile corners land in 32 different banks
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void cp_async16 (void *smem, const void *gmem) {
  const uint32_t s = uint32_t(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4 (void *smem, const void *gmem) {
  const uint32_t s = uint32_t(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" :: "r"(s), "l"(gmem));
}
// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier ----
__device__ __forceinline__ void mbar_init (uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(uint32_t(__cvta_generic_to_shared(bar))), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx (uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
               :: "r"(uint32_t(__cvta_generic_to_shared(bar))), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait (uint64_t *bar, uint32_t parity) {
  const uint32_t a = uint32_t(__cvta_generic_to_shared(bar));
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" :: "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s (void *smem, const void *gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               :: "r"(uint32_t(__cvta_generic_to_shared(smem))), "l"(gmem), "r"(bytes),
                  "r"(uint32_t(__cvta_generic_to_shared(bar))) : "memory");
}
__device__ __forceinline__ void cp_async_commit (void) { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all (void) { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_one (void) { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }

/// Full adder on 32 bit lanes: one LOP3 each (truth tables 0x96 = a ^ b ^ c, 0xE8 = majority)
__device__ __forceinline__ uint32_t xor3 (uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t maj3 (uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

/// a * b + c as an integer multiply-add (FMA pipe)
__device__ __forceinline__ uint32_t imad (uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

/// Hamming distance of two 128-bit blocks with THREE population counts instead of four: one carry-save
/// step, popc(x) + popc(y) + popc(z) = popc(x ^ y ^ z) + 2 popc(maj(x, y, z)), moves a quarter of the
/// work from the 16-lane POPC pipe (the one this kernel is bound by) to LOP3 on the 64-lane ALU pipe.
__device__ __forceinline__ uint32_t hamming4 (uint4 a, uint4 b) {
  const uint32_t x = a.x ^ b.x, y = a.y ^ b.y, z = a.z ^ b.z, w = a.w ^ b.w;
  return __popc(xor3(x, y, z)) + __popc(w) + 2u * __popc(maj3(x, y, z));
}

__device__ __forceinline__ double compat_generic (float4 m, double d) {     // crossover.h:89-92
  const double mu = double(m.x);
  const double s = d < mu ? double(m.y) : double(m.z);
  const double t = d - mu;
  return exp(-(t * t) / (2. * s * s));
}

/// {mu, sigma_in, sigma_out, class} -> {mu, -log2(e)/(2 sigma_in^2), -log2(e)/(2 sigma_out^2), class}: the
/// gaussian of crossover.h:89-92 becomes one multiply chain and one EX2 (float-vector genomes only,
/// whose distances carry float rounding already; tolerance 1e-5, BASELINE.json north_star)
__device__ __forceinline__ float4 fast_meta (float4 m) {
  return make_float4(m.x, -LOG2E / (2.f * m.y * m.y), -LOG2E / (2.f * m.z * m.z), m.w);
}
__device__ __forceinline__ float compat_fast (float4 fm, float d) {
  const float t = d - fm.x;
  const float e = t * t * (d < fm.x ? fm.y : fm.z);
  float c;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(e));
  return c;
}

/// sqrt.approx (1 ulp class): the distances of float-vector genomes carry float rounding already
__device__ __forceinline__ float sqrt_fast (float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ uint4 lds128 (uint32_t addr, int off) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr + off) : "memory");
  return v;
}

/// One 16-byte step of the float tile: 4 x 4 (birth, representative) pairs, two packed adds and two packed FMAs each
template <int OFF>
__device__ __forceinline__ void f32_tile_step (const uint32_t (&aB)[BT], const uint32_t (&aR)[RT], float2 (&acc)[BT][RT]) {
  uint4 bq[BT], rq[RT];
#pragma unroll
  for (int i = 0; i < BT; i++) bq[i] = lds128(aB[i], OFF);
#pragma unroll
  for (int j = 0; j < RT; j++) rq[j] = lds128(aR[j], OFF);
#pragma unroll
  for (int i = 0; i < BT; i++)
#pragma unroll
    for (int j = 0; j < RT; j++) {
      const float2 t0 = __fadd2_rn(make_float2(__uint_as_float(bq[i].x), __uint_as_float(bq[i].y)),
                                   make_float2(__uint_as_float(rq[j].x), __uint_as_float(rq[j].y)));
      const float2 t1 = __fadd2_rn(make_float2(__uint_as_float(bq[i].z), __uint_as_float(bq[i].w)),
                                   make_float2(__uint_as_float(rq[j].z), __uint_as_float(rq[j].w)));
      acc[i][j] = __ffma2_rn(t1, t1, __ffma2_rn(t0, t0, acc[i][j]));
    }
}

struct Winners {
  float bestScore;
  uint32_t best, first, count;
};

__device__ __forceinline__ void fold (Winners &w, float score, uint32_t species) {
  if (w.best == 0xFFFFFFFFu || w.bestScore < score || (w.bestScore == score && species < w.best)) {
    w.best = species;  w.bestScore = score;
  }
  if (score > 0.f) { w.count++; w.first = min(w.first, species); }
}

/// Running state of one (birth, species) score
template <int KIND>
struct Chain {
  double avg = 0.;        // bit-strings: always holds a float value
  float favg = 0.f;       // float vectors
  uint32_t matable = 0;
};

template <int KIND, bool LUT, bool CONT, int KT>
__global__ void __launch_bounds__(256, 2)
classify_dense_tiled_kernel (PoolView pool, ScoreCfg cfg, const uint32_t *__restrict__ b_payload,
                             const float4 *__restrict__ b_cmeta, uint32_t n_births, uint32_t spc, int lut_in_smem, int dbuf, int bulk,
                             uint32_t *__restrict__ out_best, float *__restrict__ out_best_score,
                             uint32_t *__restrict__ out_first, uint32_t *__restrict__ out_count, ShardSink sink) {
  extern __shared__ __align__(16) unsigned char smem[];
  const uint32_t K = KT ? uint32_t(KT) : pool.K, wp4 = pool.wp >> 2, stride4 = wp4 | 1u;    // KT: enveloppe size known at compile time
  // [births][dist][chunk buffer 0][chunk buffer 1 (dbuf)][table (lut_in_smem)]; a chunk buffer = rows + metadata + fills
  uint64_t *bars = reinterpret_cast<uint64_t*>(smem);                     // two mbarriers (bulk-copy staging), one per chunk buffer
  uint4 *births_s = reinterpret_cast<uint4*>(smem + 16);                  // [TB][stride4]
  float4 *bmeta_s = reinterpret_cast<float4*>(births_s + size_t(TB) * stride4);   // [TB]
  uint32_t *dist_s = reinterpret_cast<uint32_t*>(bmeta_s + TB);           // [TR][DPAD]: u32 (bit-strings) / f32 bits
  unsigned char *chunk0 = reinterpret_cast<unsigned char*>(dist_s + size_t(TR) * DPAD);
  const size_t chunkBytes = size_t(TR) * stride4 * 16 + size_t(TR) * 16 + size_t(TR) * 4;
  const double *lut = cfg.lut;
  double *lut_s = reinterpret_cast<double*>(chunk0 + (dbuf ? 2 : 1) * chunkBytes);   // 16-aligned: every block is a multiple of 16 bytes
  if (LUT && lut_in_smem) {
    const uint32_t nd = cfg.lut_classes * cfg.lut_stride;
    for (uint32_t i = threadIdx.x; i < nd; i += 256) lut_s[i] = __ldg(cfg.lut + i);
    lut = lut_s;
  }

  // thread -> tile corner: a warp covers 8 representative rows x 4 birth rows, so that a 128-bit operand load of a
  // warp touches 8 (4) distinct rows = one shared-memory wavefront each, the rest being broadcast
  const int tid = threadIdx.x;
  const int tx = ((tid >> 5) & 1) * 8 + (tid & 7), ty = (tid >> 6) * 4 + ((tid >> 3) & 3);
  const uint32_t b0 = blockIdx.x * TB;

  // ---- stage the birth tile (rows past the end are zero).  Long genomes: one TMA bulk copy per row (rows are
  //      contiguous in HBM, padded in shared memory), completing on the first chunk's mbarrier; short ones
  //      (< 128 B per row) keep 16-byte cp.async ----
  const uint32_t rowBytes = pool.wp * 4;
  const uint32_t birthRows = min(uint32_t(TB), n_births - b0);
  if (bulk) {
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
    __syncthreads();
    if (tid < int(birthRows))
      bulk_g2s(&births_s[tid * stride4], reinterpret_cast<const uint4*>(b_payload) + size_t(b0 + tid) * wp4, rowBytes, &bars[0]);
    for (uint32_t i = tid; i < (TB - birthRows) * wp4; i += 256)
      births_s[(birthRows + i / wp4) * stride4 + i % wp4] = make_uint4(0, 0, 0, 0);
  } else {
    for (uint32_t i = tid; i < TB * wp4; i += 256) {
      const uint32_t row = i / wp4, v = i % wp4;
      if (b0 + row < n_births)
        cp_async16(&births_s[row * stride4 + v], reinterpret_cast<const uint4*>(b_payload) + size_t(b0 + row) * wp4 + v);
      else births_s[row * stride4 + v] = make_uint4(0, 0, 0, 0);
    }
  }
  if (tid < TB) {
    const float4 m = b0 + tid < n_births ? __ldg(b_cmeta + b0 + tid) : make_float4(0.f, 1.f, 1.f, 0.f);
    bmeta_s[tid] = KIND == 1 ? fast_meta(m) : m;
  }

  // epilogue role: birth eb, species eg, eg + 4, ... of every chunk
  const int eb = tid & 63, eg = tid >> 6;
  Winners win { -CUDART_INF_F, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u };

  // ---- the pool streams through shared memory in chunks of whole species; with two chunk buffers
  //      (short genomes: a chunk is consumed in about the time its load takes) chunk c+1 is in flight
  //      while chunk c is computed ----
  auto stage = [&] (uint32_t s0, uint32_t buf) {
    unsigned char *cb = chunk0 + buf * chunkBytes;
    uint4 *reps = reinterpret_cast<uint4*>(cb);
    float4 *rmeta = reinterpret_cast<float4*>(reps + size_t(TR) * stride4);
    uint32_t *fill = reinterpret_cast<uint32_t*>(rmeta + TR);
    const uint32_t ns = min(spc, pool.n_slots - s0), rows = ns * K;
    const uint4 *src = reinterpret_cast<const uint4*>(pool.payload) + size_t(s0) * K * wp4;   // `rows` rows, contiguous
    if (bulk) {
      // thread 0 announces the bytes of the chunk (plus the birth tile on the very first one), one thread per row issues its copy
      if (tid == 0) mbar_expect_tx(&bars[buf], rows * (rowBytes + 16u) + (s0 == 0 ? birthRows * rowBytes : 0u));
      if (tid < int(rows)) bulk_g2s(&reps[tid * stride4], src + size_t(tid) * wp4, rowBytes, &bars[buf]);
      if (tid == 32) bulk_g2s(rmeta, pool.cmeta + size_t(s0) * K, rows * 16u, &bars[buf]);
      for (uint32_t i = tid; i < (TR - rows) * wp4; i += 256) reps[(rows + i / wp4) * stride4 + i % wp4] = make_uint4(0, 0, 0, 0);
    } else {
      for (uint32_t i = tid; i < TR * wp4; i += 256) {
        const uint32_t row = i / wp4, v = i % wp4;
        if (row < rows) cp_async16(&reps[row * stride4 + v], src + size_t(row) * wp4 + v);
        else reps[row * stride4 + v] = make_uint4(0, 0, 0, 0);
      }
      if (tid < int(rows)) cp_async16(&rmeta[tid], pool.cmeta + size_t(s0) * K + tid);
    }
    if (tid < int(ns)) cp_async4(&fill[tid], pool.fill + s0 + tid);
    cp_async_commit();
  };
  stage(0, 0);
  for (uint32_t s0 = 0, c = 0; s0 < pool.n_slots; s0 += spc, c++) {
    const uint32_t ns = min(spc, pool.n_slots - s0), rows = ns * K;
    const uint32_t buf = dbuf ? (c & 1u) : 0u;
    if (dbuf && s0 + spc < pool.n_slots) { stage(s0 + spc, buf ^ 1u); cp_async_wait_but_one(); }
    else cp_async_wait_all();
    if (bulk) mbar_wait(&bars[buf], (dbuf ? (c >> 1) : c) & 1u);      // the buffer's mbarrier flips phase once per use
    __syncthreads();
    uint4 *reps_s = reinterpret_cast<uint4*>(chunk0 + buf * chunkBytes);
    float4 *rmeta_s = reinterpret_cast<float4*>(reps_s + size_t(TR) * stride4);
    uint32_t *fill_s = reinterpret_cast<uint32_t*>(rmeta_s + TR);
    // metadata in the form the epilogue wants (published by the barrier after the tile)
    if (KIND == 1 && tid < int(rows)) rmeta_s[tid] = fast_meta(rmeta_s[tid]);
    if (tid < int(ns)) fill_s[tid] = min(fill_s[tid], K);
    if (KIND == 1 && c == 0) {
      // float vectors: the births are kept NEGATED, so that a difference is one packed add (FADD2) in the tile
      for (uint32_t i = tid; i < TB * wp4; i += 256) {
        uint4 &q = births_s[(i / wp4) * stride4 + i % wp4];
        q.x ^= 0x80000000u;  q.y ^= 0x80000000u;  q.z ^= 0x80000000u;  q.w ^= 0x80000000u;
      }
      __syncthreads();
    }
    // bit-strings with the table in shared memory: one compatibility class for the whole chunk is the common
    // case and lets the epilogue skip the per-representative class lookups (block-wide vote rides on a barrier)
    uint32_t chunkCls = 0xFFFFFFFFu;
    if (LUT && KT > 0) {
      const uint32_t c0 = __float_as_uint(rmeta_s[0].w);
      const bool same = tid >= int(rows) || __float_as_uint(rmeta_s[tid].w) == c0;
      if (__syncthreads_and(same) && lut_in_smem) chunkCls = c0;
    }
    // float vectors: same idea on the gaussian's parameters -- when the whole chunk shares them and the birth
    // has them too, min(c, c) = c and one EX2 per representative is enough
    bool chunkUni = false;
    float4 chunkMeta = make_float4(0.f, 0.f, 0.f, 0.f);
    if (KIND == 1) {
      __syncthreads();                                   // rmeta_s was just rewritten in place
      chunkMeta = rmeta_s[0];
      const float4 m = rmeta_s[tid < int(rows) ? tid : 0];
      chunkUni = __syncthreads_and(m.x == chunkMeta.x && m.y == chunkMeta.y && m.z == chunkMeta.z) != 0;
    }

    // ---- 4 x 4 register tile of distances ----
    if (KIND == 0) {
      // carry-save accumulation: per (birth, representative) pair a running word of weight-1 bits (`ones`) absorbs
      // the four XOR words of every 128-bit step through two full-adder steps; only the two carry words (weight 2)
      // are population-counted.  Two POPC + eight LOP3 per step instead of four POPC + four LOP3: the 16-lane POPC
      // pipe and the 64-lane ALU pipe end up evenly loaded.  dist = 2 * acc + popc(ones).
      uint32_t acc[BT][RT], ones[BT][RT];
      const uint32_t one = min(pool.n_slots, 1u), uno = min(pool.K, 1u);     // both == 1; two values, or the adds get merged
#pragma unroll
      for (int i = 0; i < BT; i++)
#pragma unroll
        for (int j = 0; j < RT; j++) { acc[i][j] = 0; ones[i][j] = 0; }
      for (uint32_t v = 0; v < wp4; v++) {
        uint4 bq[BT], rq[RT];
#pragma unroll
        for (int i = 0; i < BT; i++) bq[i] = births_s[(ty + 16 * i) * stride4 + v];
#pragma unroll
        for (int j = 0; j < RT; j++) rq[j] = reps_s[(tx + 16 * j) * stride4 + v];
#pragma unroll
        for (int i = 0; i < BT; i++)
#pragma unroll
          for (int j = 0; j < RT; j++) {
            const uint32_t x = bq[i].x ^ rq[j].x, y = bq[i].y ^ rq[j].y, z = bq[i].z ^ rq[j].z, w = bq[i].w ^ rq[j].w;
            const uint32_t s1 = xor3(x, y, z), c1 = maj3(x, y, z);
            const uint32_t o = ones[i][j], c2 = maj3(s1, w, o);
            ones[i][j] = xor3(s1, w, o);
            // the two counts are added on the (idle) FMA pipe, not as a ninth ALU operation: IMAD with a multiplier
            // ptxas cannot see through
            acc[i][j] = imad(__popc(c1), one, imad(__popc(c2), uno, acc[i][j]));
          }
      }
#pragma unroll
      for (int i = 0; i < BT; i++)
#pragma unroll
        for (int j = 0; j < RT; j++) dist_s[(tx + 16 * j) * DPAD + ty + 16 * i] = 2u * acc[i][j] + __popc(ones[i][j]);
    } else {
      // packed FP32 (FADD2 / FFMA2, sm_100): two coordinates per instruction -- this loop is bound by issue
      // slots, not by the FMA datapath; births are stored negated, so t = rep - birth is one add
      float2 acc[BT][RT];
#pragma unroll
      for (int i = 0; i < BT; i++)
#pragma unroll
        for (int j = 0; j < RT; j++) acc[i][j] = make_float2(0.f, 0.f);
      // operand rows by shared-window address, four 16-byte steps per iteration with immediate offsets: the FMA
      // pipe is the one this loop is bound by, so the address arithmetic (IMADs otherwise) is kept off it
      uint32_t aB[BT], aR[RT];
#pragma unroll
      for (int i = 0; i < BT; i++) aB[i] = uint32_t(__cvta_generic_to_shared(&births_s[(ty + 16 * i) * stride4]));
#pragma unroll
      for (int j = 0; j < RT; j++) aR[j] = uint32_t(__cvta_generic_to_shared(&reps_s[(tx + 16 * j) * stride4]));
      uint32_t v = 0;
      for (; v + 4 <= wp4; v += 4) {
        f32_tile_step<0>(aB, aR, acc);   f32_tile_step<16>(aB, aR, acc);
        f32_tile_step<32>(aB, aR, acc);  f32_tile_step<48>(aB, aR, acc);
#pragma unroll
        for (int i = 0; i < BT; i++) aB[i] += 64;
#pragma unroll
        for (int j = 0; j < RT; j++) aR[j] += 64;
      }
      for (; v < wp4; v++) {
        f32_tile_step<0>(aB, aR, acc);
#pragma unroll
        for (int i = 0; i < BT; i++) aB[i] += 16;
#pragma unroll
        for (int j = 0; j < RT; j++) aR[j] += 16;
      }
#pragma unroll
      for (int i = 0; i < BT; i++)
#pragma unroll
        for (int j = 0; j < RT; j++)
          dist_s[(tx + 16 * j) * DPAD + ty + 16 * i] = __float_as_uint(sqrt_fast(acc[i][j].x + acc[i][j].y));
    }
    __syncthreads();

    // ---- scores: one thread per (birth, species), representatives in order; two species per
    //      thread are walked in lockstep (the warp's 32 births share both species: no divergence) ----
    const float4 bm = bmeta_s[eb];
    const uint32_t bcls = __float_as_uint(bm.w);
    const double *lutB = LUT ? lut + size_t(bcls) * cfg.lut_stride : nullptr;
    const bool sameMeta = KIND == 1 && chunkUni && bm.x == chunkMeta.x && bm.y == chunkMeta.y && bm.z == chunkMeta.z;
    auto step = [&] (Chain<KIND> &ch, uint32_t sl, uint32_t r) {
      const uint32_t dw = dist_s[(sl * K + r) * DPAD + eb];
      if (KIND == 1) {
        const float df = __uint_as_float(dw);
        const float c = sameMeta ? compat_fast(bm, df) : fminf(compat_fast(bm, df), compat_fast(rmeta_s[sl * K + r], df));
        if (CONT) ch.favg += c; else ch.matable += (double(c) >= cfg.compat_thr);
        return;
      }
      double c;
      if (LUT) {
        const uint32_t rcls = __float_as_uint(rmeta_s[sl * K + r].w);
        const double ca = lutB[dw];
        c = ca;
        if (rcls != bcls) { const double cb = lut[size_t(rcls) * cfg.lut_stride + dw]; c = cb < ca ? cb : ca; }
      } else {
        const double d = double(dw), ca = compat_generic(bm, d), cb = compat_generic(rmeta_s[sl * K + r], d);
        c = cb < ca ? cb : ca;
      }
      if (CONT) ch.avg = round_to_float(__dadd_rn(ch.avg, c)); else ch.matable += (c >= cfg.compat_thr);
    };
    auto finish = [&] (const Chain<KIND> &ch, uint32_t k) -> float {
      if (!CONT) return float(__dsub_rn(double(ch.matable), __dmul_rn(cfg.sim_thr, double(k))));
      const float avg = KIND == 1 ? ch.favg : float(ch.avg);
      return float(__dsub_rn(double(__fdiv_rn(avg, float(k))), cfg.avg_thr));
    };
    for (uint32_t sa = eg; sa < ns; sa += 8) {
      const uint32_t sb = sa + 4;
      const uint32_t ka = fill_s[sa], kb = sb < ns ? fill_s[sb] : 0u;
      if (KIND == 0 && LUT && CONT && KT > 0) {
        // the common case, unrolled: both enveloppes full, one compatibility class (min(c, c) = c), table in
        // shared memory -- per representative one LDS of the distance, one LDS.64 of the table, one DADD and
        // the rounding; the two chains interleave.  All conditions are warp-uniform but the birth's class.
        if (chunkCls == bcls && ka == uint32_t(KT) && kb == uint32_t(KT)) {
          const double *tab = lut_s + size_t(bcls) * cfg.lut_stride;
          const uint32_t *da = dist_s + size_t(sa) * KT * DPAD + eb, *db = dist_s + size_t(sb) * KT * DPAD + eb;
          double A = 0., B = 0.;
#pragma unroll
          for (int r = 0; r < (KT ? KT : 1); r++) {
            A = round_to_float(__dadd_rn(A, tab[da[r * DPAD]]));
            B = round_to_float(__dadd_rn(B, tab[db[r * DPAD]]));
          }
          fold(win, float(__dsub_rn(double(__fdiv_rn(float(A), float(KT))), cfg.avg_thr)), s0 + sa);
          fold(win, float(__dsub_rn(double(__fdiv_rn(float(B), float(KT))), cfg.avg_thr)), s0 + sb);
          continue;
        }
      }
      if (KIND == 1 && CONT && KT > 0) {
        // float vectors, the common case, unrolled: both enveloppes full and one set of gaussian parameters for
        // the chunk and the birth -- per representative one LDS, five FP32 operations and one EX2
        if (sameMeta && ka == uint32_t(KT) && kb == uint32_t(KT)) {
          const uint32_t *da = dist_s + size_t(sa) * KT * DPAD + eb, *db = dist_s + size_t(sb) * KT * DPAD + eb;
          float A = 0.f, B = 0.f;
#pragma unroll
          for (int r = 0; r < (KT ? KT : 1); r++) {
            A += compat_fast(bm, __uint_as_float(da[r * DPAD]));
            B += compat_fast(bm, __uint_as_float(db[r * DPAD]));
          }
          fold(win, float(__dsub_rn(double(__fdiv_rn(A, float(KT))), cfg.avg_thr)), s0 + sa);
          fold(win, float(__dsub_rn(double(__fdiv_rn(B, float(KT))), cfg.avg_thr)), s0 + sb);
          continue;
        }
      }
      Chain<KIND> ca, cb;
      const uint32_t kmin = min(ka, kb);
      uint32_t r = 0;
      for (; r < kmin; r++) { step(ca, sa, r); step(cb, sb, r); }
      for (; r < ka; r++) step(ca, sa, r);
      for (; r < kb; r++) step(cb, sb, r);
      if (ka) fold(win, finish(ca, ka), s0 + sa);
      if (kb) fold(win, finish(cb, kb), s0 + sb);
    }
    __syncthreads();                                  // chunk fully consumed: its buffer and the distance tile are free
    if (!dbuf && s0 + spc < pool.n_slots) stage(s0 + spc, 0);
  }

  // ---- combine the 4 partial winners of every birth ----
  __syncthreads();
  Winners *w_s = reinterpret_cast<Winners*>(dist_s);       // [4][TB]
  w_s[eg * TB + eb] = win;
  __syncthreads();
  if (tid < TB && b0 + tid < n_births) {
    Winners w = w_s[tid];
    for (int g = 1; g < 4; g++) {
      const Winners o = w_s[g * TB + tid];
      if (o.best != 0xFFFFFFFFu &&
          (w.best == 0xFFFFFFFFu || w.bestScore < o.bestScore || (w.bestScore == o.bestScore && o.best < w.best))) {
        w.best = o.best;  w.bestScore = o.bestScore;
      }
      w.first = min(w.first, o.first);
      w.count += o.count;
    }
    if (sink.world) {
      // sharded sweep: the record goes to every rank's gather buffer (peer stores over NVLink)
      const uint32_t off = sink.species_offset;
      const uint4 rec = make_uint4(w.best == 0xFFFFFFFFu ? w.best : w.best + off, __float_as_uint(w.bestScore),
                                   w.first == 0xFFFFFFFFu ? w.first : w.first + off, w.count);
      #pragma unroll
      for (uint32_t p = 0; p < MAX_PEERS; p++)       // static indices: the pointers stay in the parameter bank
        if (p < sink.world) sink.peer[p][size_t(sink.rank) * n_births + b0 + tid] = rec;
    } else {
      out_best[b0 + tid] = w.best;  out_best_score[b0 + tid] = w.bestScore;
      out_first[b0 + tid] = w.first;  out_count[b0 + tid] = w.count;
    }
  }
}

size_t dense_chunk_bytes (uint32_t wp) {
  const size_t stride4 = (wp >> 2) | 1u;
  return size_t(TR) * stride4 * 16 + size_t(TR) * 16 + size_t(TR) * 4;
}
size_t dense_smem (uint32_t wp) {
  const size_t stride4 = (wp >> 2) | 1u;
  return 16 + size_t(TB) * stride4 * 16 + size_t(TB) * 16 + size_t(TR) * DPAD * 4 + dense_chunk_bytes(wp);
}

} // namespace

bool dense_tiled_supports (uint32_t wp, uint32_t K) {
  return K <= uint32_t(TR) && dense_smem(wp) <= (size_t(227) << 10);
}

void launch_classify_dense_tiled (const PoolView &pool, const ScoreCfg &cfg, const uint32_t *b_payload,
                                  const float4 *b_cmeta, uint32_t n_births, uint32_t *out_best,
                                  float *out_best_score, uint32_t *out_first, uint32_t *out_count,
                                  cudaStream_t stream, const ShardSink *sink) {
  if (!n_births) return;
  const ShardSink snk = sink ? *sink : ShardSink{};
  const uint32_t spc = uint32_t(TR) / pool.K;
  size_t smem = dense_smem(pool.wp);
  // a second chunk buffer when at least two CTAs per SM still fit with it
  const int dbuf = smem + dense_chunk_bytes(pool.wp) <= (size_t(110) << 10);
  if (dbuf) smem += dense_chunk_bytes(pool.wp);
  // the compatibility table rides in shared memory when two CTAs per SM still fit with it
  const size_t lutBytes = size_t(cfg.lut_classes) * cfg.lut_stride * sizeof(double);
  const int lutInSmem = cfg.use_lut && lutBytes > 0 && lutBytes <= (size_t(32) << 10) && smem + lutBytes <= (size_t(110) << 10);
  if (lutInSmem) smem += lutBytes;
  const int bulk = pool.wp * 4 >= 128;        // TMA bulk copies for rows of at least 128 bytes
  const uint32_t grid = (n_births + TB - 1) / TB;
  auto go = [&] (auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    kernel<<<grid, 256, smem, stream>>>(pool, cfg, b_payload, b_cmeta, n_births, spc, lutInSmem, dbuf, bulk, out_best, out_best_score,
                                        out_first, out_count, snk);
  };
  const bool cont = cfg.full_continuous != 0;
  auto pick = [&] (auto kt) {
    constexpr int KT = decltype(kt)::value;
    if (cfg.kind == 0) {
      if (cfg.use_lut) { if (cont) go(classify_dense_tiled_kernel<0, true, true, KT>); else go(classify_dense_tiled_kernel<0, true, false, KT>); }
      else { if (cont) go(classify_dense_tiled_kernel<0, false, true, KT>); else go(classify_dense_tiled_kernel<0, false, false, KT>); }
    } else { if (cont) go(classify_dense_tiled_kernel<1, false, true, KT>); else go(classify_dense_tiled_kernel<1, false, false, KT>); }
  };
  // the reference's default rsetSize (ptreeconfig.cpp:10) and BASELINE.json's K get unrolled epilogues
  if (pool.K == 10) pick(std::integral_constant<int, 10>{});
  else if (pool.K == 5) pick(std::integral_constant<int, 5>{});
  else pick(std::integral_constant<int, 0>{});
}

} // namespace apt
