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
//     XOR+POPC (bit-strings) or 64 x 3 FP32 ops (float vectors) -- 8 operand bytes per popc32
//     instead of 8 per popc32 *per pair* -- rows are padded to an odd number of uint4 so the
//     16 row addresses of a warp fall in different bank groups;
//   * distances go to a padded shared tile; one thread per (birth, species) then walks the k
//     representatives IN ORDER (the float accumulation is order sensitive), applies the
//     compatibility table and the threshold, and folds the score into per-birth running winners.
#include "kernels.cuh"

#include <math_constants.h>

namespace apt {

namespace {

constexpr int TB = 64;          // births per CTA
constexpr int TR = 64;          // representative rows per chunk
constexpr int BT = 4, RT = 4;   // register tile
constexpr int DPAD = TB + 1;    // distance tile row pitch (floats)

__device__ __forceinline__ void cp_async16 (void *smem, const void *gmem) {
  const uint32_t s = uint32_t(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all (void) { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ double compat_generic (float4 m, double d) {     // crossover.h:89-92
  const double mu = double(m.x);
  const double s = d < mu ? double(m.y) : double(m.z);
  const double t = d - mu;
  return exp(-(t * t) / (2. * s * s));
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

template <int KIND, bool LUT>
__global__ void __launch_bounds__(256, 2)
classify_dense_tiled_kernel (PoolView pool, ScoreCfg cfg, const uint32_t *__restrict__ b_payload,
                             const float4 *__restrict__ b_cmeta, uint32_t n_births, uint32_t spc,
                             uint32_t *__restrict__ out_best, float *__restrict__ out_best_score,
                             uint32_t *__restrict__ out_first, uint32_t *__restrict__ out_count) {
  extern __shared__ __align__(16) unsigned char smem[];
  const uint32_t K = pool.K, wp4 = pool.wp >> 2, stride4 = wp4 | 1u;
  uint4 *births_s = reinterpret_cast<uint4*>(smem);                       // [TB][stride4]
  uint4 *reps_s = births_s + size_t(TB) * stride4;                        // [TR][stride4]
  float4 *rmeta_s = reinterpret_cast<float4*>(reps_s + size_t(TR) * stride4);   // [TR]
  float4 *bmeta_s = rmeta_s + TR;                                         // [TB]
  float *dist_s = reinterpret_cast<float*>(bmeta_s + TB);                 // [TR][DPAD]
  uint32_t *fill_s = reinterpret_cast<uint32_t*>(dist_s + size_t(TR) * DPAD);   // [TR]

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const uint32_t b0 = blockIdx.x * TB;

  // ---- stage the birth tile (rows past the end are zero) ----
  for (uint32_t i = tid; i < TB * wp4; i += 256) {
    const uint32_t row = i / wp4, v = i % wp4;
    if (b0 + row < n_births)
      cp_async16(&births_s[row * stride4 + v], reinterpret_cast<const uint4*>(b_payload) + size_t(b0 + row) * wp4 + v);
    else births_s[row * stride4 + v] = make_uint4(0, 0, 0, 0);
  }
  if (tid < TB) bmeta_s[tid] = b0 + tid < n_births ? __ldg(b_cmeta + b0 + tid) : make_float4(0.f, 1.f, 1.f, 0.f);

  // epilogue role: birth eb, species eg, eg + 4, ... of every chunk
  const int eb = tid & 63, eg = tid >> 6;
  Winners win { -CUDART_INF_F, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u };

  for (uint32_t s0 = 0; s0 < pool.n_slots; s0 += spc) {
    const uint32_t ns = min(spc, pool.n_slots - s0), rows = ns * K;
    __syncthreads();                                  // previous chunk fully consumed
    // ---- stage the chunk: `rows` representative rows are contiguous in the pool ----
    const uint4 *src = reinterpret_cast<const uint4*>(pool.payload) + size_t(s0) * K * wp4;
    for (uint32_t i = tid; i < TR * wp4; i += 256) {
      const uint32_t row = i / wp4, v = i % wp4;
      if (row < rows) cp_async16(&reps_s[row * stride4 + v], src + size_t(row) * wp4 + v);
      else reps_s[row * stride4 + v] = make_uint4(0, 0, 0, 0);
    }
    if (tid < int(rows)) rmeta_s[tid] = __ldg(pool.cmeta + size_t(s0) * K + tid);
    if (tid < int(ns)) fill_s[tid] = min(__ldg(pool.fill + s0 + tid), K);
    cp_async_wait_all();
    __syncthreads();

    // ---- 4 x 4 register tile of distances ----
    if (KIND == 0) {
      uint32_t acc[BT][RT];
#pragma unroll
      for (int i = 0; i < BT; i++)
#pragma unroll
        for (int j = 0; j < RT; j++) acc[i][j] = 0;
      for (uint32_t v = 0; v < wp4; v++) {
        uint4 bq[BT], rq[RT];
#pragma unroll
        for (int i = 0; i < BT; i++) bq[i] = births_s[(ty + 16 * i) * stride4 + v];
#pragma unroll
        for (int j = 0; j < RT; j++) rq[j] = reps_s[(tx + 16 * j) * stride4 + v];
#pragma unroll
        for (int i = 0; i < BT; i++)
#pragma unroll
          for (int j = 0; j < RT; j++)
            acc[i][j] += __popc(bq[i].x ^ rq[j].x) + __popc(bq[i].y ^ rq[j].y)
                       + __popc(bq[i].z ^ rq[j].z) + __popc(bq[i].w ^ rq[j].w);
      }
#pragma unroll
      for (int i = 0; i < BT; i++)
#pragma unroll
        for (int j = 0; j < RT; j++) dist_s[(tx + 16 * j) * DPAD + ty + 16 * i] = float(acc[i][j]);
    } else {
      float acc[BT][RT];
#pragma unroll
      for (int i = 0; i < BT; i++)
#pragma unroll
        for (int j = 0; j < RT; j++) acc[i][j] = 0.f;
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
            const float t0 = __uint_as_float(bq[i].x) - __uint_as_float(rq[j].x);
            const float t1 = __uint_as_float(bq[i].y) - __uint_as_float(rq[j].y);
            const float t2 = __uint_as_float(bq[i].z) - __uint_as_float(rq[j].z);
            const float t3 = __uint_as_float(bq[i].w) - __uint_as_float(rq[j].w);
            acc[i][j] = fmaf(t0, t0, fmaf(t1, t1, fmaf(t2, t2, fmaf(t3, t3, acc[i][j]))));
          }
      }
#pragma unroll
      for (int i = 0; i < BT; i++)
#pragma unroll
        for (int j = 0; j < RT; j++) dist_s[(tx + 16 * j) * DPAD + ty + 16 * i] = sqrtf(acc[i][j]);
    }
    __syncthreads();

    // ---- scores: one thread per (birth, species), representatives in order ----
    const float4 bm = bmeta_s[eb];
    const double *lutB = LUT ? cfg.lut + size_t(__float_as_uint(bm.w)) * cfg.lut_stride : nullptr;
    for (uint32_t sl = eg; sl < ns; sl += 4) {
      const uint32_t k = fill_s[sl];
      if (k == 0) continue;
      float avg = 0.f;
      uint32_t matable = 0;
      for (uint32_t r = 0; r < k; r++) {
        const float df = dist_s[(sl * K + r) * DPAD + eb];
        const float4 rm = rmeta_s[sl * K + r];
        double c;
        if (LUT) {
          const uint32_t di = uint32_t(df);
          const double ca = __ldg(lutB + di), cb = __ldg(cfg.lut + size_t(__float_as_uint(rm.w)) * cfg.lut_stride + di);
          c = cb < ca ? cb : ca;
        } else {
          const double d = double(df), ca = compat_generic(bm, d), cb = compat_generic(rm, d);
          c = cb < ca ? cb : ca;
        }
        avg = float(__dadd_rn(double(avg), c));
        matable += (c >= cfg.compat_thr);
      }
      const float score = cfg.full_continuous
          ? float(__dsub_rn(double(__fdiv_rn(avg, float(k))), cfg.avg_thr))
          : float(__dsub_rn(double(matable), __dmul_rn(cfg.sim_thr, double(k))));
      fold(win, score, s0 + sl);
    }
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
    out_best[b0 + tid] = w.best;  out_best_score[b0 + tid] = w.bestScore;
    out_first[b0 + tid] = w.first;  out_count[b0 + tid] = w.count;
  }
}

size_t dense_smem (uint32_t wp) {
  const size_t stride4 = (wp >> 2) | 1u;
  return size_t(TB + TR) * stride4 * 16 + size_t(TR + TB) * 16 + size_t(TR) * DPAD * 4 + size_t(TR) * 4;
}

} // namespace

bool dense_tiled_supports (uint32_t wp, uint32_t K) {
  return K <= uint32_t(TR) && dense_smem(wp) <= (size_t(227) << 10);
}

void launch_classify_dense_tiled (const PoolView &pool, const ScoreCfg &cfg, const uint32_t *b_payload,
                                  const float4 *b_cmeta, uint32_t n_births, uint32_t *out_best,
                                  float *out_best_score, uint32_t *out_first, uint32_t *out_count,
                                  cudaStream_t stream) {
  if (!n_births) return;
  const uint32_t spc = uint32_t(TR) / pool.K;
  const size_t smem = dense_smem(pool.wp);
  const uint32_t grid = (n_births + TB - 1) / TB;
  auto go = [&] (auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    kernel<<<grid, 256, smem, stream>>>(pool, cfg, b_payload, b_cmeta, n_births, spc, out_best, out_best_score,
                                        out_first, out_count);
  };
  if (cfg.kind == 0) { if (cfg.use_lut) go(classify_dense_tiled_kernel<0, true>); else go(classify_dense_tiled_kernel<0, false>); }
  else go(classify_dense_tiled_kernel<1, false>);
}

} // namespace apt
