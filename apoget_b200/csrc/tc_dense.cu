// classify_dense_tc_kernel -- the dense sweep (every birth x every species) for long bit-string genomes
// on the 5th-generation tensor cores: Hamming distance as an exact integer GEMM.
//
// With every genome bit b expanded to the signed byte 2b - 1, the dot product of two genomes of L bits is
// (#agreeing - #differing positions) = L - 2 * hamming, i.e. hamming = (L - a.b) / 2 EXACTLY (int8 x int8
// products accumulated in int32, `tcgen05.mma.kind::i8`).  The reference's arithmetic per (birth, species)
// -- compatibility of the distance, sequential float sum, threshold (phylogenetictree.hpp:456-478 / :430-452,
// crossover.h:89-92) -- is unchanged and runs in the epilogue on the accumulators, so the winners are
// bit-identical to the popcount sweep of dense.cu and to the oracle.
//
// Schedule (one persistent CTA per SM, warp specialised):
//   * operands are pre-expanded in HBM into "tile images": 128 rows x 128 bytes of K, already in the
//     shared-memory order the UMMA descriptors describe (K-major, no swizzle: core matrices of 8 rows x 16 B,
//     128 B apart along K, 1 KB apart along M/N).  A pipeline stage is then ONE TMA bulk copy per operand
//     (16 KB of births, 32 KB of representatives), no tensor map and no swizzle to keep consistent;
//   * warp 0: producer (cp.async.bulk -> mbarrier), warp 1: one thread issues tcgen05.mma M=128 N=256 K=32,
//     accumulators in TMEM (2 x 256 columns, double buffered), warps 2-9: epilogue -- a thread owns one birth
//     (TMEM lane), reads the distances of whole species with tcgen05.ld, walks the k representatives IN ORDER
//     and folds the score into the birth's running winners while the next tile multiplies;
//   * a CTA keeps one tile of 128 births and streams all representative tiles past it (the expanded pool is
//     re-read from L2), then takes the next birth tile.
// Representative rows are laid out per tile as whole species: tile t holds species [t*spc, (t+1)*spc),
// spc = 256 / K, species s at rows (s - t*spc)*K ..+K, unused rows zero.
#include "kernels.cuh"
#include "round24.h"

#include <math_constants.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace apt {

namespace {

constexpr uint32_t TM = 128;            // births per tile (UMMA M)
constexpr uint32_t TN = 256;            // representative rows per tile (UMMA N)
constexpr uint32_t KBYTES = 128;        // bytes of K (= genome bits) per pipeline stage
constexpr uint32_t A_STAGE = TM * KBYTES, B_STAGE = TN * KBYTES, STAGE = A_STAGE + B_STAGE;
constexpr uint32_t UMMA_K = 32;         // K per tcgen05.mma.kind::i8
constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr int EPI_WARPS = 8, TC_THREADS = (2 + EPI_WARPS) * 32;

__device__ __forceinline__ uint32_t smem_u32 (const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init (uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx (uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive (uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait (uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s (uint32_t dst, const void *gmem, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               :: "r"(dst), "l"(gmem), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_after (void) { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before (void) { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit (uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(bar) : "memory");
}
/// D[tmem] (+)= A[smem] . B[smem]^T, int8 x int8 -> int32, M = 128, N = 256, K = 32
__device__ __forceinline__ void tc_mma_i8 (uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
               :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// ---- CTA pair (cta_group::2): the two SMs of a TPC multiply one 256 x 256 tile, each holding half of either operand ----
__device__ __forceinline__ uint32_t cluster_rank (void) { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all (void) {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
/// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote (uint32_t bar, uint32_t rank) {
  asm volatile("{\n.reg .b32 ra;\nmapa.shared::cluster.u32 ra, %0, %1;\n"
               "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n}\n" :: "r"(bar), "r"(rank) : "memory");
}
/// completion of the pair's MMAs arrives on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void tc_commit_pair (uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
               :: "r"(bar), "h"(uint16_t(3)) : "memory");
}
/// D[tmem of both CTAs] (+)= A . B^T, M = 256 (128 rows per CTA), N = 256 (128 rows of B in either CTA's shared memory), K = 32
__device__ __forceinline__ void tc_mma_i8_pair (uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n"
               :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
constexpr uint32_t TC_IDESC_PAIR = (2u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);

/// K-major operand without swizzle: core matrices (8 rows x 16 B, rows 16 B apart) 128 B apart along K
/// (leading byte offset) and 1 KB apart along M/N (stride byte offset); descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t tc_smem_desc (uint32_t saddr) {
  return uint64_t((saddr & 0x3FFFFu) >> 4) | (uint64_t(128u >> 4) << 16) | (uint64_t(1024u >> 4) << 32) | (uint64_t(1) << 46);
}
/// instruction descriptor: D = S32 (bits 4-5 = 2), A = B = signed 8 bit (bits 7-9, 10-12 = 1), both K-major, N >> 3 at 17, M >> 4 at 24
constexpr uint32_t TC_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((TN >> 3) << 17) | ((TM >> 4) << 24);

// ---- TMEM -> registers: the 32 lanes of the warp's quarter x N consecutive columns (N in {1, 2, 4}, start aligned to N) ----
__device__ __forceinline__ void tmem_ld1 (uint32_t taddr, uint32_t *v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n" : "=r"(v[0]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld2 (uint32_t taddr, uint32_t *v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n" : "=r"(v[0]), "=r"(v[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4 (uint32_t taddr, uint32_t *v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
/// N columns from a start column known to be a multiple of ALIGN (compile-time chunking: widest naturally aligned loads)
template <int N, int ALIGN, int OFF = 0>
__device__ __forceinline__ void tmem_ld_cols (uint32_t taddr, uint32_t *v) {
  if constexpr (OFF < N) {
    constexpr int rem = N - OFF;
    constexpr int al = (OFF % 4 == 0 && ALIGN % 4 == 0) ? 4 : ((OFF % 2 == 0 && ALIGN % 2 == 0) ? 2 : 1);
    constexpr int c = (al >= 4 && rem >= 4) ? 4 : ((al >= 2 && rem >= 2) ? 2 : 1);
    if constexpr (c == 4) tmem_ld4(taddr + OFF, v + OFF);
    else if constexpr (c == 2) tmem_ld2(taddr + OFF, v + OFF);
    else tmem_ld1(taddr + OFF, v + OFF);
    tmem_ld_cols<N, ALIGN, OFF + c>(taddr, v);
  }
}
/// waits for the loads above; the registers pass through the statement so that no use is scheduled ahead of it
template <int N>
__device__ __forceinline__ void tmem_ld_wait (uint32_t *v) {
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < N; i++) asm volatile("" : "+r"(v[i]));
}

struct Winners {
  float bestScore;
  uint32_t best, first, count;
};
__device__ __forceinline__ void fold (Winners &w, float score, uint32_t species) {
  if (w.best == NONE || w.bestScore < score || (w.bestScore == score && species < w.best)) { w.best = species;  w.bestScore = score; }
  if (score > 0.f) { w.count++; w.first = min(w.first, species); }
}
__device__ __forceinline__ void merge (Winners &w, const Winners &o) {
  if (o.best != NONE && (w.best == NONE || w.bestScore < o.bestScore || (w.bestScore == o.bestScore && o.best < w.best))) {
    w.best = o.best;  w.bestScore = o.bestScore;
  }
  w.first = min(w.first, o.first);
  w.count += o.count;
}

} // namespace

// Per-tile metadata of the expanded pool (device, u32 words): header {species in the tile, their common
// compatibility class or NONE, 1 when all their enveloppes are full, first species}, fill[256], class per row[256]
constexpr uint32_t TC_META_WORDS = 4 + 256 + 256;

struct TcArgs {
  const uint8_t *a_img;         ///< births: [m_tiles][nkb][16 KB]
  const uint8_t *b_img;         ///< pool:   [n_tiles][nkb][32 KB]
  const uint32_t *tile_meta;    ///< [n_tiles][TC_META_WORDS]
  const float4 *b_cmeta;        ///< births' {mu, s_in, s_out, class}
  uint32_t n_births, m_tiles, n_tiles, nkb, K, spc, lbits;
  ScoreCfg cfg;
  int lut_in_smem, stages;
  uint32_t *out_best;  float *out_best_score;  uint32_t *out_first, *out_count;
  uint32_t *out_matrix;         ///< debug / distance-matrix mode: [n_births][n_tiles * 256] Hamming distances, or null
  ShardSink sink;
  uint32_t rec_stride, rec_offset;   ///< sharded sweep: records of one rank are rec_stride apart, this launch starts at rec_offset

};

namespace {

/// CG = 1: one CTA per 128-birth tile.  CG = 2: a CTA PAIR (cluster of 2, tcgen05 cta_group::2) per 256 births -- either
/// CTA stages its own 128 births and HALF of the representative tile, the leader issues M = 256 MMAs that read both shared
/// memories: a third less shared-memory and L2 traffic per MAC (32 KB instead of 48 KB per CTA and K block).
template <int KT, bool CONT, int CG>
__global__ void __launch_bounds__(TC_THREADS, 1)
classify_dense_tc_kernel (TcArgs a) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t K = KT ? uint32_t(KT) : a.K;
  const uint32_t S = uint32_t(a.stages);
  constexpr uint32_t B_PART = B_STAGE / CG, STG = A_STAGE + B_PART;      // this CTA's share of a stage
  const uint32_t rank = CG == 2 ? cluster_rank() : 0u;
  const uint32_t firstTile = (blockIdx.x / CG) * CG + rank, tileStep = gridDim.x;
  // [stages x (A 16 KB | B 32 or 16 KB)][table][winner exchange 4 x 32 x 16 B][barriers][tmem base]
  unsigned char *stage0 = tc_smem;
  const uint32_t lutBytes = a.lut_in_smem ? ((a.cfg.lut_classes * a.cfg.lut_stride * 8u + 15u) & ~15u) : 0u;
  double *lut_s = reinterpret_cast<double*>(tc_smem + size_t(S) * STG);
  Winners *xw = reinterpret_cast<Winners*>(tc_smem + size_t(S) * STG + lutBytes);
  uint64_t *bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(xw) + 4 * 32 * sizeof(Winners));
  uint32_t *tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * 8 + 4);
  const uint32_t barFull = smem_u32(bars), barEmpty = smem_u32(bars + 8), barPeer = smem_u32(bars + 16),
                 barTFull = smem_u32(bars + 24), barTEmpty = smem_u32(bars + 26);

  if (tid == 0) {
    for (uint32_t s = 0; s < S; s++) { mbar_init(barFull + 8 * s, 1); mbar_init(barEmpty + 8 * s, 1); mbar_init(barPeer + 8 * s, 1); }
    for (uint32_t b = 0; b < 2; b++) { mbar_init(barTFull + 8 * b, 1); mbar_init(barTEmpty + 8 * b, EPI_WARPS * CG); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (a.lut_in_smem) {
    const uint32_t nd = a.cfg.lut_classes * a.cfg.lut_stride;
    for (uint32_t i = tid; i < nd; i += TC_THREADS) lut_s[i] = __ldg(a.cfg.lut + i);
  }
  if (warp == 1) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();          // the peer's barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t nkb = a.nkb;

  if (warp == 0) {
    // ===== producer: one bulk copy per operand and stage =====
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (uint32_t mt = firstTile; mt < a.m_tiles; mt += tileStep)
        for (uint32_t nt = 0; nt < a.n_tiles; nt++)
          for (uint32_t kb = 0; kb < nkb; kb++) {
            mbar_wait(barEmpty + 8 * s, ph ^ 1u);
            mbar_expect_tx(barFull + 8 * s, STG);
            const uint32_t dst = smem_u32(stage0 + size_t(s) * STG);
            bulk_g2s(dst, a.a_img + (size_t(mt) * nkb + kb) * A_STAGE, A_STAGE, barFull + 8 * s);
            bulk_g2s(dst + A_STAGE, a.b_img + (size_t(nt) * nkb + kb) * B_STAGE + size_t(rank) * B_PART, B_PART, barFull + 8 * s);
            if (++s == S) { s = 0; ph ^= 1u; }
          }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread =====
    if (lane == 0 && rank == 0) {
      uint32_t s = 0, ph = 0, it = 0;
      for (uint32_t mt = firstTile; mt < a.m_tiles; mt += tileStep)
        for (uint32_t nt = 0; nt < a.n_tiles; nt++, it++) {
          const uint32_t buf = it & 1u;
          // the epilogue (of both CTAs) has drained this accumulator
          mbar_wait(barTEmpty + 8 * buf, ((it >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem + buf * TN;
          for (uint32_t kb = 0; kb < nkb; kb++) {
            mbar_wait(barFull + 8 * s, ph);
            if (CG == 2) mbar_wait(barPeer + 8 * s, ph);                      // ... and the peer's half of the stage has landed
            tc_fence_after();
            const uint32_t sa = smem_u32(stage0 + size_t(s) * STG);
            const uint64_t da = tc_smem_desc(sa), db = tc_smem_desc(sa + A_STAGE);
#pragma unroll
            for (uint32_t k = 0; k < KBYTES / UMMA_K; k++) {    // two core matrices (256 B) further along K per step
              if (CG == 2) tc_mma_i8_pair(d_tmem, da + uint64_t(k * 16u), db + uint64_t(k * 16u), TC_IDESC_PAIR, (kb | k) != 0u);
              else tc_mma_i8(d_tmem, da + uint64_t(k * 16u), db + uint64_t(k * 16u), TC_IDESC, (kb | k) != 0u);
            }
            // stage free (in both CTAs) once these MMAs have read it
            if (CG == 2) tc_commit_pair(barEmpty + 8 * s); else tc_commit(barEmpty + 8 * s);
            if (++s == S) { s = 0; ph ^= 1u; }
          }
          if (CG == 2) tc_commit_pair(barTFull + 8 * buf); else tc_commit(barTFull + 8 * buf);      // accumulator complete
        }
    } else if (CG == 2 && lane == 0) {
      // the peer: tells the leader when ITS half of every stage has landed (bulk copies complete on local barriers only)
      uint32_t s = 0, ph = 0;
      for (uint32_t mt = firstTile; mt < a.m_tiles; mt += tileStep)
        for (uint32_t nt = 0; nt < a.n_tiles; nt++)
          for (uint32_t kb = 0; kb < nkb; kb++) {
            mbar_wait(barFull + 8 * s, ph);
            mbar_arrive_remote(barPeer + 8 * s, 0u);
            if (++s == S) { s = 0; ph ^= 1u; }
          }
    }
  } else {
    // ===== epilogue: thread = birth (TMEM lane); the two warps of a lane quarter take alternate species pairs =====
    const int ew = warp - 2, q = warp & 3, h = ew >> 2;
    const uint32_t row = uint32_t(q) * 32u + lane;
    const uint32_t tlane = tmem + ((uint32_t(q) * 32u) << 16);
    const double *lutAny = a.lut_in_smem ? lut_s : a.cfg.lut;
    const int L = int(a.lbits);
    uint32_t it = 0;
    for (uint32_t mt = firstTile; mt < a.m_tiles; mt += tileStep) {
      const uint32_t b = mt * TM + row;
      const bool live = b < a.n_births;
      const float4 bm = live ? __ldg(a.b_cmeta + b) : make_float4(0.f, 1.f, 1.f, 0.f);
      const uint32_t bcls = __float_as_uint(bm.w);
      const double *lutB = lutAny + size_t(bcls) * a.cfg.lut_stride;
      Winners win { -CUDART_INF_F, NONE, NONE, 0u };
      for (uint32_t nt = 0; nt < a.n_tiles; nt++, it++) {
        const uint32_t buf = it & 1u;
        const uint32_t *meta = a.tile_meta + size_t(nt) * TC_META_WORDS;
        const uint32_t ns = __ldg(meta), uni = __ldg(meta + 1), allFull = __ldg(meta + 2), s0 = __ldg(meta + 3);
        mbar_wait(barTFull + 8 * buf, (it >> 1) & 1u);
        tc_fence_after();
        const uint32_t tcol = tlane + buf * TN;
        if (a.out_matrix) {
          // distance-matrix mode: every column of the tile, as Hamming distances
          for (uint32_t c = uint32_t(h) * (TN / 2); c < uint32_t(h + 1) * (TN / 2); c += 4) {
            uint32_t v[4];
            tmem_ld4(tcol + c, v);
            tmem_ld_wait<4>(v);
            if (live)
#pragma unroll
              for (int j = 0; j < 4; j++) a.out_matrix[size_t(b) * (size_t(a.n_tiles) * TN) + size_t(nt) * TN + c + j] = uint32_t((L - int(v[j])) >> 1);
          }
        } else if constexpr (KT > 0) {
          const bool fast = CONT && a.lut_in_smem && allFull && uni == bcls;
          // score of one species from its KT loaded columns (general form: fills, classes, vote counting)
          auto general = [&] (const uint32_t *v, uint32_t s) -> float {
            const uint32_t k = min(__ldg(meta + 4 + s), K);
            double avg = 0.;
            uint32_t matable = 0;
#pragma unroll
            for (int r = 0; r < (KT ? KT : 1); r++) {
              if (uint32_t(r) < k) {
                const uint32_t d = uint32_t((L - int(v[r])) >> 1);
                const uint32_t rcls = __ldg(meta + 4 + 256 + s * K + r);
                double c = lutB[d];
                if (rcls != bcls) { const double cb = lutAny[size_t(rcls) * a.cfg.lut_stride + d]; c = cb < c ? cb : c; }
                if (CONT) avg = round_to_float(__dadd_rn(avg, c)); else matable += (c >= a.cfg.compat_thr);
              }
            }
            if (!CONT) return float(__dsub_rn(double(matable), __dmul_rn(a.cfg.sim_thr, double(k))));
            return float(__dsub_rn(double(__fdiv_rn(float(avg), float(k))), a.cfg.avg_thr));
          };
          for (uint32_t sa = 2u * uint32_t(h); sa < ns; sa += 4) {
            const uint32_t sb = sa + 1;
            uint32_t v[2 * (KT ? KT : 1)];
            if (sb < ns) {
              tmem_ld_cols<2 * KT, (2 * KT) % 4 == 0 ? 4 : 2>(tcol + sa * K, v);
              tmem_ld_wait<2 * KT>(v);
              if (fast) {
                double A = 0., B = 0.;
#pragma unroll
                for (int r = 0; r < (KT ? KT : 1); r++) {
                  A = round_to_float(__dadd_rn(A, lutB[(L - int(v[r])) >> 1]));
                  B = round_to_float(__dadd_rn(B, lutB[(L - int(v[KT + r])) >> 1]));
                }
                fold(win, float(__dsub_rn(double(__fdiv_rn(float(A), float(KT))), a.cfg.avg_thr)), s0 + sa);
                fold(win, float(__dsub_rn(double(__fdiv_rn(float(B), float(KT))), a.cfg.avg_thr)), s0 + sb);
              } else {
                if (__ldg(meta + 4 + sa)) fold(win, general(v, sa), s0 + sa);
                if (__ldg(meta + 4 + sb)) fold(win, general(v + KT, sb), s0 + sb);
              }
            } else {
              tmem_ld_cols<KT, (2 * KT) % 4 == 0 ? 4 : 2>(tcol + sa * K, v);
              tmem_ld_wait<KT>(v);
              if (__ldg(meta + 4 + sa)) fold(win, general(v, sa), s0 + sa);
            }
          }
        } else {
          // any enveloppe size: column by column
          for (uint32_t s = uint32_t(h); s < ns; s += 2) {
            const uint32_t k = min(__ldg(meta + 4 + s), K);
            double avg = 0.;
            uint32_t matable = 0;
            for (uint32_t r = 0; r < K; r++) {
              uint32_t v[1];
              tmem_ld1(tcol + s * K + r, v);
              tmem_ld_wait<1>(v);
              if (r < k) {
                const uint32_t d = uint32_t((L - int(v[0])) >> 1);
                const uint32_t rcls = __ldg(meta + 4 + 256 + s * K + r);
                double c = lutB[d];
                if (rcls != bcls) { const double cb = lutAny[size_t(rcls) * a.cfg.lut_stride + d]; c = cb < c ? cb : c; }
                if (CONT) avg = round_to_float(__dadd_rn(avg, c)); else matable += (c >= a.cfg.compat_thr);
              }
            }
            if (k) {
              const float score = CONT ? float(__dsub_rn(double(__fdiv_rn(float(avg), float(k))), a.cfg.avg_thr))
                                       : float(__dsub_rn(double(matable), __dmul_rn(a.cfg.sim_thr, double(k))));
              fold(win, score, s0 + s);
            }
          }
        }
        // accumulator drained: hand it back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CG == 2) mbar_arrive_remote(barTEmpty + 8 * buf, 0u); else mbar_arrive(barTEmpty + 8 * buf); }
      }
      // ---- the two halves of every birth's species are merged and the winners leave the chip ----
      if (h == 1) xw[q * 32 + lane] = win;
      asm volatile("bar.sync 1, %0;\n" :: "n"(EPI_WARPS * 32) : "memory");
      if (h == 0 && live && !a.out_matrix) {
        merge(win, xw[q * 32 + lane]);
        if (a.sink.world) {
          const uint32_t off = a.sink.species_offset;
          const uint4 rec = make_uint4(win.best == NONE ? win.best : win.best + off, __float_as_uint(win.bestScore),
                                       win.first == NONE ? win.first : win.first + off, win.count);
#pragma unroll
          for (uint32_t p = 0; p < MAX_PEERS; p++)
            if (p < a.sink.world) a.sink.peer[p][size_t(a.sink.rank) * a.rec_stride + a.rec_offset + b] = rec;
        } else {
          a.out_best[b] = win.best;  a.out_best_score[b] = win.bestScore;
          a.out_first[b] = win.first;  a.out_count[b] = win.count;
        }
      }
      asm volatile("bar.sync 1, %0;\n" :: "n"(EPI_WARPS * 32) : "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();          // the leader's MMAs read the peer's shared memory and write its tensor memory
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(512u) : "memory");
  }
}

/// Bits -> signed bytes (bit 1 -> +1, bit 0 -> -1), written straight into tile images.  One thread = 16 bits of one
/// destination row -> one 16-byte store; 64 consecutive threads fill 1 KB of contiguous image (8 rows x 8 chunks).
/// spc == 0: destination row = source row (births, images of `rpt` = 128 rows); otherwise the pool mapping
/// (images of 256 rows: tile t = species [t*spc, (t+1)*spc), row (s - t*spc)*K + r; rows past a fill stay zero).
/// Bits past the source row (K padding up to a multiple of 128) are ZERO bytes: they add nothing to a dot product.
__global__ void tc_expand_kernel (const uint32_t *__restrict__ src, uint32_t n_src_rows, uint32_t wp, uint32_t nkb,
                                  uint32_t rpt, uint32_t K, uint32_t spc, const uint32_t *__restrict__ fill,
                                  uint32_t n_dst_rows, uint8_t *__restrict__ img) {
  const uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint32_t r8 = uint32_t(t & 7u), kc = uint32_t((t >> 3) & 7u);
  const uint64_t rest = t >> 6;
  const uint32_t kb = uint32_t(rest % nkb);
  const uint64_t rg = rest / nkb;
  const uint64_t R = rg * 8 + r8;
  if (R >= n_dst_rows) return;
  int64_t q = -1;
  if (spc == 0) { if (R < n_src_rows) q = int64_t(R); }
  else {
    const uint32_t tile = uint32_t(R / TN), rr = uint32_t(R % TN);
    if (rr < spc * K) {
      const uint32_t s = tile * spc + rr / K, r = rr % K;
      if (uint64_t(s) * K + r < n_src_rows && r < min(fill[s], K)) q = int64_t(s) * K + r;
    }
  }
  const uint32_t word = kb * 4 + (kc >> 1);
  uint4 out = make_uint4(0, 0, 0, 0);
  if (q >= 0 && word < wp) {
    const uint32_t bits = (__ldg(src + size_t(q) * wp + word) >> ((kc & 1u) * 16u)) & 0xFFFFu;
    auto spread = [] (uint32_t x) {       // 4 bits -> 4 bytes, 0x01 for a set bit and 0xFF for a clear one
      const uint32_t y = ((x & 0xFu) * 0x00204081u) & 0x01010101u;
      return y | ((y ^ 0x01010101u) * 0xFFu);
    };
    out = make_uint4(spread(bits), spread(bits >> 4), spread(bits >> 8), spread(bits >> 12));
  }
  const size_t off = (size_t(R / rpt) * nkb + kb) * (size_t(rpt) * KBYTES) + (size_t((R % rpt) >> 3) * 8 + kc) * 128 + size_t(r8) * 16;
  *reinterpret_cast<uint4*>(img + off) = out;
}

/// Per-tile metadata of the expanded pool
__global__ void tc_tile_meta_kernel (const float4 *__restrict__ cmeta, const uint32_t *__restrict__ fill, uint32_t n_species,
                                     uint32_t K, uint32_t spc, uint32_t *__restrict__ meta) {
  const uint32_t nt = blockIdx.x;
  uint32_t *m = meta + size_t(nt) * TC_META_WORDS;
  const uint32_t s0 = nt * spc, ns = min(spc, n_species - s0);
  __shared__ uint32_t sh[3];      // first class seen, mixed flag, not-full flag
  if (threadIdx.x == 0) { sh[0] = NONE; sh[1] = 0; sh[2] = 0; }
  __syncthreads();
  for (uint32_t s = threadIdx.x; s < 256; s += blockDim.x) {
    const uint32_t k = s < ns ? min(fill[s0 + s], K) : 0u;
    m[4 + s] = k;
    if (s < ns && k != K) sh[2] = 1;
  }
  for (uint32_t rr = threadIdx.x; rr < 256; rr += blockDim.x) {
    const uint32_t s = rr / K, r = rr % K;
    uint32_t cls = 0;
    if (s < ns && r < min(fill[s0 + s], K)) {
      cls = __float_as_uint(cmeta[size_t(s0 + s) * K + r].w);
      atomicCAS(&sh[0], NONE, cls);
    }
    m[4 + 256 + rr] = cls;
  }
  __syncthreads();
  for (uint32_t rr = threadIdx.x; rr < 256; rr += blockDim.x) {
    const uint32_t s = rr / K, r = rr % K;
    if (s < ns && r < min(fill[s0 + s], K) && m[4 + 256 + rr] != sh[0]) sh[1] = 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) { m[0] = ns;  m[1] = sh[1] ? NONE : sh[0];  m[2] = sh[2] ? 0u : 1u;  m[3] = s0; }
}

} // namespace

bool dense_tc_supports (const ScoreCfg &cfg, uint32_t wp, uint32_t K) {
  // bit-strings with the exact compatibility table; genomes long enough for the tile to be the work (>= 1024 bits)
  return cfg.kind == 0 && cfg.use_lut && K >= 1 && K <= 32 && wp * 32u >= 1024u && wp * 32u <= 32768u;
}

uint32_t tc_nkb (uint32_t wp) { return (wp * 32u + KBYTES - 1) / KBYTES; }
uint32_t tc_spc (uint32_t K) { return TN / K; }
uint32_t tc_pool_tiles (uint32_t n_species, uint32_t K) { const uint32_t spc = tc_spc(K); return (n_species + spc - 1) / spc; }
size_t tc_pool_image_bytes (uint32_t n_species, uint32_t K, uint32_t wp) { return size_t(tc_pool_tiles(n_species, K)) * tc_nkb(wp) * B_STAGE; }
size_t tc_births_image_bytes (uint32_t n_births, uint32_t wp) { return size_t(((n_births + TM - 1) / TM + 1) & ~1u) * tc_nkb(wp) * A_STAGE; }
size_t tc_pool_meta_bytes (uint32_t n_species, uint32_t K) { return size_t(tc_pool_tiles(n_species, K)) * TC_META_WORDS * 4; }

void launch_tc_expand_pool (const PoolView &pool, uint8_t *img, uint32_t *meta, cudaStream_t st) {
  const uint32_t nkb = tc_nkb(pool.wp), spc = tc_spc(pool.K), nt = tc_pool_tiles(pool.n_slots, pool.K);
  if (!nt) return;
  const uint64_t rows = uint64_t(nt) * TN, threads = rows / 8 * nkb * 64;
  tc_expand_kernel<<<unsigned((threads + 255) / 256), 256, 0, st>>>(pool.payload, pool.n_slots * pool.K, pool.wp, nkb, TN, pool.K, spc,
                                                                    pool.fill, uint32_t(rows), img);
  tc_tile_meta_kernel<<<nt, 256, 0, st>>>(pool.cmeta, pool.fill, pool.n_slots, pool.K, spc, meta);
}

void launch_tc_expand_births (const uint32_t *b_payload, uint32_t n_births, uint32_t wp, uint8_t *img, cudaStream_t st) {
  const uint32_t nkb = tc_nkb(wp);
  const uint64_t rows = uint64_t(((n_births + TM - 1) / TM + 1) & ~1u) * TM, threads = rows / 8 * nkb * 64;     // (an even number of tiles: CTA pairs)
  if (!threads) return;
  tc_expand_kernel<<<unsigned((threads + 255) / 256), 256, 0, st>>>(b_payload, n_births, wp, nkb, TM, 1, 0, nullptr, uint32_t(rows), img);
}

/// The sweep proper over pre-expanded operands.  out_matrix != null: distance-matrix mode (no scoring).
void launch_classify_dense_tc (const PoolView &pool, const ScoreCfg &cfg, const uint8_t *a_img, const uint8_t *b_img,
                               const uint32_t *tile_meta, const float4 *b_cmeta, uint32_t n_births,
                               uint32_t *out_best, float *out_best_score, uint32_t *out_first, uint32_t *out_count,
                               uint32_t *out_matrix, cudaStream_t stream, const ShardSink *sink,
                               uint32_t rec_stride, uint32_t rec_offset) {
  if (!n_births || !pool.n_slots) return;
  TcArgs a;
  a.a_img = a_img;  a.b_img = b_img;  a.tile_meta = tile_meta;  a.b_cmeta = b_cmeta;
  a.n_births = n_births;  a.m_tiles = (n_births + TM - 1) / TM;  a.n_tiles = tc_pool_tiles(pool.n_slots, pool.K);
  a.nkb = tc_nkb(pool.wp);  a.K = pool.K;  a.spc = tc_spc(pool.K);  a.lbits = pool.wp * 32u;
  a.cfg = cfg;
  a.out_best = out_best;  a.out_best_score = out_best_score;  a.out_first = out_first;  a.out_count = out_count;
  a.out_matrix = out_matrix;
  a.sink = sink ? *sink : ShardSink{};
  a.rec_stride = rec_stride ? rec_stride : n_births;  a.rec_offset = rec_offset;
  // One CTA per birth tile by default.  CTA pairs (tcgen05 cta_group::2, APT_TC_CG=2) stage a third less per CTA but measured
  // 6-12 % SLOWER on B200 (131 072 x 100 000 x 4096 bits: 72 vs 63-68 ms): bulk copies complete on barriers of their own
  // CTA only, so the peer's half of every stage is announced to the leader by a relay thread, one more hop per stage.
  const char *cgEnv = getenv("APT_TC_CG");
  const int CGv = (cgEnv && cgEnv[0] == '2' && a.m_tiles >= 2) ? 2 : 1;
  // shared memory: as many stages (48 KB; 32 KB per CTA of a pair) as fit next to the table
  const size_t stg = A_STAGE + B_STAGE / CGv;
  const size_t fixed = 4 * 32 * sizeof(Winners) + (3 * 8 + 4) * 8 + 16;
  const size_t lutBytes = (size_t(cfg.lut_classes) * cfg.lut_stride * 8 + 15) & ~size_t(15);
  const size_t cap = size_t(227) << 10;
  a.lut_in_smem = cfg.use_lut && lutBytes > 0 && 3 * stg + lutBytes + fixed <= cap;
  size_t avail = cap - fixed - (a.lut_in_smem ? lutBytes : 0);
  a.stages = int(std::min<size_t>(8, avail / stg));
  const size_t smem = size_t(a.stages) * stg + (a.lut_in_smem ? lutBytes : 0) + fixed;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // with pairs the births' image holds an even number of tiles (launch_tc_expand_births pads): either CTA has its own
  if (CGv == 2) a.m_tiles = (a.m_tiles + 1u) & ~1u;
  const uint32_t grid = CGv == 2 ? 2u * std::min<uint32_t>(a.m_tiles / 2u, uint32_t(sms) / 2u) : std::min<uint32_t>(a.m_tiles, uint32_t(sms));
  auto go = [&] (auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaLaunchConfig_t lc {};
    lc.gridDim = dim3(grid);  lc.blockDim = dim3(TC_THREADS);  lc.dynamicSmemBytes = smem;  lc.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = unsigned(CGv);  at[0].val.clusterDim.y = 1;  at[0].val.clusterDim.z = 1;
    lc.attrs = at;  lc.numAttrs = 1;
    cudaLaunchKernelEx(&lc, kernel, a);
  };
  const bool cont = cfg.full_continuous != 0;
  auto pick = [&] (auto kt) {
    constexpr int KT = decltype(kt)::value;
    if (CGv == 2) { if (cont) go(classify_dense_tc_kernel<KT, true, 2>); else go(classify_dense_tc_kernel<KT, false, 2>); }
    else { if (cont) go(classify_dense_tc_kernel<KT, true, 1>); else go(classify_dense_tc_kernel<KT, false, 1>); }
  };
  if (pool.K == 10) pick(std::integral_constant<int, 10>{});
  else if (pool.K == 5) pick(std::integral_constant<int, 5>{});
  else pick(std::integral_constant<int, 0>{});
}

} // namespace apt
