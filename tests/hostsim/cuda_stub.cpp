// TEST DOUBLE -- lets the HOST logic of the engine (hierarchy bookkeeping, ordered resolution,
// pool refresh planning) run in the GPU-less build container.
//
// tests/hostsim/build.py links apoget_b200/csrc/engine.cpp + abi.cpp against this file INSTEAD of
// kernels.cu / dense.cu and the CUDA runtime: "device" memory is host memory, streams are
// synchronous, and each launch_* below is a plain-loop stand-in for the kernel of the same name.
// Nothing here is part of libapt_b200.so, and nothing here is timed or shipped; GPU parity of the
// real kernels is established by the -m gpu tests through the real library.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "kernels.cuh"
#include "criteria.h"

extern "C" {
cudaError_t cudaGetDeviceCount (int *c) { *c = 1; return cudaSuccess; }
cudaError_t cudaSetDevice (int) { return cudaSuccess; }
cudaError_t cudaGetDeviceProperties_v2 (cudaDeviceProp *p, int) { memset(p, 0, sizeof *p); p->major = 10; return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags (cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
cudaError_t cudaStreamDestroy (cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags (cudaEvent_t *e, unsigned) { *e = nullptr; return cudaSuccess; }
cudaError_t cudaEventDestroy (cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventCreate (cudaEvent_t *e) { *e = nullptr; return cudaSuccess; }
cudaError_t cudaEventElapsedTime (float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
cudaError_t cudaEventRecord (cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize (cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent (cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaMalloc (void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaMallocAsync (void **p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
cudaError_t cudaFree (void *p) { free(p); return cudaSuccess; }
cudaError_t cudaFreeAsync (void *p, cudaStream_t) { free(p); return cudaSuccess; }
cudaError_t cudaMallocHost (void **p, size_t n) { return cudaMalloc(p, n); }
cudaError_t cudaHostAlloc (void **p, size_t n, unsigned) { return cudaMalloc(p, n); }
cudaError_t cudaFreeHost (void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy (void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync (void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync (void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaStreamSynchronize (cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize (void) { return cudaSuccess; }
cudaError_t cudaGetLastError (void) { return cudaSuccess; }
cudaError_t cudaPointerGetAttributes (cudaPointerAttributes *a, const void*) { memset(a, 0, sizeof *a); a->type = cudaMemoryTypeUnregistered; return cudaSuccess; }
const char* cudaGetErrorString (cudaError_t) { return "stub"; }
}

extern "C" int apt_microbench (int, double out[4]) { out[0] = out[1] = out[2] = out[3] = 0; return 0; }
extern "C" int apt_microbench_cvt (int, double out[1]) { out[0] = 0; return 0; }

namespace apt {

namespace {
double dist (int kind, uint32_t wp, const uint32_t *a, const uint32_t *b) {
  if (kind == 1) {
    // pairwise (tree-like) order on purpose: the real kernel does not accumulate sequentially either
    double s0 = 0, s1 = 0;
    for (uint32_t i = 0; i < wp; i += 2) {
      float fa, fb;
      memcpy(&fa, a + i, 4);  memcpy(&fb, b + i, 4);
      s0 += (double(fa) - double(fb)) * (double(fa) - double(fb));
      if (i + 1 < wp) {
        memcpy(&fa, a + i + 1, 4);  memcpy(&fb, b + i + 1, 4);
        s1 += (double(fa) - double(fb)) * (double(fa) - double(fb));
      }
    }
    return std::sqrt(s0 + s1);
  }
  uint32_t s = 0;
  for (uint32_t i = 0; i < wp; i++) s += __builtin_popcount(a[i] ^ b[i]);
  return s;
}

double compat (const ScoreCfg &cfg, float4 m, double d) {
  if (cfg.use_lut) {
    uint32_t cls;  memcpy(&cls, &m.w, 4);
    return cfg.lut[size_t(cls) * cfg.lut_stride + uint32_t(d)];
  }
  const double mu = m.x, s = d < mu ? m.y : m.z;
  // perturbed by one ulp now and then: the real device exp() is not bit-identical to libm either
  double c = std::exp(-((d - mu) * (d - mu)) / (2. * s * s));
  return (uint64_t(d * 7919) & 3) == 0 ? std::nextafter(c, 2.) : c;
}

float score (const PoolView &pool, const ScoreCfg &cfg, const uint32_t *b, float4 bm, uint32_t slot, float *outD) {
  const uint32_t k = pool.fill[slot] < pool.K ? pool.fill[slot] : pool.K;
  float avg = 0;  uint32_t matable = 0;
  for (uint32_t r = 0; r < k; r++) {
    const size_t row = size_t(slot) * pool.K + r;
    const double d = dist(cfg.kind, pool.wp, b, pool.payload + row * pool.wp);
    const double ca = compat(cfg, bm, d), cb = compat(cfg, pool.cmeta[row], d), c = cb < ca ? cb : ca;
    avg += c;  matable += (c >= cfg.compat_thr);
    if (outD) outD[r] = float(d);
  }
  return cfg.full_continuous ? float(avg / float(k) - cfg.avg_thr) : float(matable - cfg.sim_thr * k);
}
}

void launch_score_pairs (const PoolView &pool, const ScoreCfg &cfg, const uint32_t *bp, const float4 *bm,
                         uint32_t n, const uint32_t *pb, const uint32_t *ps, float *os, float *od, cudaStream_t) {
  for (uint32_t p = 0; p < n; p++)
    os[p] = score(pool, cfg, bp + size_t(pb[p]) * pool.wp, bm[pb[p]], ps[p], od + size_t(p) * pool.K);
}

void launch_pack_births (const uint32_t *src, const float *compat, const uint32_t *cls, uint32_t n, uint32_t words,
                         uint32_t wp, uint32_t *dst, float4 *meta, cudaStream_t) {
  for (uint32_t i = 0; i < n; i++) {
    memset(dst + size_t(i) * wp, 0, wp * 4);
    memcpy(dst + size_t(i) * wp, src + size_t(i) * words, words * 4);
    float fc = 0;  if (cls) memcpy(&fc, cls + i, 4);
    meta[i] = make_float4(compat[3 * i], compat[3 * i + 1], compat[3 * i + 2], fc);
  }
}

void launch_fill_f4 (float4 *dst, uint32_t n, float4 value, cudaStream_t) { for (uint32_t i = 0; i < n; i++) dst[i] = value; }

void launch_pad_rows (const uint32_t *src, uint32_t n, uint32_t words, uint32_t wp, uint32_t *dst, cudaStream_t) {
  for (uint32_t i = 0; i < n; i++) {
    memset(dst + size_t(i) * wp, 0, wp * 4);
    memcpy(dst + size_t(i) * wp, src + size_t(i) * words, words * 4);
  }
}

void launch_gather_rows (const uint32_t *src, const uint32_t *idx, uint32_t n, uint32_t words, uint32_t *dst, cudaStream_t) {
  for (uint32_t i = 0; i < n; i++) memcpy(dst + size_t(i) * words, src + size_t(idx[i]) * words, words * 4);
}

uint32_t list_unsettled_blocks (uint32_t n) { return (n + 2047) / 2048; }
void launch_list_unsettled (const uint32_t *act, const uint32_t *fam_act, uint32_t n, uint32_t *, uint32_t *list,
                            uint32_t *total, cudaStream_t) {
  uint32_t at = 0;
  for (uint32_t i = 0; i < n; i++) {
    const bool settled = act[i] == ACT_ACCEPTED
                         || (fam_act && act[i] == ACT_FALLTHROUGH && (fam_act[i] & FAM_CODE_MASK) == FAM_JOINED);
    if (!settled) { list[2 * size_t(at)] = i;  list[2 * size_t(at) + 1] = act[i];  at++; }
  }
  if (n) *total = at;
}

void launch_group_births (const uint32_t *msid, const uint32_t *fsid, uint32_t n, const uint32_t *sid_slot, uint32_t ns,
                          uint32_t root_sid, uint32_t *key, uint32_t *total, uint32_t *total_next, uint32_t *start, uint32_t *cursor,
                          uint32_t *gidx, uint32_t *info, uint32_t *info_next, uint32_t *order, uint32_t *rest, uint32_t *h_total,
                          uint32_t *h_info, uint32_t *h_rest, uint32_t *h_order, int32_t *h_birth_group, cudaStream_t) {
  if (!n || !ns) return;
  uint32_t nr = 0;
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t m = msid[i], f = fsid[i];
    uint32_t k = 0xFFFFFFFFu;
    if (f == 0xFFFFFFFFu || m == f) {
      if (m == 0xFFFFFFFFu) k = root_sid;
      else if (m < ns && sid_slot[m] != 0xFFFFFFFFu) k = m | ((m == f ? 2u : 1u) << ORDER_CLASS_SHIFT);
    }
    key[i] = k;
    if (k != 0xFFFFFFFFu) total[k & ORDER_INDEX_MASK]++;
    else rest[nr++] = i;
  }
  // (the device appends with atomics: any order.  Reversed here, so that the host's sort is exercised)
  for (uint32_t e = 0; e < nr; e++) h_rest[e] = rest[nr - 1 - e];
  uint32_t run = 0, cnt = 0, mx = 0;
  for (uint32_t s = 0; s < ns; s++) {
    start[s] = cursor[s] = run;  gidx[s] = cnt;  h_total[s] = total[s];
    run += total[s];  cnt += total[s] ? 1u : 0u;  mx = total[s] > mx ? total[s] : mx;
    total_next[s] = 0;
  }
  info[0] = cnt;  info[1] = mx;  info[2] = run;  info[3] = mx > 2048 ? 1u : 0u;      // (the device kernel's sorting limit, so that the hand-back is exercised)
  info[4] = nr;
  for (int q = 0; q < 5; q++) h_info[q] = info[q];
  for (int q = 0; q < 8; q++) info_next[q] = 0;
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t k = key[i];
    if (k == 0xFFFFFFFFu) { h_birth_group[i] = -1;  continue; }
    const uint32_t sid = k & ORDER_INDEX_MASK;
    const uint32_t at = cursor[sid]++;
    order[at] = h_order[at] = i | (k & ~ORDER_INDEX_MASK);
    h_birth_group[i] = int32_t(gidx[sid]);
  }
}

void launch_gather_listed_rows (const uint32_t *src, const uint32_t *list, const uint32_t *count, uint32_t max_entries,
                                uint32_t words, uint32_t *dst, cudaStream_t) {
  const uint32_t n = *count < max_entries ? *count : max_entries;
  for (uint32_t e = 0; e < n; e++) memcpy(dst + size_t(e) * words, src + size_t(list[2 * e]) * words, words * 4);
}

void launch_scatter_rows (const uint32_t *sp, const float4 *sm, const uint32_t *dst, uint32_t n, uint32_t wp,
                          uint32_t *pp, float4 *pm, cudaStream_t) {
  for (uint32_t i = 0; i < n; i++) {
    memcpy(pp + size_t(dst[i]) * wp, sp + size_t(i) * wp, wp * 4);
    pm[dst[i]] = sm[i];
  }
}

void launch_scatter_u32 (const uint32_t *idx, const uint32_t *val, uint32_t n, uint32_t *dst, cudaStream_t) {
  for (uint32_t i = 0; i < n; i++) dst[idx[i]] = val[i];
}

void launch_scatter_blocks_f32 (const float *src, const uint32_t *idx, uint32_t n, uint32_t len, float *dst, cudaStream_t) {
  for (uint32_t i = 0; i < n; i++) memcpy(dst + size_t(idx[i]) * len, src + size_t(i) * len, len * sizeof(float));
}

bool stream_kernel_supports (uint32_t wp, uint32_t, int crit) {
  const uint32_t wq = wp >> 2;
  return (wq == 1 || wq == 2 || wq == 4 || wq == 8 || wq == 16 || wq == 32) && crit >= 0 && crit <= 3;
}

// Plain sequential stand-in for resolve_streams_kernel (same inputs, outputs and write-back)
void launch_resolve_streams (const StreamArgs &a, cudaStream_t) {
  if (a.preload) return;
  const uint32_t K = a.pool.K, wp = a.pool.wp;
  const float FMAX = std::numeric_limits<float>::max();
  for (uint32_t g = 0; g < a.n_groups; g++) {
    const uint32_t slot = a.grp_slot[g];
    uint32_t k = a.pool.fill[slot] < K ? a.pool.fill[slot] : K;
    std::vector<uint32_t> reps (a.pool.payload + size_t(slot) * K * wp, a.pool.payload + size_t(slot + 1) * K * wp);
    std::vector<float4> meta (a.pool.cmeta + size_t(slot) * K, a.pool.cmeta + size_t(slot + 1) * K);
    std::vector<float> D (a.pool_dist + size_t(slot) * K * K, a.pool_dist + size_t(slot + 1) * K * K);
    uint32_t changed = 0, nAcc[3] = {0, 0, 0}, nRej = 0;
    bool aborted = false;
    for (uint32_t t = a.grp_off[g]; t < a.grp_off[g + 1]; t++) {
      const uint32_t b = a.order[t] & ORDER_INDEX_MASK, ccls = a.order[t] >> ORDER_CLASS_SHIFT;
      const uint32_t *bw = a.b_payload + size_t(b) * wp;
      const float4 bm = a.b_cmeta[b];
      float d[KMAX];
      float avg = 0;  uint32_t matable = 0;
      for (uint32_t r = 0; r < k; r++) {
        const double dd = dist(a.cfg.kind, wp, bw, &reps[size_t(r) * wp]);
        const double ca = compat(a.cfg, bm, dd), cb = compat(a.cfg, meta[r], dd), c = cb < ca ? cb : ca;
        avg += c;  matable += (c >= a.cfg.compat_thr);
        d[r] = float(dd);
      }
      const float score = a.cfg.full_continuous ? float(avg / float(k) - a.cfg.avg_thr) : float(matable - a.cfg.sim_thr * k);
      if (!a.cfg.use_lut && std::fabs(score) <= a.guard) { a.act[b] = ACT_UNCERTAIN; aborted = true; break; }
      if (!(score > 0)) {
        a.act[b] = ACT_FALLTHROUGH;
        if (a.ft_list) { const uint32_t at = (*a.ft_counter)++; a.ft_list[2 * at] = b; a.ft_list[2 * at + 1] = g; }
        if (a.rej_order) { a.rej_order[a.grp_off[g] + nRej] = a.order[t]; a.rej_fill[a.grp_off[g] + nRej] = k; nRej++; }
        continue;
      }
      uint32_t than = 0xFFFFFFFFu;
      bool modifies = false, append = false;
      if (k < K) { modifies = append = true; than = k; }
      else if (K > 1 && a.crit >= 2) {
        // criteria 2 / 3: the very functions the kernel runs (criteria.h), exp() perturbed like a device's
        auto Dm = [&] (uint32_t i, uint32_t j) { return D[i * K + j]; };
        auto gv = [&] (uint32_t j) { return d[j]; };
        auto dexp = [] (double x) { const double e = std::exp(x); return (uint64_t(x * -7919.) & 3) == 0 ? std::nextafter(e, 2.) : e; };
        const CritResult r = a.crit == 2 ? crit_maxAvgMinStdDev(Dm, gv, K) : crit_maxWeightedDist2Avg<KMAX>(Dm, gv, K, dexp);
        if (r.uncertain) { a.act[b] = ACT_UNCERTAIN; aborted = true; break; }
        than = r.than;  modifies = r.better;
      }
      else if (K > 1) {
        float value = -std::numeric_limits<float>::infinity();
        for (uint32_t i = 0; i < K; i++) {
          double c;
          if (a.crit == 1) {
            float minBase = FMAX, minNew = FMAX;
            for (uint32_t j = 0; j < K; j++) { if (i == j) continue; minBase = std::min(minBase, D[i * K + j]); minNew = std::min(minNew, d[j]); }
            c = - minBase + minNew;
          } else {
            c = 0;
            for (uint32_t j = 0; j < K; j++) { if (i == j) continue; c += - D[i * K + j] + d[j]; }
          }
          if (value < c) { value = float(c); than = i; }
        }
        modifies = value > 0;
      }
      a.out_sid[b] = a.grp_sid[g];
      if (!modifies) { a.act[b] = ACT_ACCEPTED; nAcc[ccls]++; continue; }
      a.out_slot[b] = int32_t(than);
      const uint32_t row = (*a.row_counter)++;
      for (uint32_t r = 0; r < k; r++) {
        a.rows[size_t(row) * K + r] = d[r];
        if (r != than) D[r * K + than] = D[than * K + r] = d[r];
      }
      memcpy(&reps[size_t(than) * wp], bw, wp * 4);
      if (a.row_payload) memcpy(a.row_payload + size_t(row) * wp, bw, wp * 4);
      meta[than] = bm;
      a.act[b] = ACT_ENTERED | (append ? 0u : ACT_REPLACED) | (than << ACT_SLOT_SHIFT) | (row << ACT_ROW_SHIFT);
      if (append) k++;
      changed |= 1u << than;
    }
    a.grp_sum[g] = make_uint4(nAcc[0], nAcc[1], nAcc[2], aborted ? 1u : 0u);
    if (a.rej_order) a.rej_count[g] = nRej;
    if (changed && !aborted) {
      for (uint32_t r = 0; r < k; r++) {
        if (!((changed >> r) & 1u)) continue;
        memcpy(a.pool_payload_w + (size_t(slot) * K + r) * wp, &reps[size_t(r) * wp], wp * 4);
        a.pool_cmeta_w[size_t(slot) * K + r] = meta[r];
      }
      memcpy(a.pool_dist + size_t(slot) * K * K, D.data(), size_t(K) * K * sizeof(float));
      a.pool_fill_w[slot] = k;
    }
  }
}

bool family_kernel_supports (uint32_t wp, uint32_t K, int crit) {
  const uint32_t wq = wp >> 2;
  return (wq == 1 || wq == 2 || wq == 4 || wq == 8 || wq == 16 || wq == 32) && (crit == 0 || crit == 1) && K >= 1 && K <= KMAX;
}
uint32_t family_capacity (uint32_t, uint32_t, const ScoreCfg&) { return 96; }     // small on purpose: tests reach the hand-back paths

// Plain sequential stand-in for resolve_families_kernel (same inputs, outputs and write-back): the reference's
// findBestDerived / insertInto / makeNode for the rejected births of a family, in order, on local copies.
void launch_resolve_families (const FamilyArgs &a, cudaStream_t) {
  if (a.preload) return;
  const uint32_t K = a.pool.K, wp = a.pool.wp;
  const float FMAX = std::numeric_limits<float>::max();
  struct Child { std::vector<uint32_t> reps; std::vector<float4> meta; std::vector<float> D; uint32_t k, sid, changed, join[2]; };
  for (uint32_t f = 0; f < a.n_fam; f++) {
    const uint32_t g = a.fam_group[f], g0 = a.grp_off[g], nRej = a.rej_count[g];
    const uint32_t c0 = a.child_off[g], Cs = a.child_off[g + 1] - c0;
    std::vector<Child> ch;
    for (uint32_t c = 0; c < Cs; c++) {
      const uint32_t slot = a.child_slot[c0 + c];
      Child x;
      x.reps.assign(a.pool.payload + size_t(slot) * K * wp, a.pool.payload + size_t(slot + 1) * K * wp);
      x.meta.assign(a.pool.cmeta + size_t(slot) * K, a.pool.cmeta + size_t(slot + 1) * K);
      x.D.assign(a.pool_dist + size_t(slot) * K * K, a.pool_dist + size_t(slot + 1) * K * K);
      x.k = a.pool.fill[slot] < K ? a.pool.fill[slot] : K;
      x.sid = a.child_sid[c0 + c];  x.changed = 0;  x.join[0] = x.join[1] = 0;
      ch.push_back(std::move(x));
    }
    uint32_t nNew = 0, nDone = 0;
    bool handedBack = false;
    unsigned long long branching = 0, comparisons = 0;
    auto scoreOf = [&] (const Child &x, const uint32_t *bw, float4 bm, float *d) {
      float avg = 0;  uint32_t matable = 0;
      for (uint32_t r = 0; r < x.k; r++) {
        const double dd = dist(a.cfg.kind, wp, bw, &x.reps[size_t(r) * wp]);
        const double ca = compat(a.cfg, bm, dd), cb = compat(a.cfg, x.meta[r], dd), c = cb < ca ? cb : ca;
        avg += c;  matable += (c >= a.cfg.compat_thr);
        d[r] = float(dd);
      }
      return a.cfg.full_continuous ? float(avg / float(x.k) - a.cfg.avg_thr) : float(matable - a.cfg.sim_thr * x.k);
    };
    for (uint32_t t = 0; t < nRej && !handedBack; t++) {
      const uint32_t ord = a.rej_order[g0 + t], b = ord & ORDER_INDEX_MASK, ccls = ord >> ORDER_CLASS_SHIFT;
      const uint32_t *bw = a.b_payload + size_t(b) * wp;
      const float4 bm = a.b_cmeta[b];
      // visiting order: founded here newest first, then the snapshot list
      std::vector<uint32_t> visit;
      for (uint32_t c = uint32_t(ch.size()); c-- > Cs; ) visit.push_back(c);
      for (uint32_t c = 0; c < Cs; c++) visit.push_back(c);
      uint32_t j = 0xFFFFFFFFu, vb = 0, vc = 0;
      float d[KMAX];
      bool uncertain = false;
      for (uint32_t c: visit) {
        vb++;  vc += ch[c].k;
        const float sc = scoreOf(ch[c], bw, bm, d);
        if (!a.cfg.use_lut && std::fabs(sc) <= a.guard) { uncertain = true; break; }
        if (sc > 0) { j = c; break; }
      }
      if (uncertain) { handedBack = true; break; }
      uint32_t act, slot = 0, kOld = 0;
      if (j == 0xFFFFFFFFu) {
        if (nNew == a.max_new || ch.size() == a.cap) { handedBack = true; break; }
        const uint32_t ps = (*a.slot_counter)++;
        if (ps >= a.slot_limit) { handedBack = true; break; }
        a.new_slot[size_t(f) * a.max_new + nNew] = a.slot_base + ps;
        Child x;
        x.reps.assign(size_t(K) * wp, 0u);  x.meta.assign(K, make_float4(0, 1, 1, 0));  x.D.assign(size_t(K) * K, 0.f);
        x.k = 0;  x.sid = FAM_NEW_SID | (f * a.max_new + nNew);  x.changed = 0;  x.join[0] = x.join[1] = 0;
        a.new_founder[size_t(f) * a.max_new + nNew] = b;
        j = uint32_t(ch.size());
        ch.push_back(std::move(x));
        nNew++;
        act = FAM_FOUNDED;
      } else {
        Child &x = ch[j];
        kOld = x.k;
        if (x.k < K) { slot = x.k; act = FAM_ENTERED; }
        else {
          act = FAM_JOINED;
          if (K > 1) {
            float value = -std::numeric_limits<float>::infinity();
            uint32_t than = 0xFFFFFFFFu;
            for (uint32_t i = 0; i < K; i++) {
              double c;
              if (a.crit == 1) {
                float minBase = FMAX, minNew = FMAX;
                for (uint32_t q = 0; q < K; q++) { if (i == q) continue; minBase = std::min(minBase, x.D[i * K + q]); minNew = std::min(minNew, d[q]); }
                c = - minBase + minNew;
              } else {
                c = 0;
                for (uint32_t q = 0; q < K; q++) { if (i == q) continue; c += - x.D[i * K + q] + d[q]; }
              }
              if (value < c) { value = float(c); than = i; }
            }
            if (value > 0) { act = FAM_ENTERED | FAM_REPLACED; slot = than; }
          }
        }
      }
      branching += vb;  comparisons += vc + a.rej_fill[g0 + t];
      Child &x = ch[j];
      a.out_sid[b] = x.sid;
      if ((act & FAM_CODE_MASK) == FAM_JOINED) { x.join[ccls == 2 ? 1 : 0]++; a.fam_act[b] = act | (j << FAM_CHILD_SHIFT); nDone++; continue; }
      const uint32_t row = (*a.row_counter)++;
      for (uint32_t r = 0; r < kOld; r++) {
        a.rows[size_t(row) * K + r] = d[r];
        if (r != slot) x.D[r * K + slot] = x.D[slot * K + r] = d[r];
      }
      memcpy(&x.reps[size_t(slot) * wp], bw, wp * 4);
      if (a.row_payload) memcpy(a.row_payload + size_t(row) * wp, bw, wp * 4);
      x.meta[slot] = bm;
      if (slot == x.k) x.k++;
      x.changed |= 1u << slot;
      a.out_slot[b] = int32_t(slot);
      a.fam_row[b] = row;
      a.fam_act[b] = act | (slot << FAM_SLOT_SHIFT) | (j << FAM_CHILD_SHIFT);
      nDone++;
    }
    a.fam_info[f] = make_uint4(nDone, nNew, handedBack ? 1u : 0u, uint32_t(ch.size()));
    a.fam_stats[2 * size_t(f)] = branching;  a.fam_stats[2 * size_t(f) + 1] = comparisons;
    for (uint32_t c = 0; c < ch.size(); c++) {
      a.child_join[a.fam_join_off[f] + 2 * c] = ch[c].join[0];  a.child_join[a.fam_join_off[f] + 2 * c + 1] = ch[c].join[1];
      if (!ch[c].changed) continue;
      const uint32_t slot = c < Cs ? a.child_slot[c0 + c] : a.new_slot[size_t(f) * a.max_new + (c - Cs)];
      for (uint32_t r = 0; r < ch[c].k; r++) {
        if (!((ch[c].changed >> r) & 1u)) continue;
        memcpy(a.pool_payload_w + (size_t(slot) * K + r) * wp, &ch[c].reps[size_t(r) * wp], wp * 4);
        a.pool_cmeta_w[size_t(slot) * K + r] = ch[c].meta[r];
      }
      memcpy(a.pool_dist + size_t(slot) * K * K, ch[c].D.data(), size_t(K) * K * sizeof(float));
      a.pool_fill_w[slot] = ch[c].k;
    }
  }
}

void launch_classify_children (const ChildArgs &a, cudaStream_t) {
  const uint32_t n = *a.ft_counter < a.max_ft ? *a.ft_counter : a.max_ft;
  for (uint32_t e = 0; e < n; e++) {
    const uint32_t b = a.ft_list[2 * e], g = a.ft_list[2 * e + 1];
    uint32_t first = 0xFFFFFFFFu, flags = 0;
    float sc = 0;
    for (uint32_t c = a.child_off[g]; c < a.child_off[g + 1]; c++) {
      const uint32_t slot = a.child_slot[c], rank = c - a.child_off[g];
      if (slot == 0xFFFFFFFFu) { first = rank; flags = 1; break; }
      sc = score(a.pool, a.cfg, a.b_payload + size_t(b) * a.pool.wp, a.b_cmeta[b], slot, a.out_dist + size_t(e) * a.pool.K);
      if (!a.cfg.use_lut && (std::fabs(sc) <= a.guard || !a.cfg.full_continuous)) { first = rank; flags = 1; break; }
      if (sc > 0) { first = rank; break; }
    }
    a.out_first[e] = first;  a.out_flags[e] = flags;  a.out_score[e] = sc;
  }
}

void launch_env_contribution (int crit, uint32_t K, uint32_t n, const float *edist, const float *gdist, uint8_t *ob, uint32_t *ot,
                              float *ov, cudaStream_t) {
  for (uint32_t i = 0; i < n; i++) {
    const float *e = edist + size_t(i) * K * K, *g = gdist + size_t(i) * K;
    auto D = [&] (uint32_t a, uint32_t b) { return e[a * K + b]; };
    auto G = [&] (uint32_t j) { return g[j]; };
    CritResult r { false, 0xFFFFFFFFu, -std::numeric_limits<float>::infinity(), false };
    if (crit == 2) r = crit_maxAvgMinStdDev(D, G, K);
    else if (crit == 3) r = crit_maxWeightedDist2Avg<KMAX>(D, G, K, [] (double x) { return std::exp(x); });
    else {
      for (uint32_t a = 0; a < K; a++) {
        double c = 0;
        if (crit == 1) {
          float mb = std::numeric_limits<float>::max(), mn = mb;
          for (uint32_t j = 0; j < K; j++) { if (j == a) continue; mb = std::min(mb, e[a * K + j]); mn = std::min(mn, g[j]); }
          c = - mb + mn;
        } else for (uint32_t j = 0; j < K; j++) { if (j == a) continue; c += - e[a * K + j] + g[j]; }
        if (r.value < c) { r.value = float(c); r.than = a; }
      }
      r.better = r.value > 0;
    }
    ob[i] = r.better;  ot[i] = r.than;  ov[i] = r.value;
  }
}
void launch_pairwise (int kind, uint32_t wp, const uint32_t *payload, uint32_t n, const uint32_t *a, const uint32_t *b,
                      float *out, cudaStream_t) {
  for (uint32_t p = 0; p < n; p++) out[p] = float(dist(kind, wp, payload + size_t(a[p]) * wp, payload + size_t(b[p]) * wp));
}
void launch_mate_compat (int, uint32_t, const uint32_t*, const float4*, uint32_t, const uint32_t*, const uint32_t*, float*, double*, cudaStream_t) {}
void launch_classify_dense (const PoolView&, const ScoreCfg&, const uint32_t*, const float4*, uint32_t, uint32_t*, float*,
                            uint32_t*, uint32_t*, cudaStream_t) {}
bool dense_tiled_supports (uint32_t, uint32_t) { return false; }
void launch_classify_dense_tiled (const PoolView&, const ScoreCfg&, const uint32_t*, const float4*, uint32_t, uint32_t*, float*,
                                  uint32_t*, uint32_t*, cudaStream_t, const ShardSink*) {}
void launch_publish_records (const uint32_t*, const float*, const uint32_t*, const uint32_t*, uint32_t, const ShardSink&, cudaStream_t) {}
void launch_combine_shards (const uint4*, uint32_t, uint32_t, uint32_t*, float*, uint32_t*, uint32_t*, cudaStream_t) {}
// tensor-core sweep: never selected under the stub
bool dense_tc_supports (const ScoreCfg&, uint32_t, uint32_t) { return false; }
size_t tc_pool_image_bytes (uint32_t, uint32_t, uint32_t) { return 16; }
size_t tc_pool_meta_bytes (uint32_t, uint32_t) { return 16; }
size_t tc_births_image_bytes (uint32_t, uint32_t) { return 16; }
void launch_tc_expand_pool (const PoolView&, uint8_t*, uint32_t*, cudaStream_t) {}
void launch_tc_expand_births (const uint32_t*, uint32_t, uint32_t, uint8_t*, cudaStream_t) {}
void launch_classify_dense_tc (const PoolView&, const ScoreCfg&, const uint8_t*, const uint8_t*, const uint32_t*, const float4*, uint32_t,
                               uint32_t*, float*, uint32_t*, uint32_t*, uint32_t*, cudaStream_t, const ShardSink*, uint32_t, uint32_t) {}

} // namespace apt
