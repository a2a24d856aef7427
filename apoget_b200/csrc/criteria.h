// Enveloppe criteria 2 and 3 of the reference (src/core/tree/enveloppecriteria.cpp:136-202 maxAvgMinStdDev,
// :205-272 maxWeightedDist2Avg) in a form that runs on the device AND on the host: every floating-point
// operation is spelled out with its rounding (no contraction into FMAs), in the reference's order, so the
// device result is the reference's bit for bit wherever the reference uses IEEE operations only.
//
//  * maxAvgMinStdDev: additions, divisions, a square (`std::pow(x, 2)`, which g++ folds into x * x at the
//    reference's -O2) and std::sqrt -- all correctly rounded on both sides: EXACT on the device.
//  * maxWeightedDist2Avg calls exp(): libm's and the device's differ by an ulp now and then.  The decision
//    (which representative, better or not) is therefore reported together with `uncertain`: set when any
//    comparison the decision rests on is closer than 1e-13 of the terms' scale (the accumulated error is
//    ~1e-15 of it) -- the caller then lets the host decide with libm, as it does for scores in the guard band.
//
// D(i, j): intra-enveloppe distance of representatives i != j (symmetric); g(j): distance of the candidate
// genome to representative j; k = enveloppe size (full enveloppe).  The reference iterates its distance map
// in key order, i.e. over pairs (a, b), a < b, lexicographically; so do the loops below.
#ifndef APT_B200_CRITERIA_H
#define APT_B200_CRITERIA_H

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define APT_CRIT_HD __host__ __device__ inline
#else
#define APT_CRIT_HD inline
#endif

namespace apt {

namespace crit_detail {
#if defined(__CUDA_ARCH__)
__device__ inline double add (double a, double b) { return __dadd_rn(a, b); }
__device__ inline double mul (double a, double b) { return __dmul_rn(a, b); }
__device__ inline double div (double a, double b) { return __ddiv_rn(a, b); }
__device__ inline double root (double a) { return __dsqrt_rn(a); }
#else
inline double add (double a, double b) { volatile double r = a + b; return r; }     // (volatile: no contraction on the host either)
inline double mul (double a, double b) { volatile double r = a * b; return r; }
inline double div (double a, double b) { return a / b; }
inline double root (double a) { return sqrt(a); }
#endif
constexpr float NEG_INF = -__builtin_huge_valf();
}

struct CritResult {
  bool better;        ///< :199 / :269
  uint32_t than;      ///< index of the representative to replace, 0xFFFFFFFF when none
  float value;
  bool uncertain;     ///< (criterion 3) the decision is too close to call with the device's exp()
};

/// enveloppecriteria.cpp:29-37 over the map D with row/column `swap` replaced by g (swap >= k: the map itself)
template <class DM, class GV>
APT_CRIT_HD void crit_avg_stddev (const DM &D, const GV &g, uint32_t k, uint32_t swap, double &avg, double &stdDev) {
  using namespace crit_detail;
  const double n = double(k * (k - 1) / 2);
  auto at = [&] (uint32_t a, uint32_t b) -> double {
    return a == swap ? double(g(b)) : (b == swap ? double(g(a)) : double(D(a, b)));
  };
  avg = 0.;
  for (uint32_t a = 0; a < k; a++) for (uint32_t b = a + 1; b < k; b++) avg = add(avg, at(a, b));
  avg = div(avg, n);
  stdDev = 0.;
  for (uint32_t a = 0; a < k; a++)
    for (uint32_t b = a + 1; b < k; b++) { const double t = add(avg, -at(a, b)); stdDev = add(stdDev, mul(t, t)); }
  stdDev = root(div(stdDev, n));
}

/// enveloppecriteria.cpp:136-202
template <class DM, class GV>
APT_CRIT_HD CritResult crit_maxAvgMinStdDev (const DM &D, const GV &g, uint32_t k) {
  using namespace crit_detail;
  CritResult ec { false, 0xFFFFFFFFu, NEG_INF, false };
  double baseAVG, baseStdDev;
  crit_avg_stddev(D, g, k, 0xFFFFFFFFu, baseAVG, baseStdDev);
  for (uint32_t i = 0; i < k; i++) {
    double newAVG, newStdDev;
    crit_avg_stddev(D, g, k, i, newAVG, newStdDev);
    const double c = add(add(add(-baseAVG, newAVG), baseStdDev), -newStdDev);        // :162-163
    if (double(ec.value) < c) { ec.value = float(c); ec.than = i; }                   // :191-194
  }
  ec.better = ec.value > 0.f;
  return ec;
}

/// enveloppecriteria.cpp:205-272.  `expfn` = exp(); d_i / d_g sorted in decreasing order (:248, the k zeros the
/// vectors start with (:231) sort behind every distance and are never reached by the k - 1 terms of :249).
template <uint32_t KCAP, class DM, class GV, class EXP>
APT_CRIT_HD CritResult crit_maxWeightedDist2Avg (const DM &D, const GV &g, uint32_t k, const EXP &expfn) {
  using namespace crit_detail;
  CritResult ec { false, 0xFFFFFFFFu, NEG_INF, false };
  double A = 0.;
  for (uint32_t a = 0; a < k; a++) for (uint32_t b = a + 1; b < k; b++) A = add(A, double(D(a, b)));
  A = div(A, double(k * (k - 1) / 2));
  const double denom = div(mul(mul(2., A), A), 16.);                                  // 2. * A * A / 16.
  for (uint32_t i = 0; i < k; i++) {
    double di[KCAP], dg[KCAP];
    uint32_t m = 0;
    for (uint32_t j = 0; j < k; j++) {
      if (j == i) continue;
      // insertion into decreasing order (equal values are interchangeable)
      const double vi = double(D(i, j)), vg = double(g(j));
      uint32_t p = m;
      while (p > 0 && di[p - 1] < vi) { di[p] = di[p - 1]; p--; }
      di[p] = vi;
      p = m;
      while (p > 0 && dg[p - 1] < vg) { dg[p] = dg[p - 1]; p--; }
      dg[p] = vg;
      m++;
    }
    double c = 0., scale = 0.;
    for (uint32_t j = 0; j + 1 < k; j++) {
      const double nc = -di[j], pc = dg[j];
      const double t = add(pc, -A);
      const double w = add(1., -expfn(div(-mul(t, t), denom)));                       // 1 - exp(-(d-A)*(d-A) / (2.*A*A/16.))
      const double s = add(nc, pc);
      c = add(c, mul(w, s));
      scale += fabs(s);
    }
    const double eps = 1e-13 * scale + 1e-300;
    const double v = double(ec.value);
    if (fabs(v - c) <= eps) ec.uncertain = true;                                      // `ec.value < c` too close to call
    if (v < c) {
      if (float(c - eps) != float(c + eps)) ec.uncertain = true;                      // ... or the rounding to float
      ec.value = float(c);  ec.than = i;
      if (fabs(c) <= eps) ec.uncertain = true;                                        // ... or the sign, should this one win
    }
  }
  ec.better = ec.value > 0.f;
  return ec;
}

} // namespace apt

#endif
