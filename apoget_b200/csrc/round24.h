// round_to_float: a double rounded to float precision (round-to-nearest-even, float's exponent
// range incl. subnormals) WITHOUT leaving the double domain -- i.e. double(float(s)), bit for bit.
//
// Adding M = 1.5 * 2^(e+29), e = max(exponent(s), -126), makes the double addition round s at
// 2^(e-23), float's ulp in that binade, with the same ties-to-even rule (M / 2^(e-23) is even);
// subtracting M again is exact.  Used where the reference's `float avg += double c`
// (phylogenetictree.hpp:465-472) sits in an inner loop on the device: F2F conversions share the
// quarter-rate pipe with POPC there, these three operations do not.  Host + device; the host copy
// is checked against the hardware conversion in tests/test_host_logic.py.
#ifndef APT_B200_ROUND24_H
#define APT_B200_ROUND24_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define APT_HD __host__ __device__ __forceinline__
#else
#define APT_HD inline
#endif

namespace apt {

/// |s| < 2^127 (no float overflow); NaN stays NaN; a negative s that underflows gives +0 (the float is -0)
APT_HD double round_to_float (double s) {
#if defined(__CUDA_ARCH__)
  // exponent field, clamped; its low 20 bits are zero, so "+ 29 exponents | top mantissa bit" is one add
  const int e = min(max(__double2hiint(s) & 0x7FF00000, 897 << 20), 1150 << 20);
  const double M = __hiloint2double(e + ((29 << 20) | 0x00080000), 0);
  return __dadd_rn(__dadd_rn(s, M), -M);
#else
  uint64_t b;
  memcpy(&b, &s, 8);
  int64_t e = int64_t((b >> 52) & 0x7FF);
  e = e < 897 ? 897 : (e > 1150 ? 1150 : e);
  const uint64_t mb = (uint64_t(e + 29) << 52) | (uint64_t(1) << 51);
  double M;
  memcpy(&M, &mb, 8);
  volatile double t = s + M;      // keep the compiler from folding (s + M) - M
  return t - M;
#endif
}

} // namespace apt

#endif
