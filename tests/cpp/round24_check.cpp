// Checks apt::round_to_float (apoget_b200/csrc/round24.h) against the hardware conversion
// double -> float -> double: random bit patterns over float's exponent range (incl. subnormal
// results), sums shaped like the scoring loop's (float + double in [0,1]), ties and near-ties.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>

#include "round24.h"

static uint64_t bad = 0, checked = 0;

static void check (double s) {
  const double want = double(float(s)), got = apt::round_to_float(s);
  checked++;
  // bit-identical for s >= 0 (the scoring loop's domain); negative inputs that underflow give +0 instead of -0
  if (!std::signbit(s) ? memcmp(&want, &got, 8) != 0 : want != got) {
    if (bad++ < 10) printf("MISMATCH s=%a want=%a got=%a\n", s, want, got);
  }
}

int main (int argc, char **argv) {
  const uint64_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 20000000ull;
  std::mt19937_64 rng (12345);
  // 1. random doubles with exponents from below float's subnormals to 2^100, both signs
  for (uint64_t i = 0; i < n; i++) {
    uint64_t b = rng();
    const uint64_t e = 1023 - 160 + (b >> 52) % 261;
    b = (b & 0x800FFFFFFFFFFFFFull) | (e << 52);
    double s;  memcpy(&s, &b, 8);
    check(s);
  }
  // 2. the scoring loop's shape: float running sum + double compatibility
  std::uniform_real_distribution<double> u (0., 1.);
  for (uint64_t i = 0; i < n; i++) {
    float avg = 0.f;
    double shadow = 0.;
    for (int r = 0; r < 10; r++) {
      double c = u(rng);
      if (i % 7 == 0) c = std::ldexp(c, -int(rng() % 160));      // tiny compatibilities (far tails)
      check(double(avg) + c);
      avg = float(double(avg) + c);
      shadow = apt::round_to_float(shadow + c);
      if (double(avg) != shadow) { if (bad++ < 10) printf("CHAIN MISMATCH %a %a\n", double(avg), shadow); }
    }
  }
  // 3. exact ties and their neighbours in every binade (float ulp / 2 +- one double ulp)
  for (int e = -152; e <= 100; e++)
    for (uint32_t m = 0; m < 4096; m++) {
      const uint32_t mant = uint32_t(rng()) & 0x7FFFFF;
      const double base = std::ldexp(1. + mant / 8388608., e);
      const double half = std::ldexp(1., e - 24);
      for (int k = -2; k <= 2; k++) {
        double s = base + half;
        uint64_t b;  memcpy(&b, &s, 8);  b += int64_t(k);  memcpy(&s, &b, 8);
        check(s);  check(-s);
      }
    }
  check(0.);  check(-0.);
  check(std::ldexp(1., -149));  check(std::ldexp(1., -150));  check(std::ldexp(1.5, -150));  check(std::ldexp(1., -1074));
  printf("checked %llu, mismatches %llu\n", (unsigned long long) checked, (unsigned long long) bad);
  return bad ? 1 : 0;
}
