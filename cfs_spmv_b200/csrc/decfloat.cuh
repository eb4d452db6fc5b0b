// decfloat.cuh -- number parsing for the GPU Matrix Market ingest, usable from
// host and device code.
//
// The reference's loader converts tokens with atoi / atof (include/io/mmf.hpp:
// 309-343 -> glibc strtol / strtod): the value of an entry is the CORRECTLY
// ROUNDED binary64 of the decimal string. To be bit-exact on the GPU the same
// function has to be computed there:
//   * Clinger's fast path: at most 19 digits, mantissa <= 2^53 and |exponent|
//     <= 22 -> one IEEE multiplication or division of two exact doubles;
//   * otherwise the Eisel-Lemire algorithm (D. Lemire, "Number parsing at a
//     gigabyte per second", SPE 2021): a 64x128-bit product with a truncated
//     power of five (pow5_table.h) decides the rounding in all but a few cases;
//   * anything it cannot decide (or that is outside the plain decimal grammar:
//     hex floats, inf/nan, more than 19 digits with an ambiguous tail) is
//     REPORTED, not guessed: the line goes to the host's strtod
//     (mmf_ingest.cu patches it in), so results are always those of the
//     reference.
// tests/test_decfloat.py runs the host build of this header against strtod.
#pragma once

#include <stdint.h>

#include "pow5_table.h"

#if defined(__CUDACC__)
#define CFS_HD __host__ __device__ inline
#else
#define CFS_HD inline
#endif

namespace cfsb {
namespace dec {

enum { kParsed = 0, kNeedHost = 1 };

struct Product {
  uint64_t lo, hi;
};

CFS_HD Product mul_64x64(uint64_t a, uint64_t b) {
  Product p;
#if defined(__CUDA_ARCH__)
  p.lo = a * b;
  p.hi = __umul64hi(a, b);
#else
  const unsigned __int128 w = (unsigned __int128)a * b;
  p.lo = (uint64_t)w;
  p.hi = (uint64_t)(w >> 64);
#endif
  return p;
}

CFS_HD int leading_zeros(uint64_t w) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)w);
#else
  return __builtin_clzll(w);
#endif
}

CFS_HD bool is_space(char c) { // isspace() of the "C" locale
  return c == ' ' || (c >= '\t' && c <= '\r');
}
CFS_HD bool is_digit(char c) { return c >= '0' && c <= '9'; }

// binary64 fields of w * 10^q (w != 0), or undecided.
struct Rounded {
  uint64_t mantissa;
  int power2;
  bool decided;
};

CFS_HD Rounded eisel_lemire(int64_t q, uint64_t w, const uint64_t *pow5) {
  Rounded r;
  r.decided = true;
  r.mantissa = 0;
  r.power2 = 0;
  if (q < CFS_POW5_MIN_Q) // below half the smallest subnormal even for 19 digits
    return r;
  if (q > CFS_POW5_MAX_Q) {
    r.power2 = 0x7ff;
    return r;
  }
  const int lz = leading_zeros(w);
  w <<= lz;
  // w * 5^q to 55 bits of precision: the second table word only when the
  // first product leaves the rounding open
  const int idx = 2 * (int)(q - CFS_POW5_MIN_Q);
  Product p = mul_64x64(w, pow5[idx]);
  if ((p.hi & 0x1ffu) == 0x1ffu) {
    const Product p2 = mul_64x64(w, pow5[idx + 1]);
    p.lo += p2.hi;
    if (p2.hi > p.lo)
      ++p.hi;
  }
  if (p.lo == 0xffffffffffffffffULL && !(q >= -27 && q <= 55)) {
    r.decided = false; // the truncated table cannot tell: exact arithmetic needed
    return r;
  }
  const int upper = (int)(p.hi >> 63);
  const int shift = upper + 64 - 52 - 3;
  r.mantissa = p.hi >> shift;
  // floor(log2(5^q)) + q + 63, the published integer approximation
  const int log2_10q = (int)(((int64_t)(152170 + 65536) * q) >> 16) + 63;
  r.power2 = log2_10q + upper - lz + 1023;
  if (r.power2 <= 0) { // subnormal result
    if (-r.power2 + 1 >= 64) {
      r.power2 = 0;
      r.mantissa = 0;
      return r;
    }
    r.mantissa >>= -r.power2 + 1;
    r.mantissa += r.mantissa & 1;
    r.mantissa >>= 1;
    r.power2 = r.mantissa < (1ULL << 52) ? 0 : 1;
    return r;
  }
  // exactly half way between two doubles: round to even
  if (p.lo <= 1 && q >= -4 && q <= 23 && (r.mantissa & 3) == 1 &&
      (r.mantissa << shift) == p.hi)
    r.mantissa &= ~1ULL;
  r.mantissa += r.mantissa & 1;
  r.mantissa >>= 1;
  if (r.mantissa >= (2ULL << 52)) {
    r.mantissa = 1ULL << 52;
    ++r.power2;
  }
  r.mantissa &= ~(1ULL << 52);
  if (r.power2 >= 0x7ff) {
    r.power2 = 0x7ff;
    r.mantissa = 0;
  }
  return r;
}

CFS_HD double from_bits(uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)b);
#else
  union {
    uint64_t u;
    double d;
  } c;
  c.u = b;
  return c.d;
#endif
}

// strtod(begin, 0) for a token that ends at `end` (the reference hands atof a
// std::string token). Longest-valid-prefix semantics: trailing characters are
// ignored. Returns kNeedHost when this parser must not decide.
CFS_HD int parse_double(const char *p, const char *end, const uint64_t *pow5,
                        double *out) {
  while (p < end && is_space(*p))
    ++p;
  bool negative = false;
  if (p < end && (*p == '-' || *p == '+')) {
    negative = *p == '-';
    ++p;
  }
  // inf / nan / no number at all: leave it to strtod
  if (p >= end || !(is_digit(*p) || *p == '.'))
    return kNeedHost;
  if (*p == '0' && p + 1 < end && (p[1] == 'x' || p[1] == 'X'))
    return kNeedHost; // hexadecimal floating constant
  uint64_t w = 0;
  int ndigits = 0;      // significant digits held by w (<= 19)
  int64_t exp10 = 0;
  bool any = false, truncated = false;
  for (; p < end && is_digit(*p); ++p) {
    const unsigned d = (unsigned)(*p - '0');
    any = true;
    if (ndigits < 19) {
      if (w != 0 || d != 0) {
        w = w * 10 + d;
        ++ndigits;
      }
    } else {
      ++exp10;
      truncated |= d != 0;
    }
  }
  if (p < end && *p == '.') {
    ++p;
    for (; p < end && is_digit(*p); ++p) {
      const unsigned d = (unsigned)(*p - '0');
      any = true;
      if (ndigits < 19) {
        if (w != 0 || d != 0) {
          w = w * 10 + d;
          ++ndigits;
        }
        --exp10;
      } else {
        truncated |= d != 0;
      }
    }
  }
  if (!any)
    return kNeedHost; // a lone '.'
  if (p < end && (*p == 'e' || *p == 'E')) {
    const char *q = p + 1;
    bool eneg = false;
    if (q < end && (*q == '-' || *q == '+')) {
      eneg = *q == '-';
      ++q;
    }
    if (q < end && is_digit(*q)) {
      int64_t e = 0;
      for (; q < end && is_digit(*q); ++q)
        if (e < 100000)
          e = e * 10 + (*q - '0');
      exp10 += eneg ? -e : e;
    }
  }
  const uint64_t sign = negative ? 0x8000000000000000ULL : 0;
  if (w == 0) {
    *out = from_bits(sign);
    return kParsed;
  }
  if (!truncated && w <= (1ULL << 53) && exp10 >= -22 && exp10 <= 22) {
    // both operands are exact doubles: one correctly rounded operation
    const double p10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,
                            1e8,  1e9,  1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                            1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    double v = (double)w;
#if defined(__CUDA_ARCH__)
    v = exp10 < 0 ? __ddiv_rn(v, p10[-exp10]) : __dmul_rn(v, p10[exp10]);
#else
    v = exp10 < 0 ? v / p10[-exp10] : v * p10[exp10];
#endif
    *out = negative ? -v : v;
    return kParsed;
  }
  const Rounded r = eisel_lemire(exp10, w, pow5);
  if (!r.decided)
    return kNeedHost;
  if (truncated) {
    // digits were dropped: the true value lies in (w, w+1) * 10^exp10; decided
    // only if both ends round to the same double
    const Rounded r1 = eisel_lemire(exp10, w + 1, pow5);
    if (!r1.decided || r1.mantissa != r.mantissa || r1.power2 != r.power2)
      return kNeedHost;
  }
  *out = from_bits(sign | ((uint64_t)r.power2 << 52) | r.mantissa);
  return kParsed;
}

// (int)strtol(begin, 0, 10) for a token that ends at `end` (atoi of the
// reference). No digits -> 0, like strtol. More than 18 digits -> host.
CFS_HD int parse_int(const char *p, const char *end, int32_t *out) {
  while (p < end && is_space(*p))
    ++p;
  bool negative = false;
  if (p < end && (*p == '-' || *p == '+')) {
    negative = *p == '-';
    ++p;
  }
  int64_t v = 0;
  int nd = 0;
  for (; p < end && is_digit(*p); ++p) {
    if (++nd > 18)
      return kNeedHost;
    v = v * 10 + (*p - '0');
  }
  if (negative)
    v = -v;
  *out = (int32_t)(uint32_t)(uint64_t)v; // the int conversion of a long
  return kParsed;
}

// One entry line [begin, end) (without its '\n'): trimmed of ' ' and '\t' at
// both ends, cut at single ' ' with empty tokens dropped (src/mmf.cpp:6-44).
// ntokens counts all tokens; tok[k] / tok_end[k] delimit the first three.
struct LineTokens {
  const char *tok[3];
  const char *tok_end[3];
  int ntokens;
};

CFS_HD void split_line(const char *begin, const char *end, LineTokens *lt) {
  while (begin < end && (*begin == ' ' || *begin == '\t'))
    ++begin;
  while (end > begin && (end[-1] == ' ' || end[-1] == '\t'))
    --end;
  lt->ntokens = 0;
  const char *p = begin;
  while (p < end) {
    const char *q = p;
    while (q < end && *q != ' ')
      ++q;
    if (q > p) {
      if (lt->ntokens < 3) {
        lt->tok[lt->ntokens] = p;
        lt->tok_end[lt->ntokens] = q;
      }
      ++lt->ntokens;
    }
    p = q + 1;
  }
}

} // namespace dec
} // namespace cfsb
