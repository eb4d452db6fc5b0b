// platform.hpp -- enumerations and small helpers of cfs::util
// (API of reference include/utils/platform.hpp:20-37).
#ifndef PLATFORM_HPP
#define PLATFORM_HPP

#include <cmath>

#include "cfs_config.hpp"

namespace cfs {
namespace util {

using namespace std;

// The reference knows only `cpu`. Here `cpu` means "host pointers in, host
// pointers out": the arithmetic still runs on the B200 (there is no CPU path).
enum class Platform { cpu, gpu };
enum class Kernel { SpDMV };
enum class Tuning { None, Aggressive };
enum class Format { none, csr, sss, hyb };

inline int iceildiv(const int num, const int den) {
  return num / den + (num % den != 0);
}

// relative comparison used by test_spmv_mmf (eps 1e-4 single, 1e-8 double)
template <typename Real> inline bool approx_equal(Real a, Real b, Real eps) {
  return std::fabs(a - b) <= eps * std::fabs(a);
}
inline bool isEqual(float a, float b) { return approx_equal(a, b, 1e-4f); }
inline bool isEqual(double a, double b) { return approx_equal(a, b, 1e-8); }

} // namespace util
} // namespace cfs

#endif
