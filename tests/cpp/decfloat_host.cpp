// Host build of cfs_spmv_b200/csrc/decfloat.cuh for tests/test_decfloat.py:
// the same source the GPU ingest kernel compiles, checked against strtod.
// Test infrastructure only -- nothing in the product links this.
#include <stddef.h>

#include "decfloat.cuh"

static const uint64_t kPow5[CFS_POW5_TABLE_WORDS] = CFS_POW5_TABLE_INIT;

extern "C" {

// parses n tokens laid out back to back (offsets[n+1]); status[i] = 0 parsed,
// 1 = "host must decide"
void cfs_test_parse_doubles(const char *text, const long long *offsets, long n,
                            double *out, int *status) {
  for (long i = 0; i < n; ++i)
    status[i] = cfsb::dec::parse_double(text + offsets[i], text + offsets[i + 1],
                                        kPow5, &out[i]);
}

void cfs_test_parse_ints(const char *text, const long long *offsets, long n,
                         int *out, int *status) {
  for (long i = 0; i < n; ++i)
    status[i] = cfsb::dec::parse_int(text + offsets[i], text + offsets[i + 1],
                                     &out[i]);
}

// token count and the [begin,end) offsets of the first three tokens of a line
int cfs_test_split_line(const char *line, long len, long *bounds) {
  cfsb::dec::LineTokens lt;
  cfsb::dec::split_line(line, line + len, &lt);
  for (int k = 0; k < 3 && k < lt.ntokens; ++k) {
    bounds[2 * k] = lt.tok[k] - line;
    bounds[2 * k + 1] = lt.tok_end[k] - line;
  }
  return lt.ntokens;
}
}
