/*
 * cfs_oracle.c -- CPU restatement of the reference's LIVE symmetric-SpMV chain.
 *
 * TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load this. The product (cfs_spmv_b200/, include/)
 * never links, imports or executes anything under oracle/.
 *
 * Parity is PINNED: tests/test_oracle_golden.py checks this file bit-for-bit
 * (metadata AND y) against dumps of the compiled, unmodified reference
 * (oracle/_ref/ref_tool, fixtures under tests/golden/).
 *
 * What is restated (all citations relative to /root/reference):
 *   partition_by_nrows            include/matrix/csr_matrix.tpp:404-435
 *   serial() / per-partition lower-triangle extraction
 *                                 csr_matrix.tpp:642-706, 1216-1348
 *   block-row conflict graph      csr_matrix.tpp:1364-1366, 1427-1477
 *   color_greedy + 2-colour balancing
 *                                 csr_matrix.tpp:2010-2213
 *   per-colour consecutive-row ranges
 *                                 csr_matrix.tpp:1544-1627
 *   size()                        csr_matrix.tpp:191-228
 *   cpu_mv_sym_serial             csr_matrix.tpp:2707-2729
 *   cpu_mv_sym_conflict_free_v2   csr_matrix.tpp:2966-3028
 *   cpu_mv_serial / cpu_mv        csr_matrix.tpp:2665-2704
 *
 * The reference runs its partitions on P OpenMP threads; here they run one
 * after the other, colour by colour. That yields the same bits: inside one
 * colour every y entry is touched by exactly one partition, and colours are
 * separated by barriers ("Determinism").
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BLK_BITS 4 /* csr_matrix.hpp:89-90 */
#define BLK_FACTOR (1 << BLK_BITS)

typedef struct cfs_oracle {
  int nrows, P, is_double;
  long long nnz_full;
  int nnz_low, nnz_diag, ncolors, nranges;
  int *row_split;    /* P+1                                               */
  int *part_nnz_low; /* P                                                 */
  int *lower_rowptr; /* per partition local rowptr, concatenated: N+P      */
  int *lower_colind; /* nnz_low, GLOBAL column ids                        */
  void *lower_values;
  void *diagonal;    /* N                                                 */
  int nblk;          /* V = ceil(N/16)                                    */
  int *weight;       /* V                                                 */
  int *adj_ptr;      /* V+1: conflict graph, neighbours ascending         */
  int *adj;
  int *color_first;  /* V: colours straight after first-fit               */
  int *color;        /* V: colours after balancing                        */
  int *range_ptr;    /* P*(ncolors+1), per partition                      */
  int *part_nranges; /* P                                                 */
  int *range_start;  /* nranges, local row ids                            */
  int *range_end;    /* nranges, inclusive                                */
} cfs_oracle;

/* ---- csr_matrix.tpp:418-423 ------------------------------------------- */
void cfs_oracle_partition_by_nrows(int nrows, int P, int *row_split) {
  int per_split = ((nrows / P - 1) | (BLK_FACTOR - 1)) + 1;
  row_split[0] = 0;
  for (int t = 0; t < P - 1; ++t)
    row_split[t + 1] = row_split[t] + per_split;
  row_split[P] = nrows;
}

/* ---- csr_matrix.tpp:438-541, the branch of a NON-symmetric matrix
 * (:504-515) and the tail (:517-530). The reference's "fill the last split"
 * step may write row_split_[nthreads+1] (one int past its allocation,
 * :518-520); only entries 0..P exist here. */
void cfs_oracle_partition_by_nnz(int nrows, int P, const int *rowptr,
                                 int *row_split) {
  if (P == 1) {
    row_split[0] = 0;
    row_split[1] = nrows;
    return;
  }
  const int nnz_per_split = rowptr[nrows] / P;
  int curr_nnz = 0, split_cnt = 0;
  row_split[0] = 0;
  for (int i = 0; i < nrows; ++i) {
    curr_nnz += rowptr[i + 1] - rowptr[i];
    if (curr_nnz >= nnz_per_split && (i + 1) % BLK_FACTOR == 0) {
      ++split_cnt;
      if (split_cnt <= P)
        row_split[split_cnt] = i + 1;
      curr_nnz = 0;
    }
  }
  if (curr_nnz < nnz_per_split && split_cnt <= P) {
    ++split_cnt;
    if (split_cnt <= P)
      row_split[split_cnt] = nrows;
  }
  if (split_cnt > P)
    row_split[P] = nrows;
  for (int i = split_cnt + 1; i <= P; ++i)
    row_split[i] = nrows;
}

static int cmp_u64(const void *a, const void *b) {
  uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
  return x < y ? -1 : x > y;
}

typedef struct {
  uint64_t *v;
  size_t n, cap;
} edge_list;

static void edge_push(edge_list *e, int a, int b) {
  if (e->n + 2 > e->cap) {
    e->cap = e->cap ? e->cap * 2 : 1024;
    e->v = (uint64_t *)realloc(e->v, e->cap * sizeof(uint64_t));
  }
  e->v[e->n++] = ((uint64_t)(uint32_t)a << 32) | (uint32_t)b;
  e->v[e->n++] = ((uint64_t)(uint32_t)b << 32) | (uint32_t)a;
}

static void edge_compact(edge_list *e) {
  /* set semantics of the reference's unordered_set adjacency */
  if (!e->n)
    return;
  qsort(e->v, e->n, sizeof(uint64_t), cmp_u64);
  size_t m = 1;
  for (size_t i = 1; i < e->n; ++i)
    if (e->v[i] != e->v[m - 1])
      e->v[m++] = e->v[i];
  e->n = m;
}

static int part_of_row(const cfs_oracle *o, int row) {
  /* row_part_[row] (csr_matrix.tpp:427-432), by search instead of a table */
  int lo = 0, hi = o->P - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (o->row_split[mid] <= row)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

/* ---- csr_matrix.tpp:1216-1348 (and serial(), 642-706, when P == 1) ---- */
static void extract_lower(cfs_oracle *o, const int *rowptr, const int *colind,
                          const void *values) {
  const int N = o->nrows, P = o->P;
  const size_t vs = o->is_double ? 8 : 4;
  int nlow = 0;
  for (int i = 0; i < N; ++i)
    for (int j = rowptr[i]; j < rowptr[i + 1]; ++j)
      nlow += colind[j] < i;
  o->lower_rowptr = (int *)calloc((size_t)N + P, sizeof(int));
  o->lower_colind = (int *)malloc((size_t)(nlow ? nlow : 1) * sizeof(int));
  o->lower_values = malloc((size_t)(nlow ? nlow : 1) * vs);
  o->diagonal = calloc((size_t)(N ? N : 1), vs); /* zero-filled, see B1 */
  o->part_nnz_low = (int *)calloc(P, sizeof(int));
  int out = 0, ndiag = 0;
  for (int t = 0; t < P; ++t) {
    int *lrp = o->lower_rowptr + o->row_split[t] + t;
    int begin = out;
    lrp[0] = 0;
    for (int i = o->row_split[t]; i < o->row_split[t + 1]; ++i) {
      for (int j = rowptr[i]; j < rowptr[i + 1]; ++j) {
        int col = colind[j];
        if (col < i) {
          o->lower_colind[out] = col; /* global id kept */
          memcpy((char *)o->lower_values + vs * out,
                 (const char *)values + vs * j, vs);
          ++out;
        } else if (col == i) {
          memcpy((char *)o->diagonal + vs * i, (const char *)values + vs * j,
                 vs);
          ++ndiag;
        }
      }
      lrp[i - o->row_split[t] + 1] = out - begin;
    }
    o->part_nnz_low[t] = out - begin;
  }
  o->nnz_low = out;
  o->nnz_diag = ndiag;
}

/* ---- csr_matrix.tpp:1364-1366, 1427-1477 ------------------------------ */
static void build_conflict_graph(cfs_oracle *o, const int *rowptr,
                                 const int *colind) {
  const int N = o->nrows, P = o->P;
  const int V = (N + BLK_FACTOR - 1) / BLK_FACTOR;
  o->nblk = V;
  o->weight = (int *)calloc(V ? V : 1, sizeof(int));
  edge_list e = {0, 0, 0};
  int *part = (int *)malloc((size_t)(N ? N : 1) * sizeof(int));
  for (int i = 0; i < N; ++i)
    part[i] = part_of_row(o, i);
  int base = 0;
  for (int t = 0; t < P; ++t) {
    const int off = o->row_split[t];
    const int *lrp = o->lower_rowptr + off + t;
    for (int i = off; i < o->row_split[t + 1]; ++i) {
      const int blk_row = i >> BLK_BITS;
      o->weight[blk_row] += lrp[i - off + 1] - lrp[i - off];
      /* direct conflicts: a lower entry whose column belongs to an earlier
       * partition (col < row_offset, :1447) */
      for (int j = lrp[i - off]; j < lrp[i - off + 1]; ++j) {
        int col = o->lower_colind[base + j];
        if (col < off)
          edge_push(&e, blk_row, col >> BLK_BITS);
      }
      /* indirect conflicts: rows r1,r2 of different partitions that both own
       * a lower entry in column i. They are the upper-triangle part of row i
       * of the full matrix (:1453-1475) */
      int first_upper = rowptr[i + 1];
      for (int j = rowptr[i]; j < rowptr[i + 1]; ++j)
        if (colind[j] > i) {
          first_upper = j;
          break;
        }
      for (int j = first_upper + 1; j < rowptr[i + 1]; ++j)
        for (int k = first_upper; k < j; ++k)
          if (part[colind[k]] != part[colind[j]])
            edge_push(&e, colind[k] >> BLK_BITS, colind[j] >> BLK_BITS);
      if (e.n > ((size_t)1 << 24))
        edge_compact(&e);
    }
    base += o->part_nnz_low[t];
  }
  free(part);
  edge_compact(&e);
  o->adj_ptr = (int *)calloc((size_t)V + 1, sizeof(int));
  o->adj = (int *)malloc((e.n ? e.n : 1) * sizeof(int));
  for (size_t k = 0; k < e.n; ++k) {
    o->adj_ptr[(e.v[k] >> 32) + 1]++;
    o->adj[k] = (int)(uint32_t)e.v[k];
  }
  for (int v = 0; v < V; ++v)
    o->adj_ptr[v + 1] += o->adj_ptr[v];
  free(e.v);
}

/* ---- csr_matrix.tpp:2027-2077 ----------------------------------------- */
static void first_fit_coloring(cfs_oracle *o) {
  const int V = o->nblk;
  int *color = (int *)malloc((size_t)(V ? V : 1) * sizeof(int));
  int *mark = (int *)malloc((size_t)(V ? V : 1) * sizeof(int));
  for (int v = 0; v < V; ++v) {
    color[v] = V - 1; /* initial value of color_map (:1497) */
    mark[v] = INT32_MAX;
  }
  int max_color = 0;
  for (int i = 0; i < V; ++i) {
    for (int k = o->adj_ptr[i]; k < o->adj_ptr[i + 1]; ++k)
      mark[color[o->adj[k]]] = i;
    int j = 0;
    while (j < max_color && mark[j] == i)
      ++j;
    if (j == max_color)
      ++max_color;
    color[i] = j;
  }
  free(mark);
  o->ncolors = max_color;
  o->color = color;
  o->color_first = (int *)malloc((size_t)(V ? V : 1) * sizeof(int));
  memcpy(o->color_first, color, (size_t)V * sizeof(int));
}

/* ---- csr_matrix.tpp:2080-2213 (balance == true) ----------------------- */
static void balance_two_colors(cfs_oracle *o) {
  const int k_color = 2;
  const int nc = o->ncolors;
  int *color = o->color;
  int *fifo = (int *)malloc((size_t)(o->nblk ? o->nblk : 1) * sizeof(int));
  for (int t = 0; t < o->P; ++t) {
    const int off = o->row_split[t];
    const int nrows_t = o->row_split[t + 1] - off;
    const int nb = (nrows_t + BLK_FACTOR - 1) / BLK_FACTOR;
    const int b0 = off >> BLK_BITS;
    int total = 0;
    int *load = (int *)calloc(nc > k_color ? nc : k_color, sizeof(int));
    for (int i = 0; i < nb; ++i) {
      int v = b0 + i;
      if (color[v] < k_color)
        total += o->weight[v];
      load[color[v]] += o->weight[v];
    }
    const int mean = total / k_color;
    for (int step = 0; step < nc - 1; ++step) {
      /* colour with the largest positive deviation, first wins ties */
      int max_c = (load[1] - mean) > (load[0] - mean) ? 1 : 0;
      int target = (max_c + 1) % k_color;
      /* bin[max_c]: this partition's vertices of that colour, ascending */
      int nq = 0;
      for (int i = 0; i < nb; ++i)
        if (color[b0 + i] == max_c)
          fifo[nq++] = b0 + i;
      for (int q = 0; q < nq && load[max_c] - mean > 0; ++q) {
        int v = fifo[q];
        int used = 0;
        for (int k = o->adj_ptr[v]; k < o->adj_ptr[v + 1]; ++k)
          if (color[o->adj[k]] == target) {
            used = 1;
            break;
          }
        if (!used) {
          color[v] = target;
          load[max_c] -= o->weight[v];
          load[target] += o->weight[v];
        }
      }
    }
    free(load);
  }
  free(fifo);
}

/* ---- csr_matrix.tpp:1544-1627 ----------------------------------------- */
static void detect_ranges(cfs_oracle *o) {
  const int P = o->P, nc = o->ncolors;
  o->range_ptr = (int *)calloc((size_t)P * (nc + 1), sizeof(int));
  o->part_nranges = (int *)calloc((size_t)P, sizeof(int));
  size_t cap = 1024, n = 0;
  o->range_start = (int *)malloc(cap * sizeof(int));
  o->range_end = (int *)malloc(cap * sizeof(int));
  for (int t = 0; t < P; ++t) {
    const int off = o->row_split[t], end = o->row_split[t + 1];
    int *rp = o->range_ptr + (size_t)t * (nc + 1);
    size_t begin = n;
    for (int c = 0; c < nc; ++c) {
      int i = off;
      while (i < end) {
        if (o->color[i >> BLK_BITS] != c) {
          ++i;
          continue;
        }
        int s = i;
        while (i + 1 < end && o->color[(i + 1) >> BLK_BITS] == c)
          ++i;
        if (n + 1 > cap) {
          cap *= 2;
          o->range_start = (int *)realloc(o->range_start, cap * sizeof(int));
          o->range_end = (int *)realloc(o->range_end, cap * sizeof(int));
        }
        o->range_start[n] = s - off;
        o->range_end[n] = i - off;
        ++n;
        ++i;
      }
      rp[c + 1] = (int)(n - begin);
    }
    o->part_nranges[t] = (int)(n - begin);
  }
  o->nranges = (int)n;
}

/* ---- tune(): csr_matrix.tpp:231-310 + compress_symmetry() 1684-1716 ---- */
cfs_oracle *cfs_oracle_build(int nrows, const int *rowptr, const int *colind,
                             const void *values, int is_double, int P) {
  cfs_oracle *o = (cfs_oracle *)calloc(1, sizeof(cfs_oracle));
  o->nrows = nrows;
  o->P = P < 1 ? 1 : P;
  o->is_double = is_double;
  o->nnz_full = rowptr[nrows];
  o->row_split = (int *)calloc((size_t)o->P + 1, sizeof(int));
  cfs_oracle_partition_by_nrows(nrows, o->P, o->row_split);
  extract_lower(o, rowptr, colind, values);
  if (o->P > 1) {
    build_conflict_graph(o, rowptr, colind);
    first_fit_coloring(o);
    balance_two_colors(o);
    detect_ranges(o);
  }
  return o;
}

void cfs_oracle_free(cfs_oracle *o) {
  if (!o)
    return;
  free(o->row_split);
  free(o->part_nnz_low);
  free(o->lower_rowptr);
  free(o->lower_colind);
  free(o->lower_values);
  free(o->diagonal);
  free(o->weight);
  free(o->adj_ptr);
  free(o->adj);
  free(o->color_first);
  free(o->color);
  free(o->range_ptr);
  free(o->part_nranges);
  free(o->range_start);
  free(o->range_end);
  free(o);
}

/* ---- size(): csr_matrix.tpp:191-228, symmetric branch ------------------ */
long long cfs_oracle_size_bytes(const cfs_oracle *o) {
  const long long vs = o->is_double ? 8 : 4;
  long long s = ((long long)o->nrows + 1LL * o->P) * 4;
  s += (long long)o->nnz_low * 4 + (long long)o->nnz_low * vs;
  s += (long long)o->nnz_diag * vs;
  if (o->P > 1) {
    s += ((long long)o->ncolors + 1) * 4;
    s += 2LL * o->nranges * 4;
  }
  return s;
}

/* accessors for ctypes */
#define GETTER(type, name)                                                     \
  type cfs_oracle_##name(const cfs_oracle *o) { return o->name; }
GETTER(int, nrows)
GETTER(int, P)
GETTER(int, nnz_low)
GETTER(int, nnz_diag)
GETTER(int, ncolors)
GETTER(int, nranges)
GETTER(int, nblk)
GETTER(const int *, row_split)
GETTER(const int *, part_nnz_low)
GETTER(const int *, lower_rowptr)
GETTER(const int *, lower_colind)
GETTER(const void *, lower_values)
GETTER(const void *, diagonal)
GETTER(const int *, weight)
GETTER(const int *, adj_ptr)
GETTER(const int *, adj)
GETTER(const int *, color_first)
GETTER(const int *, color)
GETTER(const int *, range_ptr)
GETTER(const int *, part_nranges)
GETTER(const int *, range_start)
GETTER(const int *, range_end)

/* ---- kernels ----------------------------------------------------------- */
#define DEFINE_KERNELS(T, SUFFIX)                                              \
  /* cpu_mv_sym_serial (:2707-2729) for P == 1,                            */ \
  /* cpu_mv_sym_conflict_free_v2 (:2966-3028) otherwise.                   */ \
  static void sym_spmv_##SUFFIX(const cfs_oracle *o, T *y, const T *x) {       \
    const T *val = (const T *)o->lower_values;                                 \
    const T *diag = (const T *)o->diagonal;                                    \
    if (o->P == 1) {                                                           \
      const int *rp = o->lower_rowptr;                                         \
      for (int i = 0; i < o->nrows; ++i) {                                     \
        T y_tmp = diag[i] * x[i];                                              \
        for (int j = rp[i]; j < rp[i + 1]; ++j) {                              \
          int col = o->lower_colind[j];                                        \
          T v = val[j];                                                        \
          y_tmp += v * x[col];                                                 \
          y[col] += v * x[i];                                                  \
        }                                                                      \
        y[i] = y_tmp;                                                          \
      }                                                                        \
      return;                                                                  \
    }                                                                          \
    for (int i = 0; i < o->nrows; ++i)                                         \
      y[i] = diag[i] * x[i];                                                   \
    for (int c = 0; c < o->ncolors; ++c) {                                     \
      int base = 0, rbase = 0;                                                 \
      for (int t = 0; t < o->P; ++t) {                                         \
        const int off = o->row_split[t];                                       \
        const int *rp = o->lower_rowptr + off + t;                             \
        const int *rptr = o->range_ptr + (size_t)t * (o->ncolors + 1);         \
        for (int r = rptr[c]; r < rptr[c + 1]; ++r) {                          \
          for (int i = o->range_start[rbase + r];                              \
               i <= o->range_end[rbase + r]; ++i) {                            \
            T y_tmp = 0;                                                       \
            for (int j = rp[i]; j < rp[i + 1]; ++j) {                          \
              int col = o->lower_colind[base + j];                             \
              T v = val[base + j];                                             \
              y_tmp += v * x[col];                                             \
              y[col] += v * x[i + off];                                        \
            }                                                                  \
            y[i + off] += y_tmp;                                               \
          }                                                                    \
        }                                                                      \
        base += o->part_nnz_low[t];                                            \
        rbase += o->part_nranges[t];                                           \
      }                                                                        \
    }                                                                          \
  }                                                                            \
  /* cpu_mv_serial / cpu_mv (:2665-2704): plain CSR, the test's comparator */ \
  static void csr_spmv_##SUFFIX(int nrows, const int *rowptr,                  \
                                const int *colind, const T *values, T *y,      \
                                const T *x) {                                  \
    for (int i = 0; i < nrows; ++i) {                                          \
      T y_tmp = 0;                                                             \
      for (int j = rowptr[i]; j < rowptr[i + 1]; ++j)                          \
        y_tmp += values[j] * x[colind[j]];                                     \
      y[i] = y_tmp;                                                            \
    }                                                                          \
  }

DEFINE_KERNELS(double, f64)
DEFINE_KERNELS(float, f32)

void cfs_oracle_spmv(const cfs_oracle *o, void *y, const void *x) {
  if (o->is_double)
    sym_spmv_f64(o, (double *)y, (const double *)x);
  else
    sym_spmv_f32(o, (float *)y, (const float *)x);
}

void cfs_oracle_csr_spmv(int nrows, const int *rowptr, const int *colind,
                         const void *values, int is_double, void *y,
                         const void *x) {
  if (is_double)
    csr_spmv_f64(nrows, rowptr, colind, (const double *)values, (double *)y,
                 (const double *)x);
  else
    csr_spmv_f32(nrows, rowptr, colind, (const float *)values, (float *)y,
                 (const float *)x);
}
