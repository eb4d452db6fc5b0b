/*
 * cfs_gen.h -- deterministic synthetic symmetric matrices and vectors.
 *
 * One definition shared by host C/C++, CUDA device code and (restated in
 * numpy, tests/gen.py) the Python tests, so that the GPU path, the CPU oracle
 * and the compiled reference are always fed the SAME matrix and the SAME x.
 * The matrices are the BASELINE.json configs: 3-D 7-/27-point Laplacians and a
 * banded FEM-like SPD matrix. Every row carries a stored diagonal (the
 * reference reads uninitialised memory otherwise).
 *
 * Rows are produced independently (row -> sorted column list of the FULL
 * matrix), which is what lets a GPU build its own row shard in place.
 */
#ifndef CFS_GEN_H
#define CFS_GEN_H

#include <stdint.h>

#if defined(__CUDACC__)
#define CFS_GEN_FN __host__ __device__ static inline
#else
#define CFS_GEN_FN static inline
#endif

enum { CFS_GEN_LAP7 = 1, CFS_GEN_LAP27 = 2, CFS_GEN_BANDED = 3 };

typedef struct cfs_gen_spec {
  int32_t kind;       /* CFS_GEN_*                                           */
  int32_t nx, ny, nz; /* Laplacian grid (row = x + nx*(y + ny*z))            */
  int64_t nrows;      /* number of rows (= nx*ny*nz for the Laplacians)      */
  int32_t bw;         /* BANDED: half bandwidth                              */
  int32_t per_row;    /* BANDED: expected lower entries per row, times 16    */
  uint64_t seed;      /* BANDED: pattern/value seed. Laplacians: 0 = the
                         constant-coefficient stencil (off-diagonals -1); else
                         every edge {i,j} gets its own coefficient in
                         [-1, -0.5) (a heterogeneous diffusion problem: ~nnz/2
                         distinct values) and the diagonal is 1/16 + sum |a|  */
} cfs_gen_spec;

/* splitmix64 finaliser: the only random source used anywhere */
CFS_GEN_FN uint64_t cfs_mix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}

CFS_GEN_FN uint64_t cfs_hash3(uint64_t seed, uint64_t a, uint64_t b) {
  return cfs_mix64(cfs_mix64(cfs_mix64(seed) ^ a) ^ b);
}

/* uniform double in [0,1) from the top 53 bits */
CFS_GEN_FN double cfs_u01(uint64_t h) {
  return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

/* x vector of the reference's bench: U(0.01, 0.42)
 * (bench/bench_spmv_mmf.cpp:125), but reproducible */
CFS_GEN_FN double cfs_gen_x(uint64_t seed, int64_t i) {
  const double u = cfs_u01(cfs_hash3(seed, (uint64_t)i, 0x78ULL));
#if defined(__CUDA_ARCH__)
  /* no FMA contraction: bit-identical to the host and numpy versions */
  return __dadd_rn(0.01, __dmul_rn(0.41, u));
#else
  return 0.01 + 0.41 * u;
#endif
}

/* BANDED: is (row, row-d) a stored lower entry?  Bernoulli(per_row/16/bw). */
CFS_GEN_FN int cfs_banded_has(const cfs_gen_spec *g, int64_t row, int32_t d) {
  uint64_t h = cfs_hash3(g->seed, (uint64_t)row, (uint64_t)d);
  return (h % ((uint64_t)g->bw * 16u)) < (uint64_t)g->per_row;
}

/* BANDED: value of lower entry (row, row-d): in [-1, -1/1024] */
CFS_GEN_FN double cfs_banded_val(const cfs_gen_spec *g, int64_t row,
                                 int32_t d) {
  uint64_t h = cfs_hash3(g->seed ^ 0x5eedULL, (uint64_t)row, (uint64_t)d);
  return -(double)(1 + (h & 1023u)) * (1.0 / 1024.0);
}

/*
 * Emit one row of the FULL matrix, columns ascending, diagonal included.
 * cols/vals may be NULL (count only). Returns the entry count.
 */
CFS_GEN_FN int cfs_gen_row(const cfs_gen_spec *g, int64_t row, int32_t *cols,
                           double *vals) {
  int n = 0;
  if (g->kind == CFS_GEN_LAP7 || g->kind == CFS_GEN_LAP27) {
    const int64_t nx = g->nx, ny = g->ny, nz = g->nz;
    const int x = (int)(row % nx);
    const int y = (int)((row / nx) % ny);
    const int z = (int)(row / (nx * ny));
    const int full = (g->kind == CFS_GEN_LAP27);
    double absum = 0.0;
    int diag_at = -1;
    for (int dz = -1; dz <= 1; ++dz) {
      if (z + dz < 0 || z + dz >= nz)
        continue;
      for (int dy = -1; dy <= 1; ++dy) {
        if (y + dy < 0 || y + dy >= ny)
          continue;
        for (int dx = -1; dx <= 1; ++dx) {
          if (x + dx < 0 || x + dx >= nx)
            continue;
          const int manhattan = (dx != 0) + (dy != 0) + (dz != 0);
          if (!full && manhattan > 1)
            continue;
          if (cols) {
            const int64_t col = row + dx + nx * (dy + ny * (int64_t)dz);
            cols[n] = (int32_t)col;
            if (g->seed == 0) {
              vals[n] = manhattan == 0 ? (full ? 26.0 : 6.0) : -1.0;
            } else if (manhattan == 0) {
              diag_at = n;
            } else {
              const uint64_t lo = (uint64_t)(col < row ? col : row);
              const uint64_t hi = (uint64_t)(col < row ? row : col);
              const double a = 0.5 + 0.5 * cfs_u01(cfs_hash3(g->seed, lo, hi));
              vals[n] = -a;
              absum += a; /* ascending columns: one order everywhere */
            }
          }
          ++n;
        }
      }
    }
    if (cols && diag_at >= 0)
      vals[diag_at] = 0.0625 + absum;
    return n;
  }
  /* BANDED */
  {
    const int64_t N = g->nrows;
    double absum = 0.0;
    int diag_at;
    int32_t dmax = (int32_t)(row < g->bw ? row : g->bw);
    for (int32_t d = dmax; d >= 1; --d) {
      if (cfs_banded_has(g, row, d)) {
        const double v = cfs_banded_val(g, row, d);
        if (cols) {
          cols[n] = (int32_t)(row - d);
          vals[n] = v;
        }
        absum += -v;
        ++n;
      }
    }
    diag_at = n++;
    dmax = (int32_t)((N - 1 - row) < g->bw ? (N - 1 - row) : g->bw);
    for (int32_t d = 1; d <= dmax; ++d) {
      if (cfs_banded_has(g, row + d, d)) {
        const double v = cfs_banded_val(g, row + d, d);
        if (cols) {
          cols[n] = (int32_t)(row + d);
          vals[n] = v;
        }
        absum += -v;
        ++n;
      }
    }
    if (cols) {
      cols[diag_at] = (int32_t)row;
      vals[diag_at] = 1.0 + absum;
    }
    return n;
  }
}

/* spec constructors */
CFS_GEN_FN cfs_gen_spec cfs_gen_laplacian_seeded(int points, int nx, int ny,
                                                 int nz, uint64_t seed);
CFS_GEN_FN cfs_gen_spec cfs_gen_laplacian(int points, int nx, int ny, int nz) {
  return cfs_gen_laplacian_seeded(points, nx, ny, nz, 0);
}
CFS_GEN_FN cfs_gen_spec cfs_gen_laplacian_seeded(int points, int nx, int ny,
                                                 int nz, uint64_t seed) {
  cfs_gen_spec g;
  g.kind = points == 7 ? CFS_GEN_LAP7 : CFS_GEN_LAP27;
  g.nx = nx;
  g.ny = ny;
  g.nz = nz;
  g.nrows = (int64_t)nx * ny * nz;
  g.bw = 0;
  g.per_row = 0;
  g.seed = seed;
  return g;
}

CFS_GEN_FN cfs_gen_spec cfs_gen_banded(int64_t nrows, int bw, int per_row_x16,
                                       uint64_t seed) {
  cfs_gen_spec g;
  g.kind = CFS_GEN_BANDED;
  g.nx = g.ny = g.nz = 0;
  g.nrows = nrows;
  g.bw = bw;
  g.per_row = per_row_x16;
  g.seed = seed;
  return g;
}

#endif /* CFS_GEN_H */
