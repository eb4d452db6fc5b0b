"""Synthetic matrices that need a global sort (R-MAT); the row-local generators
(Laplacians, banded) live in csrc/cfs_gen.h and are reached through capi."""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def mix64(z):
    """splitmix64 finaliser, vectorised; identical to cfs_mix64 in cfs_gen.h"""
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def hash3(seed, a, b):
    s = mix64(np.uint64(seed))
    return mix64(mix64(s ^ np.asarray(a, np.uint64)) ^ np.asarray(b, np.uint64))


def u01(h):
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def gen_x(seed, n, dtype=np.float64, begin=0):
    """numpy restatement of cfs_gen_x (tests check it against the C version)"""
    i = np.arange(begin, begin + n, dtype=np.uint64)
    return (0.01 + 0.41 * u01(hash3(seed, i, 0x78))).astype(dtype)


def rmat(scale, edge_factor=8, seed=1, dtype=np.float64,
         abc=(0.57, 0.19, 0.19)):
    """Graph500-style R-MAT, symmetrised, self loops dropped, deduplicated,
    full diagonal added. Returns FULL CSR."""
    n = 1 << scale
    m = edge_factor * n
    a, b, c = abc
    e = np.arange(m, dtype=np.uint64)
    u = np.zeros(m, dtype=np.int64)
    v = np.zeros(m, dtype=np.int64)
    for level in range(scale):
        r = u01(hash3(seed, e, level))
        # quadrants: [0,a) -> (0,0); [a,a+b) -> (0,1); [a+b,a+b+c) -> (1,0)
        ubit = (r >= a + b).astype(np.int64)
        vbit = (((r >= a) & (r < a + b)) | (r >= a + b + c)).astype(np.int64)
        u |= ubit << level
        v |= vbit << level
    hi = np.maximum(u, v)
    lo = np.minimum(u, v)
    keep = hi != lo
    key = np.unique(hi[keep] * n + lo[keep])
    hi = key // n
    lo = key % n
    val = (2.0 * u01(hash3(seed ^ 0xA5, hi.astype(np.uint64),
                           lo.astype(np.uint64))) - 1.0)
    rows = np.concatenate([hi, lo, np.arange(n)])
    cols = np.concatenate([lo, hi, np.arange(n)])
    vals = np.concatenate([val, val, np.full(n, 4.0 * edge_factor)])
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.int32)
    return rowptr, cols.astype(np.int32), vals.astype(dtype)


def random_symmetric(n, avg_lower, seed, dtype=np.float64, empty_rows=True):
    """small irregular symmetric matrix with ragged / empty rows, for edge-case
    parity tests. Full diagonal."""
    rng = np.random.default_rng(seed)
    nl = rng.poisson(avg_lower, size=n)
    if empty_rows:
        nl[rng.random(n) < 0.2] = 0
    his, los = [], []
    for i in range(1, n):
        k = min(int(nl[i]), i)
        if k:
            c = rng.choice(i, size=k, replace=False)
            his.append(np.full(k, i))
            los.append(c)
    hi = np.concatenate(his) if his else np.zeros(0, np.int64)
    lo = np.concatenate(los) if los else np.zeros(0, np.int64)
    val = rng.uniform(-1, 1, size=len(hi))
    rows = np.concatenate([hi, lo, np.arange(n)])
    cols = np.concatenate([lo, hi, np.arange(n)])
    vals = np.concatenate([val, val, rng.uniform(1, 2, size=n) + 4])
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    return (np.cumsum(rowptr).astype(np.int32), cols.astype(np.int32),
            vals.astype(dtype))


def write_mtx(path, rowptr, colind, values, symmetric=True, trailing_newline=True):
    """Matrix Market coordinate file, lower triangle only when symmetric
    (diagonal first per row), 1-based, last line newline-terminated
    (SURVEY.md B11)."""
    n = len(rowptr) - 1
    lines = []
    for i in range(n):
        for j in range(rowptr[i], rowptr[i + 1]):
            c = colind[j]
            if symmetric and c > i:
                continue
            lines.append("%d %d %.17g" % (i + 1, c + 1, values[j]))
    with open(path, "w") as f:
        f.write("%%%%MatrixMarket matrix coordinate real %s\n"
                % ("symmetric" if symmetric else "general"))
        f.write("%d %d %d\n" % (n, n, len(lines)))
        f.write("\n".join(lines))
        if trailing_newline:
            f.write("\n")


# ---------------------------------------------------------------------------
# device-side R-MAT (torch is only the array library here): bit-identical to
# rmat() above, but fast enough for BASELINE config 3 (scale 24)
# ---------------------------------------------------------------------------
def _t_mix64(z):
    import torch

    def lsr(v, s):  # logical shift right on int64
        return (v >> s) & ((1 << (64 - s)) - 1)
    c1 = -7046029254386353131   # 0x9E3779B97F4A7C15 as int64
    c2 = -4658895280553007687   # 0xBF58476D1CE4E5B9
    c3 = -7723592293110705685   # 0x94D049BB133111EB
    z = z + c1
    z = (z ^ lsr(z, 30)) * c2
    z = (z ^ lsr(z, 27)) * c3
    return z ^ lsr(z, 31)


def _t_hash3(seed, a, b):
    import torch
    s = _t_mix64(torch.tensor(seed, dtype=torch.int64, device=a.device))
    return _t_mix64(_t_mix64(s ^ a) ^ b)


def _t_u01(h):
    import torch
    top = (h >> 11) & ((1 << 53) - 1)
    return top.to(torch.float64) * (1.0 / 9007199254740992.0)


def rmat_torch(scale, edge_factor=8, seed=1, is_double=True,
               abc=(0.57, 0.19, 0.19), device="cuda", chunk=1 << 24):
    """same matrix as rmat(), built on `device`; returns torch tensors
    (rowptr int32, colind int32, values)"""
    import torch
    n = 1 << scale
    m = edge_factor * n
    a, b, c = abc
    keys = []
    for e0 in range(0, m, chunk):
        e = torch.arange(e0, min(m, e0 + chunk), dtype=torch.int64, device=device)
        u = torch.zeros_like(e)
        v = torch.zeros_like(e)
        for level in range(scale):
            r = _t_u01(_t_hash3(seed, e, torch.tensor(level, dtype=torch.int64,
                                                      device=device)))
            ubit = (r >= a + b).to(torch.int64)
            vbit = (((r >= a) & (r < a + b)) | (r >= a + b + c)).to(torch.int64)
            u |= ubit << level
            v |= vbit << level
        hi = torch.maximum(u, v)
        lo = torch.minimum(u, v)
        keep = hi != lo
        keys.append(torch.unique(hi[keep] * n + lo[keep]))
    key = torch.unique(torch.cat(keys))
    del keys
    hi = key // n
    lo = key % n
    val = 2.0 * _t_u01(_t_hash3(seed ^ 0xA5, hi, lo)) - 1.0
    ar = torch.arange(n, dtype=torch.int64, device=device)
    rows = torch.cat([hi, lo, ar])
    cols = torch.cat([lo, hi, ar])
    vals = torch.cat([val, val, torch.full((n,), 4.0 * edge_factor,
                                           dtype=torch.float64, device=device)])
    del hi, lo, val, key
    order = torch.argsort(rows * n + cols)
    rows, cols, vals = rows[order], cols[order], vals[order]
    counts = torch.bincount(rows, minlength=n)
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    rowptr[1:] = torch.cumsum(counts, 0)
    return (rowptr.to(torch.int32), cols.to(torch.int32),
            vals if is_double else vals.to(torch.float32))
