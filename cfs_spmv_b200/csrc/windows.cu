// windows.cu -- preprocessing of kernel variant 3: per-tile x / y windows and
// slot ownership.
//
// For every tile (<= 8 slices of 32 rows) one CTA
//   1. collects the 32-column blocks the tile's entries touch (hash set in
//      shared memory), sorts them and turns runs of consecutive blocks into
//      windows (<= kMaxWindows windows, <= kWindowBlocks blocks in total,
//      lowest columns first; blocks that do not fit stay on the global path);
//   2. maps every entry to a window slot and gives every slot ONE owner warp
//      (slice) of the tile: the kernel lets only the owner update the slot, with
//      a plain shared-memory read-modify-write. Entries of other slices that
//      hit the slot, or that collide with another lane of their own warp in the
//      same step, are rewritten to "far" entries (global gather + global RED). This is the
//      reference's conflict-free idea (no two concurrent workers write the same
//      y entry, csr_matrix.tpp:1427-1477) applied inside a CTA, with ownership
//      instead of colouring because a warp's lanes and steps are ordered;
//   3. writes the rewritten index stream (sell_slot) and the window table into
//      the tile record.
//
// Index codes written to sell_slot:
//   slot >= 0                      owned window slot
//   -1                             padding
//   -(2 + col)                     "far": the slot belongs to another warp (or
//                                  lane), or the column is outside every
//                                  window: global gather + global RED
// The record also gets row_lo: the first row of the tile when its 32-row
// slices hold consecutive, unsplit rows (then x[row] and diagonal[row] are
// bulk-copied with the tile instead of gathered).
#include "common.cuh"

namespace cfsb {

namespace {

constexpr int kHash = 1024;              // slots of the block hash set
constexpr int kMaxUniq = 512;            // distinct blocks handled per tile
constexpr int kSlots = kWindowBlocks * 32;

__device__ __forceinline__ unsigned hash_block(int b) {
  return ((unsigned)b * 2654435761u) >> 22; // 10 bits
}

__global__ void __launch_bounds__(kTileSlices * 32)
    build_windows_kernel(int ntiles, int col_limit, int row_begin,
                         TileRec *__restrict__ tile_rec,
                         const int *__restrict__ vrow_row,
                         const int *__restrict__ sell_col,
                         int *__restrict__ sell_slot,
                         unsigned long long *__restrict__ far_count) {
  __shared__ int table[kHash];
  __shared__ int uniq[kMaxUniq];
  __shared__ int owner[kSlots];
  __shared__ int win_blk[kMaxWindows], win_n[kMaxWindows], win_off[kMaxWindows];
  __shared__ int nuniq, nwin, overflow, ragged;
  __shared__ unsigned int far_local;

  const int tile = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const TileRec rec = tile_rec[tile];
  const size_t e0 = (size_t)rec.step_begin * kSliceRows;
  const bool has_slice = warp < rec.nslices;
  const int k0 = has_slice ? rec.slice_step[warp] : 0;
  const int k1 = has_slice ? rec.slice_step[warp + 1] : 0;

  for (int i = tid; i < kHash; i += blockDim.x)
    table[i] = -1;
  for (int i = tid; i < kSlots; i += blockDim.x)
    owner[i] = -1;
  if (tid == 0) {
    nuniq = 0;
    nwin = 0;
    overflow = 0;
    far_local = 0;
    ragged = 0;
  }
  __syncthreads();

  // ---- 0. are the rows of the tile consecutive and unsplit?
  const int tag0 = vrow_row[(size_t)rec.slice_begin * kSliceRows];
  if (has_slice) {
    const int tag = vrow_row[((size_t)rec.slice_begin + warp) * kSliceRows + lane];
    // every lane live and holding the next row, no continuation chunks
    // (bit 30 of a tag breaks the equality)
    const bool ok_lane = tag == tag0 + warp * kSliceRows + lane;
    if (!__all_sync(0xffffffffu, ok_lane))
      ragged = 1;
  }
  // bulk copies need 16-byte aligned, 16-byte multiple row ranges
  if (tid == 0 && (tag0 < 0 || (tag0 & kVrowCont) || (tag0 & 3) ||
                   ((tag0 - row_begin) & 3)))
    ragged = 1;

  // ---- 1. distinct 32-column blocks of the tile
  for (int k = k0; k < k1; ++k) {
    const int c = sell_col[e0 + (size_t)k * kSliceRows + lane];
    if (c >= 0 && c < col_limit) {
      const int b = c >> 5;
      unsigned h = hash_block(b);
      for (int probe = 0; probe < kHash; ++probe) {
        const int old = atomicCAS(&table[h], -1, b);
        if (old == -1 || old == b)
          break;
        h = (h + 1) & (kHash - 1);
        if (probe == kHash - 1)
          overflow = 1;
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < kHash; i += blockDim.x) {
    if (table[i] >= 0) {
      const int at = atomicAdd(&nuniq, 1);
      if (at < kMaxUniq)
        uniq[at] = table[i];
      else
        overflow = 1;
    }
  }
  __syncthreads();
  const int nu = overflow ? 0 : nuniq; // too irregular: no windows at all
  for (int i = tid; i < kMaxUniq; i += blockDim.x)
    if (i >= nu)
      uniq[i] = INT_MAX;
  __syncthreads();
  // bitonic sort of uniq[0..kMaxUniq)
  for (int size = 2; size <= kMaxUniq; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < kMaxUniq / 2; i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const int a = uniq[lo], b = uniq[hi];
        if ((a > b) == up) {
          uniq[lo] = b;
          uniq[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  // runs of consecutive blocks -> windows, lowest columns first
  if (tid == 0) {
    int total = 0, i = 0, w = 0;
    while (i < nu && w < kMaxWindows && total < kWindowBlocks) {
      int j = i;
      while (j + 1 < nu && uniq[j + 1] == uniq[j] + 1)
        ++j;
      const int len = j - i + 1;
      const int take = min(len, kWindowBlocks - total);
      win_blk[w] = uniq[i];
      win_n[w] = take;
      win_off[w] = total * 32;
      total += take;
      ++w;
      i = j + 1;
    }
    nwin = w;
  }
  __syncthreads();

  // ---- 2. slots and ownership
  auto slot_of = [&](int c) -> int {
    if (c < 0 || c >= col_limit)
      return -1;
    const int b = c >> 5;
    for (int w = 0; w < nwin; ++w)
      if (b >= win_blk[w] && b < win_blk[w] + win_n[w])
        return win_off[w] + (c - win_blk[w] * 32);
    return -1;
  };
  for (int k = k0; k < k1; ++k) {
    const int s = slot_of(sell_col[e0 + (size_t)k * kSliceRows + lane]);
    if (s >= 0)
      atomicCAS(&owner[s], -1, warp);
  }
  __syncthreads();
  // ---- 3. rewrite the index stream
  unsigned my_far = 0;
  for (int k = k0; k < k1; ++k) {
    const size_t at = e0 + (size_t)k * kSliceRows + lane;
    const int c = sell_col[at];
    const int s = slot_of(c);
    // a slot belongs to the first slice that claimed it; the other slices
    // reach that y entry through a global RED
    bool mine = s >= 0 && owner[s] == warp;
    // two lanes of one warp on the same slot in the same step would lose an
    // update: keep the lowest lane, send the others to the RED path
    const unsigned peers = __match_any_sync(0xffffffffu, mine ? s : -1 - lane);
    if (mine && (peers & ((1u << lane) - 1)))
      mine = false;
    int code = -1;
    if (c >= 0) {
      code = mine ? s : -(2 + c);
      my_far += !mine;
    }
    sell_slot[at] = code;
  }
  if (my_far)
    atomicAdd(&far_local, my_far);
  __syncthreads();
  if (tid == 0) {
    TileRec out = rec;
    out.nwin = nwin;
    out.row_lo = ragged ? -1 : (tag0 & kVrowRowMask);
    for (int w = 0; w < kMaxWindows; ++w) {
      out.win_lo[w] = w < nwin ? win_blk[w] * 32 : 0;
      out.win_nblk[w] = (unsigned short)(w < nwin ? win_n[w] : 0);
    }
    tile_rec[tile] = out;
    if (far_local)
      atomicAdd(far_count, (unsigned long long)far_local);
  }
}

} // namespace

int build_windows(cfs_matrix_s *m, cudaStream_t s) {
  m->far_entries = 0;
  if (m->ntiles == 0)
    return CFS_OK;
  // windows consist of whole 32-column blocks inside the (extended) vectors;
  // halo_begin is a multiple of 32, so block starts are 16-byte aligned there
  const int vec_end = m->row_begin + m->nrows;
  const int col_limit = vec_end & ~31;
  CFS_TRY(m->sell_slot.alloc((size_t)m->padded_entries));
  DevArray<unsigned long long> far;
  CFS_TRY(far.alloc(1));
  CFS_CUDA_TRY(cudaMemsetAsync(far.p, 0, 8, s));
  build_windows_kernel<<<(unsigned)m->ntiles, kTileSlices * 32, 0, s>>>(
      (int)m->ntiles, col_limit, m->row_begin, m->tile_rec.p, m->vrow_row.p, m->sell_col.p,
      m->sell_slot.p,
      far.p);
  CFS_CUDA_TRY(cudaGetLastError());
  unsigned long long h = 0;
  CFS_CUDA_TRY(cudaMemcpyAsync(&h, far.p, 8, cudaMemcpyDeviceToHost, s));
  CFS_CUDA_TRY(cudaStreamSynchronize(s));
  m->far_entries = (int64_t)h;
  return CFS_OK;
}

} // namespace cfsb
