// K2: CSR sparsity pattern on the device.  Replaces DOLFIN's SparsityPatternBuilder + tensor init that run
// inside dolfin.assemble (/root/reference/sucemfem/SystemMatrices.py:54).  Contract: row i
// holds the sorted-ascending union of the cell-dof sets of all cells containing dof i, zeros included.
//
// Because dof = q*entity + k, the dof-level pattern is the q x q block expansion of an
// ENTITY-level pattern; only that one is built and stored (nnz/q^2 column ids):
//   one warp per row entity: gather the entities of its incident cells, rank-sort + unique in shared
//   memory, count -> scan (int64) -> fill; the same pass records pos[incidence][u] = slot of the cell's
//   u-th entity in the row (the "precomputed CSR value slots" the value kernels scatter into).
// For tetrahedra it also writes the packed 32-byte item records the canonical-frame kernels stream (faces here,
// edges -- in ring order -- in k_edge_ring).
#include <cub/cub.cuh>

#include <cstdlib>

#include "sfem_internal.h"

#define SFEM_TABLE static __device__ const
#include "generated/canonical_lut.inc"

namespace sfem {
namespace {

// SFEM_EDGE_KERNEL=1 selects the lane-pair edge kernel of round 1 (A/B measurements); default: lane per item
int edge_kernel_version() {
  const char* e = getenv("SFEM_EDGE_KERNEL");
  return (e && e[0] == '1') ? 1 : 2;
}

constexpr int kWarpsPerBlock = 4;
constexpr int kCap = 1024;      // candidates per row entity (valence * entities per cell) in the normal launch
constexpr int kCapBig = 24576;  // ... in the one-warp-per-block launch a mesh with a longer row falls back to (192 KB)

// WPB warps per block, each with two int32 arrays of `cap` candidates in dynamic shared memory
template <bool FILL, int WPB>
__global__ void __launch_bounds__(WPB * 32)
k_pattern(SpaceView sv, const int32_t* __restrict__ cells, const int64_t* __restrict__ inc_ptr,
          const int32_t* __restrict__ inc_cell,
          const uint8_t* __restrict__ inc_loc, bool canon, int64_t ent_lo, int64_t n_rows, int64_t inc_base,
          int32_t* __restrict__ row_len,
          const int64_t* __restrict__ ent_rowptr, int32_t* __restrict__ ent_cols, uint16_t* __restrict__ pos,
          uint8_t* __restrict__ rec, int* __restrict__ overflow, int cap) {
  extern __shared__ int32_t s_pat[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = blockIdx.x * (int64_t)WPB + w;
  if (r >= n_rows) return;
  const int64_t g = ent_lo + r;
  const int64_t i0 = inc_ptr[g], i1 = inc_ptr[g + 1];
  const int nec = sv.n_ent_cell;
  const int64_t L64 = (i1 - i0) * nec;
  if (L64 > cap) {
    if (lane == 0) *overflow = 1;
    return;
  }
  const int L = (int)L64;
  int32_t* cand = s_pat + (size_t)w * 2 * cap;
  int32_t* sorted = cand + cap;
  for (int idx = lane; idx < L; idx += 32) {
    int i = idx / nec, u = idx - i * nec;
    cand[idx] = sv.ent(inc_cell[i0 + i], u);
  }
  __syncwarp();
  // rank sort (stable): sorted[rank] = x
  for (int idx = lane; idx < L; idx += 32) {
    int32_t x = cand[idx];
    int rk = 0;
    for (int j = 0; j < L; ++j) {
      int32_t y = cand[j];
      rk += (y < x) || (y == x && j < idx);
    }
    sorted[rk] = x;
  }
  __syncwarp();
  // unique-compact in place into cand[0..nu)
  int nu = 0;
  for (int base = 0; base < L; base += 32) {
    int idx = base + lane;
    bool keep = idx < L && (idx == 0 || sorted[idx] != sorted[idx - 1]);
    unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) cand[nu + __popc(m & ((1u << lane) - 1))] = sorted[idx];
    nu += __popc(m);
  }
  __syncwarp();
  if (!FILL) {
    if (lane == 0) row_len[r] = nu;
    return;
  }
  const int64_t rp = ent_rowptr[r];
  for (int k = lane; k < nu; k += 32) ent_cols[rp + k] = cand[k];
  // packed 32-byte item records (order-2 tets): bytes 0-15 the cell's vertex ids in RELABELLED order, 16-19 cell,
  // 20 descriptor index into kRingDesc, 21 flags, 22-31 slots of the 10 canonical column entities.
  // Face rows are written here (descriptor 12 + local face); edge rows by k_edge_ring in ring order.
  const bool rec_here = rec && g >= sv.E;
  if (rec_here)
    for (int i = lane; i < (int)(i1 - i0); i += 32) {
      uint8_t* q = rec + (i0 - inc_base + i) * 32;
      const int64_t c = inc_cell[i0 + i];
      const int di = inc_loc[i0 + i] + 6;
      const unsigned long long desc = kRingDesc[di];
      for (int k = 0; k < 4; ++k)
        reinterpret_cast<int32_t*>(q)[k] = cells[c * 4 + (int)((desc >> (2 * k)) & 3ull)];
      reinterpret_cast<int32_t*>(q)[4] = (int32_t)c;
      q[20] = (uint8_t)di;
      q[21] = 0;
    }
  // slot of every (incidence, column entity) in the unique list.  Column entities are stored in the order the
  // value kernels visit them: for order-2 tets the CANONICAL order of the cell relabelled around the row entity
  // (u = umap[t][u'], generated/canonical_lut.inc), otherwise the cell-local order.
  for (int idx = lane; idx < L; idx += 32) {
    int i = idx / nec, u = idx - i * nec;
    const int u_canon = u;
    if (canon) u = (int)((kCanonDesc[inc_loc[i0 + i]] >> (8 + 4 * u)) & 15ull);
    int32_t x = sv.ent(inc_cell[i0 + i], u);
    int lo = 0, hi = nu - 1;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (cand[mid] < x) lo = mid + 1; else hi = mid;
    }
    pos[(i0 - inc_base) * nec + idx] = (uint16_t)lo;
    if (rec_here) rec[(i0 - inc_base + i) * 32 + 22 + u_canon] = (uint8_t)lo;
  }
}


// Edge rows of order-2 tets: the cells around an edge are re-ordered into RING order (consecutive cells share a
// face containing the edge), so that the value kernel can carry the contributions two neighbouring cells make to
// their three common column entities in registers and store every CSR slot exactly once.  One thread per edge.
// Output: packed 32-byte records in ring order.  Descriptor index = 2*t + swap (see canonical_lut.inc: the
// cell is relabelled (a, b, entry, exit)), flags: 1 has_prev (shares its entry face with the previous
// record), 2 has_next, 4 wrap (last cell of a closed ring: its exit face is the first cell's entry face).
constexpr int kMaxRing = 64;
__global__ void k_edge_ring(SpaceView sv, const int32_t* __restrict__ cells, const int64_t* __restrict__ inc_ptr,
                            const int32_t* __restrict__ inc_cell, const uint8_t* __restrict__ inc_loc, int64_t ent_lo,
                            int64_t n_edge_rows, int64_t inc_base, const int64_t* __restrict__ ent_rowptr,
                            const int32_t* __restrict__ ent_cols, uint8_t* __restrict__ rec, int* __restrict__ bad,
                            const int64_t* __restrict__ e2_slot0, const uint32_t* __restrict__ e2_tileoff,
                            const uint2* __restrict__ e2_hdr, uint4* __restrict__ aux_e2) {
  // e2_slot0 != nullptr: batched layout of the lane-per-item kernel -- the records of edge row r start at item slot
  // e2_slot0[r] of `rec` (batch = slot / 32), and every record gets an aux entry
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_edge_rows) return;
  if (e2_slot0 && *reinterpret_cast<volatile int*>(bad) != 0) return;  // the batch builder gave up: slots are invalid
  if (e2_slot0 && e2_slot0[r] < 0) return;                              // this edge's row goes to the generic kernel
  const int64_t g = ent_lo + r;
  const int64_t i0 = inc_ptr[g];
  const int m = (int)(inc_ptr[g + 1] - i0);
  if (m > kMaxRing) {
    atomicMax(bad, 1);
    return;
  }
  int32_t X[kMaxRing], Y[kMaxRing];
  for (int j = 0; j < m; ++j) {
    const int64_t c = inc_cell[i0 + j];
    const int t = inc_loc[i0 + j];
    const int e0 = tet_edge_vertex(t, 0), e1 = tet_edge_vertex(t, 1);
    int o0 = -1, o1 = -1;
    for (int v = 0; v < 4; ++v)
      if (v != e0 && v != e1) {
        if (o0 < 0) o0 = v; else o1 = v;
      }
    X[j] = cells[c * 4 + o0];
    Y[j] = cells[c * 4 + o1];
  }
  auto degree = [&](int32_t v) {
    int d = 0;
    for (int j = 0; j < m; ++j) d += (X[j] == v) + (Y[j] == v);
    return d;
  };
  for (int j = 0; j < m; ++j)
    if (degree(X[j]) > 2 || degree(Y[j]) > 2) {
      atomicMax(bad, 2);  // a face shared by more than two cells
      return;
    }
  const int64_t a = ent_rowptr[r];
  const int len = (int)(ent_rowptr[r + 1] - a);
  const int64_t slot0 = e2_slot0 ? e2_slot0[r] : i0 - inc_base;
  uint2 hdr = make_uint2(0u, 0u);
  unsigned tileoff = 0u;
  if (e2_slot0) { hdr = e2_hdr[slot0 >> 5]; tileoff = e2_tileoff[r]; }
  unsigned long long visited = 0ull;
  int pos = 0;
  while (pos < m) {
    // start of a chain: the lowest unvisited cell that has an open side (fan end); none -> closed ring
    int start = -1;
    int32_t entry = 0;
    for (int j = 0; j < m && start < 0; ++j) {
      if ((visited >> j) & 1ull) continue;
      if (degree(X[j]) == 1) { start = j; entry = X[j]; }
      else if (degree(Y[j]) == 1) { start = j; entry = Y[j]; }
    }
    const bool closed = start < 0;
    if (closed)
      for (int j = 0; j < m; ++j)
        if (!((visited >> j) & 1ull)) { start = j; entry = X[j]; break; }
    const int first_pos = pos;
    int cur = start;
    while (true) {
      const int32_t exitv = entry == X[cur] ? Y[cur] : X[cur];
      const int swap = entry == Y[cur];
      visited |= 1ull << cur;
      int nxt = -1;
      for (int j = 0; j < m; ++j)
        if (!((visited >> j) & 1ull) && (X[j] == exitv || Y[j] == exitv)) { nxt = j; break; }
      unsigned flags = (pos > first_pos ? 1u : 0u) | (nxt >= 0 ? 2u : 0u) | (nxt < 0 && closed && pos > first_pos ? 4u : 0u);
      // record
      const int64_t c = inc_cell[i0 + cur];
      const int t = inc_loc[i0 + cur];
      const unsigned long long desc = kRingDesc[2 * t + swap];
      uint8_t* q = rec + (slot0 + pos) * 32;
      if (aux_e2)  // packed: tile offset (9 bits), row length (8), position in the edge (5), items after it (5),
                   // lane distance to the last cell of a closed ring (5, first cell only; patched below)
        aux_e2[slot0 + pos] = make_uint4((unsigned)r, tileoff | ((unsigned)len << 9) | ((unsigned)pos << 17) |
                                                          ((unsigned)(m - 1 - pos) << 22), hdr.x, hdr.y);
      for (int k = 0; k < 4; ++k)
        reinterpret_cast<int32_t*>(q)[k] = cells[c * 4 + (int)((desc >> (2 * k)) & 3ull)];
      reinterpret_cast<int32_t*>(q)[4] = (int32_t)c;
      q[20] = (uint8_t)(2 * t + swap);
      for (int u = 0; u < sv.n_ent_cell; ++u) {  // 10 at order 2, the 6 edges at order 1
        const int32_t x = sv.ent(c, (int)((desc >> (8 + 4 * u)) & 15ull));
        int lo = 0, hi = len - 1;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (ent_cols[a + mid] < x) lo = mid + 1; else hi = mid;
        }
        q[22 + u] = (uint8_t)lo;
      }
      q[21] = (uint8_t)flags;
      ++pos;
      if (nxt < 0) break;
      entry = exitv;
      cur = nxt;
    }
    if (aux_e2 && closed && pos - first_pos > 1) aux_e2[slot0 + first_pos].y |= (unsigned)(pos - first_pos - 1) << 27;
  }
}

// Batches of the lane-per-item edge kernel: consecutive whole edges with at most 32 (edge, cell) items and at most
// kE2TileChunks chunks of output.  Greedy packing is a serial recurrence, so the edge rows are cut into independent
// segments of kE2Seg edges (a batch never crosses a segment boundary); one thread walks a segment.  Pass 1 counts the
// batches of each segment; pass 2 (after the scan) writes, per edge, its first item slot and its tile offset and, per
// batch, {first entity chunk of the span, span in 16-byte chunks}.
constexpr int kE2Seg = 512;
template <bool WRITE>
__global__ void k_e2_batches(const int64_t* __restrict__ inc_ptr, const int64_t* __restrict__ ent_rowptr, int64_t ent_lo,
                             int64_t n_edge_rows, int vmax, int64_t nb_max, int32_t* __restrict__ seg_nb,
                             const int64_t* __restrict__ seg_b0,
                             int64_t* __restrict__ slot0, uint32_t* __restrict__ tileoff, uint2* __restrict__ hdr,
                             int* __restrict__ bad, int64_t* __restrict__ gen_rows, int gen_cap) {
  const int64_t seg = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t r0 = seg * kE2Seg;
  if (r0 >= n_edge_rows) return;
  const int64_t r1 = min(r0 + (int64_t)kE2Seg, n_edge_rows);
  int items = 0, chunks = 0;
  int64_t nb = 0, b = WRITE ? seg_b0[seg] - 1 : -1;
  int64_t a_first = 0;
  int64_t ip = inc_ptr[ent_lo + r0], rp = ent_rowptr[r0];
  for (int64_t r = r0; r < r1; ++r) {
    const int64_t ip1 = inc_ptr[ent_lo + r + 1], rp1 = ent_rowptr[r + 1];
    const int m = (int)(ip1 - ip), c2 = 2 * (int)(rp1 - rp);
    if (m > vmax || m < 1 || c2 > kE2TileChunks) {
      // an edge the lane-per-item kernel does not take (more cells than lanes, or a row longer than the tile): it gets no
      // batch slots; the batch before it is closed, the row goes to the generic kernel (per-entity fallback, bad[1] counts)
      if (WRITE) {
        if (items > 0) hdr[b] = make_uint2((unsigned)a_first, (unsigned)chunks);
        slot0[r] = -1;
        const int k = atomicAdd(bad + 1, 1);
        if (k < gen_cap) gen_rows[k] = r; else atomicMax(bad, 1);
      }
      items = 0; chunks = 0;
      ip = ip1; rp = rp1;
      continue;
    }
    if (items == 0 || items + m > 32 || chunks + c2 > kE2TileChunks) {
      if (WRITE && items > 0) hdr[b] = make_uint2((unsigned)a_first, (unsigned)chunks);
      ++nb; ++b;
      if (WRITE && b >= nb_max) {  // more batches than the host-side bound allowed for: the caller falls back
        atomicMax(bad, 1);
        return;
      }
      items = 0; chunks = 0;
      a_first = rp;
    }
    if (WRITE) {
      slot0[r] = b * 32 + items;
      tileoff[r] = (uint32_t)chunks;
    }
    items += m;
    chunks += c2;
    ip = ip1; rp = rp1;
  }
  if (WRITE) {
    if (items > 0) hdr[b] = make_uint2((unsigned)a_first, (unsigned)chunks);
  } else {
    seg_nb[seg] = (int32_t)nb;
  }
}

__global__ void k_len_to_i64(const int32_t* __restrict__ len, int64_t* __restrict__ out, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i <= n) out[i] = i < n ? len[i] : 0;
}

// dof-level CSR from the entity-level one: dof row q*g+k; columns q*h+k'
__global__ void k_export(const int64_t* __restrict__ ent_rowptr, const int32_t* __restrict__ ent_cols, int q,
                         int64_t n_rows_ent, int64_t nnz_base, int64_t* __restrict__ rowptr,
                         int32_t* __restrict__ cols) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + w;
  if (r > n_rows_ent) return;
  if (r == n_rows_ent) {
    if (lane == 0) rowptr[q * r] = nnz_base + (int64_t)q * q * ent_rowptr[r];
    return;
  }
  const int64_t a = ent_rowptr[r];
  const int len = (int)(ent_rowptr[r + 1] - a);
  for (int k = 0; k < q; ++k) {
    const int64_t rs = (int64_t)q * q * a + (int64_t)k * q * len;
    if (lane == 0) rowptr[q * r + k] = nnz_base + rs;
    if (cols)
      for (int idx = lane; idx < q * len; idx += 32) {
        int s = idx / q, kp = idx - s * q;
        cols[rs + idx] = q * ent_cols[a + s] + kp;
      }
  }
}

inline unsigned grid_for(int64_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

__global__ void k_rebase(const int64_t* __restrict__ src, int64_t* __restrict__ dst, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i <= n) dst[i] = src[i] - src[0];
}

}  // namespace

// Launch of k_pattern: the normal configuration, or -- once a row of this space has overflowed it -- one warp per block
// with the large candidate arrays (slow for the long rows: the rank sort is quadratic; a mesh with an edge in more than
// 102 cells is the only way to get there).
template <bool FILL>
static int launch_pattern(const sfem_space_s* sp, SpaceView sv, bool canon, int64_t ent_lo, int64_t n, int64_t inc_base,
                          int32_t* row_len, const int64_t* ent_rowptr, int32_t* ent_cols, uint16_t* pos, uint8_t* rec,
                          int* overflow, cudaStream_t st) {
  if (n <= 0) return 0;
  if (!sp->pattern_big) {
    const unsigned grid = (unsigned)((n + kWarpsPerBlock - 1) / kWarpsPerBlock);
    k_pattern<FILL, kWarpsPerBlock><<<grid, kWarpsPerBlock * 32, (size_t)kWarpsPerBlock * 2 * kCap * 4, st>>>(
        sv, sp->mesh->cells, sp->inc_ptr, sp->inc_cell, sp->inc_loc, canon, ent_lo, n, inc_base, row_len, ent_rowptr,
        ent_cols, pos, rec, overflow, kCap);
  } else {
    SFEM_CUDA(cudaFuncSetAttribute(k_pattern<FILL, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kCapBig * 4));
    k_pattern<FILL, 1><<<(unsigned)n, 32, (size_t)2 * kCapBig * 4, st>>>(
        sv, sp->mesh->cells, sp->inc_ptr, sp->inc_cell, sp->inc_loc, canon, ent_lo, n, inc_base, row_len, ent_rowptr,
        ent_cols, pos, rec, overflow, kCapBig);
  }
  SFEM_CUDA(cudaGetLastError());
  return 0;
}

// Count pass over the row entities [ent_lo, ent_hi): allocates *rowptr_out [n+1] (exclusive scan of the row lengths,
// offsets relative to ent_lo) and returns the longest edge row / other row.  Synchronises `st`.
int count_pattern_rows(const sfem_space_s* sp, int64_t ent_lo, int64_t ent_hi, cudaStream_t st, int64_t** rowptr_out,
                       int64_t* nnz_ent, int* max_len_edges, int* max_len_other) {
  const int64_t n = ent_hi - ent_lo;
  SpaceView sv = make_view(sp);
  const bool canon = pos_is_canonical(sp);
  int32_t* row_len = nullptr;
  int64_t* len64 = nullptr;
  int* overflow = nullptr;
  int64_t* rowptr = nullptr;
  SFEM_CUDA(dev_malloc(&row_len, std::max<int64_t>(n, 1) * 4, st));
  SFEM_CUDA(dev_malloc(&len64, (n + 1) * 8, st));
  SFEM_CUDA(dev_malloc(&overflow, 4, st));
  SFEM_CUDA(cudaMemsetAsync(overflow, 0, 4, st));
  SFEM_CUDA(dev_malloc(&rowptr, (n + 1) * 8, st));
  SFEM_TRY(launch_pattern<false>(sp, sv, canon, ent_lo, n, 0, row_len, nullptr, nullptr, nullptr, nullptr, overflow, st));
  k_len_to_i64<<<grid_for(n + 1), 256, 0, st>>>(row_len, len64, n);
  size_t tb = 0;
  SFEM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, len64, rowptr, (int)(n + 1), st));
  size_t tb2 = 0;
  SFEM_CUDA(cub::DeviceReduce::Max(nullptr, tb2, row_len, (int32_t*)nullptr, (int)std::max<int64_t>(n, 1), st));
  void* tmp = nullptr;
  SFEM_CUDA(dev_malloc(&tmp, std::max(tb, tb2) + 16, st));
  SFEM_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, len64, rowptr, (int)(n + 1), st));
  int32_t* d_max = (int32_t*)len64;  // len64 is dead after the scan
  const int64_t n_edge_rows = std::max<int64_t>(0, std::min(ent_hi, sp->E) - ent_lo);
  if (n_edge_rows > 0) SFEM_CUDA(cub::DeviceReduce::Max(tmp, tb2, row_len, d_max, (int)n_edge_rows, st));
  if (n > n_edge_rows)
    SFEM_CUDA(cub::DeviceReduce::Max(tmp, tb2, row_len + n_edge_rows, d_max + 1, (int)(n - n_edge_rows), st));
  int h_over = 0, h_max2[2] = {0, 0};
  SFEM_CUDA(cudaMemcpyAsync(nnz_ent, rowptr + n, 8, cudaMemcpyDeviceToHost, st));
  SFEM_CUDA(cudaMemcpyAsync(&h_over, overflow, 4, cudaMemcpyDeviceToHost, st));
  if (n_edge_rows > 0) SFEM_CUDA(cudaMemcpyAsync(&h_max2[0], d_max, 4, cudaMemcpyDeviceToHost, st));
  if (n > n_edge_rows) SFEM_CUDA(cudaMemcpyAsync(&h_max2[1], d_max + 1, 4, cudaMemcpyDeviceToHost, st));
  SFEM_CUDA(cudaStreamSynchronize(st));
  dev_free(tmp, st);
  dev_free(len64, st);
  dev_free(row_len, st);
  dev_free(overflow, st);
  if (h_over) {
    dev_free(rowptr, st);
    if (!sp->pattern_big) {  // a row with more than 1024 candidates: redo the count with the large-row launch
      const_cast<sfem_space_s*>(sp)->pattern_big = true;
      return count_pattern_rows(sp, ent_lo, ent_hi, st, rowptr_out, nnz_ent, max_len_edges, max_len_other);
    }
    set_error("an entity has more incident cells than the pattern builder supports (valence * entities per cell > 24576)");
    return 3;
  }
  *rowptr_out = rowptr;
  *max_len_edges = h_max2[0];
  *max_len_other = h_max2[1];
  return 0;
}

// Fill pass: expects ent_rowptr, nnz_ent, max_row_len*, inc_base and n_inc_block of `pat` to be set.
static int fill_pattern_impl(sfem_pattern_s* pat, cudaStream_t st, bool allow_e2) {
  const sfem_space_s* sp = pat->space;
  const int64_t n = pat->n_rows_ent;
  SpaceView sv = make_view(sp);
  const bool canon = pos_is_canonical(sp);
  const int64_t n_edge_rows = std::max<int64_t>(0, std::min(pat->ent_hi, sp->E) - pat->ent_lo);
  // incidences of the block's edge rows: inc_ptr[E] = (cells) x (edges per cell), no read-back needed
  const int64_t inc_at_E = sp->mesh->C * sp->ne_loc;
  const int64_t n_edge_inc = std::min(std::max(inc_at_E, pat->inc_base), pat->inc_base + pat->n_inc_block) - pat->inc_base;
  // slots are bytes: face rows hold at most 16 entities, edge rows longer than the tile (224 entities) are excluded from the
  // batched layout one by one (k_e2_batches); the lane-pair layout needs every row to fit
  const bool e2_possible = allow_e2 && canon && sp->order == 2 && edge_kernel_version() == 2 && pat->nnz_ent < (1ll << 32) &&
                           pat->max_row_len_other <= 16;
  const bool want_rec = canon && (pat->max_row_len <= 255 || e2_possible);
  // lane-per-item edge kernel (order-2 tets): every size below is an upper bound known on the host, so that the block
  // pipelines of host_pipeline.cu see the same allocation sequence block after block and no extra synchronisation
  const int vmax = (int)std::min<int64_t>(32, std::max<int64_t>(1, (pat->max_row_len_edges - 1) / 6));  // ring of m cells: >= 6m+1 entities
  const bool want_e2 = e2_possible && n_edge_rows > 0;
  int* overflow = nullptr;  // [0] a builder gave up (whole-pattern fallback), [1] rows handed to the generic kernel
  SFEM_CUDA(dev_malloc(&overflow, 8, st));
  SFEM_CUDA(cudaMemsetAsync(overflow, 0, 8, st));
  SFEM_CUDA(dev_malloc(&pat->ent_cols, std::max<int64_t>(pat->nnz_ent, 1) * 4, st));
  SFEM_CUDA(dev_malloc(&pat->pos, std::max<int64_t>(pat->n_inc_block * sp->n_ent_cell, 1) * 2, st));
  if (want_rec) {
    // with the batched edge layout only the FACE records live in `rec` (indexed by incidence - inc_base as before:
    // the pointer is shifted, its first n_edge_inc entries are never touched)
    const int64_t skip = want_e2 ? n_edge_inc : 0;
    SFEM_CUDA(dev_malloc(&pat->rec_alloc, std::max<int64_t>(pat->n_inc_block - skip, 1) * 32, st));
    pat->rec = pat->rec_alloc - 2 * skip;
  }
  // (a row that overflows here has overflowed the count pass before: sp->pattern_big is already set)
  SFEM_TRY(launch_pattern<true>(sp, sv, canon, pat->ent_lo, n, pat->inc_base, nullptr, pat->ent_rowptr, pat->ent_cols, pat->pos,
                                reinterpret_cast<uint8_t*>(pat->rec), overflow, st));
  pat->ring_ok = false;
  int64_t* e2_slot0 = nullptr;
  uint32_t* e2_tileoff = nullptr;
  uint2* e2_hdr = nullptr;
  if (want_e2) {
    const int64_t nseg = (n_edge_rows + kE2Seg - 1) / kE2Seg;
    // a batch closed because the next edge (at most vmax items) did not fit holds more than 32 - vmax items; batches
    // closed by the tile limit (many low-valence boundary edges) get 1/64 of the items as slack, and a mesh that needs
    // more than this bound makes k_e2_batches raise `bad` (fallback below)
    const int64_t nb_max = n_edge_inc / (33 - std::min(vmax, 16)) + n_edge_inc / 64 + nseg + 16;
    const int gen_cap = (int)std::min<int64_t>(n_edge_rows, 1 << 20);
    SFEM_CUDA(dev_malloc(&pat->gen_rows, (size_t)gen_cap * 8, st));
    int32_t* seg_nb = nullptr;
    int64_t* seg_b0 = nullptr;
    SFEM_CUDA(dev_malloc(&seg_nb, (nseg + 1) * 4, st));
    SFEM_CUDA(dev_malloc(&seg_b0, (nseg + 1) * 8, st));
    SFEM_CUDA(cudaMemsetAsync(seg_nb, 0, (nseg + 1) * 4, st));
    k_e2_batches<false><<<grid_for(nseg, 64), 64, 0, st>>>(sp->inc_ptr, pat->ent_rowptr, pat->ent_lo, n_edge_rows, 32, nb_max,
                                                           seg_nb, nullptr, nullptr, nullptr, nullptr, overflow, nullptr, 0);
    size_t tb = 0;
    SFEM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, seg_nb, seg_b0, (int)(nseg + 1), st));
    void* tmp = nullptr;
    SFEM_CUDA(dev_malloc(&tmp, tb + 16, st));
    SFEM_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, seg_nb, seg_b0, (int)(nseg + 1), st));
    dev_free(tmp, st);
    pat->n_batch_e2 = nb_max;
    SFEM_CUDA(dev_malloc(&pat->rec_e2, nb_max * 32 * 32, st));
    SFEM_CUDA(dev_malloc(&pat->aux_e2, nb_max * 32 * 16, st));
    SFEM_CUDA(dev_malloc(&pat->nbatch_e2_dev, 8, st));
    SFEM_CUDA(cudaMemsetAsync(pat->rec_e2, 0xff, nb_max * 32 * 32, st));  // padding slots: cell = -1
    SFEM_CUDA(cudaMemsetAsync(pat->aux_e2, 0, nb_max * 32 * 16, st));
    SFEM_CUDA(cudaMemcpyAsync(pat->nbatch_e2_dev, seg_b0 + nseg, 8, cudaMemcpyDeviceToDevice, st));
    SFEM_CUDA(dev_malloc(&e2_slot0, n_edge_rows * 8, st));
    SFEM_CUDA(dev_malloc(&e2_tileoff, n_edge_rows * 4, st));
    SFEM_CUDA(dev_malloc(&e2_hdr, nb_max * 8, st));
    k_e2_batches<true><<<grid_for(nseg, 64), 64, 0, st>>>(sp->inc_ptr, pat->ent_rowptr, pat->ent_lo, n_edge_rows, 32, nb_max,
                                                          nullptr, seg_b0, e2_slot0, e2_tileoff, e2_hdr, overflow,
                                                          pat->gen_rows, gen_cap);
    dev_free(seg_nb, st);
    dev_free(seg_b0, st);
  }
  bool e2_failed = false;
  if (want_rec && n_edge_rows > 0) {
    k_edge_ring<<<grid_for(n_edge_rows, 128), 128, 0, st>>>(
        sv, sp->mesh->cells, sp->inc_ptr, sp->inc_cell, sp->inc_loc, pat->ent_lo, n_edge_rows, pat->inc_base,
        pat->ent_rowptr, pat->ent_cols, reinterpret_cast<uint8_t*>(want_e2 ? pat->rec_e2 : pat->rec), overflow, e2_slot0,
        e2_tileoff, e2_hdr, pat->aux_e2);
    int h_flags[2] = {0, 0};
    SFEM_CUDA(cudaMemcpyAsync(h_flags, overflow, 8, cudaMemcpyDeviceToHost, st));
    SFEM_CUDA(cudaStreamSynchronize(st));
    const int h_bad = h_flags[0];
    pat->ring_ok = h_bad == 0;
    pat->n_gen_rows = want_e2 && h_bad == 0 ? h_flags[1] : 0;
    e2_failed = want_e2 && h_bad != 0;
    dev_free(e2_slot0, st);
    dev_free(e2_tileoff, st);
    dev_free(e2_hdr, st);
  } else if (want_rec) {
    pat->ring_ok = true;
  }
  SFEM_CUDA(cudaGetLastError());
  dev_free(overflow, st);
  if (e2_failed) {
    // an edge with more cells than the batched layout takes (or a non-manifold one): redo the fill with the lane-pair
    // layout, which decides for itself whether the canonical kernels can run at all
    dev_free(pat->ent_cols, st); pat->ent_cols = nullptr;
    dev_free(pat->pos, st); pat->pos = nullptr;
    dev_free(pat->rec_alloc, st); pat->rec_alloc = nullptr; pat->rec = nullptr;
    dev_free(pat->rec_e2, st); pat->rec_e2 = nullptr;
    dev_free(pat->aux_e2, st); pat->aux_e2 = nullptr;
    dev_free(pat->nbatch_e2_dev, st); pat->nbatch_e2_dev = nullptr;
    dev_free(pat->gen_rows, st); pat->gen_rows = nullptr;
    pat->n_batch_e2 = 0;
    pat->n_gen_rows = 0;
    return fill_pattern_impl(pat, st, false);
  }
  return 0;
}

int fill_pattern(sfem_pattern_s* pat, cudaStream_t st) { return fill_pattern_impl(pat, st, true); }

int build_pattern(sfem_pattern_s* pat, cudaStream_t st) {
  const sfem_space_s* sp = pat->space;
  int64_t bounds[2];
  SFEM_CUDA(cudaMemcpyAsync(&bounds[0], sp->inc_ptr + pat->ent_lo, 8, cudaMemcpyDeviceToHost, st));
  SFEM_CUDA(cudaMemcpyAsync(&bounds[1], sp->inc_ptr + pat->ent_hi, 8, cudaMemcpyDeviceToHost, st));
  SFEM_TRY(count_pattern_rows(sp, pat->ent_lo, pat->ent_hi, st, &pat->ent_rowptr, &pat->nnz_ent, &pat->max_row_len_edges,
                              &pat->max_row_len_other));  // synchronises st: bounds are valid
  pat->inc_base = bounds[0];
  pat->n_inc_block = bounds[1] - bounds[0];
  pat->max_row_len = std::max(pat->max_row_len_edges, pat->max_row_len_other);
  SFEM_TRY(fill_pattern(pat, st));
  SFEM_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// A block [ent_lo, ent_hi) of a pattern whose count pass was done for a superset of rows: rowptr_all is the (device)
// row pointer of the row entities starting at entity all_lo.  Host-side sizes come from the caller.
int build_pattern_block(sfem_pattern_s* pat, const int64_t* rowptr_all, int64_t all_lo, int64_t nnz_ent,
                        int64_t inc_base, int64_t n_inc_block, int max_len_edges, int max_len_other, cudaStream_t st) {
  const int64_t n = pat->n_rows_ent;
  SFEM_CUDA(dev_malloc(&pat->ent_rowptr, (n + 1) * 8, st));
  k_rebase<<<grid_for(n + 1), 256, 0, st>>>(rowptr_all + (pat->ent_lo - all_lo), pat->ent_rowptr, n);
  pat->nnz_ent = nnz_ent;
  pat->inc_base = inc_base;
  pat->n_inc_block = n_inc_block;
  pat->max_row_len_edges = max_len_edges;
  pat->max_row_len_other = max_len_other;
  pat->max_row_len = std::max(max_len_edges, max_len_other);
  return fill_pattern(pat, st);
}

void free_pattern_buffers(sfem_pattern_s* pat, cudaStream_t st) {
  dev_free(pat->ent_rowptr, st); pat->ent_rowptr = nullptr;
  dev_free(pat->ent_cols, st); pat->ent_cols = nullptr;
  dev_free(pat->pos, st); pat->pos = nullptr;
  dev_free(pat->rec_alloc, st); pat->rec_alloc = nullptr; pat->rec = nullptr;
  dev_free(pat->nbatch_e2_dev, st); pat->nbatch_e2_dev = nullptr;
  dev_free(pat->gen_rows, st); pat->gen_rows = nullptr;
  dev_free(pat->rec_e2, st); pat->rec_e2 = nullptr;
  dev_free(pat->aux_e2, st); pat->aux_e2 = nullptr;
  dev_free(pat->abc_rows, st); pat->abc_rows = nullptr;
  dev_free(pat->abc_rowptr, st); pat->abc_rowptr = nullptr;
}

int export_pattern(const sfem_pattern_s* pat, int64_t* d_rowptr, int32_t* d_cols, cudaStream_t st, int64_t nnz_base) {
  const int64_t n = pat->n_rows_ent;
  const unsigned grid = (unsigned)((n + 1 + 7) / 8);
  k_export<<<grid, 256, 0, st>>>(pat->ent_rowptr, pat->ent_cols, pat->space->order, n, nnz_base, d_rowptr, d_cols);
  SFEM_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sfem
