// K3 (canonical-frame variant, tetrahedra order 2): the benchmarked S + M(=T) value kernels.
// Replaces the cell loop of dolfin.assemble + DirichletBC.apply called from
// /root/reference/sucemfem/SystemMatrices.py:54-55 for the forms of /root/reference/sucemfem/Forms.py:76,84.
//
// Work decomposition (row-owner gather, no atomics, bit-reproducible):
//   * one THREAD owns one row entity (edge or face) = 2 CSR rows of S and 2 of T; a warp owns 32 consecutive
//     entities, i.e. one contiguous span of the value arrays.
//   * for every incident cell the thread relabels the cell's vertices so that its entity sits at the canonical
//     position (edge (v0,v1) / face (v0,v1,v2)).  The two element-matrix rows are then straight-line DFMAs whose
//     reference-tensor coefficients are constant-bank operands (generated/canonical_rows.inc), identical for all
//     lanes; structural zeros are gone at compile time.  A 64-bit descriptor per local position maps canonical
//     column entities back to the true ones (orientation fix-up); the slots they land in come precomputed, in
//     canonical order, in one packed 16-byte record per (entity, cell) (pattern.cu).
//   * contributions are staged in a per-warp shared-memory tile that is the warp's output span in its FINAL
//     layout; the span leaves the SM as ONE bulk asynchronous copy (cp.async.bulk shared -> global, the TMA
//     engine), so global stores cost no LSU wavefronts and are perfectly coalesced.  Faces (<= 2 cells; the 4
//     shared entities are summed in registers) only store into the tile.  Edges visit their cells in RING order
//     (consecutive cells share a face through the edge; order precomputed in pattern.cu), carry the three column
//     entities two neighbouring cells have in common in registers and therefore also store every slot exactly once
//     (the one exception: where a closed ring closes, three chunks per row are added to the first cell's).
//   * gathers are the other LSU consumer: one 16-byte record, one 16-byte vertex quadruple, four 32-byte padded
//     coordinates and one coefficient per (entity, cell); for edges the three dependent hops are software-pipelined
//     three items deep because the tile limits residency to a few warps per SM.
//   * PEC rows are overwritten with unit rows in the tile before it is stored (row-only ident fused).
// Geometry (J, detJ, J^-1 J^-T, J^T J) is recomputed per (entity, cell) from the gathered coordinates: ~70 flops
// against ~170 DFMAs of tensor contraction per matrix, cheaper than staging per-cell geometry through HBM.
#include "sfem_internal.h"

#define SFEM_TABLE static __device__ const
#include "generated/canonical_lut.inc"
#include "generated/canonical_rows.inc"

namespace sfem {
namespace {

struct CanonArgs {
  const double* __restrict__ coords4;      // [V][4]
  const int32_t* __restrict__ cells;
  const double* __restrict__ eps;
  const double* __restrict__ mu;
  const int64_t* __restrict__ inc_ptr;
  const int64_t* __restrict__ ent_rowptr;  // block-relative, indexed by r = g - ent_lo
  const uint4* __restrict__ rec;           // packed {cell, t, slots[10]} per (entity, cell), block-relative
  const uint8_t* __restrict__ pec;
  double* __restrict__ S;
  double* __restrict__ T;
  int64_t ent_lo;    // first entity of the pattern block
  int64_t r_begin;   // first block-relative row entity handled by this launch
  int64_t r_end;
  int64_t inc_base;
  int tile_chunks;   // 16-byte chunks per warp tile = 32 * 2 * (longest entity row in this launch)
};

// per-(entity, cell) inputs
struct Item {
  double X[12];             // coordinates of the relabelled vertices
  double coef;              // eps_r (mass pass) or mu_r (stiffness pass) of the cell
  unsigned long long desc;  // generated/canonical_lut.inc
  unsigned fsel;
  uint32_t s0, s1, s2;      // bytes 4..15 of the record: t, slots[10]
};

template <int U>
__device__ __forceinline__ int slot_of(const Item& it) {
  // record byte 5+U
  if constexpr (U < 3) return (it.s0 >> (8 * (U + 1))) & 0xffu;
  else if constexpr (U < 7) return (it.s1 >> (8 * (U - 3))) & 0xffu;
  else return (it.s2 >> (8 * (U - 7))) & 0xffu;
}

// stage 3 of the gather chain: coordinates (needs the record and the cell's vertex quadruple)
__device__ __forceinline__ void gather_coords(const CanonArgs& p, const unsigned long long* __restrict__ s_desc,
                                              const unsigned* __restrict__ s_fsel, uint4 rc, int4 cv, double coef,
                                              Item& it) {
  const int t = rc.y & 0xffu;
  it.desc = s_desc[t];
  it.fsel = s_fsel[t];
  it.s0 = rc.y; it.s1 = rc.z; it.s2 = rc.w;
  it.coef = coef;
  const int32_t a[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const unsigned s = (unsigned)(it.desc >> (2 * k)) & 3u;
    const uint32_t v = (uint32_t)(s == 0 ? a[0] : (s == 1 ? a[1] : (s == 2 ? a[2] : a[3])));
    const double2* x = reinterpret_cast<const double2*>(p.coords4 + (size_t)v * 4);
    const double2 xy = x[0], z = x[1];
    it.X[3 * k] = xy.x;
    it.X[3 * k + 1] = xy.y;
    it.X[3 * k + 2] = z.x;
  }
}

// geometry factors of the relabelled cell in symmetric order (00,11,22,01,02,12):
//   MASS: |detJ| eps (J^-1 J^-T)      else: (J^T J) / (|detJ| mu)
template <bool MASS>
__device__ __forceinline__ void canon_geometry(const Item& it, double (&g)[6]) {
  double J[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int d = 0; d < 3; ++d) J[d][a] = it.X[3 * (a + 1) + d] - it.X[d];
  const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
  const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
  const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
  const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
  const double ad = fabs(det);
  if (MASS) {
    // adj(J) = det J^-1;  J^-1 J^-T |det| eps = adj adj^T eps / |det|
    double A[3][3];
    A[0][0] = c00; A[1][0] = c01; A[2][0] = c02;
    A[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
    A[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
    A[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
    A[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
    A[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
    A[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double s = it.coef / ad;
    g[0] = (A[0][0] * A[0][0] + A[0][1] * A[0][1] + A[0][2] * A[0][2]) * s;
    g[1] = (A[1][0] * A[1][0] + A[1][1] * A[1][1] + A[1][2] * A[1][2]) * s;
    g[2] = (A[2][0] * A[2][0] + A[2][1] * A[2][1] + A[2][2] * A[2][2]) * s;
    g[3] = (A[0][0] * A[1][0] + A[0][1] * A[1][1] + A[0][2] * A[1][2]) * s;
    g[4] = (A[0][0] * A[2][0] + A[0][1] * A[2][1] + A[0][2] * A[2][2]) * s;
    g[5] = (A[1][0] * A[2][0] + A[1][1] * A[2][1] + A[1][2] * A[2][2]) * s;
  } else {
    const double s = 1.0 / (ad * it.coef);
    g[0] = (J[0][0] * J[0][0] + J[1][0] * J[1][0] + J[2][0] * J[2][0]) * s;
    g[1] = (J[0][1] * J[0][1] + J[1][1] * J[1][1] + J[2][1] * J[2][1]) * s;
    g[2] = (J[0][2] * J[0][2] + J[1][2] * J[1][2] + J[2][2] * J[2][2]) * s;
    g[3] = (J[0][0] * J[0][1] + J[1][0] * J[1][1] + J[2][0] * J[2][1]) * s;
    g[4] = (J[0][0] * J[0][2] + J[1][0] * J[1][2] + J[2][0] * J[2][2]) * s;
    g[5] = (J[0][1] * J[0][2] + J[1][1] * J[1][2] + J[2][1] * J[2][2]) * s;
  }
}

__device__ __forceinline__ double sel3(unsigned code, double v0, double v1, double w) {
  return code == 0 ? v0 : (code == 1 ? v1 : w);
}

// true-frame values of the 2 columns of canonical column entity U, from canonical-frame values (v0, v1)
template <int U>
__device__ __forceinline__ double2 fix_columns(const Item& it, double v0, double v1) {
  if constexpr (U < 6) {
    const bool flip = (it.desc >> (48 + U)) & 1ull;
    return flip ? make_double2(-v1, -v0) : make_double2(v0, v1);
  } else {
    const unsigned nib = (it.fsel >> (4 * (U - 6))) & 15u;
    const double w = -(v0 + v1);
    return make_double2(sel3(nib & 3u, v0, v1, w), sel3(nib >> 2, v0, v1, w));
  }
}

// ---- the warp's output span leaves the SM as one bulk asynchronous copy -----------------------------------------
__device__ __forceinline__ void tile_store_issue(const double2* tile, double* gdst, unsigned bytes, int lane) {
  // make the generic-proxy writes of every lane visible to the async proxy, then one lane issues the copy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0 && bytes > 0) {
    const unsigned src = (unsigned)__cvta_generic_to_shared(tile);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(src), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}
// the tile may be overwritten once the engine has finished READING it
__device__ __forceinline__ void tile_store_wait(int lane) {
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
}

__device__ __forceinline__ void pec_overwrite(const CanonArgs& p, double2* __restrict__ row0, int64_t g, int len, int self) {
#pragma unroll
  for (int k = 0; k < 2; ++k)
    if (p.pec[2 * g + k]) {
      double2* row = row0 + k * len;
      for (int s = 0; s < len; ++s) row[s] = make_double2(0.0, 0.0);
      row[self] = k == 0 ? make_double2(1.0, 0.0) : make_double2(0.0, 1.0);
    }
}

__device__ __forceinline__ void stage_lut(unsigned long long* s_desc, unsigned* s_fsel) {
  if (threadIdx.x < 16) {
    s_desc[threadIdx.x] = kRingDesc[threadIdx.x];
    s_fsel[threadIdx.x] = kRingFaceSel[threadIdx.x];
  }
  __syncthreads();
}

constexpr int kThreadsEdges = 64;
constexpr int kThreadsFaces = 128;

// ---- edges: cells visited in ring order, every slot stored exactly once ------------------------------------------
// Relabelled cell = (a, b, entry, exit).  Canonical column entities by role:
//   self   5 = (0,1)                         accumulated in registers over all cells, stored at the end
//   prev   4 = (0,2), 2 = (1,2), 9 = (0,1,2)  shared with the previous cell of the ring: carry + own, then stored
//   own    0 = (2,3), 7 = (0,2,3), 6 = (1,2,3) this cell only: stored
//   next   3 = (0,3), 1 = (1,3), 8 = (0,1,3)  shared with the next cell: kept as the carry; stored if the fan ends
//                                              here, added to the first cell's slots if the ring closes here
struct RingState {
  double2 carry[2][3];
  double2 self[2];
};

template <bool MASS>
__device__ __forceinline__ void edge_item(const Item& it, double2* __restrict__ row0, int len, RingState& st) {
  double gm[6], o[20];
  canon_geometry<MASS>(it, gm);
  const unsigned flags = it.s2 >> 24;
  const bool has_prev = flags & 1u, has_next = flags & 2u, wrap = flags & 4u;
#define SFEM_RING_ROW(K, FN)                                                                       \
  {                                                                                                \
    FN(gm, o);                                                                                     \
    double2* row = row0 + K * len;                                                                 \
    double2 r;                                                                                     \
    r = fix_columns<5>(it, o[10], o[11]); st.self[K].x += r.x; st.self[K].y += r.y;                \
    r = fix_columns<4>(it, o[8], o[9]);                                                            \
    if (has_prev) { r.x += st.carry[K][0].x; r.y += st.carry[K][0].y; }                            \
    row[slot_of<4>(it)] = r;                                                                       \
    r = fix_columns<2>(it, o[4], o[5]);                                                            \
    if (has_prev) { r.x += st.carry[K][1].x; r.y += st.carry[K][1].y; }                            \
    row[slot_of<2>(it)] = r;                                                                       \
    r = fix_columns<9>(it, o[18], o[19]);                                                          \
    if (has_prev) { r.x += st.carry[K][2].x; r.y += st.carry[K][2].y; }                            \
    row[slot_of<9>(it)] = r;                                                                       \
    row[slot_of<0>(it)] = fix_columns<0>(it, o[0], o[1]);                                          \
    row[slot_of<7>(it)] = fix_columns<7>(it, o[14], o[15]);                                        \
    row[slot_of<6>(it)] = fix_columns<6>(it, o[12], o[13]);                                        \
    st.carry[K][0] = fix_columns<3>(it, o[6], o[7]);                                               \
    st.carry[K][1] = fix_columns<1>(it, o[2], o[3]);                                               \
    st.carry[K][2] = fix_columns<8>(it, o[16], o[17]);                                             \
    if (!has_next) {                                                                               \
      double2* q3 = row + slot_of<3>(it);                                                          \
      double2* q1 = row + slot_of<1>(it);                                                          \
      double2* q8 = row + slot_of<8>(it);                                                          \
      double2 c3 = st.carry[K][0], c1 = st.carry[K][1], c8 = st.carry[K][2];                       \
      if (wrap) {                                                                                  \
        const double2 t3 = *q3, t1 = *q1, t8 = *q8;                                                \
        c3.x += t3.x; c3.y += t3.y; c1.x += t1.x; c1.y += t1.y; c8.x += t8.x; c8.y += t8.y;        \
      }                                                                                            \
      *q3 = c3; *q1 = c1; *q8 = c8;                                                                \
    }                                                                                              \
  }
  if (MASS) {
    SFEM_RING_ROW(0, canon_edge_mass_k0)
    SFEM_RING_ROW(1, canon_edge_mass_k1)
  } else {
    SFEM_RING_ROW(0, canon_edge_curl_k0)
    SFEM_RING_ROW(1, canon_edge_curl_k1)
  }
#undef SFEM_RING_ROW
}

// The gather chain of an item is three dependent global loads (record -> cell vertices -> coordinates).  With only
// a few resident warps (the tile bounds occupancy) it is hidden by issuing the hops three, two and one items ahead
// of the item being computed; the two Item buffers ping-pong.
template <bool MASS>
__device__ __forceinline__ void edge_pass(const CanonArgs& p, const unsigned long long* __restrict__ s_desc,
                                          const unsigned* __restrict__ s_fsel, double2* __restrict__ tile, int lane,
                                          bool valid, int64_t g, int len, int off, int span, int64_t i0, int64_t i1,
                                          double* __restrict__ out_first) {
  const double* __restrict__ coefs = MASS ? p.eps : p.mu;
  const uint4* __restrict__ rec = p.rec - p.inc_base;
  // prologue of the chain (independent of the tile): records of items 0..2, vertices of items 0..1
  uint4 rc0 = make_uint4(0, 0, 0, 0), rc1 = rc0, rc2 = rc0;
  int4 cv1 = make_int4(0, 0, 0, 0);
  double cf1 = 1.0;
  Item bufA, bufB;
  if (valid) {
    rc0 = rec[i0];
    if (i0 + 1 < i1) rc1 = rec[i0 + 1];
    if (i0 + 2 < i1) rc2 = rec[i0 + 2];
    const int4 cv0 = *reinterpret_cast<const int4*>(p.cells + (int64_t)(int)rc0.x * 4);
    const double cf0 = coefs ? coefs[(int)rc0.x] : 1.0;
    if (i0 + 1 < i1) {
      cv1 = *reinterpret_cast<const int4*>(p.cells + (int64_t)(int)rc1.x * 4);
      cf1 = coefs ? coefs[(int)rc1.x] : 1.0;
    }
    gather_coords(p, s_desc, s_fsel, rc0, cv0, cf0, bufA);
  }
  // the previous pass's bulk store must have finished reading the tile before it is overwritten
  tile_store_wait(lane);
  double2* row0 = tile + off;
  if (valid) {
    RingState st;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      st.self[k] = make_double2(0.0, 0.0);
#pragma unroll
      for (int j = 0; j < 3; ++j) st.carry[k][j] = make_double2(0.0, 0.0);
    }
    // STEP: hop 1 of item I+3, hop 2 of item I+2, hop 3 of item I+1 (into NXT), then consume item I (CUR)
#define SFEM_EDGE_STEP(CUR, NXT, I)                                                                       \
    {                                                                                                     \
      uint4 rc3 = make_uint4(0, 0, 0, 0);                                                                 \
      int4 cv2 = make_int4(0, 0, 0, 0);                                                                   \
      double cf2 = 1.0;                                                                                   \
      if ((I) + 3 < i1) rc3 = rec[(I) + 3];                                                               \
      if ((I) + 2 < i1) {                                                                                 \
        cv2 = *reinterpret_cast<const int4*>(p.cells + (int64_t)(int)rc2.x * 4);                          \
        cf2 = coefs ? coefs[(int)rc2.x] : 1.0;                                                            \
      }                                                                                                   \
      if ((I) + 1 < i1) gather_coords(p, s_desc, s_fsel, rc1, cv1, cf1, NXT);                             \
      edge_item<MASS>(CUR, row0, len, st);                                                                \
      rc1 = rc2; rc2 = rc3; cv1 = cv2; cf1 = cf2;                                                         \
    }
    const int self = slot_of<5>(bufA);  // canonical entity 5 = the row edge itself
#pragma unroll 1
    for (int64_t i = i0; i < i1; i += 2) {
      SFEM_EDGE_STEP(bufA, bufB, i)
      if (i + 1 < i1) SFEM_EDGE_STEP(bufB, bufA, i + 1)
    }
#undef SFEM_EDGE_STEP
    row0[self] = st.self[0];
    row0[len + self] = st.self[1];
    if (p.pec) pec_overwrite(p, row0, g, len, self);
  }
  tile_store_issue(tile, out_first, (unsigned)span * 16u, lane);
}

__global__ void __launch_bounds__(kThreadsEdges) k_canon_edges(CanonArgs p) {
  extern __shared__ double2 smem[];
  __shared__ unsigned long long s_desc[16];
  __shared__ unsigned s_fsel[16];
  stage_lut(s_desc, s_fsel);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double2* tile = smem + (size_t)w * p.tile_chunks;
  const int64_t r = p.r_begin + blockIdx.x * (int64_t)kThreadsEdges + threadIdx.x;
  const bool valid = r < p.r_end;
  int64_t g = 0, a = 0, a1 = 0, i0 = 0, i1 = 0;
  if (valid) {
    g = p.ent_lo + r;
    a = p.ent_rowptr[r];
    a1 = p.ent_rowptr[r + 1];
    i0 = p.inc_ptr[g];
    i1 = p.inc_ptr[g + 1];
  }
  const int len = (int)(a1 - a);
  // the warp's span: chunks [2 a_first, 2 a_last_end) of the value arrays (lanes are consecutive entities)
  const unsigned vmask = __ballot_sync(0xffffffffu, valid);
  if (vmask == 0) return;
  const int last = 31 - __clz((int)vmask);
  const int64_t a_first = __shfl_sync(0xffffffffu, a, 0);
  const int64_t a_end = __shfl_sync(0xffffffffu, a1, last);
  const int span = (int)(2 * (a_end - a_first));
  const int off = (int)(2 * (a - a_first));
  edge_pass<false>(p, s_desc, s_fsel, tile, lane, valid, g, len, off, span, i0, i1, p.S + 4 * a_first);
  edge_pass<true>(p, s_desc, s_fsel, tile, lane, valid, g, len, off, span, i0, i1, p.T + 4 * a_first);
  tile_store_wait(lane);
}

// ---- faces: at most two cells; shared entities summed in registers, tile is store-only --------------------------
template <bool MASS>
__device__ __forceinline__ void face_pass(const CanonArgs& p, double2* __restrict__ tile, int lane, bool valid,
                                          int64_t g, int len, int off, int span, int ncell, Item& A, Item& B,
                                          double* __restrict__ out_first) {
  tile_store_wait(lane);
  double2* row0 = tile + off;
  if (valid) {
    double gA[6], gB[6], oA[20];
    canon_geometry<MASS>(A, gA);
    if (ncell > 1) canon_geometry<MASS>(B, gB);
#define SFEM_FACE_PUT(U, K, IT, O) row0[K * len + slot_of<U>(IT)] = fix_columns<U>(IT, O[2 * U], O[2 * U + 1]);
#define SFEM_FACE_SHARED(U) oA[2 * U] += oB[2 * U]; oA[2 * U + 1] += oB[2 * U + 1];
#define SFEM_FACE_ROW(K, FN)                                                                              \
    {                                                                                                     \
      if (ncell > 1) {                                                                                    \
        double oB[20];                                                                                    \
        FN(gB, oB);                                                                                       \
        SFEM_FACE_PUT(0, K, B, oB) SFEM_FACE_PUT(1, K, B, oB) SFEM_FACE_PUT(3, K, B, oB)                  \
        SFEM_FACE_PUT(6, K, B, oB) SFEM_FACE_PUT(7, K, B, oB) SFEM_FACE_PUT(8, K, B, oB)                  \
        FN(gA, oA);                                                                                       \
        /* entities of the shared face (canonical 2,4,5 = its edges, 9 = itself) look alike from both cells */ \
        SFEM_FACE_SHARED(2) SFEM_FACE_SHARED(4) SFEM_FACE_SHARED(5) SFEM_FACE_SHARED(9)                   \
      } else {                                                                                            \
        FN(gA, oA);                                                                                       \
      }                                                                                                   \
      SFEM_FACE_PUT(0, K, A, oA) SFEM_FACE_PUT(1, K, A, oA) SFEM_FACE_PUT(2, K, A, oA)                    \
      SFEM_FACE_PUT(3, K, A, oA) SFEM_FACE_PUT(4, K, A, oA) SFEM_FACE_PUT(5, K, A, oA)                    \
      SFEM_FACE_PUT(6, K, A, oA) SFEM_FACE_PUT(7, K, A, oA) SFEM_FACE_PUT(8, K, A, oA)                    \
      SFEM_FACE_PUT(9, K, A, oA)                                                                          \
    }
    if (MASS) {
      SFEM_FACE_ROW(0, canon_face_mass_k0)
      SFEM_FACE_ROW(1, canon_face_mass_k1)
    } else {
      SFEM_FACE_ROW(0, canon_face_curl_k0)
      SFEM_FACE_ROW(1, canon_face_curl_k1)
    }
#undef SFEM_FACE_ROW
#undef SFEM_FACE_SHARED
#undef SFEM_FACE_PUT
    if (p.pec) pec_overwrite(p, row0, g, len, slot_of<9>(A));  // canonical entity 9 = the row face itself
  }
  tile_store_issue(tile, out_first, (unsigned)span * 16u, lane);
}

__global__ void __launch_bounds__(kThreadsFaces, 3) k_canon_faces(CanonArgs p) {
  extern __shared__ double2 smem[];
  __shared__ unsigned long long s_desc[16];
  __shared__ unsigned s_fsel[16];
  stage_lut(s_desc, s_fsel);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double2* tile = smem + (size_t)w * p.tile_chunks;
  const int64_t r = p.r_begin + blockIdx.x * (int64_t)kThreadsFaces + threadIdx.x;
  const bool valid = r < p.r_end;
  int64_t g = 0, a = 0, a1 = 0;
  int ncell = 0;
  Item A, B;
  double mu_A = 1.0, mu_B = 1.0, eps_A = 1.0, eps_B = 1.0;
  if (valid) {
    g = p.ent_lo + r;
    a = p.ent_rowptr[r];
    a1 = p.ent_rowptr[r + 1];
    const int64_t i0 = p.inc_ptr[g];
    ncell = (int)(p.inc_ptr[g + 1] - i0);
    // both gather chains are issued before anything is consumed
    const uint4* __restrict__ rec = p.rec - p.inc_base;
    const uint4 rA = rec[i0];
    const uint4 rB = ncell > 1 ? rec[i0 + 1] : rA;
    const int4 cvA = *reinterpret_cast<const int4*>(p.cells + (int64_t)(int)rA.x * 4);
    const int4 cvB = *reinterpret_cast<const int4*>(p.cells + (int64_t)(int)rB.x * 4);
    if (p.mu) { mu_A = p.mu[(int)rA.x]; mu_B = p.mu[(int)rB.x]; }
    if (p.eps) { eps_A = p.eps[(int)rA.x]; eps_B = p.eps[(int)rB.x]; }
    gather_coords(p, s_desc, s_fsel, rA, cvA, mu_A, A);
    gather_coords(p, s_desc, s_fsel, rB, cvB, mu_B, B);
  }
  const int len = (int)(a1 - a);
  const unsigned vmask = __ballot_sync(0xffffffffu, valid);
  if (vmask == 0) return;
  const int last = 31 - __clz((int)vmask);
  const int64_t a_first = __shfl_sync(0xffffffffu, a, 0);
  const int64_t a_end = __shfl_sync(0xffffffffu, a1, last);
  const int span = (int)(2 * (a_end - a_first));
  const int off = (int)(2 * (a - a_first));
  face_pass<false>(p, tile, lane, valid, g, len, off, span, ncell, A, B, p.S + 4 * a_first);
  A.coef = eps_A;
  B.coef = eps_B;
  face_pass<true>(p, tile, lane, valid, g, len, off, span, ncell, A, B, p.T + 4 * a_first);
  tile_store_wait(lane);
}

}  // namespace

bool canonical_supported(const sfem_pattern_s* pat) {
  return pat->space->mesh->gdim == 3 && pat->space->order == 2 && pat->rec != nullptr && pat->ring_ok;
}

int canonical_launches(const sfem_pattern_s* pat) {
  const int64_t E = pat->space->E;
  return (pat->ent_lo < E ? 1 : 0) + (pat->ent_hi > E ? 1 : 0);
}

int assemble_canonical_ST(const sfem_pattern_s* pat, const double* eps, const double* mu, const uint8_t* pec,
                          double* S, double* T, cudaStream_t st) {
  const sfem_space_s* sp = pat->space;
  const sfem_mesh_s* m = sp->mesh;
  CanonArgs p;
  p.coords4 = sp->coords4; p.cells = m->cells; p.eps = eps; p.mu = mu;
  p.inc_ptr = sp->inc_ptr; p.ent_rowptr = pat->ent_rowptr; p.rec = pat->rec; p.pec = pec; p.S = S; p.T = T;
  p.ent_lo = pat->ent_lo; p.inc_base = pat->inc_base;
  const int64_t E = sp->E;
  const int64_t n_edge_rows = std::max<int64_t>(0, std::min(pat->ent_hi, E) - pat->ent_lo);
  const int64_t n_rows = pat->n_rows_ent;
  if (n_edge_rows > 0) {
    p.r_begin = 0; p.r_end = n_edge_rows;
    p.tile_chunks = 32 * 2 * pat->max_row_len_edges;
    const size_t smem = (size_t)(kThreadsEdges / 32) * p.tile_chunks * sizeof(double2);
    SFEM_CHECK(smem <= 227 * 1024, "edge valence too high for the canonical kernel's shared-memory tile; use the generic kernel");
    SFEM_CUDA(cudaFuncSetAttribute(k_canon_edges, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((n_edge_rows + kThreadsEdges - 1) / kThreadsEdges);
    k_canon_edges<<<grid, kThreadsEdges, smem, st>>>(p);
  }
  if (n_rows > n_edge_rows) {
    // a face shared by more than two cells is not a valid tetrahedral mesh
    SFEM_CHECK(pat->max_row_len_other <= 16, "a face is shared by more than two cells: not a valid tetrahedral mesh");
    p.r_begin = n_edge_rows; p.r_end = n_rows;
    p.tile_chunks = 32 * 2 * 16;
    const size_t smem = (size_t)(kThreadsFaces / 32) * p.tile_chunks * sizeof(double2);
    SFEM_CUDA(cudaFuncSetAttribute(k_canon_faces, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((n_rows - n_edge_rows + kThreadsFaces - 1) / kThreadsFaces);
    k_canon_faces<<<grid, kThreadsFaces, smem, st>>>(p);
  }
  SFEM_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sfem
