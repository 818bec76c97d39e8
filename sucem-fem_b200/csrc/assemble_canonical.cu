// K3 (canonical-frame variant, tetrahedra order 2): the benchmarked S + M(=T) value kernels.
// Replaces the cell loop of dolfin.assemble + DirichletBC.apply called from
// /root/reference/sucemfem/SystemMatrices.py:54-55 for the forms of /root/reference/sucemfem/Forms.py:76,84.
//
// Work decomposition (row-owner gather, no atomics, bit-reproducible):
//   * one THREAD owns one row entity (edge or face) = 2 CSR rows of S and 2 of T; a warp owns 32 consecutive
//     entities, i.e. one contiguous span of the value arrays.
//   * for every incident cell the thread relabels the cell's vertices so that its entity sits at the canonical
//     position (edge (v0,v1) / face (v0,v1,v2)).  The two element-matrix rows are then straight-line DFMAs whose
//     reference-tensor coefficients are immediates / constant-bank operands (generated/canonical_rows.inc),
//     identical for all lanes; structural zeros are gone at compile time.  A 64-bit descriptor per local position
//     maps canonical column entities back to the true ones (slot look-up, orientation fix-up).
//   * contributions are staged in a per-warp shared-memory tile laid out [chunk][lane] (16-byte chunk = the two
//     columns of one column entity): a lane only ever touches its own 16-byte column, so the scatter into
//     data-dependent slots is bank-conflict free by construction.  Faces (<= 2 cells, the 4 shared entities are
//     summed in registers) only store; edges accumulate (read-modify-write) over their cells in ascending order.
//   * the tile is written out cooperatively: 8 entities x 4 consecutive chunks per instruction, i.e. 64-byte
//     contiguous segments (two full sectors) per entity, conflict-free on the shared-memory side.
//   * PEC rows are overwritten with unit rows in the tile before write-out (row-only ident fused).
// Geometry (J, detJ, J^-1 J^-T, J^T J) is recomputed per (entity, cell) from gathered vertex coordinates: ~100
// flops against ~300 DFMAs of tensor contraction, cheaper than staging per-cell geometry through HBM.
#include "sfem_internal.h"

#define SFEM_TABLE static __device__ const
#include "generated/canonical_lut.inc"
#include "generated/canonical_rows.inc"

namespace sfem {
namespace {

struct CanonArgs {
  const double* __restrict__ coords;
  const int32_t* __restrict__ cells;
  const double* __restrict__ eps;
  const double* __restrict__ mu;
  const int64_t* __restrict__ inc_ptr;
  const int32_t* __restrict__ inc_cell;
  const uint8_t* __restrict__ inc_loc;
  const int64_t* __restrict__ ent_rowptr;  // block-relative, indexed by r = g - ent_lo
  const uint32_t* __restrict__ posw;       // slot table, canonical column order, 5 x (2 x uint16) per incidence
  const uint8_t* __restrict__ pec;
  double* __restrict__ S;
  double* __restrict__ T;
  int64_t ent_lo;    // first entity of the pattern block
  int64_t r_begin;   // first block-relative row entity handled by this launch
  int64_t r_end;
  int64_t inc_base;
  int max_chunks;    // 2 * (longest entity row in this launch)
};

// per-(entity, cell) inputs gathered from global memory
struct Item {
  double X[12];             // coordinates of the relabelled vertices
  double eps, mu;
  unsigned long long desc;  // generated/canonical_lut.inc
  unsigned fsel;
  uint32_t w[5];            // slots of the 10 canonical column entities (uint16 pairs)
};

__device__ __forceinline__ int slot_of(const Item& it, int U) { return (it.w[U >> 1] >> (16 * (U & 1))) & 0xffffu; }

__device__ __forceinline__ void gather_item(const CanonArgs& p, const unsigned long long* __restrict__ s_desc,
                                            const unsigned* __restrict__ s_fsel, int64_t i, int64_t c, int t, int4 cv,
                                            Item& it) {
  it.desc = s_desc[t];
  it.fsel = s_fsel[t];
  const int32_t a[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const unsigned s = (unsigned)(it.desc >> (2 * k)) & 3u;
    const uint32_t v = (uint32_t)(s == 0 ? a[0] : (s == 1 ? a[1] : (s == 2 ? a[2] : a[3])));
    const double* x = p.coords + (size_t)(v * 3u);
    it.X[3 * k] = x[0];
    it.X[3 * k + 1] = x[1];
    it.X[3 * k + 2] = x[2];
  }
  it.eps = p.eps ? p.eps[c] : 1.0;
  it.mu = p.mu ? p.mu[c] : 1.0;
  const uint32_t* pw = p.posw + (i - p.inc_base) * 5;
#pragma unroll
  for (int k = 0; k < 5; ++k) it.w[k] = pw[k];
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
    const double s = it.eps / ad;
    g[0] = (A[0][0] * A[0][0] + A[0][1] * A[0][1] + A[0][2] * A[0][2]) * s;
    g[1] = (A[1][0] * A[1][0] + A[1][1] * A[1][1] + A[1][2] * A[1][2]) * s;
    g[2] = (A[2][0] * A[2][0] + A[2][1] * A[2][1] + A[2][2] * A[2][2]) * s;
    g[3] = (A[0][0] * A[1][0] + A[0][1] * A[1][1] + A[0][2] * A[1][2]) * s;
    g[4] = (A[0][0] * A[2][0] + A[0][1] * A[2][1] + A[0][2] * A[2][2]) * s;
    g[5] = (A[1][0] * A[2][0] + A[1][1] * A[2][1] + A[1][2] * A[2][2]) * s;
  } else {
    const double s = 1.0 / (ad * it.mu);
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
__device__ __forceinline__ double2 fix_columns(unsigned long long desc, unsigned facesel, double v0, double v1) {
  if (U < 6) {
    const bool flip = (desc >> (48 + U)) & 1ull;
    return flip ? make_double2(-v1, -v0) : make_double2(v0, v1);
  } else {
    const unsigned nib = (facesel >> (4 * (U - 6))) & 15u;
    const double w = -(v0 + v1);
    return make_double2(sel3(nib & 3u, v0, v1, w), sel3(nib >> 2, v0, v1, w));
  }
}

// cooperative write-out of a warp's tile: 8 entities x 4 consecutive chunks per instruction (64-byte segments per
// entity, conflict-free tile reads); 4 independent load->store pairs in flight; streaming stores (written once).
__device__ __forceinline__ void write_out(const double2* __restrict__ tile, double* __restrict__ out, int64_t a_mine,
                                          int len_mine, int lane) {
  const int gg = lane & 7, cq = lane >> 3;
#pragma unroll 1
  for (int gb = 0; gb < 32; gb += 8) {
    const int64_t a = __shfl_sync(0xffffffffu, a_mine, gb + gg);
    const int nch = 2 * __shfl_sync(0xffffffffu, len_mine, gb + gg);
    int nmax = nch;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
    double2* dst = reinterpret_cast<double2*>(out + 4 * a);
    const double2* src = tile + gb + gg;
#pragma unroll 1
    for (int c0 = cq; c0 < nmax; c0 += 16) {
      double2 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c0 + 4 * j < nch) v[j] = src[(c0 + 4 * j) * 32];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c0 + 4 * j < nch) __stcs(dst + c0 + 4 * j, v[j]);
    }
  }
}

__device__ __forceinline__ void pec_overwrite(const CanonArgs& p, double2* __restrict__ col, int64_t g, int len, int self) {
#pragma unroll
  for (int k = 0; k < 2; ++k)
    if (p.pec[2 * g + k]) {
      for (int s = 0; s < len; ++s) col[(k * len + s) * 32] = make_double2(0.0, 0.0);
      col[(k * len + self) * 32] = k == 0 ? make_double2(1.0, 0.0) : make_double2(0.0, 1.0);
    }
}

constexpr int kThreadsEdges = 64;   // CTAs per SM are bounded by the shared-memory tile (2*max_len*512 B per warp)
constexpr int kThreadsFaces = 128;

// ---- edges: accumulate over the incident cells, software-pipelined gathers ---------------------------------------
// The gather chain of an item is three dependent global loads (incidence -> cell vertices -> coordinates).  With
// only a handful of resident warps (the tile bounds occupancy) it is hidden by issuing it one (coordinates) and two
// (incidence, vertices) items ahead of the item being computed.
template <bool MASS>
__device__ __forceinline__ void edge_pass(const CanonArgs& p, const unsigned long long* __restrict__ s_desc,
                                          const unsigned* __restrict__ s_fsel, double2* __restrict__ tile, int lane,
                                          bool valid, int64_t g, int64_t a, int len, int64_t i0, int64_t i1) {
  double2* col = tile + lane;
  if (valid) {
    for (int cc = 0; cc < 2 * len; ++cc) col[cc * 32] = make_double2(0.0, 0.0);
    Item cur, nxt;
    int64_t c1 = 0;
    int t1 = 0;
    int4 cv1 = make_int4(0, 0, 0, 0);
    {
      const int64_t c0 = p.inc_cell[i0];
      const int t0 = p.inc_loc[i0];
      if (i0 + 1 < i1) {
        c1 = p.inc_cell[i0 + 1];
        t1 = p.inc_loc[i0 + 1];
      }
      const int4 cv0 = *reinterpret_cast<const int4*>(p.cells + c0 * 4);
      if (i0 + 1 < i1) cv1 = *reinterpret_cast<const int4*>(p.cells + c1 * 4);
      gather_item(p, s_desc, s_fsel, i0, c0, t0, cv0, cur);
    }
#pragma unroll 1
    for (int64_t i = i0; i < i1; ++i) {
      int64_t c2 = 0;
      int t2 = 0;
      int4 cv2 = make_int4(0, 0, 0, 0);
      if (i + 2 < i1) {
        c2 = p.inc_cell[i + 2];
        t2 = p.inc_loc[i + 2];
      }
      if (i + 1 < i1) gather_item(p, s_desc, s_fsel, i + 1, c1, t1, cv1, nxt);
      if (i + 2 < i1) cv2 = *reinterpret_cast<const int4*>(p.cells + c2 * 4);
      double gm[6], o[40];
      canon_geometry<MASS>(cur, gm);
      if (MASS) canon_edge_mass(gm, o); else canon_edge_curl(gm, o);
#define SFEM_EDGE_ACC(U)                                                                        \
      {                                                                                         \
        const int slot = slot_of(cur, U);                                                       \
        const double2 r0 = fix_columns<U>(cur.desc, cur.fsel, o[2 * U], o[2 * U + 1]);          \
        const double2 r1 = fix_columns<U>(cur.desc, cur.fsel, o[20 + 2 * U], o[20 + 2 * U + 1]);\
        double2* q0 = col + slot * 32;                                                          \
        double2* q1 = col + (len + slot) * 32;                                                  \
        double2 t0 = *q0, t1v = *q1;                                                            \
        t0.x += r0.x; t0.y += r0.y; t1v.x += r1.x; t1v.y += r1.y;                               \
        *q0 = t0; *q1 = t1v;                                                                    \
      }
      SFEM_EDGE_ACC(0) SFEM_EDGE_ACC(1) SFEM_EDGE_ACC(2) SFEM_EDGE_ACC(3) SFEM_EDGE_ACC(4)
      SFEM_EDGE_ACC(5) SFEM_EDGE_ACC(6) SFEM_EDGE_ACC(7) SFEM_EDGE_ACC(8) SFEM_EDGE_ACC(9)
#undef SFEM_EDGE_ACC
      if (i + 1 == i1 && p.pec) pec_overwrite(p, col, g, len, slot_of(cur, 5));  // canonical entity 5 = the row edge
      cur = nxt;
      c1 = c2; t1 = t2; cv1 = cv2;
    }
  }
  __syncwarp();
  write_out(tile, MASS ? p.T : p.S, a, valid ? len : 0, lane);
  __syncwarp();
}

__device__ __forceinline__ void stage_lut(unsigned long long* s_desc, unsigned* s_fsel) {
  if (threadIdx.x < 10) {
    s_desc[threadIdx.x] = kCanonDesc[threadIdx.x];
    s_fsel[threadIdx.x] = kCanonFaceSel[threadIdx.x];
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreadsEdges) k_canon_edges(CanonArgs p) {
  extern __shared__ double2 smem[];
  __shared__ unsigned long long s_desc[10];
  __shared__ unsigned s_fsel[10];
  stage_lut(s_desc, s_fsel);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double2* tile = smem + (size_t)w * p.max_chunks * 32;
  const int64_t r = p.r_begin + blockIdx.x * (int64_t)kThreadsEdges + threadIdx.x;
  const bool valid = r < p.r_end;
  int64_t g = 0, a = 0, i0 = 0, i1 = 0;
  int len = 0;
  if (valid) {
    g = p.ent_lo + r;
    a = p.ent_rowptr[r];
    len = (int)(p.ent_rowptr[r + 1] - a);
    i0 = p.inc_ptr[g];
    i1 = p.inc_ptr[g + 1];
  }
  edge_pass<false>(p, s_desc, s_fsel, tile, lane, valid, g, a, len, i0, i1);
  edge_pass<true>(p, s_desc, s_fsel, tile, lane, valid, g, a, len, i0, i1);
}

// ---- faces: at most two cells; shared entities summed in registers, tile is store-only --------------------------
template <bool MASS>
__device__ __forceinline__ void face_pass(const CanonArgs& p, double2* __restrict__ tile, int lane, bool valid,
                                          int64_t g, int64_t a, int len, int ncell, const Item& A, const Item& B) {
  double2* col = tile + lane;
  if (valid) {
    double gm[6], oA[40];
#define SFEM_FACE_PUT(U, IT, O)                                                                        \
    {                                                                                                  \
      const int slot = slot_of(IT, U);                                                                 \
      col[slot * 32] = fix_columns<U>(IT.desc, IT.fsel, O[2 * U], O[2 * U + 1]);                       \
      col[(len + slot) * 32] = fix_columns<U>(IT.desc, IT.fsel, O[20 + 2 * U], O[20 + 2 * U + 1]);     \
    }
    if (ncell > 1) {
      double oB[40];
      canon_geometry<MASS>(B, gm);
      if (MASS) canon_face_mass(gm, oB); else canon_face_curl(gm, oB);
      SFEM_FACE_PUT(0, B, oB) SFEM_FACE_PUT(1, B, oB) SFEM_FACE_PUT(3, B, oB)
      SFEM_FACE_PUT(6, B, oB) SFEM_FACE_PUT(7, B, oB) SFEM_FACE_PUT(8, B, oB)
      canon_geometry<MASS>(A, gm);
      if (MASS) canon_face_mass(gm, oA); else canon_face_curl(gm, oA);
      // entities of the shared face (canonical 2,4,5 = its edges, 9 = itself) are seen identically from both cells
#define SFEM_FACE_SHARED(U)                                                    \
      oA[2 * U] += oB[2 * U]; oA[2 * U + 1] += oB[2 * U + 1];                  \
      oA[20 + 2 * U] += oB[20 + 2 * U]; oA[20 + 2 * U + 1] += oB[20 + 2 * U + 1];
      SFEM_FACE_SHARED(2) SFEM_FACE_SHARED(4) SFEM_FACE_SHARED(5) SFEM_FACE_SHARED(9)
#undef SFEM_FACE_SHARED
    } else {
      canon_geometry<MASS>(A, gm);
      if (MASS) canon_face_mass(gm, oA); else canon_face_curl(gm, oA);
    }
    SFEM_FACE_PUT(0, A, oA) SFEM_FACE_PUT(1, A, oA) SFEM_FACE_PUT(2, A, oA) SFEM_FACE_PUT(3, A, oA)
    SFEM_FACE_PUT(4, A, oA) SFEM_FACE_PUT(5, A, oA) SFEM_FACE_PUT(6, A, oA) SFEM_FACE_PUT(7, A, oA)
    SFEM_FACE_PUT(8, A, oA) SFEM_FACE_PUT(9, A, oA)
#undef SFEM_FACE_PUT
    if (p.pec) pec_overwrite(p, col, g, len, slot_of(A, 9));  // canonical entity 9 = the row face
  }
  __syncwarp();
  write_out(tile, MASS ? p.T : p.S, a, valid ? len : 0, lane);
  __syncwarp();
}

__global__ void __launch_bounds__(kThreadsFaces) k_canon_faces(CanonArgs p) {
  extern __shared__ double2 smem[];
  __shared__ unsigned long long s_desc[10];
  __shared__ unsigned s_fsel[10];
  stage_lut(s_desc, s_fsel);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double2* tile = smem + (size_t)w * p.max_chunks * 32;
  const int64_t r = p.r_begin + blockIdx.x * (int64_t)kThreadsFaces + threadIdx.x;
  const bool valid = r < p.r_end;
  int64_t g = 0, a = 0;
  int len = 0, ncell = 0;
  Item A, B;
  if (valid) {
    g = p.ent_lo + r;
    a = p.ent_rowptr[r];
    len = (int)(p.ent_rowptr[r + 1] - a);
    const int64_t i0 = p.inc_ptr[g];
    ncell = (int)(p.inc_ptr[g + 1] - i0);
    // both gather chains are issued before anything is consumed
    const int64_t cA = p.inc_cell[i0];
    const int tA = p.inc_loc[i0];
    const int64_t cB = ncell > 1 ? p.inc_cell[i0 + 1] : cA;
    const int tB = ncell > 1 ? p.inc_loc[i0 + 1] : tA;
    const int4 cvA = *reinterpret_cast<const int4*>(p.cells + cA * 4);
    const int4 cvB = *reinterpret_cast<const int4*>(p.cells + cB * 4);
    gather_item(p, s_desc, s_fsel, i0, cA, tA, cvA, A);
    gather_item(p, s_desc, s_fsel, ncell > 1 ? i0 + 1 : i0, cB, tB, cvB, B);
  }
  face_pass<false>(p, tile, lane, valid, g, a, len, ncell, A, B);
  face_pass<true>(p, tile, lane, valid, g, a, len, ncell, A, B);
}

}  // namespace

bool canonical_supported(const sfem_pattern_s* pat) {
  return pat->space->mesh->gdim == 3 && pat->space->order == 2;
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
  p.coords = m->coords; p.cells = m->cells; p.eps = eps; p.mu = mu;
  p.inc_ptr = sp->inc_ptr; p.inc_cell = sp->inc_cell; p.inc_loc = sp->inc_loc;
  p.ent_rowptr = pat->ent_rowptr; p.posw = reinterpret_cast<const uint32_t*>(pat->pos); p.pec = pec; p.S = S; p.T = T;
  p.ent_lo = pat->ent_lo; p.inc_base = pat->inc_base;
  // faces meet at most two cells (a face shared by more is a broken mesh and is refused when the space is built)
  const int64_t E = sp->E;
  const int64_t n_edge_rows = std::max<int64_t>(0, std::min(pat->ent_hi, E) - pat->ent_lo);
  const int64_t n_rows = pat->n_rows_ent;
  if (n_edge_rows > 0) {
    p.r_begin = 0; p.r_end = n_edge_rows;
    p.max_chunks = 2 * pat->max_row_len_edges;
    const size_t smem = (size_t)(kThreadsEdges / 32) * p.max_chunks * 32 * sizeof(double2);
    SFEM_CHECK(smem <= 227 * 1024, "edge valence too high for the canonical kernel's shared-memory tile; use the generic kernel");
    SFEM_CUDA(cudaFuncSetAttribute(k_canon_edges, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((n_edge_rows + kThreadsEdges - 1) / kThreadsEdges);
    k_canon_edges<<<grid, kThreadsEdges, smem, st>>>(p);
  }
  if (n_rows > n_edge_rows) {
    p.r_begin = n_edge_rows; p.r_end = n_rows;
    p.max_chunks = 32;  // 2 rows x at most 16 column entities
    SFEM_CHECK(pat->max_row_len_other <= 16, "a face is shared by more than two cells: not a valid tetrahedral mesh");
    const size_t smem = (size_t)(kThreadsFaces / 32) * p.max_chunks * 32 * sizeof(double2);
    SFEM_CUDA(cudaFuncSetAttribute(k_canon_faces, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((n_rows - n_edge_rows + kThreadsFaces - 1) / kThreadsFaces);
    k_canon_faces<<<grid, kThreadsFaces, smem, st>>>(p);
  }
  SFEM_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sfem
