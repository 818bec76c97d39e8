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
  const uint16_t* __restrict__ pos;
  const uint8_t* __restrict__ pec;
  double* __restrict__ S;
  double* __restrict__ T;
  int64_t ent_lo;    // first entity of the pattern block
  int64_t r_begin;   // first block-relative row entity handled by this launch
  int64_t r_end;
  int64_t inc_base;
  int max_chunks;    // 2 * (longest entity row in this launch)
};

// geometry factors of the relabelled cell (vertex ids v[0..3]) in symmetric order (00,11,22,01,02,12):
//   MASS: |detJ| eps (J^-1 J^-T)      else: (J^T J) / (|detJ| mu)
template <bool MASS>
__device__ __forceinline__ void canon_geometry(const double* __restrict__ coords, const int32_t (&v)[4], double coef,
                                               double (&g)[6]) {
  double x0[3], J[3][3];
#pragma unroll
  for (int d = 0; d < 3; ++d) x0[d] = coords[(int64_t)v[0] * 3 + d];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int d = 0; d < 3; ++d) J[d][a] = coords[(int64_t)v[a + 1] * 3 + d] - x0[d];
  const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
  const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
  const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
  const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
  const double ad = fabs(det);
  if (MASS) {
    // adj(J)[a][d] = det * Jinv[a][d];  Ginv = adj adj^T / det^2 ; times |det| eps -> adj adj^T * eps/|det|
    double A[3][3];
    A[0][0] = c00; A[1][0] = c01; A[2][0] = c02;
    A[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
    A[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
    A[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
    A[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
    A[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
    A[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double s = coef / ad;
    g[0] = (A[0][0] * A[0][0] + A[0][1] * A[0][1] + A[0][2] * A[0][2]) * s;
    g[1] = (A[1][0] * A[1][0] + A[1][1] * A[1][1] + A[1][2] * A[1][2]) * s;
    g[2] = (A[2][0] * A[2][0] + A[2][1] * A[2][1] + A[2][2] * A[2][2]) * s;
    g[3] = (A[0][0] * A[1][0] + A[0][1] * A[1][1] + A[0][2] * A[1][2]) * s;
    g[4] = (A[0][0] * A[2][0] + A[0][1] * A[2][1] + A[0][2] * A[2][2]) * s;
    g[5] = (A[1][0] * A[2][0] + A[1][1] * A[2][1] + A[1][2] * A[2][2]) * s;
  } else {
    const double s = 1.0 / (ad * coef);
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

__device__ __forceinline__ void load_relabelled(const int32_t* __restrict__ cells, int64_t c, unsigned long long desc,
                                                int32_t (&v)[4]) {
  const int4 cv = *reinterpret_cast<const int4*>(cells + c * 4);
  const int32_t a[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const unsigned s = (unsigned)(desc >> (2 * i)) & 3u;
    v[i] = s == 0 ? a[0] : (s == 1 ? a[1] : (s == 2 ? a[2] : a[3]));
  }
}

// cooperative write-out of a warp's tile: 8 entities x 4 consecutive chunks per instruction
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
    for (int cc = cq; cc < nmax; cc += 4)
      if (cc < nch) dst[cc] = tile[cc * 32 + gb + gg];
  }
}

constexpr int kThreadsEdges = 64;   // CTAs per SM are bounded by the shared-memory tile (2*max_len*512 B per warp)
constexpr int kThreadsFaces = 128;

// ---- edges: accumulate over the incident cells ----------------------------------------------------------------
template <bool MASS>
__device__ __forceinline__ void edge_pass(const CanonArgs& p, double2* __restrict__ tile, int lane, bool valid,
                                          int64_t g, int64_t a, int len, int64_t i0, int64_t i1) {
  if (valid)
    for (int cc = 0; cc < 2 * len; ++cc) tile[cc * 32 + lane] = make_double2(0.0, 0.0);
  if (valid) {
#pragma unroll 1
    for (int64_t i = i0; i < i1; ++i) {
      const int64_t c = p.inc_cell[i];
      const int t = p.inc_loc[i];
      const unsigned long long desc = kCanonDesc_edge[t];
      const unsigned facesel = kCanonFaceSel_edge[t];
      int32_t v[4];
      load_relabelled(p.cells, c, desc, v);
      double gm[6], o[40];
      const double coef = MASS ? (p.eps ? p.eps[c] : 1.0) : (p.mu ? p.mu[c] : 1.0);
      canon_geometry<MASS>(p.coords, v, coef, gm);
      if (MASS) canon_edge_mass(gm, o); else canon_edge_curl(gm, o);
      const uint16_t* ps = p.pos + (i - p.inc_base) * 10;
#define SFEM_EDGE_ACC(U)                                                                   \
      {                                                                                    \
        const int slot = ps[(unsigned)(desc >> (8 + 4 * U)) & 15u];                        \
        const double2 r0 = fix_columns<U>(desc, facesel, o[2 * U], o[2 * U + 1]);          \
        const double2 r1 = fix_columns<U>(desc, facesel, o[20 + 2 * U], o[20 + 2 * U + 1]);\
        double2* q0 = tile + slot * 32 + lane;                                             \
        double2* q1 = tile + (len + slot) * 32 + lane;                                     \
        double2 t0 = *q0, t1 = *q1;                                                        \
        t0.x += r0.x; t0.y += r0.y; t1.x += r1.x; t1.y += r1.y;                            \
        *q0 = t0; *q1 = t1;                                                                \
      }
      SFEM_EDGE_ACC(0) SFEM_EDGE_ACC(1) SFEM_EDGE_ACC(2) SFEM_EDGE_ACC(3) SFEM_EDGE_ACC(4)
      SFEM_EDGE_ACC(5) SFEM_EDGE_ACC(6) SFEM_EDGE_ACC(7) SFEM_EDGE_ACC(8) SFEM_EDGE_ACC(9)
#undef SFEM_EDGE_ACC
    }
    if (p.pec) {
      // unit rows: the row entity's own slot = slot of canonical entity 5 (= itself) in any incident cell
      const int self = p.pos[(i0 - p.inc_base) * 10 + p.inc_loc[i0]];
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (p.pec[2 * g + k]) {
          for (int s = 0; s < len; ++s) tile[(k * len + s) * 32 + lane] = make_double2(0.0, 0.0);
          tile[(k * len + self) * 32 + lane] = k == 0 ? make_double2(1.0, 0.0) : make_double2(0.0, 1.0);
        }
    }
  }
  __syncwarp();
  write_out(tile, MASS ? p.T : p.S, a, valid ? len : 0, lane);
  __syncwarp();
}

__global__ void __launch_bounds__(kThreadsEdges) k_canon_edges(CanonArgs p) {
  extern __shared__ double2 smem[];
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
  edge_pass<false>(p, tile, lane, valid, g, a, len, i0, i1);
  edge_pass<true>(p, tile, lane, valid, g, a, len, i0, i1);
}

// ---- faces: at most two cells; shared entities summed in registers, tile is store-only --------------------------
template <bool MASS>
__device__ __forceinline__ void face_pass(const CanonArgs& p, double2* __restrict__ tile, int lane, bool valid,
                                          int64_t g, int64_t a, int len, int64_t i0, int ncell) {
  if (valid) {
    const int64_t cA = p.inc_cell[i0];
    const int tA = p.inc_loc[i0] - 6;
    const unsigned long long dA = kCanonDesc_face[tA];
    const unsigned fA = kCanonFaceSel_face[tA];
    const uint16_t* psA = p.pos + (i0 - p.inc_base) * 10;
    int32_t v[4];
    double gm[6], oA[40];
    load_relabelled(p.cells, cA, dA, v);
    canon_geometry<MASS>(p.coords, v, MASS ? (p.eps ? p.eps[cA] : 1.0) : (p.mu ? p.mu[cA] : 1.0), gm);
    if (MASS) canon_face_mass(gm, oA); else canon_face_curl(gm, oA);
    if (ncell > 1) {
      const int64_t cB = p.inc_cell[i0 + 1];
      const int tB = p.inc_loc[i0 + 1] - 6;
      const unsigned long long dB = kCanonDesc_face[tB];
      const unsigned fB = kCanonFaceSel_face[tB];
      const uint16_t* psB = psA + 10;
      double oB[40];
      load_relabelled(p.cells, cB, dB, v);
      canon_geometry<MASS>(p.coords, v, MASS ? (p.eps ? p.eps[cB] : 1.0) : (p.mu ? p.mu[cB] : 1.0), gm);
      if (MASS) canon_face_mass(gm, oB); else canon_face_curl(gm, oB);
      // entities of the shared face (canonical 2,4,5 = its edges, 9 = itself) are seen identically from both cells
#define SFEM_FACE_SHARED(U)                                                    \
      oA[2 * U] += oB[2 * U]; oA[2 * U + 1] += oB[2 * U + 1];                  \
      oA[20 + 2 * U] += oB[20 + 2 * U]; oA[20 + 2 * U + 1] += oB[20 + 2 * U + 1];
      SFEM_FACE_SHARED(2) SFEM_FACE_SHARED(4) SFEM_FACE_SHARED(5) SFEM_FACE_SHARED(9)
#undef SFEM_FACE_SHARED
#define SFEM_FACE_PUT(U, DESC, FSEL, PS, O)                                                           \
      {                                                                                               \
        const int slot = PS[(unsigned)(DESC >> (8 + 4 * U)) & 15u];                                   \
        tile[slot * 32 + lane] = fix_columns<U>(DESC, FSEL, O[2 * U], O[2 * U + 1]);                  \
        tile[(len + slot) * 32 + lane] = fix_columns<U>(DESC, FSEL, O[20 + 2 * U], O[20 + 2 * U + 1]);\
      }
      SFEM_FACE_PUT(0, dB, fB, psB, oB) SFEM_FACE_PUT(1, dB, fB, psB, oB) SFEM_FACE_PUT(3, dB, fB, psB, oB)
      SFEM_FACE_PUT(6, dB, fB, psB, oB) SFEM_FACE_PUT(7, dB, fB, psB, oB) SFEM_FACE_PUT(8, dB, fB, psB, oB)
    }
    SFEM_FACE_PUT(0, dA, fA, psA, oA) SFEM_FACE_PUT(1, dA, fA, psA, oA) SFEM_FACE_PUT(2, dA, fA, psA, oA)
    SFEM_FACE_PUT(3, dA, fA, psA, oA) SFEM_FACE_PUT(4, dA, fA, psA, oA) SFEM_FACE_PUT(5, dA, fA, psA, oA)
    SFEM_FACE_PUT(6, dA, fA, psA, oA) SFEM_FACE_PUT(7, dA, fA, psA, oA) SFEM_FACE_PUT(8, dA, fA, psA, oA)
    SFEM_FACE_PUT(9, dA, fA, psA, oA)
#undef SFEM_FACE_PUT
    if (p.pec) {
      const int self = psA[p.inc_loc[i0]];
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (p.pec[2 * g + k]) {
          for (int s = 0; s < len; ++s) tile[(k * len + s) * 32 + lane] = make_double2(0.0, 0.0);
          tile[(k * len + self) * 32 + lane] = k == 0 ? make_double2(1.0, 0.0) : make_double2(0.0, 1.0);
        }
    }
  }
  __syncwarp();
  write_out(tile, MASS ? p.T : p.S, a, valid ? len : 0, lane);
  __syncwarp();
}

__global__ void __launch_bounds__(kThreadsFaces) k_canon_faces(CanonArgs p) {
  extern __shared__ double2 smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double2* tile = smem + (size_t)w * p.max_chunks * 32;
  const int64_t r = p.r_begin + blockIdx.x * (int64_t)kThreadsFaces + threadIdx.x;
  const bool valid = r < p.r_end;
  int64_t g = 0, a = 0, i0 = 0;
  int len = 0, ncell = 0;
  if (valid) {
    g = p.ent_lo + r;
    a = p.ent_rowptr[r];
    len = (int)(p.ent_rowptr[r + 1] - a);
    i0 = p.inc_ptr[g];
    ncell = (int)(p.inc_ptr[g + 1] - i0);
  }
  face_pass<false>(p, tile, lane, valid, g, a, len, i0, ncell);
  face_pass<true>(p, tile, lane, valid, g, a, len, i0, ncell);
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
  p.ent_rowptr = pat->ent_rowptr; p.pos = pat->pos; p.pec = pec; p.S = S; p.T = T;
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
