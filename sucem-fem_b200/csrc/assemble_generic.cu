// K3 (generic variant): row-owner gather of the S / M(=T) / S_0 values into the precomputed CSR slots.
// Replaces the cell loop of dolfin.assemble (ufc tabulate_tensor + GenericMatrix::add) and
// DirichletBC.apply -> ident called from /root/reference/sucemfem/SystemMatrices.py:54-55 for the forms of
// /root/reference/sucemfem/Forms.py:76,84 and /root/reference/sucemfem/BoundaryConditions/ABC.py:57.
//
// One warp owns one row entity (an edge or a face / cell interior: its q rows share a pattern).  Lanes own
// output columns; for each owned column the lane walks the entity's incident cells in ascending order
// (DOLFIN's summation order), finds the cell-local column through the slot table and contracts the cell's
// geometry factors with the full pre-integrated reference tensor.  Every CSR value is written exactly once,
// coalesced, without atomics: results are bit-reproducible.  Works for any dim/order and any local position;
// it is the cross-check for the canonical-frame kernels in assemble_canonical.cu.
//
// Also here: the thread-per-row-entity kernels (SFEM_KERNEL_ROWS) -- k_values_tri (triangles of order 1 / 2, the 2-D fast
// path) and k_values_tet1 (order-1 tetrahedra the canonical kernel refuses) -- and the S_0 row selection.
#include <cub/cub.cuh>

#include "sfem_internal.h"

#define SFEM_TABLE static __device__ const
#include "generated/n1curl_tables.inc"
#include "generated/canonical_lut.inc"

namespace sfem {
namespace {

template <int DIM, int ORDER>
struct Elem;
template <>
struct Elem<3, 1> {
  static constexpr int NP = 6;
  __device__ static const double* mass(int i, int j) { return kMassTet1[i][j]; }
  __device__ static const double* curl(int i, int j) { return kCurlTet1[i][j]; }
  __device__ static const double* facet(int f, int i, int j) { return kFacetTet1[f][i][j]; }
};
template <>
struct Elem<3, 2> {
  static constexpr int NP = 6;
  __device__ static const double* mass(int i, int j) { return kMassTet2[i][j]; }
  __device__ static const double* curl(int i, int j) { return kCurlTet2[i][j]; }
  __device__ static const double* facet(int f, int i, int j) { return kFacetTet2[f][i][j]; }
};
template <>
struct Elem<2, 1> {
  static constexpr int NP = 3;
  __device__ static const double* mass(int i, int j) { return kMassTri1[i][j]; }
  __device__ static const double* curl(int i, int j) { return &kCurlTri1[i][j]; }
  __device__ static const double* facet(int f, int i, int j) { return kFacetTri1[f][i][j]; }
};
template <>
struct Elem<2, 2> {
  static constexpr int NP = 3;
  __device__ static const double* mass(int i, int j) { return kMassTri2[i][j]; }
  __device__ static const double* curl(int i, int j) { return &kCurlTri2[i][j]; }
  __device__ static const double* facet(int f, int i, int j) { return kFacetTri2[f][i][j]; }
};

// Per-cell geometry factors, symmetric index order (00,11,22,01,02,12) / (00,11,01):
//   gT = |detJ| eps_r (J^-1 J^-T)       gS = (J^T J) / (|detJ| mu_r)   (2-D: scalar 1/(|detJ| mu_r))
// x = x_v0 + J X, columns of J = x_vi - x_v0, detJ may be negative.
template <int DIM>
__device__ __forceinline__ void cell_geometry(const double* __restrict__ coords, const int32_t* __restrict__ cell,
                                              double eps, double mu, double* gT, double* gS, double* Jinv_out) {
  double x[DIM + 1][DIM];
#pragma unroll
  for (int v = 0; v <= DIM; ++v)
#pragma unroll
    for (int d = 0; d < DIM; ++d) x[v][d] = coords[(int64_t)cell[v] * DIM + d];
  double J[DIM][DIM];  // J[d][a] = x_{a+1}[d] - x_0[d]
#pragma unroll
  for (int d = 0; d < DIM; ++d)
#pragma unroll
    for (int a = 0; a < DIM; ++a) J[d][a] = x[a + 1][d] - x[0][d];
  double Ji[DIM][DIM];  // inverse: Ji[a][d]
  double det;
  if (DIM == 3) {
    double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    double id = 1.0 / det;
    Ji[0][0] = c00 * id;
    Ji[1][0] = c01 * id;
    Ji[2][0] = c02 * id;
    Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
    Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
    Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
    Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
  } else {
    det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    double id = 1.0 / det;
    Ji[0][0] = J[1][1] * id;
    Ji[0][1] = -J[0][1] * id;
    Ji[1][0] = -J[1][0] * id;
    Ji[1][1] = J[0][0] * id;
  }
  const double ad = fabs(det);
  const int pa[6] = {0, 1, 2, 0, 0, 1}, pb[6] = {0, 1, 2, 1, 2, 2};
  const int qa[3] = {0, 1, 0}, qb[3] = {0, 1, 1};
  constexpr int NP = DIM == 3 ? 6 : 3;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const int a = DIM == 3 ? pa[p] : qa[p], b = DIM == 3 ? pb[p] : qb[p];
    double gi = 0, gg = 0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      gi += Ji[a][d] * Ji[b][d];
      gg += J[d][a] * J[d][b];
    }
    if (gT) gT[p] = gi * ad * eps;
    if (gS && DIM == 3) gS[p] = gg / (ad * mu);
  }
  if (gS && DIM == 2) gS[0] = 1.0 / (ad * mu);
  if (Jinv_out)
#pragma unroll
    for (int a = 0; a < DIM; ++a)
#pragma unroll
      for (int d = 0; d < DIM; ++d) Jinv_out[a * DIM + d] = Ji[a][d];
}

template <int DIM>
__global__ void k_cell_geometry(const double* __restrict__ coords, const int32_t* __restrict__ cells,
                                const double* __restrict__ eps, const double* __restrict__ mu, int64_t C,
                                double* __restrict__ geo) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= C) return;
  double gT[6], gS[6];
  cell_geometry<DIM>(coords, cells + c * (DIM + 1), eps ? eps[c] : 1.0, mu ? mu[c] : 1.0, gT, gS, nullptr);
  constexpr int NP = DIM == 3 ? 6 : 3;
#pragma unroll
  for (int p = 0; p < NP; ++p) geo[c * 12 + p] = gT[p];
#pragma unroll
  for (int p = 0; p < (DIM == 3 ? 6 : 1); ++p) geo[c * 12 + 6 + p] = gS[p];
}

// ABC facet factors for local facet lf of a cell: K_ab = meas * (dim-1)! * ((J^-1 J^-T)_ab - nu_a nu_b),
// nu = J^-1 n  (so that sum_ab K_ab Bhat^{ab}_ij = int_f (n x phi_i).(n x phi_j) ds).
template <int DIM>
__device__ void facet_geometry(const double* __restrict__ coords, const int32_t* __restrict__ cell, int lf,
                               double* K) {
  double gT[6], Ji[DIM * DIM];
  cell_geometry<DIM>(coords, cell, 1.0, 1.0, gT, nullptr, Ji);
  // undo the |detJ| factor folded into gT: recompute Ginv from Ji
  double n[DIM], meas;
  if (DIM == 3) {
    const double* p0 = coords + (int64_t)cell[tet_face_vertex(lf, 0)] * 3;
    const double* p1 = coords + (int64_t)cell[tet_face_vertex(lf, 1)] * 3;
    const double* p2 = coords + (int64_t)cell[tet_face_vertex(lf, 2)] * 3;
    double u[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
    double v[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
    n[0] = u[1] * v[2] - u[2] * v[1];
    n[1] = u[2] * v[0] - u[0] * v[2];
    n[2 % DIM] = u[0] * v[1] - u[1] * v[0];
    meas = sqrt(n[0] * n[0] + n[1] * n[1] + n[2 % DIM] * n[2 % DIM]);
  } else {
    const double* p0 = coords + (int64_t)cell[tri_edge_vertex(lf, 0)] * 2;
    const double* p1 = coords + (int64_t)cell[tri_edge_vertex(lf, 1)] * 2;
    double t[2] = {p1[0] - p0[0], p1[1] - p0[1]};
    n[0] = t[1];
    n[1] = -t[0];
    meas = sqrt(n[0] * n[0] + n[1] * n[1]);
  }
  double nu[DIM];
#pragma unroll
  for (int a = 0; a < DIM; ++a) {
    double s = 0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) s += Ji[a * DIM + d] * n[d];
    nu[a] = s / meas;
  }
  const int pa[6] = {0, 1, 2, 0, 0, 1}, pb[6] = {0, 1, 2, 1, 2, 2};
  const int qa[3] = {0, 1, 0}, qb[3] = {0, 1, 1};
  constexpr int NP = DIM == 3 ? 6 : 3;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const int a = DIM == 3 ? pa[p] : qa[p], b = DIM == 3 ? pb[p] : qb[p];
    double gi = 0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) gi += Ji[a * DIM + d] * Ji[b * DIM + d];
    K[p] = (gi - nu[a] * nu[b]) * meas;
  }
}

struct PatView {
  const int64_t* inc_ptr;
  const int32_t* inc_cell;
  const uint8_t* inc_loc;
  const int64_t* ent_rowptr;
  const int32_t* ent_cols;
  const uint16_t* pos;
  int64_t ent_lo, n_rows, inc_base;
  int64_t gshift;  // global id of row entity = ent_lo + r + gshift
  int nec;
  bool canon;  // slot table in canonical column order
  const int64_t* row_list;  // optional: only these block-relative row entities are assembled (ABC: boundary rows)
  int64_t n_list;
  const int64_t* out_rowptr;  // optional (with row_list): entity-level output offset of list entry i (compact S_0)
};

constexpr int kWarps = 8;

// FORM: 0 = M only, 1 = S only, 3 = S and M fused, 2 = ABC
template <int DIM, int ORDER, int FORM>
__global__ void __launch_bounds__(kWarps * 32)
k_values_generic(PatView pv, const double* __restrict__ geo, const double* __restrict__ coords,
                 const int32_t* __restrict__ cells, const int32_t* __restrict__ cell_facets,
                 const uint8_t* __restrict__ cell_extmask, const uint8_t* __restrict__ facet_mask,
                 const uint8_t* __restrict__ pec, double* __restrict__ outA, double* __restrict__ outB) {
  using El = Elem<DIM, ORDER>;
  constexpr int Q = ORDER, NP = El::NP, NFAC = DIM + 1;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t r = blockIdx.x * (int64_t)kWarps + w;
  const int64_t li = r;
  if (pv.row_list) {
    if (r >= pv.n_list) return;
    r = pv.row_list[r];
  }
  if (r >= pv.n_rows) return;
  const int64_t g = pv.ent_lo + r;
  const int64_t a = pv.ent_rowptr[r];
  const int len = (int)(pv.ent_rowptr[r + 1] - a);
  const int64_t a_out = pv.out_rowptr ? pv.out_rowptr[li] : a;
  const int64_t i0 = pv.inc_ptr[g], i1 = pv.inc_ptr[g + 1];
  for (int k = 0; k < Q; ++k) {
    const int64_t rs = (int64_t)Q * Q * a_out + (int64_t)k * Q * len;
    const bool pecrow = pec && pec[Q * g + k];
    for (int idx = lane; idx < Q * len; idx += 32) {
      const int s = idx / Q, kp = idx - s * Q;
      double accA = 0.0, accB = 0.0;
      if (pecrow) {
        accA = accB = (pv.ent_cols[a + s] == g + pv.gshift && kp == k) ? 1.0 : 0.0;
      } else {
        for (int64_t i = i0; i < i1; ++i) {
          const uint16_t* ps = pv.pos + (i - pv.inc_base) * pv.nec;
          int u = -1;
          for (int uu = 0; uu < pv.nec; ++uu)
            if (ps[uu] == s) u = uu;
          if (u < 0) continue;
          if (pv.canon) u = (int)((kCanonDesc[pv.inc_loc[i]] >> (8 + 4 * u)) & 15ull);
          const int64_t c = pv.inc_cell[i];
          const int li = pv.inc_loc[i] * Q + k, lj = u * Q + kp;
          if (FORM == 2) {
            unsigned m = cell_extmask[c];
            if (!m) continue;
            for (int lf = 0; lf < NFAC; ++lf) {
              if (!((m >> lf) & 1)) continue;
              if (facet_mask && !facet_mask[cell_facets[c * NFAC + lf]]) continue;
              double K[6];
              facet_geometry<DIM>(coords, cells + c * (DIM + 1), lf, K);
              const double* B = El::facet(lf, li, lj);
#pragma unroll
              for (int p = 0; p < NP; ++p) accA += K[p] * B[p];
            }
          } else {
            const double* gc = geo + c * 12;
            if (FORM == 0 || FORM == 3) {
              const double* Mh = El::mass(li, lj);
              double t = 0;
#pragma unroll
              for (int p = 0; p < NP; ++p) t += gc[p] * Mh[p];
              if (FORM == 3) accB += t; else accA += t;
            }
            if (FORM == 1 || FORM == 3) {
              const double* Ch = El::curl(li, lj);
              double t = 0;
              if (DIM == 3) {
#pragma unroll
                for (int p = 0; p < 6; ++p) t += gc[6 + p] * Ch[p];
              } else {
                t = gc[6] * Ch[0];
              }
              accA += t;
            }
          }
        }
      }
      outA[rs + idx] = accA;
      if (FORM == 3) outB[rs + idx] = accB;
    }
  }
}

// ---- K4: the ABC matrix S_0, rows of the entities ON the boundary ---------------------------------------------------
// One warp per listed row entity.  Only the dofs in the closure of an exterior facet have a tangential trace on it, so a
// row of S_0 is non-zero only at the closure entities of the (usually two) exterior facets that contain the row entity:
// the warp clears the entity's rows (unit rows for Dirichlet dofs), then visits the incident cells in ascending order and
// their exterior facets in ascending local order -- the generic kernel's summation order, so both give the same bits --
// and for each facet that contains the row entity computes the facet geometry ONCE and lets one lane per (closure
// entity, k, k') add its 6-term contraction to the slot.  The generic kernel (kept as SFEM_KERNEL_GENERIC cross-check)
// recomputes the geometry for every (column, cell, facet) of the full row.
template <int DIM, int ORDER>
__global__ void __launch_bounds__(kWarps * 32)
k_abc_rows(PatView pv, const double* __restrict__ coords, const int32_t* __restrict__ cells,
           const int32_t* __restrict__ cell_facets, const uint8_t* __restrict__ cell_extmask,
           const uint8_t* __restrict__ facet_mask, const uint8_t* __restrict__ pec, double* __restrict__ out) {
  using El = Elem<DIM, ORDER>;
  constexpr int Q = ORDER, NP = El::NP, NFAC = DIM + 1;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t li = blockIdx.x * (int64_t)kWarps + w;
  if (li >= pv.n_list) return;
  const int64_t r = pv.row_list[li];
  const int64_t g = pv.ent_lo + r;
  const int64_t a = pv.ent_rowptr[r];
  const int len = (int)(pv.ent_rowptr[r + 1] - a);
  const int64_t a_out = pv.out_rowptr ? pv.out_rowptr[li] : a;
  const int64_t rs0 = (int64_t)Q * Q * a_out;
  bool pecrow[Q];
#pragma unroll
  for (int k = 0; k < Q; ++k) pecrow[k] = pec && pec[Q * g + k];
  for (int k = 0; k < Q; ++k)
    for (int idx = lane; idx < Q * len; idx += 32) {
      const int s = idx / Q, kp = idx - s * Q;
      out[rs0 + (int64_t)k * Q * len + idx] = (pecrow[k] && pv.ent_cols[a + s] == g + pv.gshift && kp == k) ? 1.0 : 0.0;
    }
  __syncwarp();
  // the lanes look at the incident cells in PARALLEL (one dependent load chain for all of them instead of one per cell);
  // the cells that do have an exterior facet containing the row entity are then visited in ascending order
  const int64_t i_begin = pv.inc_ptr[g], i_end = pv.inc_ptr[g + 1];
  for (int64_t ib = i_begin; ib < i_end; ib += 32) {
    const int64_t il = ib + lane;
    int64_t c_l = 0;
    int loc_l = 0;
    unsigned hit_l = 0u;  // exterior local facets of this lane's cell that contain the row entity
    if (il < i_end) {
      c_l = pv.inc_cell[il];
      loc_l = pv.inc_loc[il];
      const unsigned m = cell_extmask[c_l];
      unsigned on;  // local facets that contain the row entity (k_abc_row_flags)
      if (DIM == 3) on = loc_l < 6 ? ((0xCA6953u >> (4 * loc_l)) & 15u) : (1u << (loc_l - 6));
      else on = loc_l < 3 ? (1u << loc_l) : 0u;
      hit_l = m & on;
    }
    unsigned todo = __ballot_sync(0xffffffffu, hit_l != 0u);
    while (todo) {
      const int src = __ffs((int)todo) - 1;
      todo &= todo - 1;
      const int64_t c = __shfl_sync(0xffffffffu, c_l, src);
      const int loc = __shfl_sync(0xffffffffu, loc_l, src);
      const unsigned hit = __shfl_sync(0xffffffffu, hit_l, src);
      const int64_t i = ib + src;
      const uint16_t* ps = pv.pos + (i - pv.inc_base) * pv.nec;
      for (int lf = 0; lf < NFAC; ++lf) {
        if (!((hit >> lf) & 1u)) continue;
        if (facet_mask && !facet_mask[cell_facets[c * NFAC + lf]]) continue;
        double K[6];
        facet_geometry<DIM>(coords, cells + c * (DIM + 1), lf, K);
        // closure of the facet: tets -- its three edges (+ the face entity at order 2); triangles -- the edge itself
        constexpr int NCL = DIM == 3 ? (ORDER == 2 ? 4 : 3) : 1;
        if (lane < NCL * Q * Q) {
          const int ci = lane / (Q * Q), k = (lane / Q) % Q, kp = lane % Q;
          const int u = DIM == 3 ? (ci < 3 ? tet_face_edge(lf, ci) : 6 + lf) : lf;
          // slot of local entity u of this cell in the row (the slot table of tets is in canonical column order)
          int s = -1;
          for (int uu = 0; uu < pv.nec; ++uu) {
            const int ut = pv.canon ? (int)((kCanonDesc[loc] >> (8 + 4 * uu)) & 15ull) : uu;
            if (ut == u) s = ps[uu];
          }
          if (s >= 0 && !pecrow[k]) {
            const double* B = El::facet(lf, loc * Q + k, u * Q + kp);
            double* o = out + rs0 + (int64_t)k * Q * len + (int64_t)Q * s + kp;
            double acc = *o;
#pragma unroll
            for (int p = 0; p < NP; ++p) acc += K[p] * B[p];
            *o = acc;
          }
        }
        __syncwarp();  // the next facet may add to the same slot (the row entity's own block)
      }
    }
  }
}

// ---- triangles, S and T in one sweep: one THREAD per row entity -------------------------------------------------
// Two triangles that share an edge share nothing else that carries a dof, so the q rows of an entity are
// [the entity itself] + [the other entities of each incident cell]: every CSR chunk has exactly one contributing cell
// except the diagonal block, which is summed in registers over the incident cells in ascending order (DOLFIN's order).
// No atomics, bit-reproducible.  The 32 consecutive entities of a warp own one contiguous span of S and of T; the
// lanes assemble it in a shared-memory tile in its FINAL layout and the warp then streams the tile out with fully
// coalesced stores (a thread writing its own rows straight to global memory issues 32 partial sectors per store
// instruction: 0.83 TB/s at order 2 against the tile's figure in DESIGN.md).  A warp whose span does not fit the tile
// (non-manifold edges) writes directly.  Geometry is recomputed per (entity, cell) from the gathered vertices (2 x 2
// Jacobian: ~25 flops); the reference tensors sit in shared memory.  PEC rows are written as unit rows (ident fused).
// the q rows of one entity into dS / dT (a shared-memory tile or the global arrays), rs = offset of its row 0
template <int ORDER>
__device__ __forceinline__ void tri_entity_rows(const PatView& pv, const double* __restrict__ sM,
                                                const double* __restrict__ sC, const double* __restrict__ coords,
                                                const int32_t* __restrict__ cells, const double* __restrict__ eps,
                                                const double* __restrict__ mu, const uint8_t* __restrict__ pec,
                                                int64_t g, int64_t rs, int len, double* __restrict__ dS,
                                                double* __restrict__ dT) {
  constexpr int Q = ORDER, NEC = ORDER == 1 ? 3 : 4, ND = Q * NEC;
  // the tensors are stored per local row entity with a padded stride (kTriPad): lanes of a warp differ in `le` only, and the
  // unpadded strides of order 2 (384 B / 128 B) put every le on the same banks (3-way conflicts on every operand read)
  constexpr int SM_STRIDE = Q * ND * 3 + (ORDER == 2 ? 1 : 0), SC_STRIDE = Q * ND + (ORDER == 2 ? 1 : 0);
  const int64_t i0 = pv.inc_ptr[g], i1 = pv.inc_ptr[g + 1];
  bool pecrow[Q];
#pragma unroll
  for (int k = 0; k < Q; ++k) pecrow[k] = pec && pec[Q * g + k];
  double selfS[Q][Q], selfT[Q][Q];
#pragma unroll
  for (int k = 0; k < Q; ++k)
#pragma unroll
    for (int kp = 0; kp < Q; ++kp) selfS[k][kp] = selfT[k][kp] = 0.0;
  int self_slot = 0;
  // An entity has at most two incident cells on a manifold mesh: at order 2 everything the two cells need (incidence, slots,
  // vertex ids, coordinates, coefficients) is loaded for BOTH before either is used, so the two gather chains overlap.
  struct CellIn {
    int le;
    int slot[NEC];
    double x[3][2];
    double ceps, cmu;
  };
  auto load_cell = [&](int64_t i, CellIn& ci) {
    const int64_t c = pv.inc_cell[i];
    ci.le = pv.inc_loc[i];
    const uint16_t* __restrict__ ps = pv.pos + (i - pv.inc_base) * NEC;
#pragma unroll
    for (int u = 0; u < NEC; ++u) ci.slot[u] = ps[u];
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      const int64_t vid = cells[c * 3 + v];
      ci.x[v][0] = coords[vid * 2];
      ci.x[v][1] = coords[vid * 2 + 1];
    }
    ci.ceps = eps ? eps[c] : 1.0;
    ci.cmu = mu ? mu[c] : 1.0;
  };
  auto add_cell = [&](const CellIn& ci) {
    const int le = ci.le;
    // 2 x 2 geometry of the cell (cell_geometry<2> on the gathered vertices)
    const double J00 = ci.x[1][0] - ci.x[0][0], J01 = ci.x[2][0] - ci.x[0][0];
    const double J10 = ci.x[1][1] - ci.x[0][1], J11 = ci.x[2][1] - ci.x[0][1];
    const double det = J00 * J11 - J01 * J10, id = 1.0 / det, ad = fabs(det);
    const double Ji00 = J11 * id, Ji01 = -J01 * id, Ji10 = -J10 * id, Ji11 = J00 * id;
    double gT[3], gS[1];
    gT[0] = (Ji00 * Ji00 + Ji01 * Ji01) * ad * ci.ceps;
    gT[1] = (Ji10 * Ji10 + Ji11 * Ji11) * ad * ci.ceps;
    gT[2] = (Ji00 * Ji10 + Ji01 * Ji11) * ad * ci.ceps;
    gS[0] = 1.0 / (ad * ci.cmu);
#pragma unroll
    for (int u = 0; u < NEC; ++u) {
      const int s = ci.slot[u];
      double vS[Q][Q], vT[Q][Q];
#pragma unroll
      for (int k = 0; k < Q; ++k)
#pragma unroll
        for (int kp = 0; kp < Q; ++kp) {
          const int ij = k * ND + u * Q + kp;
          const double* __restrict__ mh = sM + le * SM_STRIDE + 3 * ij;
          vT[k][kp] = gT[0] * mh[0] + gT[1] * mh[1] + gT[2] * mh[2];
          vS[k][kp] = gS[0] * sC[le * SC_STRIDE + ij];
        }
      if (u == le) {
        self_slot = s;
#pragma unroll
        for (int k = 0; k < Q; ++k)
#pragma unroll
          for (int kp = 0; kp < Q; ++kp) { selfS[k][kp] += vS[k][kp]; selfT[k][kp] += vT[k][kp]; }
      } else {
#pragma unroll
        for (int k = 0; k < Q; ++k) {
          const int64_t o = rs + (int64_t)k * Q * len + Q * s;
          if constexpr (Q == 2) {  // o is even and the arrays are 16-byte aligned: one 16-byte store per (row, column entity)
            *reinterpret_cast<double2*>(dS + o) = pecrow[k] ? make_double2(0.0, 0.0) : make_double2(vS[k][0], vS[k][1]);
            *reinterpret_cast<double2*>(dT + o) = pecrow[k] ? make_double2(0.0, 0.0) : make_double2(vT[k][0], vT[k][1]);
          } else {
#pragma unroll
            for (int kp = 0; kp < Q; ++kp) {
              dS[o + kp] = pecrow[k] ? 0.0 : vS[k][kp];
              dT[o + kp] = pecrow[k] ? 0.0 : vT[k][kp];
            }
          }
        }
      }
    }
  };
  int64_t i = i0;
  if (ORDER == 2 && i1 - i0 >= 2) {  // (order 1 is faster one cell at a time: fewer registers, more resident warps)
    CellIn c0, c1;
    load_cell(i0, c0);
    load_cell(i0 + 1, c1);
    add_cell(c0);
    add_cell(c1);
    i = i0 + 2;
  }
  for (; i < i1; ++i) {  // the only cell of a boundary edge / an interior entity; non-manifold edges
    CellIn c0;
    load_cell(i, c0);
    add_cell(c0);
  }
#pragma unroll
  for (int k = 0; k < Q; ++k) {
    const int64_t o = rs + (int64_t)k * Q * len + Q * self_slot;
    if constexpr (Q == 2) {
      const double2 unit = make_double2(k == 0 ? 1.0 : 0.0, k == 1 ? 1.0 : 0.0);
      *reinterpret_cast<double2*>(dS + o) = pecrow[k] ? unit : make_double2(selfS[k][0], selfS[k][1]);
      *reinterpret_cast<double2*>(dT + o) = pecrow[k] ? unit : make_double2(selfT[k][0], selfT[k][1]);
    } else {
#pragma unroll
      for (int kp = 0; kp < Q; ++kp) {
        dS[o + kp] = pecrow[k] ? (kp == k ? 1.0 : 0.0) : selfS[k][kp];
        dT[o + kp] = pecrow[k] ? (kp == k ? 1.0 : 0.0) : selfT[k][kp];
      }
    }
  }
}

constexpr int kTriThreads = 64;
template <int ORDER>
__global__ void __launch_bounds__(kTriThreads)
k_values_tri(PatView pv, const double* __restrict__ coords, const int32_t* __restrict__ cells,
             const double* __restrict__ eps, const double* __restrict__ mu, const uint8_t* __restrict__ pec,
             double* __restrict__ outS, double* __restrict__ outT) {
  using El = Elem<2, ORDER>;
  constexpr int Q = ORDER, NEC = ORDER == 1 ? 3 : 4, ND = Q * NEC;
  constexpr int CAP = 32 * Q * Q * (2 * NEC - 1);  // values per warp tile: 32 entities x q rows x q (2 NEC - 1) columns
  constexpr int SM_ROW = Q * ND * 3, SC_ROW = Q * ND;  // entries per local row entity; padded strides as in tri_entity_rows
  constexpr int SM_STRIDE = SM_ROW + (ORDER == 2 ? 1 : 0), SC_STRIDE = SC_ROW + (ORDER == 2 ? 1 : 0);
  __shared__ double sM[NEC * SM_STRIDE];
  __shared__ double sC[NEC * SC_STRIDE];
  __shared__ __align__(16) double tiles[kTriThreads / 32][2][CAP];
  for (int t = threadIdx.x; t < ND * ND * 3; t += blockDim.x) sM[(t / SM_ROW) * SM_STRIDE + t % SM_ROW] = El::mass(0, 0)[t];
  for (int t = threadIdx.x; t < ND * ND; t += blockDim.x) sC[(t / SC_ROW) * SC_STRIDE + t % SC_ROW] = El::curl(0, 0)[t];
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool valid = r < pv.n_rows;
  const unsigned vmask = __ballot_sync(0xffffffffu, valid);
  if (vmask == 0u) return;
  const int64_t g = pv.ent_lo + r;
  const int64_t a = valid ? pv.ent_rowptr[r] : 0;
  const int64_t a1 = valid ? pv.ent_rowptr[r + 1] : 0;
  const int len = (int)(a1 - a);
  const int64_t a_first = __shfl_sync(0xffffffffu, a, 0);
  const int64_t a_end = __shfl_sync(0xffffffffu, a1, 31 - __clz((int)vmask));
  const int64_t span = (int64_t)Q * Q * (a_end - a_first);
  const bool tiled = span <= CAP;  // warp-uniform
  const int64_t rs = (int64_t)Q * Q * (a - a_first);  // first value of row 0 in the span; row k: + k * Q * len
  if (valid) {
    if (tiled)
      tri_entity_rows<ORDER>(pv, sM, sC, coords, cells, eps, mu, pec, g, rs, len, tiles[w][0], tiles[w][1]);
    else
      tri_entity_rows<ORDER>(pv, sM, sC, coords, cells, eps, mu, pec, g, rs, len, outS + (int64_t)Q * Q * a_first,
                             outT + (int64_t)Q * Q * a_first);
  }
  if (!tiled) return;
  __syncwarp();
  // stream the span out: the value arrays are written once and never read here
  double* __restrict__ gS_ = outS + (int64_t)Q * Q * a_first;
  double* __restrict__ gT_ = outT + (int64_t)Q * Q * a_first;
  const int n = (int)span;
  if (Q == 2) {  // spans start at multiples of 4 values: 16-byte stores
    const double2* tS = reinterpret_cast<const double2*>(tiles[w][0]);
    const double2* tT = reinterpret_cast<const double2*>(tiles[w][1]);
    for (int j = lane; j < n / 2; j += 32) {
      __stcs(reinterpret_cast<double2*>(gS_) + j, tS[j]);
      __stcs(reinterpret_cast<double2*>(gT_) + j, tT[j]);
    }
  } else {
    for (int j = lane; j < n; j += 32) {
      __stcs(gS_ + j, tiles[w][0][j]);
      __stcs(gT_ + j, tiles[w][1][j]);
    }
  }
}

// ---- order-1 tetrahedra (Whitney elements), S and T in one sweep: one THREAD per edge ----------------------------
// A row is short (1 + 3 m entries for an edge of m cells), so a thread sums its row over the incident cells in
// ascending order (DOLFIN's order) in a lane-private stretch of the warp's shared-memory tile; the 32 consecutive edges
// of a warp own one contiguous span of S and of T, streamed out with coalesced stores.  ~60 registers per thread,
// so many resident warps hide the gather chain (incidence -> cell -> vertices) that bounds the lane-pair kernel
// k_canon1_edges.  No atomics, bit-reproducible.  PEC rows are written as unit rows.  tile_cap = values per warp tile.
constexpr int kTet1Threads = 128;
template <bool SHARED>
__device__ __forceinline__ void tet1_edge_row(const PatView& pv, const double* __restrict__ sM,
                                              const double* __restrict__ sC, const double* __restrict__ coords,
                                              const int32_t* __restrict__ cells, const double* __restrict__ eps,
                                              const double* __restrict__ mu, const uint8_t* __restrict__ pec, int64_t g,
                                              int64_t rs, int len, double* __restrict__ dS, double* __restrict__ dT) {
  for (int s = 0; s < len; ++s) dS[rs + s] = dT[rs + s] = 0.0;
  int self_slot = 0;
  const int64_t i0 = pv.inc_ptr[g], i1 = pv.inc_ptr[g + 1];
  for (int64_t i = i0; i < i1; ++i) {
    const int64_t c = pv.inc_cell[i];
    const int le = pv.inc_loc[i];
    const uint16_t* __restrict__ ps = pv.pos + (i - pv.inc_base) * 6;
    const unsigned long long desc = kCanonDesc[le];  // slot table is in canonical column order (pattern.cu)
    double gT[6], gS[6];
    cell_geometry<3>(coords, cells + c * 4, eps ? eps[c] : 1.0, mu ? mu[c] : 1.0, gT, gS, nullptr);
#pragma unroll
    for (int uu = 0; uu < 6; ++uu) {
      const int s = ps[uu];
      const int u = (int)((desc >> (8 + 4 * uu)) & 15ull);
      const double* __restrict__ Mh = sM + (le * 6 + u) * 6;
      const double* __restrict__ Ch = sC + (le * 6 + u) * 6;
      double vT = 0.0, vS = 0.0;
#pragma unroll
      for (int p = 0; p < 6; ++p) { vT += gT[p] * Mh[p]; vS += gS[p] * Ch[p]; }
      if (u == le) self_slot = s;
      dS[rs + s] += vS;
      dT[rs + s] += vT;
    }
  }
  if (pec && pec[g])
    for (int s = 0; s < len; ++s) dS[rs + s] = dT[rs + s] = (s == self_slot) ? 1.0 : 0.0;
}

__global__ void __launch_bounds__(kTet1Threads)
k_values_tet1(PatView pv, const double* __restrict__ coords, const int32_t* __restrict__ cells,
              const double* __restrict__ eps, const double* __restrict__ mu, const uint8_t* __restrict__ pec,
              double* __restrict__ outS, double* __restrict__ outT, int tile_cap) {
  extern __shared__ double smem_t1[];
  double* sM = smem_t1;
  double* sC = smem_t1 + 216;
  for (int t = threadIdx.x; t < 216; t += blockDim.x) {
    sM[t] = Elem<3, 1>::mass(0, 0)[t];
    sC[t] = Elem<3, 1>::curl(0, 0)[t];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double* tS = smem_t1 + 432 + (size_t)w * 2 * tile_cap;
  double* tT = tS + tile_cap;
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool valid = r < pv.n_rows;
  const unsigned vmask = __ballot_sync(0xffffffffu, valid);
  if (vmask == 0u) return;
  const int64_t g = pv.ent_lo + r;
  const int64_t a = valid ? pv.ent_rowptr[r] : 0;
  const int64_t a1 = valid ? pv.ent_rowptr[r + 1] : 0;
  const int64_t a_first = __shfl_sync(0xffffffffu, a, 0);
  const int64_t a_end = __shfl_sync(0xffffffffu, a1, 31 - __clz((int)vmask));
  const int64_t span = a_end - a_first;
  const bool tiled = span <= tile_cap;  // warp-uniform
  if (valid) {
    if (tiled)
      tet1_edge_row<true>(pv, sM, sC, coords, cells, eps, mu, pec, g, a - a_first, (int)(a1 - a), tS, tT);
    else
      tet1_edge_row<false>(pv, sM, sC, coords, cells, eps, mu, pec, g, a - a_first, (int)(a1 - a), outS + a_first,
                           outT + a_first);
  }
  if (!tiled) return;
  __syncwarp();
  for (int j = lane; j < (int)span; j += 32) {
    __stcs(outS + a_first + j, tS[j]);
    __stcs(outT + a_first + j, tT[j]);
  }
}

// rows that can carry ABC contributions: row entities that LIE ON an exterior facet (the tangential trace of a basis
// function vanishes on every facet that does not contain its entity, so all other rows of S_0 are zero).
// inc_loc = local entity of the row entity in the incident cell (tets: edges 0-5, faces 6-9; triangles: edges 0-2,
// interior 3); cell_extmask bit lf = local facet lf (opposite vertex lf) is exterior.
template <int DIM>
__global__ void k_abc_row_flags(const int64_t* __restrict__ inc_ptr, const int32_t* __restrict__ inc_cell,
                                const uint8_t* __restrict__ inc_loc, const uint8_t* __restrict__ cell_extmask,
                                int64_t ent_lo, int64_t n_rows, uint8_t* __restrict__ flags) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int64_t g = ent_lo + r;
  uint8_t f = 0;
  for (int64_t i = inc_ptr[g]; i < inc_ptr[g + 1]; ++i) {
    const unsigned m = cell_extmask[inc_cell[i]];
    if (!m) continue;
    const int loc = inc_loc[i];
    unsigned on;  // local facets that contain the local entity
    if (DIM == 3) {
      // tet edge (a,b) lies on the faces opposite the other two vertices: e0=(2,3) e1=(1,3) e2=(1,2) e3=(0,3) e4=(0,2) e5=(0,1)
      on = loc < 6 ? ((0xCA6953u >> (4 * loc)) & 15u) : (1u << (loc - 6));
    } else {
      on = loc < 3 ? (1u << loc) : 0u;
    }
    if (m & on) f = 1;
  }
  flags[r] = f;
}

__global__ void k_ident_rows(const int64_t* __restrict__ ent_rowptr, const int32_t* __restrict__ ent_cols, int q,
                             int64_t ent_lo, int64_t gshift, int64_t n_rows, const uint8_t* __restrict__ flags,
                             double* __restrict__ vals) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + w;
  if (r >= n_rows) return;
  const int64_t g = ent_lo + r;
  const int64_t a = ent_rowptr[r];
  const int len = (int)(ent_rowptr[r + 1] - a);
  for (int k = 0; k < q; ++k) {
    if (!flags[q * g + k]) continue;
    const int64_t rs = (int64_t)q * q * a + (int64_t)k * q * len;
    for (int idx = lane; idx < q * len; idx += 32) {
      int s = idx / q, kp = idx - s * q;
      vals[rs + idx] = (ent_cols[a + s] == g + gshift && kp == k) ? 1.0 : 0.0;
    }
  }
}

__global__ void k_list_row_len(const int64_t* __restrict__ ent_rowptr, const int64_t* __restrict__ list, int64_t n,
                               int64_t* __restrict__ len) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i > n) return;
  len[i] = i < n ? ent_rowptr[list[i] + 1] - ent_rowptr[list[i]] : 0;
}

}  // namespace

int build_abc_rows(sfem_pattern_s* pat, cudaStream_t st) {
  if (pat->n_abc_rows >= 0) return 0;
  const sfem_space_s* sp = pat->space;
  const int64_t n = pat->n_rows_ent;
  uint8_t* flags = nullptr;
  int64_t* d_count = nullptr;
  SFEM_CUDA(dev_malloc(&flags, std::max<int64_t>(n, 1), st));
  SFEM_CUDA(dev_malloc(&d_count, 8, st));
  SFEM_CUDA(dev_malloc(&pat->abc_rows, std::max<int64_t>(n, 1) * 8, st));
  if (sp->mesh->gdim == 3)
    k_abc_row_flags<3><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sp->inc_ptr, sp->inc_cell, sp->inc_loc,
                                                                 sp->cell_extmask, pat->ent_lo, n, flags);
  else
    k_abc_row_flags<2><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sp->inc_ptr, sp->inc_cell, sp->inc_loc,
                                                                 sp->cell_extmask, pat->ent_lo, n, flags);
  cub::CountingInputIterator<int64_t> it(0);
  size_t tb = 0;
  SFEM_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, it, flags, pat->abc_rows, d_count, (int)n, st));
  void* tmp = nullptr;
  SFEM_CUDA(dev_malloc(&tmp, tb + 16, st));
  SFEM_CUDA(cub::DeviceSelect::Flagged(tmp, tb, it, flags, pat->abc_rows, d_count, (int)n, st));
  int64_t h = 0;
  SFEM_CUDA(cudaMemcpyAsync(&h, d_count, 8, cudaMemcpyDeviceToHost, st));
  SFEM_CUDA(cudaStreamSynchronize(st));
  dev_free(tmp, st); dev_free(flags, st); dev_free(d_count, st);
  // offsets of the listed rows in the compact storage: exclusive scan of their lengths
  int64_t* len = nullptr;
  SFEM_CUDA(dev_malloc(&len, (h + 1) * 8, st));
  SFEM_CUDA(dev_malloc(&pat->abc_rowptr, (h + 1) * 8, st));
  k_list_row_len<<<(unsigned)((h + 1 + 255) / 256), 256, 0, st>>>(pat->ent_rowptr, pat->abc_rows, h, len);
  tb = 0;
  SFEM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, len, pat->abc_rowptr, (int)(h + 1), st));
  SFEM_CUDA(dev_malloc(&tmp, tb + 16, st));
  SFEM_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, len, pat->abc_rowptr, (int)(h + 1), st));
  SFEM_CUDA(cudaMemcpyAsync(&pat->abc_nnz_ent, pat->abc_rowptr + h, 8, cudaMemcpyDeviceToHost, st));
  SFEM_CUDA(cudaStreamSynchronize(st));
  dev_free(tmp, st); dev_free(len, st);
  pat->n_abc_rows = h;
  return 0;
}

namespace {

template <int DIM, int ORDER>
int launch_generic(const sfem_pattern_s* pat, int form, const double* coef_a, const double* coef_b,
                   const uint8_t* facet_mask, const uint8_t* pec, double* out_a, double* out_b,
                   cudaStream_t st, bool abc_compact, bool abc_rows_kernel, const int64_t* row_list, int64_t n_list) {
  const sfem_space_s* sp = pat->space;
  const sfem_mesh_s* m = sp->mesh;
  PatView pv{sp->inc_ptr, sp->inc_cell, sp->inc_loc, pat->ent_rowptr, pat->ent_cols, pat->pos,
             pat->ent_lo, pat->n_rows_ent, pat->inc_base, pat->gshift, sp->n_ent_cell, pos_is_canonical(sp), nullptr, 0, nullptr};
  if (pat->n_rows_ent == 0) return 0;
  unsigned grid = (unsigned)((pat->n_rows_ent + kWarps - 1) / kWarps);
  if (form == SFEM_FORM_ABC) {
    // Only the rows of entities ON the boundary are non-zero.  Full form: S_0 lives on the S/M pattern -- clear
    // everything at memset speed, assemble the boundary rows, then write the unit rows of Dirichlet dofs (which may be
    // interior).  Compact form: only those rows are stored (offsets abc_rowptr), nothing to clear.
    SFEM_TRY(build_abc_rows(const_cast<sfem_pattern_s*>(pat), st));
    const int64_t nnz = (int64_t)ORDER * ORDER * pat->nnz_ent;
    if (!abc_compact) SFEM_CUDA(cudaMemsetAsync(out_a, 0, nnz * sizeof(double), st));
    pv.row_list = pat->abc_rows;
    pv.n_list = pat->n_abc_rows;
    if (abc_compact) pv.out_rowptr = pat->abc_rowptr;
    grid = (unsigned)((pat->n_abc_rows + kWarps - 1) / kWarps);
  }
  if (form != SFEM_FORM_ABC && row_list) {  // only these rows (per-entity fallback of the canonical path)
    if (n_list == 0) return 0;
    pv.row_list = row_list;
    pv.n_list = n_list;
    grid = (unsigned)((n_list + kWarps - 1) / kWarps);
  }
  const int32_t* cell_facets = DIM == 3 ? sp->cell_faces : sp->cell_edges;
  double* geo = nullptr;
  if (form != SFEM_FORM_ABC) {
    SFEM_CUDA(dev_malloc(&geo, m->C * 12 * sizeof(double), st));
    // coef_a = eps_r (mass), coef_b = mu_r (stiffness)
    k_cell_geometry<DIM><<<(unsigned)((m->C + 255) / 256), 256, 0, st>>>(m->coords, m->cells, coef_a, coef_b, m->C, geo);
  }
  switch (form) {
    case SFEM_FORM_MASS:
      k_values_generic<DIM, ORDER, 0><<<grid, kWarps * 32, 0, st>>>(pv, geo, m->coords, m->cells, cell_facets,
                                                                   sp->cell_extmask, facet_mask, pec, out_a, nullptr);
      break;
    case SFEM_FORM_STIFFNESS:
      k_values_generic<DIM, ORDER, 1><<<grid, kWarps * 32, 0, st>>>(pv, geo, m->coords, m->cells, cell_facets,
                                                                   sp->cell_extmask, facet_mask, pec, out_a, nullptr);
      break;
    case SFEM_FORM_ABC:
      if (grid > 0 && abc_rows_kernel)
        k_abc_rows<DIM, ORDER><<<grid, kWarps * 32, 0, st>>>(pv, m->coords, m->cells, cell_facets, sp->cell_extmask,
                                                            facet_mask, abc_compact ? pec : nullptr, out_a);
      else if (grid > 0)
        k_values_generic<DIM, ORDER, 2><<<grid, kWarps * 32, 0, st>>>(pv, geo, m->coords, m->cells, cell_facets,
                                                                     sp->cell_extmask, facet_mask,
                                                                     abc_compact ? pec : nullptr, out_a, nullptr);
      if (pec && !abc_compact)
        k_ident_rows<<<(unsigned)((pat->n_rows_ent + 7) / 8), 256, 0, st>>>(pat->ent_rowptr, pat->ent_cols, ORDER,
                                                                          pat->ent_lo, pat->gshift, pat->n_rows_ent, pec, out_a);
      break;
    default:  // 3: fused S (out_a) and M (out_b)
      k_values_generic<DIM, ORDER, 3><<<grid, kWarps * 32, 0, st>>>(pv, geo, m->coords, m->cells, cell_facets,
                                                                   sp->cell_extmask, facet_mask, pec, out_a, out_b);
  }
  SFEM_CUDA(cudaGetLastError());
  if (geo) dev_free(geo, st);
  return 0;
}

}  // namespace

// form: SFEM_FORM_MASS (coef_a = eps_r), SFEM_FORM_STIFFNESS (coef_b = mu_r), SFEM_FORM_ABC, or 3 = fused
// S -> out_a, M -> out_b with coef_a = eps_r, coef_b = mu_r.
int assemble_generic(const sfem_pattern_s* pat, int form, const double* coef_a, const double* coef_b,
                     const uint8_t* facet_mask, const uint8_t* pec, double* out_a, double* out_b,
                     cudaStream_t st, bool abc_compact, bool abc_rows_kernel, const int64_t* row_list, int64_t n_list) {
  const int dim = pat->space->mesh->gdim, order = pat->space->order;
  if (dim == 3 && order == 1) return launch_generic<3, 1>(pat, form, coef_a, coef_b, facet_mask, pec, out_a, out_b, st, abc_compact, abc_rows_kernel, row_list, n_list);
  if (dim == 3 && order == 2) return launch_generic<3, 2>(pat, form, coef_a, coef_b, facet_mask, pec, out_a, out_b, st, abc_compact, abc_rows_kernel, row_list, n_list);
  if (dim == 2 && order == 1) return launch_generic<2, 1>(pat, form, coef_a, coef_b, facet_mask, pec, out_a, out_b, st, abc_compact, abc_rows_kernel, row_list, n_list);
  if (dim == 2 && order == 2) return launch_generic<2, 2>(pat, form, coef_a, coef_b, facet_mask, pec, out_a, out_b, st, abc_compact, abc_rows_kernel, row_list, n_list);
  set_error("unsupported (dim, order)");
  return 2;
}

// Thread-per-row-entity kernels (SFEM_KERNEL_ROWS): triangles of order 1 / 2 and order-1 tetrahedra, S and M(=T)
// together in one launch.
bool rows_supported(const sfem_pattern_s* pat) {
  const sfem_space_s* sp = pat->space;
  return sp->mesh->gdim == 2 ? (sp->order == 1 || sp->order == 2) : (sp->mesh->gdim == 3 && sp->order == 1);
}

int assemble_rows_ST(const sfem_pattern_s* pat, const double* eps, const double* mu, const uint8_t* pec, double* S,
                     double* T, cudaStream_t st) {
  const sfem_space_s* sp = pat->space;
  const sfem_mesh_s* m = sp->mesh;
  if (!rows_supported(pat)) {
    set_error("SFEM_KERNEL_ROWS: triangles of order 1 or 2 and tetrahedra of order 1 only");
    return 2;
  }
  if (pat->n_rows_ent == 0) return 0;
  PatView pv{sp->inc_ptr, sp->inc_cell, sp->inc_loc, pat->ent_rowptr, pat->ent_cols, pat->pos,
             pat->ent_lo, pat->n_rows_ent, pat->inc_base, pat->gshift, sp->n_ent_cell, pos_is_canonical(sp), nullptr, 0, nullptr};
  if (m->gdim == 3) {
    // tile = 32 rows of the longest length, capped so that several CTAs stay resident (longer spans write directly)
    constexpr int kWarpsT1 = kTet1Threads / 32;
    int cap = 32 * std::max(pat->max_row_len, 1);
    const int cap_max = (int)((96 * 1024 - 432 * 8) / (kWarpsT1 * 2 * 8));
    cap = std::min(cap, cap_max);
    const size_t smem = (size_t)(432 + kWarpsT1 * 2 * cap) * sizeof(double);
    static bool attr_set[64] = {};
    if (m->device >= 0 && m->device < 64 && !attr_set[m->device]) {
      SFEM_CUDA(cudaFuncSetAttribute(k_values_tet1, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attr_set[m->device] = true;
    }
    const unsigned grid = (unsigned)((pat->n_rows_ent + kTet1Threads - 1) / kTet1Threads);
    k_values_tet1<<<grid, kTet1Threads, smem, st>>>(pv, m->coords, m->cells, eps, mu, pec, S, T, cap);
  } else {
    const unsigned grid = (unsigned)((pat->n_rows_ent + kTriThreads - 1) / kTriThreads);
    if (sp->order == 1)
      k_values_tri<1><<<grid, kTriThreads, 0, st>>>(pv, m->coords, m->cells, eps, mu, pec, S, T);
    else
      k_values_tri<2><<<grid, kTriThreads, 0, st>>>(pv, m->coords, m->cells, eps, mu, pec, S, T);
  }
  SFEM_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sfem
