// Internal declarations shared by the translation units of libsucem_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/sucem_b200.h"

namespace sfem {

void set_error(const std::string& msg);

#define SFEM_CUDA(call)                                                                          \
  do {                                                                                           \
    cudaError_t _e = (call);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      sfem::set_error(std::string(#call) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ +     \
                      ":" + std::to_string(__LINE__) + ")");                                     \
      return 1;                                                                                  \
    }                                                                                            \
  } while (0)

#define SFEM_CHECK(cond, msg)        \
  do {                               \
    if (!(cond)) {                   \
      sfem::set_error(msg);          \
      return 2;                      \
    }                                \
  } while (0)

#define SFEM_TRY(expr)           \
  do {                           \
    int _rc = (expr);            \
    if (_rc != 0) return _rc;    \
  } while (0)

// Device memory: stream-ordered allocations from the device's default memory pool.  The library raises the pool's
// release threshold once per device, so freed blocks stay cached instead of being returned to the driver at every
// stream synchronisation (set-up after set-up then reuses them); sfem_release_cached_memory() trims the pool.
cudaError_t dev_malloc_bytes(void** p, size_t bytes, cudaStream_t st);
template <class T>
inline cudaError_t dev_malloc(T** p, size_t bytes, cudaStream_t st) {
  return dev_malloc_bytes(reinterpret_cast<void**>(p), bytes, st);
}
inline void dev_free(void* p, cudaStream_t st) {
  if (p) cudaFreeAsync(p, st);
}

// Makes `dev` the current device for the lifetime of the guard (every entry point that launches or allocates runs on
// the device its handle was created on, whatever the caller's current device is).
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != prev && prev >= 0) cudaSetDevice(prev);
  }
};

// Local topology of the UFC reference cells.
// tet edges e0=(2,3) e1=(1,3) e2=(1,2) e3=(0,3) e4=(0,2) e5=(0,1); faces f_i opposite v_i
// triangle edges e_i opposite v_i
__host__ __device__ inline int tet_edge_vertex(int e, int k) {
  const int t[6][2] = {{2, 3}, {1, 3}, {1, 2}, {0, 3}, {0, 2}, {0, 1}};
  return t[e][k];
}
__host__ __device__ inline int tet_face_vertex(int f, int k) {
  const int t[4][3] = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};
  return t[f][k];
}
__host__ __device__ inline int tri_edge_vertex(int e, int k) {
  const int t[3][2] = {{1, 2}, {0, 2}, {0, 1}};
  return t[e][k];
}
// the three local edges of local tet face f (edges whose vertices both lie in the face)
__host__ __device__ inline int tet_face_edge(int f, int k) {
  const int t[4][3] = {{0, 1, 2}, {0, 3, 4}, {1, 3, 5}, {2, 4, 5}};
  return t[f][k];
}

}  // namespace sfem

struct sfem_mesh_s {
  int device = 0;
  int gdim = 0;   // 2 or 3
  int nv = 0;     // vertices per cell
  int64_t V = 0, C = 0;
  const double* coords = nullptr;  // [V][gdim]   caller-owned
  const int32_t* cells = nullptr;  // [C][nv]     caller-owned
};

struct sfem_space_s {
  sfem_mesh_s* mesh = nullptr;
  int device = 0;
  cudaStream_t stream = nullptr;  // stream of the creating call: the handle's buffers are freed in its order
  int order = 0;         // = dofs per entity q
  int ne_loc = 0;        // local edges per cell (6 | 3)
  int nf_loc = 0;        // local faces per cell (4 | 0)
  int n_ent_cell = 0;    // dof entities per cell: order 1: ne_loc; order 2: ne_loc + (4 faces | 1 interior)
  int ndof_cell = 0;     // order * n_ent_cell
  int64_t E = 0, F = 0;  // #edges, #faces (3-D)
  int64_t n_ent = 0;     // dof entities: E (+ F in 3-D order 2, + C in 2-D order 2)
  int64_t ndof = 0;
  int32_t* cell_edges = nullptr;  // [C][ne_loc]
  int32_t* cell_faces = nullptr;  // [C][4] (3-D only)
  // entity -> cell incidences of the dof entities, ascending cell index within an entity
  int64_t* inc_ptr = nullptr;     // [n_ent+1]
  int32_t* inc_cell = nullptr;    // [n_inc]
  uint8_t* inc_loc = nullptr;     // [n_inc] local dof-entity index t of the entity in that cell
  int64_t n_inc = 0;
  // facets (3-D faces / 2-D edges): number of incident cells and the first (cell, local facet)
  int64_t n_facets = 0;
  uint8_t* facet_ncells = nullptr;    // [n_facets]
  int32_t* facet_first_cell = nullptr;// [n_facets]
  uint8_t* facet_first_loc = nullptr; // [n_facets]
  uint8_t* cell_extmask = nullptr;    // [C] bit lf set <=> local facet lf is exterior
  double* coords4 = nullptr;          // [V][4] 32-byte padded copy of the coordinates (3-D; vector gathers)
  bool pattern_big = false;           // a row has more than 1024 candidates: the pattern kernels use their large-row launch
  double* ones = nullptr;             // [C] 1.0, the default eps_r / mu_r (allocated on first use)
  double2* coef2 = nullptr;           // [C] (mu_r, eps_r) packed by every sfem_assemble_ST call (lane-per-item edge kernel)
  // count pass over ALL rows (lazily, sfem_space_count_nnz): entity-level row pointer, longest rows
  int64_t* rowptr_all = nullptr;      // [n_ent+1]
  int64_t nnz_ent_all = 0;
  int max_len_edges_all = 0, max_len_other_all = 0;
  cudaStream_t drain = nullptr;       // device->host drain stream of sfem_space_assemble_to_host (lazily created)
  // Partitioned mesh (sfem_space_set_global_ids): this space was numbered on ONE RANK's cells (own + halo); column
  // entities are reported through these maps (local edge / face -> global combined entity id), rows of the owned
  // ranges are shifted by gshift_e / gshift_f (global id - local combined id).  nullptr / 0: the space is the mesh.
  int32_t* gid_e = nullptr;           // [E]
  int32_t* gid_f = nullptr;           // [F] (already offset by the global edge count)
  int64_t gshift_e = 0, gshift_f = 0;
  int64_t n_ent_global = 0, ndof_global = 0;
};

struct sfem_pattern_s {
  sfem_space_s* space = nullptr;
  int device = 0;
  cudaStream_t stream = nullptr;  // stream of the creating call: the handle's buffers are freed in its order
  int64_t ent_lo = 0, ent_hi = 0;
  int64_t n_rows_ent = 0;
  int64_t nnz_ent = 0;            // entity-level non-zeros of the block
  int64_t* ent_rowptr = nullptr;  // [n_rows_ent+1], offsets into this block
  int32_t* ent_cols = nullptr;    // [nnz_ent] global entity ids, ascending within a row
  uint16_t* pos = nullptr;        // [n_inc_block][n_ent_cell]: slot of the cell's u-th column entity in the row
                                  // (u in canonical order when pos_is_canonical(space))
  uint4* rec_alloc = nullptr;     // the allocation behind `rec` (rec may point before it: see fill_pattern)
  uint4* rec = nullptr;           // [n_inc_block][2] packed 32-byte item records {relabelled vertex ids, cell,
                                  // descriptor, flags, 10 uint8 slots} (order-2 tets, rows <= 255 entities)
  bool ring_ok = false;           // edge records are in ring order (manifold edges of bounded valence)
  // order-2 tets, lane-per-item edge kernel (k_canon_edges2): the edge records live in BATCHES of 32 item slots (whole
  // edges, at most kE2TileChunks 16-byte chunks of output per batch), padded with cell = -1
  uint4* rec_e2 = nullptr;        // [n_batch_e2][32][2] packed item records in ring order
  uint4* aux_e2 = nullptr;        // [n_batch_e2][32] {edge row, packed (tile offset, row length, position, remaining,
                                  // ring-closing lane distance), first entity chunk of the batch's span, span chunks}
  int64_t n_batch_e2 = 0;         // upper bound known on the host (allocation size); the count itself stays on the device:
  int64_t* nbatch_e2_dev = nullptr;
  int64_t* gen_rows = nullptr;    // block-relative edge rows the batched layout does not take (more than 32 cells, rows longer
  int64_t n_gen_rows = 0;         // than the tile): the canonical path hands exactly these to the generic kernel
  int64_t* abc_rows = nullptr;    // block-relative row entities that lie on an exterior facet (lazily built)
  int64_t n_abc_rows = -1;
  int64_t* abc_rowptr = nullptr;  // [n_abc_rows+1] entity-level offsets of those rows in the COMPACT S_0 storage
  int64_t abc_nnz_ent = 0;
  int64_t gshift = 0;             // global combined entity id of a row entity = ent_lo + r + gshift (partitioned mesh)
  int64_t inc_base = 0;           // = space->inc_ptr[ent_lo]
  int64_t n_inc_block = 0;
  int max_row_len = 0;            // longest entity row of the block
  int max_row_len_edges = 0;      // ... among its edge rows
  int max_row_len_other = 0;      // ... among its face / interior rows
};

namespace sfem {
// dof-entity u of cell c: edges first, then faces (3-D) / the cell interior (2-D)
struct SpaceView {
  const int32_t* cell_edges;
  const int32_t* cell_faces;
  const int32_t* gid_e;  // partitioned mesh: local -> global combined entity ids (else nullptr)
  const int32_t* gid_f;
  int ne_loc, n_ent_cell;
  int64_t E;
  // the id under which dof entity u of cell c appears as a COLUMN (global numbering)
  __device__ __forceinline__ int32_t ent(int64_t c, int u) const {
    if (u < ne_loc) {
      const int32_t e = cell_edges[c * ne_loc + u];
      return gid_e ? gid_e[e] : e;
    }
    if (cell_faces) {
      const int32_t f = cell_faces[c * 4 + (u - ne_loc)];
      return gid_f ? gid_f[f] : (int32_t)(E + f);
    }
    return (int32_t)(E + c);
  }
};
inline SpaceView make_view(const sfem_space_s* s) {
  SpaceView v;
  v.cell_edges = s->cell_edges;
  v.cell_faces = (s->mesh->gdim == 3 && s->order == 2) ? s->cell_faces : nullptr;
  v.gid_e = s->gid_e;
  v.gid_f = s->gid_f;
  v.ne_loc = s->ne_loc;
  v.n_ent_cell = s->n_ent_cell;
  v.E = s->E;
  return v;
}

// tets: the slot table is stored in canonical column order (see pattern.cu / assemble_canonical.cu)
inline bool pos_is_canonical(const sfem_space_s* s) { return s->mesh->gdim == 3; }

// frees every device buffer of a pattern (stream-ordered); the struct itself stays
void free_pattern_buffers(sfem_pattern_s* pat, cudaStream_t st);
// output chunks (16 B) per warp tile of the lane-per-item edge kernel: a batch never exceeds it
constexpr int kE2TileChunks = 448;
int build_numbering(sfem_space_s* sp, cudaStream_t st);
int build_pattern(sfem_pattern_s* pat, cudaStream_t st);
int count_pattern_rows(const sfem_space_s* sp, int64_t ent_lo, int64_t ent_hi, cudaStream_t st, int64_t** rowptr_out,
                       int64_t* nnz_ent, int* max_len_edges, int* max_len_other);
int build_pattern_block(sfem_pattern_s* pat, const int64_t* rowptr_all, int64_t all_lo, int64_t nnz_ent,
                        int64_t inc_base, int64_t n_inc_block, int max_len_edges, int max_len_other, cudaStream_t st);
// rowptr entries are offset by nnz_base (0 for a stand-alone block)
int export_pattern(const sfem_pattern_s* pat, int64_t* d_rowptr, int32_t* d_cols, cudaStream_t st, int64_t nnz_base = 0);
int assemble_generic(const sfem_pattern_s* pat, int form, const double* coef_a, const double* coef_b,
                     const uint8_t* facet_mask, const uint8_t* pec, double* out_a, double* out_b,
                     cudaStream_t st, bool abc_compact = false, bool abc_rows_kernel = true,
                     const int64_t* row_list = nullptr, int64_t n_list = 0);
// boundary-row list + compact offsets of the ABC matrix (lazily built, synchronises `st` the first time)
int build_abc_rows(sfem_pattern_s* pat, cudaStream_t st);
int assemble_rows_ST(const sfem_pattern_s* pat, const double* eps, const double* mu, const uint8_t* pec, double* S,
                     double* T, cudaStream_t st);
bool rows_supported(const sfem_pattern_s* pat);
int assemble_canonical_ST(const sfem_pattern_s* pat, const double* eps, const double* mu, const uint8_t* pec,
                          double* S, double* T, cudaStream_t st);
bool canonical_supported(const sfem_pattern_s* pat);
int canonical_launches(const sfem_pattern_s* pat);
}  // namespace sfem
