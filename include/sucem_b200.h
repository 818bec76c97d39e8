/*
 * sucem_b200.h -- C ABI of libsucem_b200.so: B200-native N1curl (Nedelec 1st kind, order 1-2) assembly
 * of the curl-curl stiffness S, mass M(=T) and first-order-ABC boundary matrix S_0 with PEC row
 * elimination, on tetrahedra and triangles.
 *
 * The reference (cemagg/sucem-fem) has NO FFI of its own for this path: it reaches DOLFIN through its
 * SWIG Python module at exactly two call sites,
 *     dolfin.assemble(form, tensor=mat)                 sucemfem/SystemMatrices.py:54
 *     self.boundary_conditions.apply_essential(mat)     sucemfem/SystemMatrices.py:55
 * plus dolfin.FunctionSpace(mesh, family, order)        sucemfem/ProblemConfigurations/EMProblem.py:100-105
 * Each entry point below names the reference interface (file:line, relative to /root/reference) whose
 * work it replaces.  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; sfem_last_error() returns a
 *     thread-local message for the last failing call on the calling thread.
 *   - handles are opaque, created and destroyed by the library; a handle is bound to the CUDA device
 *     that was current at creation.  No global mutable state; calls on different handles are re-entrant.
 *   - all bulk arrays are caller-owned DEVICE pointers (e.g. torch tensors' data_ptr()); the library
 *     never frees them and they must outlive the handle that was created from them.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are asynchronous with
 *     respect to the host unless stated; creation calls synchronise the stream before returning because
 *     they return sizes.
 */
#ifndef SUCEM_B200_H
#define SUCEM_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define SFEM_API __attribute__((visibility("default")))
#else
#define SFEM_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sfem_mesh_s* sfem_mesh_t;
typedef struct sfem_space_s* sfem_space_t;
typedef struct sfem_pattern_s* sfem_pattern_t;
typedef struct sfem_host_job_s* sfem_host_job_t;

/* forms: which bilinear form a value array is assembled for */
enum {
  SFEM_FORM_MASS = 0,      /* eps_r*inner(v,u)*dx              sucemfem/Forms.py:71-77            */
  SFEM_FORM_STIFFNESS = 1, /* dot(curl(v),curl(u))/mu_r*dx     sucemfem/Forms.py:79-85            */
  SFEM_FORM_ABC = 2        /* inner(cross(n,v),cross(n,u))*ds  sucemfem/BoundaryConditions/ABC.py:56-57 */
};

/* kernel variants for the S/M value kernels */
enum {
  SFEM_KERNEL_AUTO = 0,    /* fastest available for the space                                    */
  SFEM_KERNEL_GENERIC = 1, /* row-owner gather with full reference tensors (any dim/order)       */
  SFEM_KERNEL_CANONICAL = 2,/* canonical-frame straight-line kernels (tetrahedra, order 1 and 2) */
  SFEM_KERNEL_ROWS = 3     /* one thread per row entity (triangles order 1-2, tetrahedra order 1) */
};

SFEM_API const char* sfem_last_error(void);
/* "sucem_b200 <version> sm_100a" */
SFEM_API const char* sfem_version(void);

/* Library-owned device buffers are stream-ordered allocations from the current device's default memory pool, kept
 * cached when handles are destroyed (repeated set-ups reuse them).  This returns the cached memory to the driver
 * (synchronises the device).  Buffers of a handle are freed in the order of the stream the handle was created on. */
SFEM_API int sfem_release_cached_memory(void);
/* Grows the pool by `bytes` in ONE piece ahead of a large set-up (one driver mapping instead of the thousands the set-up's
 * own allocations would trigger one by one: the patterns of UnitCube(203), 65 GB, take 13 s to allocate piecemeal). */
SFEM_API int sfem_reserve_cached_memory(int64_t bytes, void* stream);
/* Bytes the pool currently holds from the driver / bytes handed out to live buffers (diagnostics). */
SFEM_API int sfem_cached_memory_info(int64_t* reserved_bytes, int64_t* used_bytes);

/* ---- mesh: replaces dolfin.Mesh / UnitCube / UnitSquare as inputs of the path -------------------
 * (examples/resonant_cavity/driver.py:34, sucemfem/Utilities/MeshGenerators.py:32-36).
 * gdim = 2 (triangles) or 3 (tetrahedra). d_coords: double [n_vertices][gdim]; d_cells: int32
 * [n_cells][gdim+1], each cell's vertex ids ascending (UFC order; checked, error otherwise). */
SFEM_API int sfem_mesh_create(int gdim, int64_t n_vertices, const double* d_coords, int64_t n_cells,
                     const int32_t* d_cells, void* stream, sfem_mesh_t* out);
SFEM_API int sfem_mesh_destroy(sfem_mesh_t mesh);

/* Fill a structured UnitCube(nx,ny,nz) / UnitSquare(nx,ny) mesh on the device (DOLFIN 1.0 layout,
 * SURVEY 8.1-M): d_coords double [(nx+1)(ny+1)(nz+1)][3], d_cells int32 [6 nx ny nz][4] (nz = 0: the
 * 2-D right-diagonal square, [2 nx ny][3]).  Coordinates are scale*(i/n) + shift per axis. */
SFEM_API int sfem_fill_unit_mesh(int nx, int ny, int nz, const double* scale3, const double* shift3,
                        double* d_coords, int32_t* d_cells, void* stream);

/* ---- function space + dof map: replaces dolfin.FunctionSpace(mesh, "Nedelec 1st kind H(curl)", order)
 * (sucemfem/ProblemConfigurations/EMProblem.py:36,100-105): numbers edges (and faces) by first
 * appearance, builds the cell->entity tables and the entity->cell incidence lists (sort/scan kernels). */
SFEM_API int sfem_space_create(sfem_mesh_t mesh, int order, void* stream, sfem_space_t* out);
SFEM_API int sfem_space_destroy(sfem_space_t space);
/* global_dimension() (EMProblem.py:44-47), #edges, #faces (0 in 2-D), dofs per cell */
SFEM_API int sfem_space_info(sfem_space_t space, int64_t* ndof, int64_t* n_edges, int64_t* n_faces, int* ndof_cell);
/* dofmap.tabulate_dofs for every cell (sucemfem/Sources/point_source.py:81): int32 [n_cells][ndof_cell] */
SFEM_API int sfem_space_cell_dofs(sfem_space_t space, int32_t* d_out, void* stream);
/* cell->edge int32 [n_cells][6|3] and (3-D) cell->face int32 [n_cells][4] tables (either may be NULL) */
SFEM_API int sfem_space_cell_entities(sfem_space_t space, int32_t* d_cell_edges, int32_t* d_cell_faces, void* stream);
/* number of facets (faces in 3-D, edges in 2-D) = size of a facet MeshFunction */
SFEM_API int sfem_space_num_facets(sfem_space_t space, int64_t* n_facets);

/* DomainBoundary().mark(mesh_function, value) (sucemfem/BoundaryConditions/essential.py:93-97):
 * d_markers uint32 [n_facets]; facets with exactly one incident cell are set to `value`, others untouched. */
SFEM_API int sfem_mark_exterior_facets(sfem_space_t space, uint32_t* d_markers, uint32_t value, void* stream);
/* dofs constrained by DirichletBC(V, zero, mesh_function, region) (essential.py:77-80): sets
 * d_row_flags[dof] = 1 (uint8 [ndof], not cleared first) for every dof in the closure of a facet whose
 * marker == region. */
SFEM_API int sfem_pec_row_flags(sfem_space_t space, const uint32_t* d_markers, uint32_t region, uint8_t* d_row_flags,
                       void* stream);

/* Materials: coef[c] = value_of_region[region[c]] (sucemfem/Materials.py:135-146).  d_regions uint32
 * [n_cells]; region_ids/values host arrays of length n_regions.  On an unknown region returns non-zero
 * and stores that region number in *bad_region (caller raises ValueError like Materials.py:147-148). */
SFEM_API int sfem_expand_material(int64_t n_cells, const uint32_t* d_regions, int n_regions, const uint32_t* region_ids,
                         const double* values, double* d_out, uint32_t* bad_region, void* stream);

/* Partitioned mesh (multi-GPU, BASELINE north_star: "the mesh is partitioned ... ghost elements duplicated"): `space`
 * was created from ONE RANK's cells -- its own contiguous range of the global cell list plus every other cell that
 * shares a vertex with them (halo), in global cell order, with locally renumbered vertices.  First-appearance numbering
 * of such a sub-mesh gives the entities that first appear in the rank's OWN cells one contiguous local index range, in
 * the global order; their global indices are that range shifted (offsets from an all-gather of the per-rank counts).
 * This call attaches the global numbering: d_edge_gid[E_local] / d_face_gid[F_local] = global edge / face index of
 * every local entity (owned: the shift; others: looked up in the owners' tables -- sucemfem/partition.py), plus the
 * first owned local / global index of each class.  From then on patterns of this space report GLOBAL column ids, sorted
 * by them, and GLOBAL row numbers (sfem_pattern_info, sfem_pattern_export, sfem_abc_export); entity ranges passed to
 * sfem_pattern_create stay LOCAL (the owned ranges), as do the PEC flag and marker arrays.  Rows of owned entities are
 * bit-identical to the same rows assembled from the whole mesh.  Tetrahedra only; a row block may not mix edge rows
 * and face rows.  No counterpart in the reference (serial). */
SFEM_API int sfem_space_set_global_ids(sfem_space_t space, const int32_t* d_edge_gid, const int32_t* d_face_gid,
                                       int64_t n_edges_global, int64_t n_faces_global, int64_t own_edge_lo_local,
                                       int64_t own_edge_lo_global, int64_t own_face_lo_local,
                                       int64_t own_face_lo_global, void* stream);

/* ---- sparsity pattern: replaces DOLFIN's SparsityPatternBuilder + tensor init inside dolfin.assemble
 * (sucemfem/SystemMatrices.py:54).  Rows are owned by DOF-entity: the pattern covers the rows of the
 * entities [ent_lo, ent_hi) of the combined entity numbering (edges, then faces / cell interiors), i.e.
 * dof rows [order*ent_lo, order*ent_hi).  ent_lo = 0, ent_hi = -1 means all rows.  Multi-GPU: each rank
 * builds the block it owns; assembly of a block needs no communication. */
SFEM_API int sfem_pattern_create(sfem_space_t space, int64_t ent_lo, int64_t ent_hi, void* stream, sfem_pattern_t* out);
SFEM_API int sfem_pattern_destroy(sfem_pattern_t pat);
SFEM_API int sfem_pattern_info(sfem_pattern_t pat, int64_t* row_begin, int64_t* n_rows, int64_t* nnz);
/* CSR arrays as returned by A.data() (sucemfem/Utilities/Converters.py:37): d_rowptr int64 [n_rows+1]
 * (offsets into this block), d_cols int32 [nnz] global column ids ascending within a row. */
SFEM_API int sfem_pattern_export(sfem_pattern_t pat, int64_t* d_rowptr, int32_t* d_cols, void* stream);
/* Split [0, n_entities) into nparts contiguous entity ranges of (nearly) equal nnz; host int64 [nparts+1]. */
SFEM_API int sfem_partition_rows(sfem_space_t space, int nparts, int64_t* ent_bounds, void* stream);
/* The same split inside each entity class: edge_bounds[p..p+1] is a range of [0, n_edges), other_bounds[p..p+1]
 * a range of [n_edges, n_entities) (faces in 3-D, cell interiors in 2-D at order 2; empty at order 1), each
 * balanced by incidence count; host int64 [nparts+1] each.  A rank that owns one range of each class (two
 * patterns) runs the same mix of edge-row and face-row work as every other rank. */
SFEM_API int sfem_partition_rows_by_class(sfem_space_t space, int nparts, int64_t* edge_bounds, int64_t* other_bounds,
                                 void* stream);

/* ---- values: replaces dolfin.assemble(form, tensor=mat) + apply_essential(mat)
 * (sucemfem/SystemMatrices.py:54-55).  d_coef: per-cell double [n_cells] (eps_r for MASS, mu_r for
 * STIFFNESS; NULL = 1.0; ignored for ABC).  d_pec_row_flags: uint8 [ndof] or NULL; flagged rows are
 * written as unit rows (row-only ident, pattern kept).  d_values: double [nnz of the pattern block].
 * For ABC, d_facet_mask uint8 [n_facets] or NULL selects the facets to integrate over (NULL = all
 * exterior facets, what ABC.py:32-33 does). */
SFEM_API int sfem_assemble(sfem_pattern_t pat, int form, const double* d_coef, const uint8_t* d_facet_mask,
                  const uint8_t* d_pec_row_flags, double* d_values, int kernel, void* stream);
/* Fused S and M(=T) in one pass over the rows (the benchmarked hot path). */
SFEM_API int sfem_assemble_ST(sfem_pattern_t pat, const double* d_eps_r, const double* d_mu_r,
                     const uint8_t* d_pec_row_flags, double* d_S, double* d_T, int kernel, void* stream);

/* number of kernel launches one sfem_assemble_ST call issues with the given kernel variant (bench accounting) */
SFEM_API int sfem_launches_per_assemble_ST(sfem_pattern_t pat, int kernel);

/* DirichletBC.apply(A) -> GenericMatrix::ident (sucemfem/BoundaryConditions/essential.py:77-81, called from
 * sucemfem/BoundaryConditions/container.py:39-42) as a separate step on already assembled values:
 * flagged rows of the block become unit rows, pattern kept, columns untouched. */
SFEM_API int sfem_apply_ident(sfem_pattern_t pat, const uint8_t* d_row_flags, double* d_values, void* stream);

/* Diagonal of sum_m c_m A_m (complex, interleaved re/im, one entry per row of the pattern block): the Jacobi
 * preconditioner of the device Krylov solvers (scipy.sparse.spdiags(1./A.diagonal()),
 * /root/reference/sucemfem/Utilities/LinalgSolvers.py:115).  n_mats / coef_re_im as in sfem_matvec. */
SFEM_API int sfem_diagonal(sfem_pattern_t pat, int n_mats, const double* const* d_values, const double* coef_re_im,
                           double* d_out_complex, void* stream);

/* ---- consumers that stay on the device (solvers of the reference, sucemfem/Utilities/LinalgSolvers.py) ----------
 * y = sum_m c_m (A_m x) for n_mats <= 4 matrices that share the pattern `pat` (d_values: HOST array of n_mats DEVICE
 * pointers, each double [nnz]); coef_re_im: HOST array [2*n_mats] (re, im) or NULL = all 1.  complex_vectors != 0:
 * d_x complex128 [ndof], d_y complex128 [n_rows of the block] (interleaved re, im); else real arrays and real c_m.
 * This is `M*x` (EMVectorWaveEigenproblem.py:86,98) and A = S - k0^2 M + j k0 S_0 (EMDrivenProblem.py:67-72) applied
 * without forming A. */
SFEM_API int sfem_matvec(sfem_pattern_t pat, int n_mats, const double* const* d_values, const double* coef_re_im,
                int complex_vectors, const double* d_x, double* d_y, void* stream);
/* BiCGStab on sum_m c_m A_m (scipy.sparse.linalg.bicgstab as called by BiCGStabSolver, LinalgSolvers.py:239-249;
 * solve_sparse_system, :303-322).  d_b complex128 [ndof]; d_x complex128 [ndof]: initial guess in, solution out.
 * preconditioner 0: none, 1: diagonal (1/diag(A), LinalgSolvers.py:115).  Stops when ||b - A x|| <= rtol ||b||.
 * Returns 0 on convergence, 6 when maxiter was reached (x = last iterate), 5 on breakdown.  Needs the pattern of
 * all rows.  Synchronises the stream every iteration (the scalars of the recurrence are reduced on the device and
 * read back). */
SFEM_API int sfem_bicgstab(sfem_pattern_t pat, int n_mats, const double* const* d_values, const double* coef_re_im,
                  const double* d_b, double* d_x, int preconditioner, double rtol, int maxiter, int* iterations,
                  double* relres, void* stream);

/* Compact storage of the ABC matrix S_0 (/root/reference/sucemfem/BoundaryConditions/ABC.py:56-57): only the rows of
 * DOF entities that lie ON an exterior facet are non-zero (a basis function has no tangential trace on a facet that
 * does not contain its entity), so only those rows are stored, each with the columns of the same row of the S/M
 * pattern: compact values == the boundary rows of the full-pattern S_0 of sfem_assemble(SFEM_FORM_ABC), bit for bit,
 * and no nnz-sized array has to be cleared.  Dirichlet rows inside the set become unit rows; Dirichlet rows outside
 * it exist only in the full form.
 *   sfem_abc_info     dof rows and non-zeros of the compact storage (builds the row list on first use)
 *   sfem_abc_export   d_rows[n_rows] global dof row numbers, d_rowptr[n_rows+1], d_cols[nnz] (or NULL)
 *   sfem_assemble_abc values into d_values_compact[nnz]; d_facet_mask as in sfem_assemble
 * Consumers: OR SFEM_MATS_LAST_IS_ABC_COMPACT into n_mats of sfem_matvec / sfem_bicgstab when the LAST value array is
 * in this storage (A = S - k0^2 M + j k0 S_0 applied without a full-pattern S_0,
 * /root/reference/sucemfem/ProblemConfigurations/EMDrivenProblem.py:67-72). */
#define SFEM_MATS_LAST_IS_ABC_COMPACT 0x100
SFEM_API int sfem_abc_info(sfem_pattern_t pat, void* stream, int64_t* n_rows, int64_t* nnz);
SFEM_API int sfem_abc_export(sfem_pattern_t pat, int64_t* d_rows, int64_t* d_rowptr, int32_t* d_cols, void* stream);
SFEM_API int sfem_assemble_abc(sfem_pattern_t pat, const uint8_t* d_facet_mask, const uint8_t* d_pec_row_flags,
                               double* d_values_compact, void* stream);

/* Sustained fp64 FMA rate of the current device in TFLOP/s (8 independent DFMA chains per thread on every SM): the
 * denominator of the fp64 roofline bench.py reports; no counterpart in the reference (measurement support). */
SFEM_API int sfem_measure_fp64_peak(double* tflops, void* stream);

/* ---- host-buffer entry points: the whole path with HOST arrays in and out --------------------------------------
 * What a host consumer gets from the reference today: FunctionSpace(mesh, family, order) (EMProblem.py:100-105),
 * assemble(M), assemble(S) [, assemble(S_0)] + apply_essential (SystemMatrices.py:54-55) and A.data()
 * (Utilities/Converters.py:37).  All pointers of these calls are HOST pointers; page-locked memory
 * (sfem_host_alloc, cudaHostRegister, torch pin_memory) lets the copies run at PCIe rate and asynchronously.
 *
 * sfem_host_job_create: copies the mesh (h_coords double [n_vertices][gdim], h_cells int32 [n_cells][gdim+1], UFC
 *   order) and the per-cell coefficients (h_eps_r / h_mu_r double [n_cells] or NULL = 1.0) to the device, numbers the
 *   entities, marks the PEC rows and counts the pattern; returns ndof and nnz so that the caller can size the outputs.
 *   pec_mode 0: no essential condition; 1: all exterior facets (PECWallsBoundaryCondition, essential.py:86-99);
 *   2: facets with h_facet_markers[f] == pec_region (uint32 [n_facets], first-appearance facet numbering).
 * sfem_host_job_run: streams the rows in n_blocks blocks of equal non-zeros (<= 0: chosen by the library): pattern
 *   fill + values + CSR expansion of block k on `stream` overlap the device->host copy of block k-1.  Outputs:
 *   h_rowptr int64 [ndof+1], h_cols int32 [nnz], h_S / h_T / h_S0 double [nnz]; any of cols/S/T/S0 may be NULL
 *   (skipped).  Returns when all outputs are complete.  Device memory in flight: two blocks. */
SFEM_API int sfem_host_job_create(int gdim, int64_t n_vertices, const double* h_coords, int64_t n_cells,
                         const int32_t* h_cells, int order, const double* h_eps_r, const double* h_mu_r, int pec_mode,
                         const uint32_t* h_facet_markers, uint32_t pec_region, void* stream, int64_t* ndof,
                         int64_t* nnz, sfem_host_job_t* out);
SFEM_API int sfem_host_job_run(sfem_host_job_t job, int64_t* h_rowptr, int32_t* h_cols, double* h_S, double* h_T,
                      double* h_S0, int n_blocks);
SFEM_API int sfem_host_job_destroy(sfem_host_job_t job);
/* Multi-GPU use of a job: every rank creates the job from the same mesh, asks the job's space for its row ranges
 * (sfem_partition_rows_by_class), sizes its outputs with sfem_space_rows_nnz and streams only the rows of the entities
 * [ent_lo, ent_hi): h_rowptr int64 [order*(ent_hi-ent_lo)+1] (offsets relative to the range), h_cols / values sized by
 * the range's nnz.  No communication between the ranks. */
SFEM_API int sfem_host_job_space(sfem_host_job_t job, sfem_space_t* space);
SFEM_API int sfem_host_job_run_rows(sfem_host_job_t job, int64_t ent_lo, int64_t ent_hi, int64_t* h_rowptr, int32_t* h_cols,
                           double* h_S, double* h_T, double* h_S0, int n_blocks);
/* The same streaming hand-off for a space that already lives on the device (what the sucemfem mirror calls when the
 * MatrixClass is the host matrix, EMDrivenProblem.py:82-87): sfem_space_count_nnz sizes the outputs (one count pass
 * over all rows, cached in the space), sfem_space_assemble_to_host streams the row blocks.  d_* are DEVICE arrays
 * (d_eps_r / d_mu_r double [n_cells] or NULL = 1.0; d_pec_row_flags uint8 [ndof] or NULL), h_* HOST arrays as above. */
SFEM_API int sfem_space_count_nnz(sfem_space_t space, void* stream, int64_t* nnz);
SFEM_API int sfem_space_assemble_to_host(sfem_space_t space, const double* d_eps_r, const double* d_mu_r,
                                const uint8_t* d_pec_row_flags, int64_t* h_rowptr, int32_t* h_cols, double* h_S,
                                double* h_T, double* h_S0, int n_blocks, void* stream);
/* the same for the rows of the entities [ent_lo, ent_hi) only (ent_hi = -1: to the end); outputs relative to the range */
SFEM_API int sfem_space_rows_nnz(sfem_space_t space, int64_t ent_lo, int64_t ent_hi, void* stream, int64_t* nnz);
SFEM_API int sfem_space_assemble_rows_to_host(sfem_space_t space, int64_t ent_lo, int64_t ent_hi, const double* d_eps_r,
                                     const double* d_mu_r, const uint8_t* d_pec_row_flags, int64_t* h_rowptr,
                                     int32_t* h_cols, double* h_S, double* h_T, double* h_S0, int n_blocks,
                                     void* stream);
/* page-locked host memory for the arrays above (cudaHostAlloc / cudaFreeHost) */
SFEM_API int sfem_host_alloc(size_t bytes, void** out);
SFEM_API int sfem_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* SUCEM_B200_H */
