// C ABI of libsucem_b200.so (declared in include/sucem_b200.h) and the small device helpers that sit
// either side of the assembly kernels: structured mesh fill, exterior-facet marking, PEC row flags,
// material expansion, row partitioning.
#include <cub/cub.cuh>

#include <vector>

#include "sfem_internal.h"

namespace sfem {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

cudaError_t dev_malloc_bytes(void** p, size_t bytes, cudaStream_t st) {
  static bool pool_ready[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && !pool_ready[dev]) {  // idempotent: a race only repeats the same attribute write
    cudaMemPool_t pool;
    e = cudaDeviceGetDefaultMemPool(&pool, dev);
    if (e != cudaSuccess) return e;
    uint64_t keep = ~0ull;
    e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    if (e != cudaSuccess) return e;
    pool_ready[dev] = true;
  }
  return cudaMallocAsync(p, bytes ? bytes : 16, st);
}

namespace {
inline unsigned grid_for(int64_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

// UnitCube / UnitSquare of DOLFIN 1.0: vertex id = ix + (nx+1) iy + (nx+1)(ny+1) iz; cube
// (ix,iy,iz), x fastest, emits six tets in a fixed order, each vertex-sorted.
__global__ void k_fill_coords(int nx, int ny, int nz, double sx, double sy, double sz, double tx, double ty,
                              double tz, double* __restrict__ coords, int64_t V) {
  int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= V) return;
  int64_t ix = v % (nx + 1), iy = (v / (nx + 1)) % (ny + 1), iz = v / ((int64_t)(nx + 1) * (ny + 1));
  if (nz > 0) {
    coords[v * 3 + 0] = ((double)ix / (double)nx) * sx + tx;
    coords[v * 3 + 1] = ((double)iy / (double)ny) * sy + ty;
    coords[v * 3 + 2] = ((double)iz / (double)nz) * sz + tz;
  } else {
    coords[v * 2 + 0] = ((double)ix / (double)nx) * sx + tx;
    coords[v * 2 + 1] = ((double)iy / (double)ny) * sy + ty;
  }
}

__device__ __forceinline__ void sort4(int32_t* v) {
#define SFEM_CSWAP(a, b) { int32_t lo = min(v[a], v[b]), hi = max(v[a], v[b]); v[a] = lo; v[b] = hi; }
  SFEM_CSWAP(0, 1) SFEM_CSWAP(2, 3) SFEM_CSWAP(0, 2) SFEM_CSWAP(1, 3) SFEM_CSWAP(1, 2)
#undef SFEM_CSWAP
}

__global__ void k_fill_cells3(int nx, int ny, int nz, int32_t* __restrict__ cells, int64_t ncube) {
  int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (q >= ncube) return;
  int64_t ix = q % nx, iy = (q / nx) % ny, iz = q / ((int64_t)nx * ny);
  const int32_t sx = 1, sy = nx + 1, sz = (nx + 1) * (ny + 1);
  const int32_t v0 = (int32_t)(ix + sy * iy + (int64_t)sz * iz);
  const int32_t cv[8] = {v0, v0 + sx, v0 + sy, v0 + sy + sx, v0 + sz, v0 + sz + sx, v0 + sz + sy, v0 + sz + sy + sx};
  const int t[6][4] = {{0, 1, 3, 7}, {0, 1, 7, 5}, {0, 5, 7, 4}, {0, 3, 2, 7}, {0, 6, 4, 7}, {0, 2, 6, 7}};
  for (int k = 0; k < 6; ++k) {
    int32_t v[4] = {cv[t[k][0]], cv[t[k][1]], cv[t[k][2]], cv[t[k][3]]};
    sort4(v);
    *reinterpret_cast<int4*>(cells + (q * 6 + k) * 4) = make_int4(v[0], v[1], v[2], v[3]);
  }
}

__global__ void k_fill_cells2(int nx, int ny, int32_t* __restrict__ cells, int64_t nsq) {
  int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (q >= nsq) return;
  int64_t ix = q % nx, iy = q / nx;
  const int32_t v0 = (int32_t)(ix + (int64_t)(nx + 1) * iy), v1 = v0 + 1, v2 = v0 + nx + 1, v3 = v2 + 1;
  int32_t* o = cells + q * 6;
  o[0] = v0; o[1] = v1; o[2] = v3;   // right diagonal: (v0,v1,v3), (v0,v2,v3) -- already ascending
  o[3] = v0; o[4] = v2; o[5] = v3;
}

__global__ void k_mark_exterior(const uint8_t* __restrict__ ncells, uint32_t* __restrict__ markers, uint32_t value,
                                int64_t n) {
  int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f < n && ncells[f] == 1) markers[f] = value;
}

// closure dofs of marked facets: tet face -> its 3 edges (q dofs each) + q(q-1) face dofs; triangle edge -> q dofs
__global__ void k_pec_flags(const uint32_t* __restrict__ markers, uint32_t region, int dim, int q, int64_t E,
                            const int32_t* __restrict__ first_cell, const uint8_t* __restrict__ first_loc,
                            const int32_t* __restrict__ cell_edges, uint8_t* __restrict__ flags, int64_t n) {
  int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= n || markers[f] != region) return;
  if (dim == 2) {
    for (int k = 0; k < q; ++k) flags[q * f + k] = 1;
    return;
  }
  const int64_t c = first_cell[f];
  const int lf = first_loc[f];
  for (int j = 0; j < 3; ++j) {
    const int64_t e = cell_edges[c * 6 + tet_face_edge(lf, j)];
    for (int k = 0; k < q; ++k) flags[q * e + k] = 1;
  }
  if (q == 2) {
    flags[2 * (E + f)] = 1;
    flags[2 * (E + f) + 1] = 1;
  }
}

__global__ void k_expand_material(const uint32_t* __restrict__ regions, int nreg, const uint32_t* __restrict__ ids,
                                  const double* __restrict__ vals, double* __restrict__ out, int64_t C,
                                  unsigned long long* __restrict__ bad) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= C) return;
  const uint32_t r = regions[c];
  for (int i = 0; i < nreg; ++i)
    if (ids[i] == r) {
      out[c] = vals[i];
      return;
    }
  atomicMin(bad, ((unsigned long long)c << 32) | r);  // first offending cell wins, like the reference's loop
}

__global__ void k_cell_dofs(SpaceView sv, int q, int64_t C, int32_t* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int nd = q * sv.n_ent_cell;
  if (i >= C * nd) return;
  int64_t c = i / nd;
  int j = (int)(i - c * nd), u = j / q, k = j - u * q;
  out[i] = q * sv.ent(c, u) + k;
}

// Row partitioning: weight of an entity row = its number of incident cells (proportional to its nnz and to the work
// of the value kernels); edge rows can be weighted kEdgeWeight16/16 relative to face rows (edges visit their cells in a
// dependent ring, faces write more bytes per incidence).
constexpr int64_t kEdgeWeight16 = 16;  // measured equal per incidence with the current kernels (r1: 1.71 ms / 6C vs 1.13 ms / 4C at n = 70)
__device__ __forceinline__ __int128 row_weight_prefix(const int64_t* __restrict__ inc_ptr, int64_t E, int64_t g) {
  const __int128 e = (__int128)(inc_ptr[g < E ? g : E] - inc_ptr[0]) * kEdgeWeight16;
  const __int128 f = g > E ? (__int128)(inc_ptr[g] - inc_ptr[E]) * 16 : 0;
  return e + f;
}
__global__ void k_find_bounds(const int64_t* __restrict__ inc_ptr, int64_t E, int64_t n, int nparts,
                              int64_t* __restrict__ out) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > nparts) return;
  const __int128 total = row_weight_prefix(inc_ptr, E, n);
  const __int128 target = total * p / nparts;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (row_weight_prefix(inc_ptr, E, mid) < target) lo = mid + 1; else hi = mid;
  }
  out[p] = p == nparts ? n : lo;
}

// the same split inside one entity class [g0, g1) (edges, or faces / cell interiors): rank p then owns one range of
// each class, so every rank runs the same mix of edge-row and face-row work
__global__ void k_find_bounds_class(const int64_t* __restrict__ inc_ptr, int64_t g0, int64_t g1, int nparts,
                                    int64_t* __restrict__ out) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > nparts) return;
  const __int128 total = inc_ptr[g1] - inc_ptr[g0];
  const int64_t target = inc_ptr[g0] + (int64_t)(total * p / nparts);
  int64_t lo = g0, hi = g1;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (inc_ptr[mid] < target) lo = mid + 1; else hi = mid;
  }
  out[p] = p == nparts ? g1 : (p == 0 ? g0 : lo);
}

// row-only ident on assembled values: one warp per row entity
__global__ void k_ident(const int64_t* __restrict__ ent_rowptr, const int32_t* __restrict__ ent_cols, int q,
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

}  // namespace
__global__ void k_offset_i32(const int32_t* __restrict__ in, int32_t* __restrict__ out, int64_t n, int32_t off) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] + off;
}

// dof-level CSR of the compact ABC storage: global dof row numbers, row pointer, (optionally) column ids
__global__ void k_abc_export(const int64_t* __restrict__ ent_rowptr, const int32_t* __restrict__ ent_cols,
                             const int64_t* __restrict__ list, const int64_t* __restrict__ out_rowptr, int64_t n_list,
                             int q, int64_t ent_lo_global, int64_t* __restrict__ rows, int64_t* __restrict__ rowptr,
                             int32_t* __restrict__ cols) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = blockIdx.x * (int64_t)(blockDim.x >> 5) + w;
  if (i > n_list) return;
  if (i == n_list) {
    if (lane == 0) rowptr[q * i] = (int64_t)q * q * out_rowptr[i];
    return;
  }
  const int64_t r = list[i], a = ent_rowptr[r], ao = out_rowptr[i];
  const int len = (int)(ent_rowptr[r + 1] - a);
  for (int k = 0; k < q; ++k) {
    const int64_t rs = (int64_t)q * q * ao + (int64_t)k * q * len;
    if (lane == 0) {
      rows[q * i + k] = q * (ent_lo_global + r) + k;
      rowptr[q * i + k] = rs;
    }
    if (cols)
      for (int idx = lane; idx < q * len; idx += 32) {
        const int s = idx / q, kp = idx - s * q;
        cols[rs + idx] = q * ent_cols[a + s] + kp;
      }
  }
}

}  // namespace sfem

using namespace sfem;

extern "C" {

const char* sfem_last_error(void) { return g_last_error.c_str(); }
const char* sfem_version(void) { return "sucem_b200 0.1 sm_100a"; }

int sfem_mesh_create(int gdim, int64_t n_vertices, const double* d_coords, int64_t n_cells,
                     const int32_t* d_cells, void* stream, sfem_mesh_t* out) {
  (void)stream;
  SFEM_CHECK(out != nullptr, "out is NULL");
  SFEM_CHECK(gdim == 2 || gdim == 3, "gdim must be 2 or 3");
  SFEM_CHECK(n_vertices > 0 && n_cells > 0 && d_coords && d_cells, "empty mesh or NULL arrays");
  SFEM_CHECK(n_vertices < (int64_t)2147483647, "too many vertices for int32 connectivity");
  int dev = 0;
  SFEM_CUDA(cudaGetDevice(&dev));
  auto* m = new sfem_mesh_s();
  m->device = dev;
  m->gdim = gdim;
  m->nv = gdim + 1;
  m->V = n_vertices;
  m->C = n_cells;
  m->coords = d_coords;
  m->cells = d_cells;
  *out = m;
  return 0;
}

int sfem_release_cached_memory(void) {
  int dev = 0;
  SFEM_CUDA(cudaGetDevice(&dev));
  cudaMemPool_t pool;
  SFEM_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
  SFEM_CUDA(cudaDeviceSynchronize());
  SFEM_CUDA(cudaMemPoolTrimTo(pool, 0));
  return 0;
}

int sfem_reserve_cached_memory(int64_t bytes, void* stream) {
  SFEM_CHECK(bytes >= 0, "negative size");
  if (bytes == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  void* p = nullptr;
  SFEM_CUDA(dev_malloc_bytes(&p, (size_t)bytes, st));   // grows the pool in one piece ...
  dev_free(p, st);                                      // ... which stays cached (release threshold: never)
  SFEM_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int sfem_cached_memory_info(int64_t* reserved_bytes, int64_t* used_bytes) {
  int dev = 0;
  SFEM_CUDA(cudaGetDevice(&dev));
  cudaMemPool_t pool;
  SFEM_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
  uint64_t r = 0, u = 0;
  SFEM_CUDA(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &r));
  SFEM_CUDA(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &u));
  if (reserved_bytes) *reserved_bytes = (int64_t)r;
  if (used_bytes) *used_bytes = (int64_t)u;
  return 0;
}

int sfem_mesh_destroy(sfem_mesh_t mesh) {
  delete mesh;
  return 0;
}

int sfem_fill_unit_mesh(int nx, int ny, int nz, const double* scale3, const double* shift3, double* d_coords,
                        int32_t* d_cells, void* stream) {
  SFEM_CHECK(nx > 0 && ny > 0 && nz >= 0 && d_coords && d_cells, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const double s[3] = {scale3 ? scale3[0] : 1.0, scale3 ? scale3[1] : 1.0, scale3 ? scale3[2] : 1.0};
  const double t[3] = {shift3 ? shift3[0] : 0.0, shift3 ? shift3[1] : 0.0, shift3 ? shift3[2] : 0.0};
  const int64_t V = (int64_t)(nx + 1) * (ny + 1) * (nz + 1);
  SFEM_CHECK(V < (int64_t)2147483647, "too many vertices for int32 connectivity");
  k_fill_coords<<<grid_for(V), 256, 0, st>>>(nx, ny, nz, s[0], s[1], s[2], t[0], t[1], t[2], d_coords, V);
  if (nz > 0) {
    const int64_t nc = (int64_t)nx * ny * nz;
    k_fill_cells3<<<grid_for(nc), 256, 0, st>>>(nx, ny, nz, d_cells, nc);
  } else {
    const int64_t nc = (int64_t)nx * ny;
    k_fill_cells2<<<grid_for(nc), 256, 0, st>>>(nx, ny, d_cells, nc);
  }
  SFEM_CUDA(cudaGetLastError());
  return 0;
}

int sfem_space_create(sfem_mesh_t mesh, int order, void* stream, sfem_space_t* out) {
  SFEM_CHECK(mesh && out, "NULL argument");
  SFEM_CHECK(order == 1 || order == 2, "only N1curl order 1 and 2 are supported");
  DeviceGuard dg(mesh->device);
  auto* sp = new sfem_space_s();
  sp->mesh = mesh;
  sp->device = mesh->device;
  sp->stream = (cudaStream_t)stream;
  sp->order = order;
  int rc = build_numbering(sp, (cudaStream_t)stream);
  if (rc != 0) {
    sfem_space_destroy(sp);
    return rc;
  }
  *out = sp;
  return 0;
}

int sfem_space_destroy(sfem_space_t sp) {
  if (!sp) return 0;
  DeviceGuard dg(sp->device);
  cudaStream_t st = sp->stream;
  dev_free(sp->cell_edges, st);
  dev_free(sp->cell_faces, st);
  dev_free(sp->inc_ptr, st);
  dev_free(sp->inc_cell, st);
  dev_free(sp->inc_loc, st);
  dev_free(sp->facet_ncells, st);
  dev_free(sp->facet_first_cell, st);
  dev_free(sp->facet_first_loc, st);
  dev_free(sp->cell_extmask, st);
  dev_free(sp->coords4, st);
  dev_free(sp->ones, st);
  dev_free(sp->coef2, st);
  dev_free(sp->gid_e, st);
  dev_free(sp->gid_f, st);
  dev_free(sp->rowptr_all, st);
  if (sp->drain) {
    cudaStreamSynchronize(sp->drain);
    cudaStreamDestroy(sp->drain);
  }
  delete sp;
  return 0;
}

int sfem_space_info(sfem_space_t sp, int64_t* ndof, int64_t* n_edges, int64_t* n_faces, int* ndof_cell) {
  SFEM_CHECK(sp, "NULL space");
  if (ndof) *ndof = sp->ndof;
  if (n_edges) *n_edges = sp->E;
  if (n_faces) *n_faces = sp->F;
  if (ndof_cell) *ndof_cell = sp->ndof_cell;
  return 0;
}

int sfem_space_num_facets(sfem_space_t sp, int64_t* n_facets) {
  SFEM_CHECK(sp && n_facets, "NULL argument");
  *n_facets = sp->n_facets;
  return 0;
}

int sfem_space_cell_dofs(sfem_space_t sp, int32_t* d_out, void* stream) {
  SFEM_CHECK(sp && d_out, "NULL argument");
  SFEM_CHECK(sp->ndof < (int64_t)2147483647, "dof ids exceed int32");
  DeviceGuard dg(sp->mesh->device);
  const int64_t n = sp->mesh->C * sp->ndof_cell;
  k_cell_dofs<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(make_view(sp), sp->order, sp->mesh->C, d_out);
  SFEM_CUDA(cudaGetLastError());
  return 0;
}

int sfem_space_cell_entities(sfem_space_t sp, int32_t* d_cell_edges, int32_t* d_cell_faces, void* stream) {
  SFEM_CHECK(sp, "NULL space");
  DeviceGuard dg(sp->mesh->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (d_cell_edges)
    SFEM_CUDA(cudaMemcpyAsync(d_cell_edges, sp->cell_edges, sp->mesh->C * sp->ne_loc * 4, cudaMemcpyDeviceToDevice, st));
  if (d_cell_faces) {
    SFEM_CHECK(sp->cell_faces, "no faces in 2-D");
    SFEM_CUDA(cudaMemcpyAsync(d_cell_faces, sp->cell_faces, sp->mesh->C * 16, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

int sfem_mark_exterior_facets(sfem_space_t sp, uint32_t* d_markers, uint32_t value, void* stream) {
  SFEM_CHECK(sp && d_markers, "NULL argument");
  DeviceGuard dg(sp->mesh->device);
  k_mark_exterior<<<grid_for(sp->n_facets), 256, 0, (cudaStream_t)stream>>>(sp->facet_ncells, d_markers, value,
                                                                             sp->n_facets);
  SFEM_CUDA(cudaGetLastError());
  return 0;
}

int sfem_pec_row_flags(sfem_space_t sp, const uint32_t* d_markers, uint32_t region, uint8_t* d_row_flags,
                       void* stream) {
  SFEM_CHECK(sp && d_markers && d_row_flags, "NULL argument");
  DeviceGuard dg(sp->mesh->device);
  k_pec_flags<<<grid_for(sp->n_facets), 256, 0, (cudaStream_t)stream>>>(
      d_markers, region, sp->mesh->gdim, sp->order, sp->E, sp->facet_first_cell, sp->facet_first_loc,
      sp->cell_edges, d_row_flags, sp->n_facets);
  SFEM_CUDA(cudaGetLastError());
  return 0;
}

int sfem_expand_material(int64_t n_cells, const uint32_t* d_regions, int n_regions, const uint32_t* region_ids,
                         const double* values, double* d_out, uint32_t* bad_region, void* stream) {
  SFEM_CHECK(d_regions && region_ids && values && d_out && n_regions > 0, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* d_ids = nullptr;
  double* d_vals = nullptr;
  unsigned long long* d_bad = nullptr;
  SFEM_CUDA(dev_malloc(&d_ids, n_regions * 4, st));
  SFEM_CUDA(dev_malloc(&d_vals, n_regions * 8, st));
  SFEM_CUDA(dev_malloc(&d_bad, 8, st));
  SFEM_CUDA(cudaMemcpyAsync(d_ids, region_ids, n_regions * 4, cudaMemcpyHostToDevice, st));
  SFEM_CUDA(cudaMemcpyAsync(d_vals, values, n_regions * 8, cudaMemcpyHostToDevice, st));
  SFEM_CUDA(cudaMemsetAsync(d_bad, 0xff, 8, st));
  k_expand_material<<<grid_for(n_cells), 256, 0, st>>>(d_regions, n_regions, d_ids, d_vals, d_out, n_cells, d_bad);
  unsigned long long h_bad = 0;
  SFEM_CUDA(cudaMemcpyAsync(&h_bad, d_bad, 8, cudaMemcpyDeviceToHost, st));
  SFEM_CUDA(cudaStreamSynchronize(st));
  dev_free(d_ids, st);
  dev_free(d_vals, st);
  dev_free(d_bad, st);
  if (h_bad != ~0ull) {
    if (bad_region) *bad_region = (uint32_t)(h_bad & 0xffffffffu);
    set_error("Material number " + std::to_string((uint32_t)(h_bad & 0xffffffffu)) + " not found");
    return 4;
  }
  return 0;
}

int sfem_space_set_global_ids(sfem_space_t sp, const int32_t* d_edge_gid, const int32_t* d_face_gid,
                              int64_t n_edges_global, int64_t n_faces_global, int64_t own_edge_lo_local,
                              int64_t own_edge_lo_global, int64_t own_face_lo_local, int64_t own_face_lo_global,
                              void* stream) {
  SFEM_CHECK(sp && d_edge_gid, "NULL argument");
  SFEM_CHECK(sp->mesh->gdim == 3, "partitioned meshes: tetrahedra only");
  SFEM_CHECK(sp->order == 1 || d_face_gid, "order 2 needs the face map");
  DeviceGuard dg(sp->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (!sp->gid_e) SFEM_CUDA(dev_malloc(&sp->gid_e, std::max<int64_t>(sp->E, 1) * 4, st));
  SFEM_CUDA(cudaMemcpyAsync(sp->gid_e, d_edge_gid, sp->E * 4, cudaMemcpyDeviceToDevice, st));
  if (sp->order == 2) {
    if (!sp->gid_f) SFEM_CUDA(dev_malloc(&sp->gid_f, std::max<int64_t>(sp->F, 1) * 4, st));
    k_offset_i32<<<grid_for(sp->F), 256, 0, st>>>(d_face_gid, sp->gid_f, sp->F, (int32_t)n_edges_global);
  }
  sp->gshift_e = own_edge_lo_global - own_edge_lo_local;
  sp->gshift_f = (n_edges_global + own_face_lo_global) - (sp->E + own_face_lo_local);
  sp->n_ent_global = n_edges_global + (sp->order == 2 ? n_faces_global : 0);
  sp->ndof_global = sp->order * sp->n_ent_global;
  SFEM_CHECK(sp->n_ent_global < (int64_t)2147483647, "global entity ids exceed int32");
  SFEM_CUDA(cudaGetLastError());
  return 0;
}

int sfem_pattern_create(sfem_space_t sp, int64_t ent_lo, int64_t ent_hi, void* stream, sfem_pattern_t* out) {
  SFEM_CHECK(sp && out, "NULL argument");
  if (ent_hi < 0) ent_hi = sp->n_ent;
  SFEM_CHECK(0 <= ent_lo && ent_lo <= ent_hi && ent_hi <= sp->n_ent, "entity range out of bounds");
  SFEM_CHECK(!sp->gid_e || ent_hi <= sp->E || ent_lo >= sp->E,
             "partitioned mesh: a row block holds edge rows or face rows, not both");
  DeviceGuard dg(sp->mesh->device);
  auto* pat = new sfem_pattern_s();
  pat->space = sp;
  pat->device = sp->device;
  pat->stream = (cudaStream_t)stream;
  pat->ent_lo = ent_lo;
  pat->ent_hi = ent_hi;
  if (sp->gid_e) pat->gshift = ent_lo >= sp->E && ent_hi > ent_lo ? sp->gshift_f : sp->gshift_e;
  pat->n_rows_ent = ent_hi - ent_lo;
  int rc = build_pattern(pat, (cudaStream_t)stream);
  if (rc != 0) {
    sfem_pattern_destroy(pat);
    return rc;
  }
  *out = pat;
  return 0;
}

int sfem_pattern_destroy(sfem_pattern_t pat) {
  if (!pat) return 0;
  DeviceGuard dg(pat->device);
  cudaStream_t st = pat->stream;
  sfem::free_pattern_buffers(pat, st);
  delete pat;
  return 0;
}

int sfem_pattern_info(sfem_pattern_t pat, int64_t* row_begin, int64_t* n_rows, int64_t* nnz) {
  SFEM_CHECK(pat, "NULL pattern");
  const int q = pat->space->order;
  if (row_begin) *row_begin = q * (pat->ent_lo + pat->gshift);
  if (n_rows) *n_rows = q * pat->n_rows_ent;
  if (nnz) *nnz = (int64_t)q * q * pat->nnz_ent;
  return 0;
}

int sfem_pattern_export(sfem_pattern_t pat, int64_t* d_rowptr, int32_t* d_cols, void* stream) {
  SFEM_CHECK(pat && d_rowptr, "NULL argument");
  SFEM_CHECK(std::max(pat->space->ndof, pat->space->ndof_global) < (int64_t)2147483647, "column ids exceed int32");
  DeviceGuard dg(pat->space->mesh->device);
  return export_pattern(pat, d_rowptr, d_cols, (cudaStream_t)stream);
}

int sfem_partition_rows(sfem_space_t sp, int nparts, int64_t* ent_bounds, void* stream) {
  SFEM_CHECK(sp && ent_bounds && nparts > 0, "bad arguments");
  DeviceGuard dg(sp->mesh->device);
  cudaStream_t st = (cudaStream_t)stream;
  int64_t* d_out = nullptr;
  SFEM_CUDA(dev_malloc(&d_out, (nparts + 1) * 8, st));
  k_find_bounds<<<grid_for(nparts + 1, 64), 64, 0, st>>>(sp->inc_ptr, std::min(sp->E, sp->n_ent), sp->n_ent, nparts, d_out);
  SFEM_CUDA(cudaMemcpyAsync(ent_bounds, d_out, (nparts + 1) * 8, cudaMemcpyDeviceToHost, st));
  SFEM_CUDA(cudaStreamSynchronize(st));
  dev_free(d_out, st);
  ent_bounds[0] = 0;
  return 0;
}

int sfem_partition_rows_by_class(sfem_space_t sp, int nparts, int64_t* edge_bounds, int64_t* other_bounds, void* stream) {
  SFEM_CHECK(sp && edge_bounds && other_bounds && nparts > 0, "bad arguments");
  DeviceGuard dg(sp->mesh->device);
  cudaStream_t st = (cudaStream_t)stream;
  int64_t* d_out = nullptr;
  SFEM_CUDA(dev_malloc(&d_out, 2 * (nparts + 1) * 8, st));
  const int64_t E = std::min(sp->E, sp->n_ent);
  k_find_bounds_class<<<grid_for(nparts + 1, 64), 64, 0, st>>>(sp->inc_ptr, 0, E, nparts, d_out);
  k_find_bounds_class<<<grid_for(nparts + 1, 64), 64, 0, st>>>(sp->inc_ptr, E, sp->n_ent, nparts, d_out + nparts + 1);
  SFEM_CUDA(cudaMemcpyAsync(edge_bounds, d_out, (nparts + 1) * 8, cudaMemcpyDeviceToHost, st));
  SFEM_CUDA(cudaMemcpyAsync(other_bounds, d_out + nparts + 1, (nparts + 1) * 8, cudaMemcpyDeviceToHost, st));
  SFEM_CUDA(cudaStreamSynchronize(st));
  dev_free(d_out, st);
  return 0;
}

// SFEM_KERNEL_AUTO (measured): tetrahedra -> canonical-frame kernels where the pattern allows them
// (UnitCube(119) order 1: 2.2 ms against 4.3 ms thread-per-edge), triangles -> thread-per-entity kernel; order-1 tets
// the canonical kernel refuses (high-valence / non-manifold edges) -> thread-per-edge; everything else -> generic.
static int auto_kernel(const sfem_pattern_s* pat) {
  if (pat->space->mesh->gdim == 3 && canonical_supported(pat)) return SFEM_KERNEL_CANONICAL;
  if (rows_supported(pat)) return SFEM_KERNEL_ROWS;
  return SFEM_KERNEL_GENERIC;
}

int sfem_launches_per_assemble_ST(sfem_pattern_t pat, int kernel) {
  if (!pat) return 0;
  if (kernel == SFEM_KERNEL_AUTO) kernel = auto_kernel(pat);
  if (kernel == SFEM_KERNEL_ROWS) return 1;
  return kernel == SFEM_KERNEL_CANONICAL ? canonical_launches(pat) : 2;  // generic: cell geometry + row gather
}

int sfem_apply_ident(sfem_pattern_t pat, const uint8_t* d_row_flags, double* d_values, void* stream) {
  SFEM_CHECK(pat && d_row_flags && d_values, "NULL argument");
  DeviceGuard dg(pat->space->mesh->device);
  if (pat->n_rows_ent == 0) return 0;
  k_ident<<<(unsigned)((pat->n_rows_ent + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      pat->ent_rowptr, pat->ent_cols, pat->space->order, pat->ent_lo, pat->gshift, pat->n_rows_ent, d_row_flags, d_values);
  SFEM_CUDA(cudaGetLastError());
  return 0;
}

int sfem_assemble(sfem_pattern_t pat, int form, const double* d_coef, const uint8_t* d_facet_mask,
                  const uint8_t* d_pec_row_flags, double* d_values, int kernel, void* stream) {
  SFEM_CHECK(pat && d_values, "NULL argument");
  SFEM_CHECK(form == SFEM_FORM_MASS || form == SFEM_FORM_STIFFNESS || form == SFEM_FORM_ABC, "unknown form");
  SFEM_CHECK(kernel != SFEM_KERNEL_CANONICAL, "the canonical kernels assemble S and M together: use sfem_assemble_ST");
  DeviceGuard dg(pat->space->mesh->device);
  const double* eps = form == SFEM_FORM_MASS ? d_coef : nullptr;
  const double* mu = form == SFEM_FORM_STIFFNESS ? d_coef : nullptr;
  // S_0: the boundary-row kernel unless the generic kernel is asked for (cross-check)
  return assemble_generic(pat, form, eps, mu, d_facet_mask, d_pec_row_flags, d_values, nullptr, (cudaStream_t)stream, false,
                          kernel != SFEM_KERNEL_GENERIC);
}

int sfem_abc_info(sfem_pattern_t pat, void* stream, int64_t* n_rows, int64_t* nnz) {
  SFEM_CHECK(pat, "NULL pattern");
  DeviceGuard dg(pat->space->mesh->device);
  SFEM_TRY(build_abc_rows(pat, (cudaStream_t)stream));
  const int q = pat->space->order;
  if (n_rows) *n_rows = q * pat->n_abc_rows;
  if (nnz) *nnz = (int64_t)q * q * pat->abc_nnz_ent;
  return 0;
}

int sfem_abc_export(sfem_pattern_t pat, int64_t* d_rows, int64_t* d_rowptr, int32_t* d_cols, void* stream) {
  SFEM_CHECK(pat && d_rows && d_rowptr, "NULL argument");
  SFEM_CHECK(!d_cols || pat->space->ndof < (int64_t)2147483647, "column ids exceed int32");
  DeviceGuard dg(pat->space->mesh->device);
  cudaStream_t st = (cudaStream_t)stream;
  SFEM_TRY(build_abc_rows(pat, st));
  const int64_t n = pat->n_abc_rows;
  k_abc_export<<<(unsigned)((n + 1 + 7) / 8), 256, 0, st>>>(pat->ent_rowptr, pat->ent_cols, pat->abc_rows, pat->abc_rowptr, n,
                                                          pat->space->order, pat->ent_lo + pat->gshift, d_rows, d_rowptr, d_cols);
  SFEM_CUDA(cudaGetLastError());
  return 0;
}

int sfem_assemble_abc(sfem_pattern_t pat, const uint8_t* d_facet_mask, const uint8_t* d_pec_row_flags,
                      double* d_values_compact, void* stream) {
  SFEM_CHECK(pat && d_values_compact, "NULL argument");
  DeviceGuard dg(pat->space->mesh->device);
  return assemble_generic(pat, SFEM_FORM_ABC, nullptr, nullptr, d_facet_mask, d_pec_row_flags, d_values_compact, nullptr,
                          (cudaStream_t)stream, true);
}

int sfem_assemble_ST(sfem_pattern_t pat, const double* d_eps_r, const double* d_mu_r,
                     const uint8_t* d_pec_row_flags, double* d_S, double* d_T, int kernel, void* stream) {
  SFEM_CHECK(pat && d_S && d_T, "NULL argument");
  DeviceGuard dg(pat->space->mesh->device);
  if (kernel == SFEM_KERNEL_AUTO) kernel = auto_kernel(pat);
  if (kernel == SFEM_KERNEL_ROWS)
    return assemble_rows_ST(pat, d_eps_r, d_mu_r, d_pec_row_flags, d_S, d_T, (cudaStream_t)stream);
  if (kernel == SFEM_KERNEL_CANONICAL) {
    SFEM_CHECK(canonical_supported(pat), "canonical kernels are unavailable for this pattern (they need order-2 tetrahedra, manifold edges of at most 64 cells and rows of at most 255 entities); use SFEM_KERNEL_GENERIC or SFEM_KERNEL_AUTO");
    return assemble_canonical_ST(pat, d_eps_r, d_mu_r, d_pec_row_flags, d_S, d_T, (cudaStream_t)stream);
  }
  return assemble_generic(pat, 3, d_eps_r, d_mu_r, nullptr, d_pec_row_flags, d_S, d_T, (cudaStream_t)stream);
}

}  // extern "C"
