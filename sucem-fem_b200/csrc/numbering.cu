// K1: device mesh/DOF builder.  Numbers edges and faces by FIRST APPEARANCE over (cell ascending, local
// entity ascending) -- the numbering DOLFIN 1.0 produces for dolfin.FunctionSpace(mesh, "Nedelec 1st kind
// H(curl)", order) (/root/reference/sucemfem/ProblemConfigurations/EMProblem.py:100-105; conventions in
// SURVEY.md 8.1-N, verified there against the reference's shipped facet-marker files) -- with sort/scan
// kernels instead of a serial hash map:
//
//   seq = cell*nloc + local          (the visiting order)
//   stable LSD radix sort of seq by the entity's vertex tuple  -> equal entities form runs, seq ascending
//   run head = first appearance; is_first[seq(head)] = 1; exclusive scan over seq  -> entity index
//   run members scatter (cell, local) into the entity's incidence list, which is therefore ascending in cell
//
// Also produced: facet -> #cells (exterior detection, DomainBoundary().mark of
// /root/reference/sucemfem/BoundaryConditions/essential.py:93-97) and per-cell exterior-facet bit masks.
#include <cub/cub.cuh>

#include "sfem_internal.h"

namespace sfem {

namespace {

template <int NLOC, int NVK>
struct Topo;
template <>
struct Topo<6, 2> {  // tet edges
  __device__ static int v(int l, int k) { return tet_edge_vertex(l, k); }
};
template <>
struct Topo<4, 3> {  // tet faces
  __device__ static int v(int l, int k) { return tet_face_vertex(l, k); }
};
template <>
struct Topo<3, 2> {  // triangle edges
  __device__ static int v(int l, int k) { return tri_edge_vertex(l, k); }
};

__global__ void k_iota(uint32_t* a, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) a[i] = (uint32_t)i;
}

template <int NLOC, int NVK>
__global__ void k_keys(const uint32_t* __restrict__ seq, const int32_t* __restrict__ cells, int nv, int comp,
                       uint32_t* __restrict__ keys, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s = seq[i];
  uint32_t c = s / NLOC, l = s - c * NLOC;
  keys[i] = (uint32_t)cells[(int64_t)c * nv + Topo<NLOC, NVK>::v(l, comp)];
}

// head[p] = tuple(seq[p]) != tuple(seq[p-1]);  is_first[seq[p]] = head[p];  headpos[p] = head ? p : 0
template <int NLOC, int NVK>
__global__ void k_heads(const uint32_t* __restrict__ seq, const int32_t* __restrict__ cells, int nv,
                        int32_t* __restrict__ is_first, int32_t* __restrict__ headpos, int64_t n) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n) return;
  uint32_t s = seq[p];
  bool head = (p == 0);
  if (!head) {
    uint32_t s0 = seq[p - 1];
    uint32_t c = s / NLOC, l = s - c * NLOC, c0 = s0 / NLOC, l0 = s0 - c0 * NLOC;
#pragma unroll
    for (int k = 0; k < NVK; ++k)
      head |= cells[(int64_t)c * nv + Topo<NLOC, NVK>::v(l, k)] != cells[(int64_t)c0 * nv + Topo<NLOC, NVK>::v(l0, k)];
  }
  is_first[s] = head ? 1 : 0;
  headpos[p] = head ? (int32_t)p : 0;
}

// ent = first_idx[seq[headpos[p]]]; cell_x[seq[p]] = ent; count[ent]++
__global__ void k_assign(const uint32_t* __restrict__ seq, const int32_t* __restrict__ headpos,
                         const int32_t* __restrict__ first_idx, int32_t* __restrict__ cell_x,
                         int32_t* __restrict__ ent_of_p, unsigned int* __restrict__ count, int64_t n) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n) return;
  int32_t e = first_idx[seq[headpos[p]]];
  cell_x[seq[p]] = e;
  ent_of_p[p] = e;
  atomicAdd(&count[e], 1u);
}

// incidence scatter: slot = ptr[ent] + (p - headpos[p])
template <int NLOC>
__global__ void k_incidence(const uint32_t* __restrict__ seq, const int32_t* __restrict__ headpos,
                            const int32_t* __restrict__ ent_of_p, const int64_t* __restrict__ ptr,
                            int32_t* __restrict__ inc_cell, uint8_t* __restrict__ inc_loc, int loc_offset,
                            int64_t n) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n) return;
  uint32_t s = seq[p];
  uint32_t c = s / NLOC, l = s - c * NLOC;
  int64_t slot = ptr[ent_of_p[p]] + (p - headpos[p]);
  inc_cell[slot] = (int32_t)c;
  inc_loc[slot] = (uint8_t)(l + loc_offset);
}

__global__ void k_count_to_i64(const unsigned int* __restrict__ cnt, int64_t* __restrict__ out, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = cnt[i];
}

__global__ void k_add_offset(int64_t* a, int64_t off, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) a[i] += off;
}

__global__ void k_check_sorted(const int32_t* __restrict__ cells, int nv, int64_t C, int64_t V, int* bad) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= C) return;
  bool ok = cells[c * nv] >= 0 && cells[c * nv + nv - 1] < V;
  for (int k = 1; k < nv; ++k) ok &= cells[c * nv + k] > cells[c * nv + k - 1];
  if (!ok) *bad = 1;
}

__global__ void k_facet_info(const int64_t* __restrict__ ptr, const int32_t* __restrict__ inc_cell,
                             const uint8_t* __restrict__ inc_loc, int loc_offset, uint8_t* __restrict__ ncells,
                             int32_t* __restrict__ first_cell, uint8_t* __restrict__ first_loc, int64_t n) {
  int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= n) return;
  int64_t a = ptr[f], b = ptr[f + 1];
  ncells[f] = (uint8_t)min((int64_t)255, b - a);
  first_cell[f] = inc_cell[a];
  first_loc[f] = (uint8_t)(inc_loc[a] - loc_offset);
}

__global__ void k_cell_extmask(const int32_t* __restrict__ cell_facets, int nloc,
                               const uint8_t* __restrict__ ncells, uint8_t* __restrict__ mask, int64_t C) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= C) return;
  unsigned m = 0;
  for (int l = 0; l < nloc; ++l)
    if (ncells[cell_facets[c * nloc + l]] == 1) m |= 1u << l;
  mask[c] = (uint8_t)m;
}

__global__ void k_pad_coords(const double* __restrict__ c3, double* __restrict__ c4, int64_t V) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= V * 4) return;
  const int64_t v = i >> 2;
  const int d = (int)(i & 3);
  c4[i] = d < 3 ? c3[v * 3 + d] : 0.0;
}

__global__ void k_interior_incidence(int64_t* __restrict__ ptr, int32_t* __restrict__ inc_cell,
                                     uint8_t* __restrict__ inc_loc, int64_t base, int loc, int64_t C) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c > C) return;
  ptr[c] = base + c;
  if (c < C) {
    inc_cell[base + c] = (int32_t)c;
    inc_loc[base + c] = (uint8_t)loc;
  }
}

inline unsigned grid_for(int64_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

struct Tmp {
  void* p = nullptr;
  cudaStream_t st;
  explicit Tmp(cudaStream_t s) : st(s) {}
  cudaError_t alloc(size_t bytes) { return dev_malloc(&p, bytes ? bytes : 16, st); }
  ~Tmp() {
    if (p) dev_free(p, st);
  }
  template <class T>
  T* as() { return (T*)p; }
};

// Number one entity class.  Outputs: cell_x [C][NLOC] (device, preallocated), *n_out, and -- written at the
// given offsets of preallocated unified arrays when inc_* != nullptr -- the incidence lists.
// ptr_out [n_out+1] is allocated here (int64, entries offset by ptr_offset).
template <int NLOC, int NVK>
int number_class(const sfem_mesh_s* m, int32_t* cell_x, int64_t* n_out, int64_t** ptr_out, int64_t ptr_offset,
                 int32_t* inc_cell, uint8_t* inc_loc, int loc_offset, cudaStream_t st) {
  const int64_t n = m->C * NLOC;
  SFEM_CHECK(n < (int64_t)2147483647, "mesh too large: cells*local_entities must be < 2^31");
  Tmp seqA(st), seqB(st), keyA(st), keyB(st), isfirst(st), firstidx(st), headpos(st), cubtmp(st);
  SFEM_CUDA(seqA.alloc(n * 4));
  SFEM_CUDA(seqB.alloc(n * 4));
  SFEM_CUDA(keyA.alloc(n * 4));
  SFEM_CUDA(keyB.alloc(n * 4));
  int bits = 1;
  while (((int64_t)1 << bits) < m->V) ++bits;
  cub::DoubleBuffer<uint32_t> dkeys(keyA.as<uint32_t>(), keyB.as<uint32_t>());
  cub::DoubleBuffer<uint32_t> dseq(seqA.as<uint32_t>(), seqB.as<uint32_t>());
  size_t tmp_bytes = 0;
  SFEM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dkeys, dseq, (int)n, 0, bits, st));
  size_t scan_bytes = 0;
  SFEM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (int32_t*)nullptr, (int32_t*)nullptr, (int)n, st));
  size_t scan2 = 0;
  SFEM_CUDA(cub::DeviceScan::InclusiveScan(nullptr, scan2, (int32_t*)nullptr, (int32_t*)nullptr, cub::Max(), (int)n, st));
  size_t scan3 = 0;
  SFEM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan3, (int64_t*)nullptr, (int64_t*)nullptr, (int)n + 1, st));
  tmp_bytes = std::max(std::max(tmp_bytes, scan_bytes), std::max(scan2, scan3));
  SFEM_CUDA(cubtmp.alloc(tmp_bytes));

  k_iota<<<grid_for(n), 256, 0, st>>>(dseq.Current(), n);
  for (int comp = NVK - 1; comp >= 0; --comp) {
    k_keys<NLOC, NVK><<<grid_for(n), 256, 0, st>>>(dseq.Current(), m->cells, m->nv, comp, dkeys.Current(), n);
    SFEM_CUDA(cub::DeviceRadixSort::SortPairs(cubtmp.p, tmp_bytes, dkeys, dseq, (int)n, 0, bits, st));
  }
  const uint32_t* seq = dseq.Current();
  // key buffers are free now: reuse them
  int32_t* ent_of_p = (int32_t*)dkeys.Current();
  SFEM_CUDA(isfirst.alloc(n * 4));
  SFEM_CUDA(firstidx.alloc(n * 4));
  SFEM_CUDA(headpos.alloc(n * 4));
  k_heads<NLOC, NVK><<<grid_for(n), 256, 0, st>>>(seq, m->cells, m->nv, isfirst.as<int32_t>(), headpos.as<int32_t>(), n);
  SFEM_CUDA(cub::DeviceScan::ExclusiveSum(cubtmp.p, tmp_bytes, isfirst.as<int32_t>(), firstidx.as<int32_t>(), (int)n, st));
  SFEM_CUDA(cub::DeviceScan::InclusiveScan(cubtmp.p, tmp_bytes, headpos.as<int32_t>(), headpos.as<int32_t>(), cub::Max(), (int)n, st));
  int32_t last_idx = 0, last_flag = 0;
  SFEM_CUDA(cudaMemcpyAsync(&last_idx, firstidx.as<int32_t>() + (n - 1), 4, cudaMemcpyDeviceToHost, st));
  SFEM_CUDA(cudaMemcpyAsync(&last_flag, isfirst.as<int32_t>() + (n - 1), 4, cudaMemcpyDeviceToHost, st));
  SFEM_CUDA(cudaStreamSynchronize(st));
  const int64_t ne = (int64_t)last_idx + last_flag;
  *n_out = ne;

  Tmp count(st);
  SFEM_CUDA(count.alloc((ne + 1) * 4));
  SFEM_CUDA(cudaMemsetAsync(count.p, 0, (ne + 1) * 4, st));
  k_assign<<<grid_for(n), 256, 0, st>>>(seq, headpos.as<int32_t>(), firstidx.as<int32_t>(), cell_x, ent_of_p,
                                        count.as<unsigned int>(), n);
  int64_t* ptr = nullptr;
  SFEM_CUDA(dev_malloc(&ptr, (ne + 1) * 8, st));
  Tmp cnt64(st);
  SFEM_CUDA(cnt64.alloc((ne + 1) * 8));
  k_count_to_i64<<<grid_for(ne + 1), 256, 0, st>>>(count.as<unsigned int>(), cnt64.as<int64_t>(), ne + 1);
  SFEM_CUDA(cub::DeviceScan::ExclusiveSum(cubtmp.p, tmp_bytes, cnt64.as<int64_t>(), ptr, (int)(ne + 1), st));
  if (inc_cell) {
    k_incidence<NLOC><<<grid_for(n), 256, 0, st>>>(seq, headpos.as<int32_t>(), ent_of_p, ptr, inc_cell, inc_loc,
                                                   loc_offset, n);
  }
  if (ptr_offset) k_add_offset<<<grid_for(ne + 1), 256, 0, st>>>(ptr, ptr_offset, ne + 1);
  *ptr_out = ptr;
  SFEM_CUDA(cudaGetLastError());
  SFEM_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // namespace

int build_numbering(sfem_space_s* sp, cudaStream_t st) {
  const sfem_mesh_s* m = sp->mesh;
  const int64_t C = m->C;
  {
    int* bad = nullptr;
    SFEM_CUDA(dev_malloc(&bad, 4, st));
    SFEM_CUDA(cudaMemsetAsync(bad, 0, 4, st));
    k_check_sorted<<<grid_for(C), 256, 0, st>>>(m->cells, m->nv, C, m->V, bad);
    int h = 0;
    SFEM_CUDA(cudaMemcpyAsync(&h, bad, 4, cudaMemcpyDeviceToHost, st));
    SFEM_CUDA(cudaStreamSynchronize(st));
    dev_free(bad, st);
    SFEM_CHECK(h == 0, "cells must list in-range vertex ids in strictly ascending order (UFC ordering)");
  }
  const bool d3 = m->gdim == 3;
  sp->ne_loc = d3 ? 6 : 3;
  sp->nf_loc = d3 ? 4 : 0;
  const int q = sp->order;
  sp->n_ent_cell = sp->ne_loc + (q == 2 ? (d3 ? 4 : 1) : 0);
  sp->ndof_cell = q * sp->n_ent_cell;
  // unified incidence arrays of the dof entities
  const int64_t inc_edges = C * sp->ne_loc;
  const int64_t inc_total = C * sp->n_ent_cell;
  SFEM_CUDA(dev_malloc(&sp->inc_cell, std::max<int64_t>(inc_total, 1) * 4, st));
  SFEM_CUDA(dev_malloc(&sp->inc_loc, std::max<int64_t>(inc_total, 1), st));
  SFEM_CUDA(dev_malloc(&sp->cell_edges, C * sp->ne_loc * 4, st));
  int64_t* ptr_e = nullptr;
  if (d3)
    SFEM_TRY((number_class<6, 2>(m, sp->cell_edges, &sp->E, &ptr_e, 0, sp->inc_cell, sp->inc_loc, 0, st)));
  else
    SFEM_TRY((number_class<3, 2>(m, sp->cell_edges, &sp->E, &ptr_e, 0, sp->inc_cell, sp->inc_loc, 0, st)));

  int64_t* ptr_f = nullptr;
  int32_t* f_inc_cell = nullptr;
  uint8_t* f_inc_loc = nullptr;
  if (d3) {
    SFEM_CUDA(dev_malloc(&sp->cell_faces, C * 4 * 4, st));
    if (q == 2) {
      f_inc_cell = sp->inc_cell + inc_edges;
      f_inc_loc = sp->inc_loc + inc_edges;
      SFEM_TRY((number_class<4, 3>(m, sp->cell_faces, &sp->F, &ptr_f, inc_edges, f_inc_cell, f_inc_loc, 6, st)));
      // facet info reads the unified arrays through the (offset) pointer table
      f_inc_cell = sp->inc_cell;
      f_inc_loc = sp->inc_loc;
    } else {
      SFEM_CUDA(dev_malloc(&f_inc_cell, C * 4 * 4, st));
      SFEM_CUDA(dev_malloc(&f_inc_loc, C * 4, st));
      SFEM_TRY((number_class<4, 3>(m, sp->cell_faces, &sp->F, &ptr_f, 0, f_inc_cell, f_inc_loc, 6, st)));
    }
  }
  // facets
  sp->n_facets = d3 ? sp->F : sp->E;
  SFEM_CUDA(dev_malloc(&sp->facet_ncells, std::max<int64_t>(sp->n_facets, 1), st));
  SFEM_CUDA(dev_malloc(&sp->facet_first_cell, std::max<int64_t>(sp->n_facets, 1) * 4, st));
  SFEM_CUDA(dev_malloc(&sp->facet_first_loc, std::max<int64_t>(sp->n_facets, 1), st));
  SFEM_CUDA(dev_malloc(&sp->cell_extmask, C, st));
  if (d3) {
    k_facet_info<<<grid_for(sp->F), 256, 0, st>>>(ptr_f, f_inc_cell, f_inc_loc, 6, sp->facet_ncells,
                                                  sp->facet_first_cell, sp->facet_first_loc, sp->F);
    k_cell_extmask<<<grid_for(C), 256, 0, st>>>(sp->cell_faces, 4, sp->facet_ncells, sp->cell_extmask, C);
  } else {
    k_facet_info<<<grid_for(sp->E), 256, 0, st>>>(ptr_e, sp->inc_cell, sp->inc_loc, 0, sp->facet_ncells,
                                                  sp->facet_first_cell, sp->facet_first_loc, sp->E);
    k_cell_extmask<<<grid_for(C), 256, 0, st>>>(sp->cell_edges, 3, sp->facet_ncells, sp->cell_extmask, C);
  }
  // unified incidence pointer table
  sp->n_ent = sp->E + (q == 2 ? (d3 ? sp->F : C) : 0);
  sp->ndof = q * sp->n_ent;
  sp->n_inc = inc_total;
  SFEM_CUDA(dev_malloc(&sp->inc_ptr, (sp->n_ent + 1) * 8, st));
  SFEM_CUDA(cudaMemcpyAsync(sp->inc_ptr, ptr_e, (sp->E + 1) * 8, cudaMemcpyDeviceToDevice, st));
  if (q == 2 && d3) {
    SFEM_CUDA(cudaMemcpyAsync(sp->inc_ptr + sp->E, ptr_f, (sp->F + 1) * 8, cudaMemcpyDeviceToDevice, st));
  } else if (q == 2) {
    k_interior_incidence<<<grid_for(C + 1), 256, 0, st>>>(sp->inc_ptr + sp->E, sp->inc_cell, sp->inc_loc,
                                                          inc_edges, 3, C);
  }
  if (d3) {
    SFEM_CUDA(dev_malloc(&sp->coords4, m->V * 32, st));
    k_pad_coords<<<grid_for(m->V * 4), 256, 0, st>>>(m->coords, sp->coords4, m->V);
  }
  SFEM_CUDA(cudaGetLastError());
  SFEM_CUDA(cudaStreamSynchronize(st));
  dev_free(ptr_e, st);
  if (ptr_f) dev_free(ptr_f, st);
  if (d3 && q == 1) {
    dev_free(f_inc_cell, st);
    dev_free(f_inc_loc, st);
  }
  return 0;
}

}  // namespace sfem
