// Host-buffer entry points of the path: mesh, materials and facet markers come in as HOST arrays and the CSR arrays
// of S, T(=M) (and optionally S_0) go out to HOST arrays -- what FunctionSpace + assemble(M) + assemble(S) +
// apply_essential + A.data() (/root/reference/sucemfem/SystemMatrices.py:54-55,
// /root/reference/sucemfem/Utilities/Converters.py:37) deliver to a host consumer such as scipy.
//
// The device->host copy of the values dominates this call (20 bytes per non-zero over PCIe against 16 bytes per
// non-zero written to HBM at > 4 TB/s), so the job is organised around keeping the copy engine busy:
//   create: H2D of the inputs, numbering (K1), PEC flags, ONE count pass over all rows (sizes for the caller)
//   run:    the rows are cut into blocks of equal non-zeros; for each block the compute stream fills the pattern,
//           assembles S/T with the same kernels as sfem_assemble_ST and expands the dof-level CSR into one of two
//           staging sets, while a second stream drains the previous set to the caller's (pinned) arrays.
// Device memory in flight is two blocks, not the matrix: a mesh whose S+T exceed HBM still streams through.
#include <algorithm>
#include <vector>

#include "sfem_internal.h"

struct sfem_host_job_s {
  int device = 0;
  cudaStream_t st = nullptr;    // compute stream (the caller's)
  sfem_mesh_t mesh = nullptr;
  sfem_space_t space = nullptr;
  double* d_coords = nullptr;
  int32_t* d_cells = nullptr;
  double* d_eps = nullptr;
  double* d_mu = nullptr;
  uint8_t* d_flags = nullptr;
};

namespace sfem {
namespace {

inline unsigned grid_for(int64_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

// block boundaries of (nearly) equal non-zeros inside the row entities [r_lo, r_hi) + the row-pointer /
// incidence-pointer values there
__global__ void k_block_bounds(const int64_t* __restrict__ rowptr_all, const int64_t* __restrict__ inc_ptr, int64_t r_lo,
                               int64_t r_hi, int nblocks, int64_t* __restrict__ out) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > nblocks) return;
  const __int128 total = rowptr_all[r_hi] - rowptr_all[r_lo];
  const int64_t target = rowptr_all[r_lo] + (int64_t)(total * p / nblocks);
  int64_t lo = r_lo, hi = r_hi;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (rowptr_all[mid] < target) lo = mid + 1; else hi = mid;
  }
  const int64_t b = p == nblocks ? r_hi : (p == 0 ? r_lo : lo);
  out[3 * p] = b;
  out[3 * p + 1] = rowptr_all[b];
  out[3 * p + 2] = inc_ptr[b];
}

}  // namespace
}  // namespace sfem

using namespace sfem;

extern "C" {

int sfem_host_alloc(size_t bytes, void** out) {
  SFEM_CHECK(out, "out is NULL");
  SFEM_CUDA(cudaHostAlloc(out, bytes ? bytes : 16, cudaHostAllocDefault));
  return 0;
}

int sfem_host_free(void* p) {
  if (p) SFEM_CUDA(cudaFreeHost(p));
  return 0;
}

int sfem_host_job_destroy(sfem_host_job_t job) {
  if (!job) return 0;
  DeviceGuard dg(job->device);
  cudaStream_t st = job->st;
  if (job->space) sfem_space_destroy(job->space);
  if (job->mesh) sfem_mesh_destroy(job->mesh);
  dev_free(job->d_coords, st);
  dev_free(job->d_cells, st);
  dev_free(job->d_eps, st);
  dev_free(job->d_mu, st);
  dev_free(job->d_flags, st);
  delete job;
  return 0;
}

int sfem_host_job_create(int gdim, int64_t n_vertices, const double* h_coords, int64_t n_cells, const int32_t* h_cells,
                         int order, const double* h_eps_r, const double* h_mu_r, int pec_mode,
                         const uint32_t* h_facet_markers, uint32_t pec_region, void* stream, int64_t* ndof, int64_t* nnz,
                         sfem_host_job_t* out) {
  SFEM_CHECK(out && h_coords && h_cells, "NULL argument");
  SFEM_CHECK(gdim == 2 || gdim == 3, "gdim must be 2 or 3");
  SFEM_CHECK(pec_mode >= 0 && pec_mode <= 2, "pec_mode must be 0 (none), 1 (all exterior facets) or 2 (markers == region)");
  SFEM_CHECK(pec_mode != 2 || h_facet_markers, "pec_mode 2 needs facet markers");
  cudaStream_t st = (cudaStream_t)stream;
  auto* job = new sfem_host_job_s();
  SFEM_CUDA(cudaGetDevice(&job->device));
  job->st = st;
  int rc = 0;
  auto fail = [&](int code) {
    sfem_host_job_destroy(job);
    return code;
  };
#define JOB_CUDA(call)                                                                                          \
  do {                                                                                                          \
    cudaError_t _e = (call);                                                                                    \
    if (_e != cudaSuccess) {                                                                                    \
      set_error(std::string(#call) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
      return fail(1);                                                                                           \
    }                                                                                                           \
  } while (0)
  const size_t nv = gdim + 1;
  JOB_CUDA(dev_malloc(&job->d_coords, n_vertices * gdim * sizeof(double), st));
  JOB_CUDA(dev_malloc(&job->d_cells, n_cells * nv * sizeof(int32_t), st));
  JOB_CUDA(cudaMemcpyAsync(job->d_coords, h_coords, n_vertices * gdim * sizeof(double), cudaMemcpyHostToDevice, st));
  JOB_CUDA(cudaMemcpyAsync(job->d_cells, h_cells, n_cells * nv * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  if (h_eps_r) {
    JOB_CUDA(dev_malloc(&job->d_eps, n_cells * sizeof(double), st));
    JOB_CUDA(cudaMemcpyAsync(job->d_eps, h_eps_r, n_cells * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  if (h_mu_r) {
    JOB_CUDA(dev_malloc(&job->d_mu, n_cells * sizeof(double), st));
    JOB_CUDA(cudaMemcpyAsync(job->d_mu, h_mu_r, n_cells * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  if ((rc = sfem_mesh_create(gdim, n_vertices, job->d_coords, n_cells, job->d_cells, stream, &job->mesh)) != 0) return fail(rc);
  if ((rc = sfem_space_create(job->mesh, order, stream, &job->space)) != 0) return fail(rc);
  sfem_space_s* sp = job->space;
  if (pec_mode != 0) {
    uint32_t* d_markers = nullptr;
    JOB_CUDA(dev_malloc(&d_markers, std::max<int64_t>(sp->n_facets, 1) * 4, st));
    if (pec_mode == 2) {
      JOB_CUDA(cudaMemcpyAsync(d_markers, h_facet_markers, sp->n_facets * 4, cudaMemcpyHostToDevice, st));
    } else {
      // PECWallsBoundaryCondition: a marker function that is `region` on every exterior facet and differs elsewhere
      JOB_CUDA(cudaMemsetAsync(d_markers, pec_region == 0 ? 0xff : 0, std::max<int64_t>(sp->n_facets, 1) * 4, st));
      rc = sfem_mark_exterior_facets(sp, d_markers, pec_region, stream);
    }
    if (rc == 0 && dev_malloc(&job->d_flags, std::max<int64_t>(sp->ndof, 1), st) != cudaSuccess) {
      set_error("out of device memory for the PEC row flags");
      rc = 1;
    }
    if (rc == 0) {
      cudaMemsetAsync(job->d_flags, 0, std::max<int64_t>(sp->ndof, 1), st);
      rc = sfem_pec_row_flags(sp, d_markers, pec_region, job->d_flags, stream);
    }
    dev_free(d_markers, st);
    if (rc != 0) return fail(rc);
  }
  if ((rc = sfem_space_count_nnz(sp, stream, nnz)) != 0) return fail(rc);
#undef JOB_CUDA
  if (ndof) *ndof = sp->ndof;
  *out = job;
  return 0;
}

int sfem_space_count_nnz(sfem_space_t sp, void* stream, int64_t* nnz) {
  SFEM_CHECK(sp, "NULL space");
  DeviceGuard dg(sp->device);
  if (!sp->rowptr_all)
    SFEM_TRY(count_pattern_rows(sp, 0, sp->n_ent, (cudaStream_t)stream, &sp->rowptr_all, &sp->nnz_ent_all,
                                &sp->max_len_edges_all, &sp->max_len_other_all));
  if (nnz) *nnz = (int64_t)sp->order * sp->order * sp->nnz_ent_all;
  return 0;
}

int sfem_host_job_run(sfem_host_job_t job, int64_t* h_rowptr, int32_t* h_cols, double* h_S, double* h_T, double* h_S0,
                      int n_blocks) {
  SFEM_CHECK(job, "NULL job");
  DeviceGuard dg(job->device);
  return sfem_space_assemble_to_host(job->space, job->d_eps, job->d_mu, job->d_flags, h_rowptr, h_cols, h_S, h_T, h_S0,
                                     n_blocks, job->st);
}

int sfem_host_job_run_rows(sfem_host_job_t job, int64_t ent_lo, int64_t ent_hi, int64_t* h_rowptr, int32_t* h_cols,
                           double* h_S, double* h_T, double* h_S0, int n_blocks) {
  SFEM_CHECK(job, "NULL job");
  DeviceGuard dg(job->device);
  return sfem_space_assemble_rows_to_host(job->space, ent_lo, ent_hi, job->d_eps, job->d_mu, job->d_flags, h_rowptr, h_cols,
                                          h_S, h_T, h_S0, n_blocks, job->st);
}

int sfem_host_job_space(sfem_host_job_t job, sfem_space_t* out) {
  SFEM_CHECK(job && out, "NULL argument");
  *out = job->space;
  return 0;
}

int sfem_space_rows_nnz(sfem_space_t sp, int64_t ent_lo, int64_t ent_hi, void* stream, int64_t* nnz) {
  SFEM_CHECK(sp && nnz, "NULL argument");
  if (ent_hi < 0) ent_hi = sp->n_ent;
  SFEM_CHECK(0 <= ent_lo && ent_lo <= ent_hi && ent_hi <= sp->n_ent, "entity range out of bounds");
  DeviceGuard dg(sp->device);
  cudaStream_t st = (cudaStream_t)stream;
  SFEM_TRY(sfem_space_count_nnz(sp, stream, nullptr));
  int64_t h[2] = {0, 0};
  SFEM_CUDA(cudaMemcpyAsync(&h[0], sp->rowptr_all + ent_lo, 8, cudaMemcpyDeviceToHost, st));
  SFEM_CUDA(cudaMemcpyAsync(&h[1], sp->rowptr_all + ent_hi, 8, cudaMemcpyDeviceToHost, st));
  SFEM_CUDA(cudaStreamSynchronize(st));
  *nnz = (int64_t)sp->order * sp->order * (h[1] - h[0]);
  return 0;
}

int sfem_space_assemble_to_host(sfem_space_t sp, const double* d_eps_r, const double* d_mu_r,
                                const uint8_t* d_pec_row_flags, int64_t* h_rowptr, int32_t* h_cols, double* h_S,
                                double* h_T, double* h_S0, int n_blocks, void* stream) {
  SFEM_CHECK(sp, "NULL space");
  return sfem_space_assemble_rows_to_host(sp, 0, sp->n_ent, d_eps_r, d_mu_r, d_pec_row_flags, h_rowptr, h_cols, h_S, h_T,
                                          h_S0, n_blocks, stream);
}

int sfem_space_assemble_rows_to_host(sfem_space_t sp, int64_t ent_lo, int64_t ent_hi, const double* d_eps_r,
                                     const double* d_mu_r, const uint8_t* d_pec_row_flags, int64_t* h_rowptr,
                                     int32_t* h_cols, double* h_S, double* h_T, double* h_S0, int n_blocks,
                                     void* stream) {
  SFEM_CHECK(sp && h_rowptr, "NULL argument");
  if (ent_hi < 0) ent_hi = sp->n_ent;
  SFEM_CHECK(0 <= ent_lo && ent_lo <= ent_hi && ent_hi <= sp->n_ent, "entity range out of bounds");
  SFEM_CHECK(h_S || h_T || h_S0 || h_cols, "nothing to do: pass at least one output array");
  SFEM_CHECK(!sp->gid_e || ent_hi <= sp->E || ent_lo >= sp->E,
             "partitioned mesh: a row range holds edge rows or face rows, not both");
  DeviceGuard dg(sp->device);
  cudaStream_t st = (cudaStream_t)stream;
  SFEM_TRY(sfem_space_count_nnz(sp, stream, nullptr));
  if (!sp->drain) SFEM_CUDA(cudaStreamCreateWithFlags(&sp->drain, cudaStreamNonBlocking));
  cudaStream_t cp = sp->drain;
  const int q = sp->order;
  const int64_t n_ent = ent_hi - ent_lo;
  int64_t nnz = 0;
  SFEM_TRY(sfem_space_rows_nnz(sp, ent_lo, ent_hi, stream, &nnz));
  SFEM_CHECK(!h_cols || sp->ndof < (int64_t)2147483647, "column ids exceed int32");
  if (n_blocks <= 0) n_blocks = (int)std::min<int64_t>(256, std::max<int64_t>(1, (nnz + (128ll << 20) - 1) / (128ll << 20)));
  n_blocks = (int)std::min<int64_t>(n_blocks, std::max<int64_t>(n_ent, 1));

  // block boundaries (entity, entity-level nnz offset, incidence offset) -> host
  std::vector<int64_t> hb(3 * (n_blocks + 1));
  int64_t* d_b = nullptr;
  SFEM_CUDA(dev_malloc(&d_b, hb.size() * 8, st));
  k_block_bounds<<<grid_for(n_blocks + 1, 64), 64, 0, st>>>(sp->rowptr_all, sp->inc_ptr, ent_lo, ent_hi, n_blocks, d_b);
  SFEM_CUDA(cudaMemcpyAsync(hb.data(), d_b, hb.size() * 8, cudaMemcpyDeviceToHost, st));
  SFEM_CUDA(cudaStreamSynchronize(st));
  dev_free(d_b, st);
  int64_t max_rows = 0, max_nnz_ent = 0;
  for (int k = 0; k < n_blocks; ++k) {
    max_rows = std::max(max_rows, hb[3 * (k + 1)] - hb[3 * k]);
    max_nnz_ent = std::max(max_nnz_ent, hb[3 * (k + 1) + 1] - hb[3 * k + 1]);
  }
  const int64_t cap = std::max<int64_t>((int64_t)q * q * max_nnz_ent, 1);
  const int64_t base_ent = hb[1];  // entity-level non-zeros before the range: outputs are relative to the range

  struct Stage {
    double *S = nullptr, *T = nullptr, *S0 = nullptr;
    int32_t* cols = nullptr;
    int64_t* rowptr = nullptr;
    cudaEvent_t filled = nullptr, drained = nullptr;
  } stage[2];
  int rc = 0;
  auto cleanup = [&]() {
    cudaStreamSynchronize(cp);
    for (auto& s : stage) {
      dev_free(s.S, st); dev_free(s.T, st); dev_free(s.S0, st); dev_free(s.cols, st); dev_free(s.rowptr, st);
      if (s.filled) cudaEventDestroy(s.filled);
      if (s.drained) cudaEventDestroy(s.drained);
    }
  };
#define RUN_CUDA(call)                                                                                          \
  do {                                                                                                          \
    cudaError_t _e = (call);                                                                                    \
    if (_e != cudaSuccess) {                                                                                    \
      set_error(std::string(#call) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
      cleanup();                                                                                                \
      return 1;                                                                                                 \
    }                                                                                                           \
  } while (0)
  const int n_stage = n_blocks > 1 ? 2 : 1;
  for (int i = 0; i < n_stage; ++i) {
    Stage& s = stage[i];
    if (h_S) RUN_CUDA(dev_malloc(&s.S, cap * 8, st));
    if (h_T) RUN_CUDA(dev_malloc(&s.T, cap * 8, st));
    if (h_S0) RUN_CUDA(dev_malloc(&s.S0, cap * 8, st));
    if (h_cols) RUN_CUDA(dev_malloc(&s.cols, cap * 4, st));
    RUN_CUDA(dev_malloc(&s.rowptr, (q * max_rows + 1) * 8, st));
    RUN_CUDA(cudaEventCreateWithFlags(&s.filled, cudaEventDisableTiming));
    RUN_CUDA(cudaEventCreateWithFlags(&s.drained, cudaEventDisableTiming));
  }

  for (int k = 0; k < n_blocks && rc == 0; ++k) {
    Stage& s = stage[k % n_stage];
    const int64_t lo = hb[3 * k], hi = hb[3 * (k + 1)];
    const int64_t off_ent = hb[3 * k + 1], nnz_ent = hb[3 * (k + 1) + 1] - off_ent;
    const int64_t inc_lo = hb[3 * k + 2], n_inc = hb[3 * (k + 1) + 2] - inc_lo;
    const int64_t off = (int64_t)q * q * (off_ent - base_ent), nz = (int64_t)q * q * nnz_ent;
    if (hi == lo) continue;
    if (k >= n_stage) RUN_CUDA(cudaStreamWaitEvent(st, s.drained, 0));
    sfem_pattern_s pat;
    pat.space = sp;
    pat.device = sp->device;
    pat.stream = st;
    pat.ent_lo = lo;
    if (sp->gid_e) pat.gshift = lo >= sp->E ? sp->gshift_f : sp->gshift_e;  // partitioned mesh: global row ids
    pat.ent_hi = hi;
    pat.n_rows_ent = hi - lo;
    rc = build_pattern_block(&pat, sp->rowptr_all, 0, nnz_ent, inc_lo, n_inc, sp->max_len_edges_all, sp->max_len_other_all, st);
    if (rc == 0 && (h_S || h_T)) {
      // the fused kernels always produce both matrices: an output the caller skipped lands in the other's staging slot
      double* dS = s.S ? s.S : s.T;
      double* dT = s.T ? s.T : s.S;
      if (s.S && s.T) {
        rc = sfem_assemble_ST(&pat, d_eps_r, d_mu_r, d_pec_row_flags, dS, dT, SFEM_KERNEL_AUTO, st);
      } else {
        rc = sfem_assemble(&pat, s.S ? SFEM_FORM_STIFFNESS : SFEM_FORM_MASS, s.S ? d_mu_r : d_eps_r, nullptr,
                           d_pec_row_flags, s.S ? dS : dT, SFEM_KERNEL_AUTO, st);
      }
    }
    if (rc == 0 && h_S0) rc = sfem_assemble(&pat, SFEM_FORM_ABC, nullptr, nullptr, d_pec_row_flags, s.S0, SFEM_KERNEL_AUTO, st);
    if (rc == 0) rc = export_pattern(&pat, s.rowptr, s.cols, st, off);
    // the block's pattern is dead once its kernels are queued (stream-ordered frees)
    free_pattern_buffers(&pat, st);
    if (rc != 0) break;
    RUN_CUDA(cudaEventRecord(s.filled, st));
    RUN_CUDA(cudaStreamWaitEvent(cp, s.filled, 0));
    if (h_S) RUN_CUDA(cudaMemcpyAsync(h_S + off, s.S, nz * 8, cudaMemcpyDeviceToHost, cp));
    if (h_T) RUN_CUDA(cudaMemcpyAsync(h_T + off, s.T, nz * 8, cudaMemcpyDeviceToHost, cp));
    if (h_S0) RUN_CUDA(cudaMemcpyAsync(h_S0 + off, s.S0, nz * 8, cudaMemcpyDeviceToHost, cp));
    if (h_cols) RUN_CUDA(cudaMemcpyAsync(h_cols + off, s.cols, nz * 4, cudaMemcpyDeviceToHost, cp));
    RUN_CUDA(cudaMemcpyAsync(h_rowptr + q * (lo - ent_lo), s.rowptr, (q * (hi - lo) + 1) * 8, cudaMemcpyDeviceToHost, cp));
    RUN_CUDA(cudaEventRecord(s.drained, cp));
  }
#undef RUN_CUDA
  if (n_ent == 0) h_rowptr[0] = 0;
  cleanup();  // waits for the drain stream
  if (rc == 0) SFEM_CUDA(cudaStreamSynchronize(st));
  return rc;
}

}  // extern "C"
