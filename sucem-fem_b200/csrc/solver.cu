// Consumers of the assembled matrices that stay on the device: the matrix-vector product of a
// linear combination of matrices that share one pattern, and the BiCGStab solver the reference runs through scipy
// for driven problems (/root/reference/sucemfem/Utilities/LinalgSolvers.py:239-249, solve_sparse_system :303-322,
// called with preconditioner_type='diagonal' at /root/reference/examples/infinitesimal_dipole/driver.py:83).
//
// The driven system matrix A = S - k0^2 M + j k0 S_0 (/root/reference/sucemfem/ProblemConfigurations/
// EMDrivenProblem.py:67-72) is never formed: S, M and S_0 share the CSR pattern, so y = A x reads the three value
// arrays once, side by side, with the ENTITY-level column ids (one int32 per q x q block of non-zeros: 1 byte per
// non-zero at order 2 instead of 4) and complex coefficients applied in registers.  HBM-bound: 8 n_mats + 4/q^2 bytes
// per non-zero, x is gathered through L2.
//   k_matvec: half a warp per row entity (q rows): lane s walks the row's column entities with stride 16, loads the
//   column entity id, the q complex x entries of that entity and, per matrix and row, q contiguous values (16 B at
//   order 2; the 16 lanes read consecutive 16-byte chunks); complex FMAs; 4 xor-shuffle steps reduce the q sums.
// BiCGStab is driven from the host side of the library (scalars come back per iteration for the convergence test);
// vectors are complex (interleaved re, im) and live in caller-owned device arrays plus a pooled work space.
#include <algorithm>
#include <cmath>
#include <complex>
#include <vector>

#include "sfem_internal.h"

namespace sfem {
namespace {

constexpr int kMaxMats = 4;

struct MatvecArgs {
  const int64_t* __restrict__ ent_rowptr;
  const int32_t* __restrict__ ent_cols;
  const double* __restrict__ vals[kMaxMats];
  double cre[kMaxMats], cim[kMaxMats];
  int n_mats;
  int64_t n_rows_ent;
  // optional last term in compact boundary-row storage (S_0, sfem_assemble_abc): y += c * S0 x over the listed rows
  const double* __restrict__ abc_vals;
  const int64_t* __restrict__ abc_rows;
  const int64_t* __restrict__ abc_rowptr;
  int64_t n_abc;
  double abc_cre, abc_cim;
};

// y[q*r + k] += c * sum_cols S0[q*r + k, col] x[col] for the row entities of the compact ABC storage: one warp per row
// entity (the boundary rows are a vanishing share of the matrix: no tuning).
template <int Q, bool CPLX>
__global__ void k_matvec_abc(MatvecArgs p, const double* __restrict__ x, double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (i >= p.n_abc) return;
  const int64_t r = p.abc_rows[i], a = p.ent_rowptr[r], ao = p.abc_rowptr[i];
  const int len = (int)(p.ent_rowptr[r + 1] - a);
  for (int k = 0; k < Q; ++k) {
    double2 acc = make_double2(0.0, 0.0);
    for (int idx = lane; idx < Q * len; idx += 32) {
      const int s = idx / Q, kp = idx - s * Q;
      const double v = p.abc_vals[(int64_t)Q * Q * ao + (int64_t)k * Q * len + idx];
      const int64_t c = (int64_t)Q * p.ent_cols[a + s] + kp;
      if (CPLX) {
        const double2 xv = reinterpret_cast<const double2*>(x)[c];
        acc.x = fma(v, xv.x, acc.x);
        acc.y = fma(v, xv.y, acc.y);
      } else {
        acc.x = fma(v, x[c], acc.x);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
      if (CPLX) acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
    }
    if (lane == 0) {
      if (CPLX) {
        double2* yy = reinterpret_cast<double2*>(y) + Q * r + k;
        double2 t = *yy;
        t.x += p.abc_cre * acc.x - p.abc_cim * acc.y;
        t.y += p.abc_cre * acc.y + p.abc_cim * acc.x;
        *yy = t;
      } else {
        y[Q * r + k] += p.abc_cre * acc.x;
      }
    }
  }
}

__device__ __forceinline__ double2 cfma(double2 acc, double2 a, double2 b) {  // acc + a*b
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
  return acc;
}

// y[q*r + k] = sum_m c_m sum_cols A_m[q*r + k, col] x[col]      (CPLX: x, y complex; else real and c_m real)
// Half a warp per row entity; two column entities per lane and iteration are in flight (rows of order-2 tets hold
// 13-45 column entities: one or two iterations).  NM is a template parameter so that the value pointers and
// coefficients live in registers / immediate constant-bank operands.
template <int Q, int NM, bool CPLX>
__global__ void __launch_bounds__(256, 3) k_matvec(MatvecArgs p, const double* __restrict__ x,
                                                                  double* __restrict__ y) {
  const int lane = threadIdx.x & 31, sub = lane & 15;
  const unsigned half_mask = 0xffffu << (lane & 16);
  const int64_t n_hw = (int64_t)gridDim.x * (blockDim.x >> 4);  // persistent half-warps, rows r, r + n_hw, ...
  int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 4;
  int64_t a = 0, a1 = 0;
  if (r < p.n_rows_ent) { a = p.ent_rowptr[r]; a1 = p.ent_rowptr[r + 1]; }
  while (r < p.n_rows_ent) {
    // row pointer of this half-warp's next row, one row ahead
    const int64_t rn = r + n_hw;
    int64_t an = 0, a1n = 0;
    if (rn < p.n_rows_ent) { an = p.ent_rowptr[rn]; a1n = p.ent_rowptr[rn + 1]; }
    double2 acc[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) acc[k] = make_double2(0.0, 0.0);
    const int len = (int)(a1 - a);
    const int64_t base = (int64_t)Q * Q * a;
    for (int s0 = sub; s0 < len; s0 += 32) {
      constexpr int U = 2;
      int sl[U];
      int64_t c[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        sl[u] = s0 + 16 * u;
        c[u] = sl[u] < len ? (int64_t)p.ent_cols[a + sl[u]] : (int64_t)-1;
      }
      // matrix entries first (independent of the column ids), combined over the matrices
      double2 w[U][Q][Q];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int k = 0; k < Q; ++k) {
#pragma unroll
          for (int j = 0; j < Q; ++j) w[u][k][j] = make_double2(0.0, 0.0);
          if (sl[u] < len) {
#pragma unroll
            for (int m = 0; m < NM; ++m) {
              const double* v = p.vals[m] + base + (int64_t)k * Q * len + (int64_t)Q * sl[u];
              if (Q == 2) {
                const double2 vv = __ldcs(reinterpret_cast<const double2*>(v));
                w[u][k][0].x = fma(p.cre[m], vv.x, w[u][k][0].x);
                w[u][k][1].x = fma(p.cre[m], vv.y, w[u][k][1].x);
                if (CPLX) {
                  w[u][k][0].y = fma(p.cim[m], vv.x, w[u][k][0].y);
                  w[u][k][1].y = fma(p.cim[m], vv.y, w[u][k][1].y);
                }
              } else {
                const double vv = __ldcs(v);
                w[u][k][0].x = fma(p.cre[m], vv, w[u][k][0].x);
                if (CPLX) w[u][k][0].y = fma(p.cim[m], vv, w[u][k][0].y);
              }
            }
          }
        }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (c[u] < 0) continue;
#pragma unroll
        for (int j = 0; j < Q; ++j) {
          double2 xv;
          if (CPLX) xv = reinterpret_cast<const double2*>(x)[Q * c[u] + j];
          else xv = make_double2(x[Q * c[u] + j], 0.0);
#pragma unroll
          for (int k = 0; k < Q; ++k) {
            if (CPLX) acc[k] = cfma(acc[k], w[u][k][j], xv);
            else acc[k].x = fma(w[u][k][j].x, xv.x, acc[k].x);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < Q; ++k)
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        acc[k].x += __shfl_xor_sync(half_mask, acc[k].x, o);
        if (CPLX) acc[k].y += __shfl_xor_sync(half_mask, acc[k].y, o);
      }
    if (sub == 0) {
#pragma unroll
      for (int k = 0; k < Q; ++k) {
        if (CPLX) reinterpret_cast<double2*>(y)[Q * r + k] = acc[k];
        else y[Q * r + k] = acc[k].x;
      }
    }
    r = rn; a = an; a1 = a1n;
  }
}

// 1 / diagonal of the combined matrix (Jacobi preconditioner; scipy.sparse.spdiags(1./A.diagonal()), LinalgSolvers.py:115)
template <int Q>
__global__ void k_inv_diag(MatvecArgs p, int64_t ent_lo, double2* __restrict__ out) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= p.n_rows_ent) return;
  const int64_t a = p.ent_rowptr[r];
  const int len = (int)(p.ent_rowptr[r + 1] - a);
  int lo = 0, hi = len - 1;
  const int32_t g = (int32_t)(ent_lo + r);
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (p.ent_cols[a + mid] < g) lo = mid + 1; else hi = mid;
  }
  for (int k = 0; k < Q; ++k) {
    double dre = 0.0, dim = 0.0;
    for (int m = 0; m < p.n_mats; ++m) {
      const double v = p.vals[m][(int64_t)Q * Q * a + (int64_t)k * Q * len + (int64_t)Q * lo + k];
      dre = fma(p.cre[m], v, dre);
      dim = fma(p.cim[m], v, dim);
    }
    out[Q * r + k] = make_double2(dre, dim);
  }
}
// + the diagonal of the compact ABC term
template <int Q>
__global__ void k_diag_abc(MatvecArgs p, int64_t ent_lo, double2* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= p.n_abc) return;
  const int64_t r = p.abc_rows[i], a = p.ent_rowptr[r], ao = p.abc_rowptr[i];
  const int len = (int)(p.ent_rowptr[r + 1] - a);
  int lo = 0, hi = len - 1;
  const int32_t g = (int32_t)(ent_lo + r);
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (p.ent_cols[a + mid] < g) lo = mid + 1; else hi = mid;
  }
  for (int k = 0; k < Q; ++k) {
    const double v = p.abc_vals[(int64_t)Q * Q * ao + (int64_t)k * Q * len + (int64_t)Q * lo + k];
    out[Q * r + k].x += p.abc_cre * v;
    out[Q * r + k].y += p.abc_cim * v;
  }
}
__global__ void k_invert_diag(double2* __restrict__ d, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2 v = d[i];
  const double n2 = v.x * v.x + v.y * v.y;
  d[i] = n2 > 0.0 ? make_double2(v.x / n2, -v.y / n2) : make_double2(1.0, 0.0);
}

// ---- complex vector kernels of the solver (deterministic two-stage reductions) ---------------------------------
constexpr int kRedBlocks = 592;  // 4 per SM
constexpr int kRedThreads = 256;

__device__ __forceinline__ double2 block_sum(double2 v) {
  __shared__ double2 s[kRedThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  double2 t = make_double2(0.0, 0.0);
  if (threadIdx.x < kRedThreads / 32) t = s[threadIdx.x];
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      t.x += __shfl_xor_sync(0xffffffffu, t.x, o);
      t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
    }
  }
  return t;  // valid in thread 0
}

// partial[blk] = sum conj(a) b  and  partial2[blk] = sum conj(c) d   (c == nullptr: only the first)
__global__ void __launch_bounds__(kRedThreads) k_dotc2(const double2* __restrict__ a, const double2* __restrict__ b,
                                                       const double2* __restrict__ c, const double2* __restrict__ d,
                                                       int64_t n, double2* __restrict__ partial) {
  double2 s1 = make_double2(0.0, 0.0), s2 = s1;
  for (int64_t i = blockIdx.x * (int64_t)kRedThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kRedThreads) {
    const double2 u = a[i], v = b[i];
    s1.x += u.x * v.x + u.y * v.y;
    s1.y += u.x * v.y - u.y * v.x;
    if (c) {
      const double2 e = c[i], f = d[i];
      s2.x += e.x * f.x + e.y * f.y;
      s2.y += e.x * f.y - e.y * f.x;
    }
  }
  const double2 t1 = block_sum(s1);
  if (threadIdx.x == 0) partial[blockIdx.x] = t1;
  if (c) {
    const double2 t2 = block_sum(s2);
    if (threadIdx.x == 0) partial[gridDim.x + blockIdx.x] = t2;
  }
}
__global__ void __launch_bounds__(kRedThreads) k_dot_final(const double2* __restrict__ partial, int nblk, int nout,
                                                           double2* __restrict__ out) {
  for (int o = 0; o < nout; ++o) {
    double2 s = make_double2(0.0, 0.0);
    for (int i = threadIdx.x; i < nblk; i += kRedThreads) {
      s.x += partial[o * nblk + i].x;
      s.y += partial[o * nblk + i].y;
    }
    const double2 t = block_sum(s);
    if (threadIdx.x == 0) out[o] = t;
    __syncthreads();
  }
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// p = r + beta (p - omega v);  phat = minv .* p
__global__ void k_update_p(double2* __restrict__ p, const double2* __restrict__ r, const double2* __restrict__ v,
                           double2 beta, double2 omega, const double2* __restrict__ minv, double2* __restrict__ phat,
                           int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2 pv = p[i], vv = v[i], rv = r[i];
  const double2 ov = cmul(omega, vv);
  const double2 t = cmul(beta, make_double2(pv.x - ov.x, pv.y - ov.y));
  const double2 np = make_double2(rv.x + t.x, rv.y + t.y);
  p[i] = np;
  phat[i] = minv ? cmul(minv[i], np) : np;
}
// s = r - alpha v;  shat = minv .* s
__global__ void k_update_s(double2* __restrict__ s, const double2* __restrict__ r, const double2* __restrict__ v,
                           double2 alpha, const double2* __restrict__ minv, double2* __restrict__ shat, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2 av = cmul(alpha, v[i]), rv = r[i];
  const double2 sv = make_double2(rv.x - av.x, rv.y - av.y);
  s[i] = sv;
  shat[i] = minv ? cmul(minv[i], sv) : sv;
}
// x += alpha phat + omega shat;  r = s - omega t
__global__ void k_update_xr(double2* __restrict__ x, const double2* __restrict__ phat, const double2* __restrict__ shat,
                            double2* __restrict__ r, const double2* __restrict__ s, const double2* __restrict__ t,
                            double2 alpha, double2 omega, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2 a = cmul(alpha, phat[i]), o = cmul(omega, shat[i]), xv = x[i];
  x[i] = make_double2(xv.x + a.x + o.x, xv.y + a.y + o.y);
  if (r) {
    const double2 ot = cmul(omega, t[i]), sv = s[i];
    r[i] = make_double2(sv.x - ot.x, sv.y - ot.y);
  }
}
// r = b - y
__global__ void k_residual(const double2* __restrict__ b, const double2* __restrict__ y, double2* __restrict__ r, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) r[i] = make_double2(b[i].x - y[i].x, b[i].y - y[i].y);
}

inline unsigned grid_for(int64_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

int fill_args(sfem_pattern_s* pat, int n_mats_flags, const double* const* d_values, const double* coef, MatvecArgs* a,
              cudaStream_t st) {
  const bool last_compact = (n_mats_flags & SFEM_MATS_LAST_IS_ABC_COMPACT) != 0;
  const int n_all = n_mats_flags & 0xff;
  SFEM_CHECK(n_all >= 1 && n_all <= kMaxMats, "n_mats must be between 1 and 4");
  const int n_mats = last_compact ? n_all - 1 : n_all;
  a->abc_vals = nullptr; a->abc_rows = nullptr; a->abc_rowptr = nullptr; a->n_abc = 0; a->abc_cre = a->abc_cim = 0.0;
  if (last_compact) {
    SFEM_CHECK(d_values[n_all - 1], "NULL value array");
    SFEM_TRY(build_abc_rows(pat, st));
    a->abc_vals = d_values[n_all - 1];
    a->abc_rows = pat->abc_rows;
    a->abc_rowptr = pat->abc_rowptr;
    a->n_abc = pat->n_abc_rows;
    a->abc_cre = coef ? coef[2 * (n_all - 1)] : 1.0;
    a->abc_cim = coef ? coef[2 * (n_all - 1) + 1] : 0.0;
  }
  a->ent_rowptr = pat->ent_rowptr;
  a->ent_cols = pat->ent_cols;
  a->n_mats = n_mats;
  a->n_rows_ent = pat->n_rows_ent;
  for (int m = 0; m < kMaxMats; ++m) {
    a->vals[m] = m < n_mats ? d_values[m] : nullptr;
    a->cre[m] = m < n_mats ? (coef ? coef[2 * m] : 1.0) : 0.0;
    a->cim[m] = m < n_mats ? (coef ? coef[2 * m + 1] : 0.0) : 0.0;
    SFEM_CHECK(m >= n_mats || d_values[m], "NULL value array");
  }
  return 0;
}

template <int Q, bool CPLX>
void launch_matvec_nm(const MatvecArgs& a, unsigned grid, const double* x, double* y, cudaStream_t st) {
  switch (a.n_mats) {
    case 1: k_matvec<Q, 1, CPLX><<<grid, 256, 0, st>>>(a, x, y); break;
    case 2: k_matvec<Q, 2, CPLX><<<grid, 256, 0, st>>>(a, x, y); break;
    case 3: k_matvec<Q, 3, CPLX><<<grid, 256, 0, st>>>(a, x, y); break;
    default: k_matvec<Q, 4, CPLX><<<grid, 256, 0, st>>>(a, x, y); break;
  }
}

int launch_matvec(const sfem_pattern_s* pat, const MatvecArgs& a, bool cplx, const double* x, double* y, cudaStream_t st) {
  if (pat->n_rows_ent == 0) return 0;
  const int q0 = pat->space->order;
  if (a.n_mats == 0)  // only the compact term: y = 0 first
    SFEM_CUDA(cudaMemsetAsync(y, 0, (size_t)q0 * pat->n_rows_ent * (cplx ? 16 : 8), st));
  int num_sms = 148;
  SFEM_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, pat->device));
  // persistent: as many CTAs as are resident (3 per SM by the kernel's launch bounds), at most one row per half-warp
  const unsigned grid = (unsigned)std::min<int64_t>((int64_t)num_sms * 3, grid_for(pat->n_rows_ent * 16));
  const int q = pat->space->order;
  if (a.n_mats > 0) {
    if (q == 2 && cplx) launch_matvec_nm<2, true>(a, grid, x, y, st);
    else if (q == 2) launch_matvec_nm<2, false>(a, grid, x, y, st);
    else if (cplx) launch_matvec_nm<1, true>(a, grid, x, y, st);
    else launch_matvec_nm<1, false>(a, grid, x, y, st);
  }
  if (a.n_abc > 0) {
    const unsigned g2 = grid_for(a.n_abc * 32);
    if (q == 2 && cplx) k_matvec_abc<2, true><<<g2, 256, 0, st>>>(a, x, y);
    else if (q == 2) k_matvec_abc<2, false><<<g2, 256, 0, st>>>(a, x, y);
    else if (cplx) k_matvec_abc<1, true><<<g2, 256, 0, st>>>(a, x, y);
    else k_matvec_abc<1, false><<<g2, 256, 0, st>>>(a, x, y);
  }
  SFEM_CUDA(cudaGetLastError());
  return 0;
}

typedef std::complex<double> cplx_t;
inline double2 d2(cplx_t z) { return make_double2(z.real(), z.imag()); }

}  // namespace
}  // namespace sfem

using namespace sfem;

extern "C" {

int sfem_matvec(sfem_pattern_t pat, int n_mats, const double* const* d_values, const double* coef_re_im,
                int complex_vectors, const double* d_x, double* d_y, void* stream) {
  SFEM_CHECK(pat && d_values && d_x && d_y, "NULL argument");
  DeviceGuard dg(pat->device);
  MatvecArgs a;
  SFEM_TRY(fill_args(pat, n_mats, d_values, coef_re_im, &a, (cudaStream_t)stream));
  if (!complex_vectors) {
    for (int m = 0; m < a.n_mats; ++m) SFEM_CHECK(a.cim[m] == 0.0, "real vectors need real coefficients");
    SFEM_CHECK(a.abc_cim == 0.0, "real vectors need real coefficients");
  }
  return launch_matvec(pat, a, complex_vectors != 0, d_x, d_y, (cudaStream_t)stream);
}

int sfem_diagonal(sfem_pattern_t pat, int n_mats, const double* const* d_values, const double* coef_re_im,
                  double* d_out_complex, void* stream) {
  SFEM_CHECK(pat && d_values && d_out_complex, "NULL argument");
  DeviceGuard dg(pat->device);
  cudaStream_t st = (cudaStream_t)stream;
  MatvecArgs a;
  SFEM_TRY(fill_args(pat, n_mats, d_values, coef_re_im, &a, st));
  if (pat->n_rows_ent == 0) return 0;
  double2* out = reinterpret_cast<double2*>(d_out_complex);
  const int q = pat->space->order;
  if (q == 2) k_inv_diag<2><<<grid_for(pat->n_rows_ent), 256, 0, st>>>(a, pat->ent_lo + pat->gshift, out);
  else k_inv_diag<1><<<grid_for(pat->n_rows_ent), 256, 0, st>>>(a, pat->ent_lo + pat->gshift, out);
  if (a.n_abc > 0) {
    if (q == 2) k_diag_abc<2><<<grid_for(a.n_abc), 256, 0, st>>>(a, pat->ent_lo + pat->gshift, out);
    else k_diag_abc<1><<<grid_for(a.n_abc), 256, 0, st>>>(a, pat->ent_lo + pat->gshift, out);
  }
  SFEM_CUDA(cudaGetLastError());
  return 0;
}

int sfem_bicgstab(sfem_pattern_t pat, int n_mats, const double* const* d_values, const double* coef_re_im,
                  const double* d_b, double* d_x, int preconditioner, double rtol, int maxiter, int* iterations,
                  double* relres, void* stream) {
  SFEM_CHECK(pat && d_values && d_b && d_x, "NULL argument");
  SFEM_CHECK(pat->ent_lo == 0 && pat->ent_hi == pat->space->n_ent, "the solver needs the pattern of all rows");
  SFEM_CHECK(preconditioner == 0 || preconditioner == 1, "preconditioner: 0 none, 1 diagonal");
  DeviceGuard dg(pat->device);
  cudaStream_t st = (cudaStream_t)stream;
  MatvecArgs a;
  SFEM_TRY(fill_args(pat, n_mats, d_values, coef_re_im, &a, st));
  const int64_t n = pat->space->ndof;
  const int q = pat->space->order;
  double2* work = nullptr;  // r, rhat, p, v, s, t, phat, shat, minv + reduction scratch
  const int64_t nvec = 9;
  SFEM_CUDA(dev_malloc(&work, (nvec * n + 2 * kRedBlocks + 2) * sizeof(double2), st));
  double2 *r = work, *rhat = r + n, *p = rhat + n, *v = p + n, *s = v + n, *t = s + n, *phat = t + n, *shat = phat + n,
          *minv = shat + n, *partial = minv + n, *dres = partial + 2 * kRedBlocks;
  const double2* b = reinterpret_cast<const double2*>(d_b);
  double2* x = reinterpret_cast<double2*>(d_x);
  int rc = 0, it = 0;
  double res = 0.0, bnorm = 0.0;
  auto dots = [&](const double2* u1, const double2* w1, const double2* u2, const double2* w2, cplx_t* o1, cplx_t* o2) {
    k_dotc2<<<kRedBlocks, kRedThreads, 0, st>>>(u1, w1, u2, w2, n, partial);
    k_dot_final<<<1, kRedThreads, 0, st>>>(partial, kRedBlocks, u2 ? 2 : 1, dres);
    double h[4] = {0, 0, 0, 0};
    cudaError_t e = cudaMemcpyAsync(h, dres, (u2 ? 2 : 1) * sizeof(double2), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      set_error(std::string("bicgstab reduction: ") + cudaGetErrorString(e));
      return 1;
    }
    *o1 = cplx_t(h[0], h[1]);
    if (o2) *o2 = cplx_t(h[2], h[3]);
    return 0;
  };
  auto matvec = [&](const double2* in, double2* out) {
    return launch_matvec(pat, a, true, reinterpret_cast<const double*>(in), reinterpret_cast<double*>(out), st);
  };
  const unsigned g = grid_for(n);
  const double2* mi = nullptr;
  do {
    if (preconditioner == 1) {
      if (q == 2) k_inv_diag<2><<<grid_for(pat->n_rows_ent), 256, 0, st>>>(a, pat->ent_lo + pat->gshift, minv);
      else k_inv_diag<1><<<grid_for(pat->n_rows_ent), 256, 0, st>>>(a, pat->ent_lo + pat->gshift, minv);
      if (a.n_abc > 0) {
        if (q == 2) k_diag_abc<2><<<grid_for(a.n_abc), 256, 0, st>>>(a, pat->ent_lo + pat->gshift, minv);
        else k_diag_abc<1><<<grid_for(a.n_abc), 256, 0, st>>>(a, pat->ent_lo + pat->gshift, minv);
      }
      k_invert_diag<<<grid_for(n), 256, 0, st>>>(minv, n);
      mi = minv;
    }
    cplx_t bb, rr;
    if ((rc = matvec(x, v)) != 0) break;
    k_residual<<<g, 256, 0, st>>>(b, v, r, n);
    if ((rc = dots(b, b, r, r, &bb, &rr)) != 0) break;
    bnorm = std::sqrt(bb.real());
    const double target = rtol * bnorm;
    res = std::sqrt(rr.real());
    if (bnorm == 0.0) {  // x = 0 solves it
      cudaMemsetAsync(x, 0, n * sizeof(double2), st);
      res = 0.0;
      break;
    }
    if (res <= target) break;
    cudaMemcpyAsync(rhat, r, n * sizeof(double2), cudaMemcpyDeviceToDevice, st);
    cudaMemsetAsync(p, 0, n * sizeof(double2), st);
    cudaMemsetAsync(v, 0, n * sizeof(double2), st);
    cplx_t rho = 1.0, alpha = 1.0, omega = 1.0;
    for (it = 1; it <= maxiter; ++it) {
      cplx_t rho_new;
      if ((rc = dots(rhat, r, nullptr, nullptr, &rho_new, nullptr)) != 0) break;
      if (rho_new == cplx_t(0.0) || omega == cplx_t(0.0)) { rc = -10; break; }  // breakdown (scipy: info = -10)
      const cplx_t beta = (rho_new / rho) * (alpha / omega);
      k_update_p<<<g, 256, 0, st>>>(p, r, v, d2(beta), d2(omega), mi, phat, n);
      if ((rc = matvec(phat, v)) != 0) break;
      cplx_t rv;
      if ((rc = dots(rhat, v, nullptr, nullptr, &rv, nullptr)) != 0) break;
      if (rv == cplx_t(0.0)) { rc = -11; break; }
      alpha = rho_new / rv;
      k_update_s<<<g, 256, 0, st>>>(s, r, v, d2(alpha), mi, shat, n);
      cplx_t ss;
      if ((rc = dots(s, s, nullptr, nullptr, &ss, nullptr)) != 0) break;
      if (std::sqrt(ss.real()) <= target) {  // converged on the half step
        k_update_xr<<<g, 256, 0, st>>>(x, phat, shat, nullptr, nullptr, nullptr, d2(alpha), make_double2(0.0, 0.0), n);
        res = std::sqrt(ss.real());
        break;
      }
      if ((rc = matvec(shat, t)) != 0) break;
      cplx_t ts, tt;
      if ((rc = dots(t, s, t, t, &ts, &tt)) != 0) break;
      if (tt.real() == 0.0 || !std::isfinite(tt.real()) || !std::isfinite(std::abs(ts))) { rc = -11; break; }
      omega = ts / tt;
      k_update_xr<<<g, 256, 0, st>>>(x, phat, shat, r, s, t, d2(alpha), d2(omega), n);
      rho = rho_new;
      if ((rc = dots(r, r, nullptr, nullptr, &rr, nullptr)) != 0) break;
      res = std::sqrt(rr.real());
      if (!std::isfinite(res)) { rc = -11; break; }  // NaN / Inf: no later `res <= target` test could ever succeed
      if (res <= target) break;
    }
  } while (false);
  if (relres) *relres = bnorm > 0.0 ? res / bnorm : 0.0;
  cudaStreamSynchronize(st);
  dev_free(work, st);
  if (iterations) *iterations = it > maxiter ? maxiter : it;
  if (rc == -10 || rc == -11) {
    set_error("BiCGStab breakdown");
    return 5;
  }
  if (rc != 0) return rc;
  SFEM_CUDA(cudaGetLastError());
  if (it > maxiter) {  // x holds the last iterate
    set_error("BiCGStab: convergence to tolerance not achieved in " + std::to_string(maxiter) + " iterations");
    return 6;
  }
  return 0;
}

}  // extern "C"
