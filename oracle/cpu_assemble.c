/* TEST INFRASTRUCTURE -- CPU oracle / CPU baseline, never linked into or called by the product path.
 *
 * Plain-C restatement of the algorithm the reference's hot path runs inside DOLFIN 1.0
 * (dolfin.assemble + DirichletBC.apply, called at /root/reference/sucemfem/SystemMatrices.py:54-55):
 *
 *   oc_number_entities  first-appearance numbering of edges/faces over (cell asc., local entity asc.)
 *                       (DOLFIN mesh.init(d) triggered by FunctionSpace, EMProblem.py:100-105)
 *   oc_pattern          per-row sorted union of the cell-dof sets (SparsityPatternBuilder)
 *   oc_assemble_cells   serial loop over cells in index order: element tensors of
 *                       eps_r*inner(v,u)*dx (Forms.py:76) and dot(curl v,curl u)/mu_r*dx (Forms.py:84)
 *                       by contraction of per-cell geometry with pre-integrated reference tensors (what an
 *                       FFC tensor-representation tabulate_tensor does), then A.add(block, dofs) as per-entry
 *                       binary-search insertion into the preallocated CSR rows (uBLAS/PETSc MatSetValues)
 *   oc_apply_ident      row-only ident (DirichletBC::apply -> GenericMatrix::ident)
 *
 * DOLFIN itself is not available here (un-vendored, un-pinned dependency; Python 2), so this is a "port"
 * in bench.py's cpu_baseline vocabulary: "CPU restatement of the reference algorithm", never "DOLFIN".
 * The reference tensors are passed in by the caller (oracle/refelem.py, exact rationals rounded to fp64),
 * so this file shares no tables with the CUDA library.  Tets only (the benchmarked configs); 2-D goes
 * through oracle/fem.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static const int TET_EDGES[6][2] = {{2, 3}, {1, 3}, {1, 2}, {0, 3}, {0, 2}, {0, 1}};
static const int TET_FACES[4][3] = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};

typedef struct { int64_t k0, k1; int64_t seq; } key_t_;

static int cmp_key(const void* a, const void* b) {
  const key_t_* x = (const key_t_*)a; const key_t_* y = (const key_t_*)b;
  if (x->k0 != y->k0) return x->k0 < y->k0 ? -1 : 1;
  if (x->k1 != y->k1) return x->k1 < y->k1 ? -1 : 1;
  return x->seq < y->seq ? -1 : (x->seq > y->seq);
}
static int cmp_i64(const void* a, const void* b) {
  int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
  return x < y ? -1 : (x > y);
}

/* number one entity class; nloc local entities of nvk vertices; out[c*nloc+l] = index; returns count */
static int64_t number_class(const int32_t* cells, int64_t C, int nloc, int nvk, const int* topo, int64_t V,
                            int32_t* out) {
  int64_t n = C * nloc, i;
  key_t_* keys = (key_t_*)malloc(sizeof(key_t_) * n);
  for (i = 0; i < n; ++i) {
    int64_t c = i / nloc; int l = (int)(i % nloc);
    const int32_t* cv = cells + c * 4;
    int64_t a = cv[topo[l * nvk]], b = cv[topo[l * nvk + 1]], d = nvk == 3 ? cv[topo[l * nvk + 2]] : 0;
    keys[i].k0 = a * V + b; keys[i].k1 = d; keys[i].seq = i;
  }
  qsort(keys, n, sizeof(key_t_), cmp_key);
  /* first appearance = smallest seq of each run (runs are sorted by seq); rank the first seqs */
  int64_t* first = (int64_t*)malloc(sizeof(int64_t) * n); int64_t nf = 0;
  for (i = 0; i < n; ++i)
    if (i == 0 || keys[i].k0 != keys[i - 1].k0 || keys[i].k1 != keys[i - 1].k1) first[nf++] = keys[i].seq;
  qsort(first, nf, sizeof(int64_t), cmp_i64);
  int64_t head_seq = -1; int64_t idx = -1;
  for (i = 0; i < n; ++i) {
    if (i == 0 || keys[i].k0 != keys[i - 1].k0 || keys[i].k1 != keys[i - 1].k1) {
      head_seq = keys[i].seq;
      int64_t* p = (int64_t*)bsearch(&head_seq, first, nf, sizeof(int64_t), cmp_i64);
      idx = p - first;
    }
    out[keys[i].seq] = (int32_t)idx;
  }
  free(first); free(keys);
  return nf;
}

int oc_number_entities(const int32_t* cells, int64_t C, int64_t V, int32_t* cell_edges, int32_t* cell_faces,
                       int64_t* nE, int64_t* nF) {
  *nE = number_class(cells, C, 6, 2, &TET_EDGES[0][0], V, cell_edges);
  *nF = number_class(cells, C, 4, 3, &TET_FACES[0][0], V, cell_faces);
  return 0;
}

/* cell dofs, SURVEY 8.1-D: order 1: edge; order 2: 2*edge+k, then 2E + 2*face + k */
void oc_cell_dofs(const int32_t* cell_edges, const int32_t* cell_faces, int64_t C, int64_t E, int order,
                  int32_t* out) {
  int64_t c; int u, k;
  if (order == 1) { memcpy(out, cell_edges, sizeof(int32_t) * 6 * C); return; }
  for (c = 0; c < C; ++c) {
    for (u = 0; u < 6; ++u) for (k = 0; k < 2; ++k) out[c * 20 + 2 * u + k] = 2 * cell_edges[c * 6 + u] + k;
    for (u = 0; u < 4; ++u) for (k = 0; k < 2; ++k) out[c * 20 + 12 + 2 * u + k] = (int32_t)(2 * E + 2 * cell_faces[c * 4 + u] + k);
  }
}

/* in-place ascending sort of int32 (median-of-three quicksort, insertion sort below 16) */
static void sort_i32(int32_t* a, int64_t n) {
  while (n > 16) {
    int32_t x = a[0], y = a[n / 2], z = a[n - 1], p;
    if (x > y) { int32_t t = x; x = y; y = t; }
    p = z < x ? x : (z > y ? y : z);
    int64_t i = 0, j = n - 1;
    while (i <= j) {
      while (a[i] < p) ++i;
      while (a[j] > p) --j;
      if (i <= j) { int32_t t = a[i]; a[i] = a[j]; a[j] = t; ++i; --j; }
    }
    if (j + 1 < n - i) { sort_i32(a, j + 1); a += i; n -= i; } else { sort_i32(a + i, n - i); n = j + 1; }
  }
  for (int64_t i = 1; i < n; ++i) {
    int32_t v = a[i]; int64_t j = i - 1;
    while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; --j; }
    a[j + 1] = v;
  }
}

/* pattern (DOLFIN SparsityPatternBuilder + tensor init): a dof -> cells incidence by counting sort, then per row the
 * sorted union of the incident cells' dof lists.  cols == NULL: count only (rowptr filled); otherwise rowptr must hold
 * the result of the count call and cols is filled.  Rows are independent: OpenMP over rows when threads > 1. */
int64_t oc_pattern_mt(const int32_t* cd, int64_t C, int nd, int64_t ndof, int64_t* rowptr, int32_t* cols, int threads) {
  int64_t* ptr = (int64_t*)calloc(ndof + 1, sizeof(int64_t));
  int64_t i, c, r; int j;
  for (i = 0; i < C * nd; ++i) ptr[cd[i] + 1]++;
  for (r = 0; r < ndof; ++r) ptr[r + 1] += ptr[r];
  int32_t* inc = (int32_t*)malloc(sizeof(int32_t) * C * nd);
  int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * ndof);
  memcpy(fill, ptr, sizeof(int64_t) * ndof);
  for (c = 0; c < C; ++c) for (j = 0; j < nd; ++j) inc[fill[cd[c * nd + j]]++] = (int32_t)c;
  int64_t maxm = 0;
  for (r = 0; r < ndof; ++r) if ((ptr[r + 1] - ptr[r]) * nd > maxm) maxm = (ptr[r + 1] - ptr[r]) * nd;
#ifdef _OPENMP
  if (threads < 1) threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads) if (threads > 1)
#endif
  {
    int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * (maxm + 1));
    int64_t rr;
#ifdef _OPENMP
#pragma omp for schedule(static, 1024)
#endif
    for (rr = 0; rr < ndof; ++rr) {
      int64_t m = (ptr[rr + 1] - ptr[rr]) * nd, k = 0, u = 0, ii; int jj;
      for (ii = ptr[rr]; ii < ptr[rr + 1]; ++ii) for (jj = 0; jj < nd; ++jj) tmp[k++] = cd[(int64_t)inc[ii] * nd + jj];
      sort_i32(tmp, m);
      if (cols) {
        int32_t* out = cols + rowptr[rr];
        for (ii = 0; ii < m; ++ii) if (ii == 0 || tmp[ii] != tmp[ii - 1]) out[u++] = tmp[ii];
      } else {
        for (ii = 0; ii < m; ++ii) if (ii == 0 || tmp[ii] != tmp[ii - 1]) ++u;
        rowptr[rr + 1] = u;
      }
    }
    free(tmp);
  }
  if (!cols) { rowptr[0] = 0; for (r = 0; r < ndof; ++r) rowptr[r + 1] += rowptr[r]; }
  int64_t nnz = rowptr[ndof];
  free(fill); free(inc); free(ptr);
  return nnz;
}

int64_t oc_pattern(const int32_t* cd, int64_t C, int nd, int64_t ndof, int64_t* rowptr, int32_t* cols /* may be NULL */) {
  return oc_pattern_mt(cd, C, nd, ndof, rowptr, cols, 1);
}

static inline int64_t find_slot(const int64_t* rowptr, const int32_t* cols, int32_t r, int32_t c) {
  int64_t lo = rowptr[r], hi = rowptr[r + 1] - 1;
  while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (cols[mid] < c) lo = mid + 1; else hi = mid; }
  return lo;
}

/* Mhat, Chat: [3][3][nd][nd] as produced by oracle/refelem.py.  threads = 1: serial, ascending cell order
 * (what DOLFIN 1.0 + sucemfem run); threads > 1: OpenMP over cells with atomic adds (order not fixed). */
int oc_assemble_cells(const double* coords, const int32_t* cells, int64_t C, const int32_t* cd, int nd,
                      const double* Mhat, const double* Chat, const double* eps, const double* mu,
                      const int64_t* rowptr, const int32_t* cols, double* S, double* M, int threads) {
  int64_t c;
#ifdef _OPENMP
  if (threads < 1) threads = omp_get_max_threads();
#pragma omp parallel for num_threads(threads) schedule(static) if (threads > 1)
#endif
  for (c = 0; c < C; ++c) {
    const int32_t* cv = cells + c * 4;
    double J[3][3], Ji[3][3], G[3][3], Gi[3][3];
    int a, b, d, i, j;
    for (d = 0; d < 3; ++d) for (a = 0; a < 3; ++a) J[d][a] = coords[(int64_t)cv[a + 1] * 3 + d] - coords[(int64_t)cv[0] * 3 + d];
    double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2],
           c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02, id = 1.0 / det;
    Ji[0][0] = c00 * id; Ji[1][0] = c01 * id; Ji[2][0] = c02 * id;
    Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
    Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
    Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
    Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
    double ad = fabs(det), sm = ad * (eps ? eps[c] : 1.0), ss = 1.0 / (ad * (mu ? mu[c] : 1.0));
    for (a = 0; a < 3; ++a) for (b = 0; b < 3; ++b) {
      double gi = 0, gg = 0;
      for (d = 0; d < 3; ++d) { gi += Ji[a][d] * Ji[b][d]; gg += J[d][a] * J[d][b]; }
      Gi[a][b] = gi * sm; G[a][b] = gg * ss;
    }
    double Ae[2][20 * 20];
    for (i = 0; i < nd * nd; ++i) {
      double m = 0, s = 0;
      for (a = 0; a < 3; ++a) for (b = 0; b < 3; ++b) {
        m += Gi[a][b] * Mhat[(a * 3 + b) * nd * nd + i];
        s += G[a][b] * Chat[(a * 3 + b) * nd * nd + i];
      }
      Ae[0][i] = s; Ae[1][i] = m;
    }
    const int32_t* dofs = cd + c * nd;
    for (i = 0; i < nd; ++i) for (j = 0; j < nd; ++j) {
      int64_t slot = find_slot(rowptr, cols, dofs[i], dofs[j]);
      if (threads > 1) {
#ifdef _OPENMP
#pragma omp atomic
#endif
        S[slot] += Ae[0][i * nd + j];
#ifdef _OPENMP
#pragma omp atomic
#endif
        M[slot] += Ae[1][i * nd + j];
      } else {
        S[slot] += Ae[0][i * nd + j];
        M[slot] += Ae[1][i * nd + j];
      }
    }
  }
  return 0;
}

void oc_apply_ident(const int64_t* rowptr, const int32_t* cols, double* vals, const int64_t* rows, int64_t nrows) {
  int64_t k, i;
  for (k = 0; k < nrows; ++k) {
    int64_t r = rows[k];
    for (i = rowptr[r]; i < rowptr[r + 1]; ++i) vals[i] = cols[i] == r ? 1.0 : 0.0;
  }
}

/* ---- first-appearance numbering in ONE serial pass with a hash map ---------------------------------------------
 * What DOLFIN 1.0's TopologyComputation does when FunctionSpace triggers mesh.init(d)
 * (/root/reference/sucemfem/ProblemConfigurations/EMProblem.py:100-105): visit the cells in index order
 * and, within a cell, the local entities in UFC order; an entity that has not been seen yet gets the next number.
 * Same result as oc_number_entities (sort-based), O(C) time: the oracle for bench-size meshes (10^7..10^8 cells). */
typedef struct { int64_t k0; int32_t k1; int32_t val; } hslot_t;

static inline uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

static int64_t number_class_hash(const int32_t* cells, int64_t C, int nloc, int nvk, const int* topo, int64_t V,
                                 int64_t expected, int32_t* out) {
  uint64_t cap = 1; while ((int64_t)cap < 2 * expected + 16) cap <<= 1;
  hslot_t* tab = (hslot_t*)malloc(sizeof(hslot_t) * cap);
  if (!tab) return -1;
  for (uint64_t i = 0; i < cap; ++i) tab[i].val = -1;
  int64_t next = 0;
  for (int64_t c = 0; c < C; ++c) {
    const int32_t* cv = cells + c * 4;
    for (int l = 0; l < nloc; ++l) {
      const int64_t k0 = (int64_t)cv[topo[l * nvk]] * V + cv[topo[l * nvk + 1]];
      const int32_t k1 = nvk == 3 ? cv[topo[l * nvk + 2]] : 0;
      uint64_t h = mix64((uint64_t)k0 * 0x9e3779b97f4a7c15ULL + (uint64_t)(uint32_t)k1) & (cap - 1);
      while (tab[h].val >= 0 && !(tab[h].k0 == k0 && tab[h].k1 == k1)) h = (h + 1) & (cap - 1);
      if (tab[h].val < 0) {
        if (next >= (int64_t)(cap >> 1) + (int64_t)(cap >> 2)) { free(tab); return -2; }  /* expected was too small */
        tab[h].k0 = k0; tab[h].k1 = k1; tab[h].val = (int32_t)next++;
      }
      out[c * nloc + l] = tab[h].val;
    }
  }
  free(tab);
  return next;
}

/* expected_edges / expected_faces: upper estimates of the entity counts (<= 0: 6 C / 4 C, always enough) */
int oc_number_entities_hash(const int32_t* cells, int64_t C, int64_t V, int64_t expected_edges, int64_t expected_faces,
                            int32_t* cell_edges, int32_t* cell_faces, int64_t* nE, int64_t* nF) {
  *nE = number_class_hash(cells, C, 6, 2, &TET_EDGES[0][0], V, expected_edges > 0 ? expected_edges : 6 * C, cell_edges);
  *nF = number_class_hash(cells, C, 4, 3, &TET_FACES[0][0], V, expected_faces > 0 ? expected_faces : 4 * C, cell_faces);
  return (*nE < 0 || *nF < 0) ? 1 : 0;
}

/* ---- the reference algorithm restricted to a SAMPLE of rows (bench-size parity) --------------------------------
 * Same arithmetic and the same order as oc_pattern + oc_assemble_cells + oc_apply_ident, but only the rows listed in
 * rows[] (global dof numbers, ascending) are kept: pattern row = sorted union of the dofs of the cells that contain
 * the row's dof; values = element-tensor rows added in ascending cell order by binary-search insertion; rows flagged
 * in pec (uint8 per sampled row, may be NULL) become unit rows.  Order 2 tetrahedra; dofs from the entity tables
 * (SURVEY 8.1-D), so no [C][20] dof array is needed at 5e7 cells.
 * Call 1 (out_cols == NULL): fills out_rowptr[nrows+1] (offsets into the sample's own arrays), returns its nnz.
 * Call 2: fills out_cols / out_S / out_M. */
static inline void cell_dofs20(const int32_t* ce, const int32_t* cf, int64_t c, int64_t E, int64_t* d) {
  int u, k;
  for (u = 0; u < 6; ++u) for (k = 0; k < 2; ++k) d[2 * u + k] = 2 * (int64_t)ce[c * 6 + u] + k;
  for (u = 0; u < 4; ++u) for (k = 0; k < 2; ++k) d[12 + 2 * u + k] = 2 * E + 2 * (int64_t)cf[c * 4 + u] + k;
}
static int cmp_i64_(const void* a, const void* b) {
  int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
  return x < y ? -1 : (x > y);
}
static inline int64_t find_row(const int64_t* rows, int64_t n, int64_t r) {
  int64_t lo = 0, hi = n - 1;
  while (lo <= hi) { int64_t mid = (lo + hi) >> 1; if (rows[mid] < r) lo = mid + 1; else if (rows[mid] > r) hi = mid - 1; else return mid; }
  return -1;
}

int64_t oc_assemble_rows(const double* coords, const int32_t* cells, int64_t C, const int32_t* cell_edges,
                         const int32_t* cell_faces, int64_t E, int64_t F, const double* Mhat, const double* Chat,
                         const double* eps, const double* mu, int64_t nrows, const int64_t* rows, const uint8_t* pec,
                         int64_t* out_rowptr, int64_t* out_cols, double* out_S, double* out_M) {
  const int nd = 20;
  const int64_t ndof = 2 * E + 2 * F;
  /* flag bytes: which dofs are sampled (a cell is looked at only if one of its 20 dofs is) */
  uint8_t* flag = (uint8_t*)calloc(ndof, 1);
  int64_t i, c; int a, b, d, j;
  for (i = 0; i < nrows; ++i) flag[rows[i]] = 1;
  if (!out_cols) {
    /* count pass: candidates per row, then sort-unique */
    int64_t* cnt = (int64_t*)calloc(nrows + 1, sizeof(int64_t));
    for (c = 0; c < C; ++c) {
      int64_t dd[20]; cell_dofs20(cell_edges, cell_faces, c, E, dd);
      for (j = 0; j < nd; ++j) if (flag[dd[j]]) cnt[find_row(rows, nrows, dd[j]) + 1] += nd;
    }
    for (i = 0; i < nrows; ++i) cnt[i + 1] += cnt[i];
    int64_t* cand = (int64_t*)malloc(sizeof(int64_t) * (cnt[nrows] + 1));
    int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * (nrows + 1));
    memcpy(fill, cnt, sizeof(int64_t) * (nrows + 1));
    for (c = 0; c < C; ++c) {
      int64_t dd[20]; cell_dofs20(cell_edges, cell_faces, c, E, dd);
      for (j = 0; j < nd; ++j) if (flag[dd[j]]) {
        int64_t r = find_row(rows, nrows, dd[j]);
        memcpy(cand + fill[r], dd, sizeof(int64_t) * nd); fill[r] += nd;
      }
    }
    out_rowptr[0] = 0;
    for (i = 0; i < nrows; ++i) {
      int64_t m = cnt[i + 1] - cnt[i], u = 0, k;
      qsort(cand + cnt[i], m, sizeof(int64_t), cmp_i64_);
      for (k = 0; k < m; ++k) if (k == 0 || cand[cnt[i] + k] != cand[cnt[i] + k - 1]) ++u;
      out_rowptr[i + 1] = out_rowptr[i] + u;
    }
    free(fill); free(cand); free(cnt); free(flag);
    return out_rowptr[nrows];
  }
  /* fill pass: columns (same sort-unique), then the cell loop */
  {
    int64_t* cnt = (int64_t*)calloc(nrows + 1, sizeof(int64_t));
    for (c = 0; c < C; ++c) {
      int64_t dd[20]; cell_dofs20(cell_edges, cell_faces, c, E, dd);
      for (j = 0; j < nd; ++j) if (flag[dd[j]]) cnt[find_row(rows, nrows, dd[j]) + 1] += nd;
    }
    for (i = 0; i < nrows; ++i) cnt[i + 1] += cnt[i];
    int64_t* cand = (int64_t*)malloc(sizeof(int64_t) * (cnt[nrows] + 1));
    int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * (nrows + 1));
    memcpy(fill, cnt, sizeof(int64_t) * (nrows + 1));
    for (c = 0; c < C; ++c) {
      int64_t dd[20]; cell_dofs20(cell_edges, cell_faces, c, E, dd);
      for (j = 0; j < nd; ++j) if (flag[dd[j]]) {
        int64_t r = find_row(rows, nrows, dd[j]);
        memcpy(cand + fill[r], dd, sizeof(int64_t) * nd); fill[r] += nd;
      }
    }
    for (i = 0; i < nrows; ++i) {
      int64_t m = cnt[i + 1] - cnt[i], u = 0, k;
      qsort(cand + cnt[i], m, sizeof(int64_t), cmp_i64_);
      for (k = 0; k < m; ++k) if (k == 0 || cand[cnt[i] + k] != cand[cnt[i] + k - 1]) out_cols[out_rowptr[i] + u++] = cand[cnt[i] + k];
    }
    free(fill); free(cand); free(cnt);
  }
  memset(out_S, 0, sizeof(double) * out_rowptr[nrows]);
  memset(out_M, 0, sizeof(double) * out_rowptr[nrows]);
  for (c = 0; c < C; ++c) {
    int64_t dd[20]; cell_dofs20(cell_edges, cell_faces, c, E, dd);
    int any = 0;
    for (j = 0; j < nd; ++j) any |= flag[dd[j]];
    if (!any) continue;
    const int32_t* cv = cells + c * 4;
    double J[3][3], Ji[3][3], G[3][3], Gi[3][3];
    for (d = 0; d < 3; ++d) for (a = 0; a < 3; ++a) J[d][a] = coords[(int64_t)cv[a + 1] * 3 + d] - coords[(int64_t)cv[0] * 3 + d];
    double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2],
           c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02, id = 1.0 / det;
    Ji[0][0] = c00 * id; Ji[1][0] = c01 * id; Ji[2][0] = c02 * id;
    Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
    Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
    Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
    Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
    double ad = fabs(det), sm = ad * (eps ? eps[c] : 1.0), ss = 1.0 / (ad * (mu ? mu[c] : 1.0));
    for (a = 0; a < 3; ++a) for (b = 0; b < 3; ++b) {
      double gi = 0, gg = 0;
      for (d = 0; d < 3; ++d) { gi += Ji[a][d] * Ji[b][d]; gg += J[d][a] * J[d][b]; }
      Gi[a][b] = gi * sm; G[a][b] = gg * ss;
    }
    for (i = 0; i < nd; ++i) {
      if (!flag[dd[i]]) continue;
      const int64_t r = find_row(rows, nrows, dd[i]);
      for (j = 0; j < nd; ++j) {
        double m = 0, s = 0;
        for (a = 0; a < 3; ++a) for (b = 0; b < 3; ++b) {
          m += Gi[a][b] * Mhat[(a * 3 + b) * nd * nd + i * nd + j];
          s += G[a][b] * Chat[(a * 3 + b) * nd * nd + i * nd + j];
        }
        int64_t lo = out_rowptr[r], hi = out_rowptr[r + 1] - 1;
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (out_cols[mid] < dd[j]) lo = mid + 1; else hi = mid; }
        out_S[lo] += s; out_M[lo] += m;
      }
    }
  }
  if (pec)
    for (i = 0; i < nrows; ++i) if (pec[i]) {
      int64_t k;
      for (k = out_rowptr[i]; k < out_rowptr[i + 1]; ++k) { out_S[k] = out_cols[k] == rows[i] ? 1.0 : 0.0; out_M[k] = out_S[k]; }
    }
  free(flag);
  return out_rowptr[nrows];
}

int oc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
