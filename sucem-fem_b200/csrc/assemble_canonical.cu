// K3 (canonical-frame variant, tetrahedra): the benchmarked S + M(=T) value kernels (order 2: k_canon_edges +
// k_canon_faces; order 1: k_canon1_edges).  Replaces the cell loop of dolfin.assemble + DirichletBC.apply called from
// /root/reference/sucemfem/SystemMatrices.py:54-55 for the forms of /root/reference/sucemfem/Forms.py:76,84.
//
// Work decomposition (row-owner gather, no atomics, bit-reproducible):
//   * EDGES (k_canon_edges2, the default since round 2): one LANE per (edge, cell) item; a warp takes a batch of whole
//     consecutive edges with at most 32 items (see the comment above that kernel).  The description below is the
//     lane-PAIR layout, which the face kernel uses and which k_canon_edges keeps as the fallback for edges:
//   * a PAIR of lanes owns one row entity (edge or face) = 2 CSR rows of S and 2 of T, and splits the entity's
//     incident cells between its two lanes; a warp owns 16 consecutive entities, i.e. one contiguous span of the
//     value arrays.  Warps are persistent over batches of 16 entities.
//   * every incident cell is seen RELABELLED so that the row entity sits at the canonical position (edge (v0,v1) /
//     face (v0,v1,v2)).  The two element-matrix rows are then straight-line DFMAs whose reference-tensor
//     coefficients are constant-bank operands (generated/canonical_rows.inc), identical for all lanes; structural
//     zeros are gone at generation time.  A 64-bit descriptor per relabelling maps canonical column entities back to
//     the true ones (orientation fix-up, branch-free: a sign flip + two adds for edges in the (d, m) basis, two-level
//     selp.f64 for faces); relabelled vertex ids, slots, descriptor index and ring flags come precomputed in one packed
//     32-byte record per (entity, cell) (pattern.cu).
//   * contributions are staged in a per-warp shared-memory tile that is the warp's output span in its FINAL
//     layout; the span leaves the SM as ONE bulk asynchronous copy (cp.async.bulk shared -> global, the TMA
//     engine), so global stores cost no LSU wavefronts and are perfectly coalesced.  Faces: one lane per cell, the 4
//     entities of the shared face are exchanged by shuffle and summed in registers.  Edges visit their cells in RING
//     order (consecutive cells share a face through the edge; order precomputed in pattern.cu), alternate cells
//     between the two lanes, hand the three column entities two neighbouring cells have in common over by shuffle and
//     therefore store every slot exactly once (the one exception: where a closed ring closes, three chunks per row
//     are added to the first cell's).
//   * gathers are the other LSU consumer: the record (2 x 16 B), four 32-byte padded coordinates and two
//     coefficients per (entity, cell).  The chain record -> coordinates is issued two / one items ahead, straight
//     into the single item buffer once the current item's geometry has been computed; at the last step of a pass the
//     buffer receives the first item of the NEXT pass (S -> T: same batch, T -> S: next batch).
//   * PEC rows are overwritten with unit rows in the tile before it is stored (row-only ident fused).
// Geometry (J, detJ, J^-1 J^-T, J^T J) is recomputed per (entity, cell) from the gathered coordinates: ~70 flops
// against ~160 DFMAs of tensor contraction per matrix, cheaper than staging per-cell geometry through HBM.
#include "sfem_internal.h"

#define SFEM_TABLE static __device__ const
#include "generated/canonical_lut.inc"
#include "generated/canonical_rows.inc"

namespace sfem {
namespace {

struct CanonArgs {
  const double* __restrict__ coords4;      // [V][4]
  const int32_t* __restrict__ cells;
  const double* __restrict__ eps;
  const double* __restrict__ mu;
  const int64_t* __restrict__ inc_ptr;
  const int64_t* __restrict__ ent_rowptr;  // block-relative, indexed by r = g - ent_lo
  const uint4* __restrict__ rec;           // packed 32-byte item records (pattern.cu), 2 x uint4 per (entity, cell)
  const uint8_t* __restrict__ pec;
  double* __restrict__ S;
  double* __restrict__ T;
  int64_t ent_lo;    // first entity of the pattern block
  int64_t r_begin;   // first block-relative row entity handled by this launch
  int64_t r_end;
  int64_t inc_base;
  int tile_chunks;   // 16-byte chunks per warp tile = 16 entities * 2 rows * (longest entity row in this launch)
};

// per-(entity, cell) inputs
struct Item {
  double X[12];         // coordinates of the relabelled vertices
  double cmu, ceps;     // the cell's coefficients
  uint32_t w1, w2, w3;  // record bytes 20..31: descriptor index, flags, slots[10]
};

// the part of an item that outlives its coordinates: slots, flags, descriptor index
struct Slots {
  uint32_t w1, w2, w3;
};
__device__ __forceinline__ Slots slots_of(const Item& it) {
  Slots s;
  s.w1 = it.w1; s.w2 = it.w2; s.w3 = it.w3;
  return s;
}
template <int U>
__device__ __forceinline__ int slot_of(const Slots& it) {
  if constexpr (U < 2) return (it.w1 >> (16 + 8 * U)) & 0xffu;
  else if constexpr (U < 6) return (it.w2 >> (8 * (U - 2))) & 0xffu;
  else return (it.w3 >> (8 * (U - 6))) & 0xffu;
}
__device__ __forceinline__ unsigned flags_of(const Slots& it) { return (it.w1 >> 8) & 0xffu; }
__device__ __forceinline__ unsigned desc_index(const Slots& it) { return it.w1 & 0xffu; }

// last hop of the gather chain: coordinates + coefficients, from the record (vertex ids already relabelled)
// SFEM_WHATIF_* macros: timing-only experiments (wrong results) that bound what a lever could give; never defined in the
// shipped library (tools/whatif_gpu.sh builds them into lib/variants/).
template <bool EDGE = false>
__device__ __forceinline__ void gather_item(const CanonArgs& p, uint4 ra, uint4 rb, Item& it) {
  it.w1 = rb.y; it.w2 = rb.z; it.w3 = rb.w;
#if defined(SFEM_WHATIF_COORDS)
  if (EDGE) { ra.x &= 1023u; ra.y &= 1023u; ra.z &= 1023u; ra.w &= 1023u; rb.x &= 1023u; }
#endif
  const uint32_t v[4] = {ra.x, ra.y, ra.z, ra.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double2* x = reinterpret_cast<const double2*>(p.coords4 + (size_t)v[k] * 4);
    const double2 xy = x[0], z = x[1];
    it.X[3 * k] = xy.x;
    it.X[3 * k + 1] = xy.y;
    it.X[3 * k + 2] = z.x;
  }
  it.cmu = p.mu[(int)rb.x];
  it.ceps = p.eps[(int)rb.x];
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
    const double s = it.ceps / ad;
    g[0] = (A[0][0] * A[0][0] + A[0][1] * A[0][1] + A[0][2] * A[0][2]) * s;
    g[1] = (A[1][0] * A[1][0] + A[1][1] * A[1][1] + A[1][2] * A[1][2]) * s;
    g[2] = (A[2][0] * A[2][0] + A[2][1] * A[2][1] + A[2][2] * A[2][2]) * s;
    g[3] = (A[0][0] * A[1][0] + A[0][1] * A[1][1] + A[0][2] * A[1][2]) * s;
    g[4] = (A[0][0] * A[2][0] + A[0][1] * A[2][1] + A[0][2] * A[2][2]) * s;
    g[5] = (A[1][0] * A[2][0] + A[1][1] * A[2][1] + A[1][2] * A[2][2]) * s;
  } else {
    const double s = 1.0 / (ad * it.cmu);
    g[0] = (J[0][0] * J[0][0] + J[1][0] * J[1][0] + J[2][0] * J[2][0]) * s;
    g[1] = (J[0][1] * J[0][1] + J[1][1] * J[1][1] + J[2][1] * J[2][1]) * s;
    g[2] = (J[0][2] * J[0][2] + J[1][2] * J[1][2] + J[2][2] * J[2][2]) * s;
    g[3] = (J[0][0] * J[0][1] + J[1][0] * J[1][1] + J[2][0] * J[2][1]) * s;
    g[4] = (J[0][0] * J[0][2] + J[1][0] * J[1][2] + J[2][0] * J[2][2]) * s;
    g[5] = (J[0][1] * J[0][2] + J[1][1] * J[1][2] + J[2][1] * J[2][2]) * s;
  }
}

// branch-free select on one bit of a word: (word >> BIT) & 1 ? a : b   (the compiler turns the equivalent ?: on
// doubles into short divergent branches with predicated moves; selp.f64 is two SEL and no BSSY/BSYNC)
template <int BIT>
__device__ __forceinline__ double sel_bit(unsigned word, double a, double b) {
  double r;
  asm("{ .reg .pred q; .reg .b32 t; and.b32 t, %3, %4; setp.ne.u32 q, t, 0; selp.f64 %0, %1, %2, q; }"
      : "=d"(r) : "d"(a), "d"(b), "r"(word), "n"(1u << BIT));
  return r;
}
__device__ __forceinline__ double sel_flag(bool p, double a, double b) {
  double r;
  asm("{ .reg .pred q; setp.ne.u32 q, %3, 0; selp.f64 %0, %1, %2, q; }" : "=d"(r) : "d"(a), "d"(b), "r"((unsigned)p));
  return r;
}

// orientation fix-up context of one item (generated/canonical_lut.inc)
struct Fix {
  unsigned long long desc;
  unsigned fsel;
};

// true-frame values of the 2 columns of canonical column entity U, from the canonical-frame values the generated row
// functions produce: (v0, v1) for faces and for edges that are never seen reversed (DM = false); for the other edges
// (d, m) = ((v0 - v1)/2, (v0 + v1)/2), where a reversed edge only flips the sign of m (tools/gen_canonical.py).
// Faces: an ordered pair out of (v0, v1, w = -(v0 + v1)), three selector bits per face (swap, w0, w1).
template <int U, bool DM = false>
__device__ __forceinline__ double2 fix_columns(const Fix& f, double v0, double v1) {
#if defined(SFEM_WHATIF_NOFIX)
  return make_double2(v0, v1);
#endif
  if constexpr (U < 6 && DM) {
    const unsigned sign = ((unsigned)(f.desc >> 32) << (15 - U)) & 0x80000000u;  // descriptor bit 48 + U -> bit 31
    const double sm = __hiloint2double(__double2hiint(v1) ^ (int)sign, __double2loint(v1));
    return make_double2(sm + v0, sm - v0);
  } else if constexpr (U < 6) {
    return make_double2(v0, v1);
  } else {
    const double w = -(v0 + v1);
    const double u0 = sel_bit<4 * (U - 6)>(f.fsel, v1, v0);
    const double u1 = sel_bit<4 * (U - 6)>(f.fsel, v0, v1);
    return make_double2(sel_bit<4 * (U - 6) + 1>(f.fsel, w, u0), sel_bit<4 * (U - 6) + 2>(f.fsel, w, u1));
  }
}

// ---- the warp's output span leaves the SM as one bulk asynchronous copy -----------------------------------------
__device__ __forceinline__ void tile_store_issue(const double2* tile, double* gdst, unsigned bytes, int lane) {
  // make the generic-proxy writes of every lane visible to the async proxy, then one lane issues the copy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
#if defined(SFEM_WHATIF_NOBULK)
  bytes = 0;
#endif
  if (lane == 0 && bytes > 0) {
    const unsigned src = (unsigned)__cvta_generic_to_shared(tile);
#ifndef SFEM_NO_STORE_HINT
    // the value arrays are written once and never read by these kernels: evict-first keeps the 126 MB L2 for the
    // gathered inputs (records are read again by the T pass, vertex data by every neighbouring entity)
    unsigned long long policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(src),
                 "r"(bytes), "l"(policy)
                 : "memory");
#else
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(src), "r"(bytes)
                 : "memory");
#endif
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}
// the tile may be overwritten once the engine has finished READING it
__device__ __forceinline__ void tile_store_wait(int lane) {
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
}

// Loads that are issued well ahead of their use must not be consumed early: the compiler likes to hoist the cheap
// arithmetic that depends on them (address formation, subtraction of a base, byte extraction) right behind the load,
// where it then waits for the full memory latency.  An empty asm with the register as in/out operand pins the first
// use to the place where it is written.
__device__ __forceinline__ void keep_opaque(unsigned& v) { asm volatile("" : "+r"(v)); }
__device__ __forceinline__ void keep_opaque(int64_t& v) { asm volatile("" : "+l"(v)); }
__device__ __forceinline__ void keep_opaque(uint4& v) {
  asm volatile("" : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w));
}

// pecbits: bit k set <=> dof row (entity, k) is Dirichlet-constrained (loaded at the start of the batch)
__device__ __forceinline__ unsigned load_pecbits(const CanonArgs& p, int64_t g, bool valid) {
  if (!p.pec || !valid) return 0u;
  const uchar2 f = *reinterpret_cast<const uchar2*>(p.pec + 2 * g);
  return (f.x ? 1u : 0u) | (f.y ? 2u : 0u);
}
// the same in two halves: the raw two flag bytes (loaded at the start of the batch) and their decoding (at the use)
__device__ __forceinline__ unsigned load_pecraw(const CanonArgs& p, int64_t g, bool valid) {
  if (!p.pec || !valid) return 0u;
  return *reinterpret_cast<const unsigned short*>(p.pec + 2 * g);
}
__device__ __forceinline__ unsigned decode_pecraw(unsigned raw) {
  return ((raw & 0xffu) ? 1u : 0u) | ((raw >> 8) ? 2u : 0u);
}
__device__ __forceinline__ void pec_overwrite(unsigned pecbits, double2* __restrict__ row0, int len, int self) {
#pragma unroll
  for (int k = 0; k < 2; ++k)
    if ((pecbits >> k) & 1u) {
      double2* row = row0 + k * len;
      for (int s = 0; s < len; ++s) row[s] = make_double2(0.0, 0.0);
      row[self] = k == 0 ? make_double2(1.0, 0.0) : make_double2(0.0, 1.0);
    }
}

__device__ __forceinline__ void stage_lut(unsigned long long* s_desc, unsigned* s_fsel) {
  if (threadIdx.x < 16) {
    s_desc[threadIdx.x] = kRingDesc[threadIdx.x];
    s_fsel[threadIdx.x] = kRingFaceSel[threadIdx.x];
  }
  __syncthreads();
}

#ifndef SFEM_EDGE_CTAS
#define SFEM_EDGE_CTAS 2
#endif
#ifndef SFEM_FACE_CTAS
#define SFEM_FACE_CTAS 2
#endif
#ifndef SFEM_EDGE_THREADS
#define SFEM_EDGE_THREADS 128
#endif
#ifndef SFEM_FACE_THREADS
#define SFEM_FACE_THREADS 128
#endif
// tiles per warp of the face kernel: 2 = S and T in separate tiles (a pass never waits for the bulk copy before it)
#ifndef SFEM_FACE_TILES
#define SFEM_FACE_TILES 2
#endif
constexpr int kThreadsEdges = SFEM_EDGE_THREADS;  // warps x 16 edges; CTAs per SM are bounded by tiles and registers
constexpr int kThreadsFaces = SFEM_FACE_THREADS;  // warps x 16 faces

__device__ __forceinline__ double2 shfl_xor1(double2 v) {
  return make_double2(__shfl_xor_sync(0xffffffffu, v.x, 1), __shfl_xor_sync(0xffffffffu, v.y, 1));
}

// ---- edges: a PAIR of lanes per edge; cells visited in ring order, every slot stored exactly once ---------------
// The two lanes of a pair take alternate cells of the ring (lane h: cells h, h+2, ...), so an edge's dependent
// chain of cells is half as long and a warp's tile covers 16 edges: twice the resident warps for the same shared
// memory, no redundant work.  Relabelled cell = (a, b, entry, exit).  Canonical column entities by role:
//   self   5 = (0,1)                          partial sums per lane, combined and stored at the end
//   prev   4 = (0,2), 2 = (1,2), 9 = (0,1,2)   shared with the previous cell of the ring: carry + own, then stored
//   own    0 = (2,3), 7 = (0,2,3), 6 = (1,2,3)  this cell only: stored
//   next   3 = (0,3), 1 = (1,3), 8 = (0,1,3)   shared with the next cell, which the PARTNER lane is working on:
//                                               handed over by shuffle; stored if the fan ends here, added to the
//                                               first cell's slots if the ring closes here
// Gathers: the chain record -> coordinates of a lane's next item is issued one item ahead (record two ahead); at the
// LAST step of a pass the free buffer instead receives the first item of the NEXT pass (same batch for S -> T, the
// next batch for T -> S), so no pass starts with an exposed chain.
struct EdgeMeta {
  int64_t a, a1, i0, i1;  // entity-row range [a, a1) and record range [i0, i1) of this lane's edge (i0 > i1: none)
};

__device__ __forceinline__ EdgeMeta load_edge_meta(const CanonArgs& p, int64_t batch, int lane) {
  EdgeMeta m;
  const int64_t r = p.r_begin + batch * 16 + (lane >> 1);
  if (r < p.r_end) {
    const int64_t g = p.ent_lo + r;
    m.a = p.ent_rowptr[r];
    m.a1 = p.ent_rowptr[r + 1];
    m.i0 = p.inc_ptr[g] - p.inc_base;
    m.i1 = p.inc_ptr[g + 1] - p.inc_base;
  } else {
    m.a = m.a1 = 0;
    m.i0 = 1; m.i1 = 0;
  }
  return m;
}
// the same in two halves: the raw loads (i0, i1 still count from the start of the space's incidence array) ...
__device__ __forceinline__ EdgeMeta load_edge_meta_raw(const CanonArgs& p, int64_t batch, int lane) {
  EdgeMeta m;
  const int64_t r = p.r_begin + batch * 16 + (lane >> 1);
  if (r < p.r_end) {
    const int64_t g = p.ent_lo + r;
    m.a = p.ent_rowptr[r];
    m.a1 = p.ent_rowptr[r + 1];
    m.i0 = p.inc_ptr[g];
    m.i1 = p.inc_ptr[g + 1];
  } else {
    m.a = m.a1 = 0;
    m.i0 = p.inc_base + 1; m.i1 = p.inc_base;
  }
  return m;
}
// ... and the first use, placed by the caller where the loads have had time to arrive
__device__ __forceinline__ void finish_edge_meta(const CanonArgs& p, EdgeMeta& m) {
  keep_opaque(m.i0);
  keep_opaque(m.i1);
  m.i0 -= p.inc_base;
  m.i1 -= p.inc_base;
}

template <bool MASS>
__device__ __forceinline__ void edge_pass(const CanonArgs& p, const unsigned long long* __restrict__ s_desc,
                                          const unsigned* __restrict__ s_fsel, double2* __restrict__ tile, int lane,
                                          const EdgeMeta& m, unsigned pecraw, int off, int span, int nsteps,
                                          Item& cur, uint4& rc1a, uint4& rc1b,
                                          int64_t nf, int64_t nf_end, double* __restrict__ out_first) {
  // cur: this lane's first item (loaded or in flight); rc1: record of its second item; [nf, nf_end): this lane's
  // record range in the next pass (nf = first item; nf >= nf_end: none)
  const int h = lane & 1;
  const bool valid = m.i0 <= m.i1 && m.a1 > m.a;
  const int len = (int)(m.a1 - m.a);
  const uint4* __restrict__ rec = p.rec;
#if defined(SFEM_WHATIF_REC)
#define SFEM_RI(x) ((x) & 8191)
#else
#define SFEM_RI(x) (x)
#endif
  uint4 rfa = make_uint4(0, 0, 0, 0), rfb = rfa;
  if (nf < nf_end) { rfa = rec[2 * SFEM_RI(nf)]; rfb = rec[2 * SFEM_RI(nf) + 1]; }
  // the previous pass's bulk store must have finished reading the tile before it is overwritten
  tile_store_wait(lane);
  double2* row0 = tile + off;
  double2 self[2] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
  double2 saved[2][3];  // even lane: what the partner handed over at the end of the previous step
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int j = 0; j < 3; ++j) saved[k][j] = make_double2(0.0, 0.0);
  int self_slot = 0;
  int64_t i = m.i0 + h;
#pragma unroll 1
  for (int s = 0; s < nsteps; ++s, i += 2) {
    const bool active = valid && i < m.i1;
    // everything that is needed from the current item; its buffer is then free for the next one
    const Slots sl = slots_of(cur);
    const unsigned flags = active ? flags_of(sl) : 0u;
    const bool has_prev = flags & 1u, has_next = flags & 2u, wrap = flags & 4u;
    double gm[6], o[20];
    Fix fx;
    fx.desc = 0ull; fx.fsel = 0u;
    if (active) {
      canon_geometry<MASS>(cur, gm);
      fx.desc = s_desc[desc_index(sl)];
      fx.fsel = s_fsel[desc_index(sl)];
      if (s == 0 && h == 0) self_slot = slot_of<5>(sl);
    }
    uint4 rc2a = make_uint4(0, 0, 0, 0), rc2b = rc2a;
    if (s + 1 < nsteps) {
      // this lane's next item of the pass (record two ahead)
      if (i + 2 < m.i1) {
        if (i + 4 < m.i1) { rc2a = rec[2 * SFEM_RI(i + 4)]; rc2b = rec[2 * SFEM_RI(i + 4) + 1]; }
        gather_item<true>(p, rc1a, rc1b, cur);
      }
    } else {
      // last step: the first item of the next pass, and the record of its second item
      if (nf < nf_end) {
        if (nf + 2 < nf_end) { rc2a = rec[2 * SFEM_RI(nf + 2)]; rc2b = rec[2 * SFEM_RI(nf + 2) + 1]; }
        keep_opaque(rfa);  // loaded at the start of the pass; first use here
        keep_opaque(rfb);
        gather_item<true>(p, rfa, rfb, cur);
      }
    }
    __syncwarp();
    // 0 / 1 weights of the two possible carries (exact: x * 1 + y = x + y, x * 0 + y = y for finite x)
    const double mg = (has_prev && h) ? 1.0 : 0.0, ms = (has_prev && !h) ? 1.0 : 0.0;
#if defined(SFEM_WHATIF_NOCONFLICT)
#define SFEM_SLOT(K, U) ((U) * 32 + lane - off - (K) * len + (K) * 320)
#else
#define SFEM_SLOT(K, U) slot_of<U>(sl)
#endif
#define SFEM_RING_ROW(K, FN)                                                                       \
    {                                                                                              \
      double2 n3 = make_double2(0.0, 0.0), n1 = n3, n8 = n3;                                       \
      if (active) {                                                                                \
        FN(gm, o);                                                                                 \
        n3 = fix_columns<3, true>(fx, o[6], o[7]);                                                       \
        n1 = fix_columns<1, true>(fx, o[2], o[3]);                                                       \
        n8 = fix_columns<8>(fx, o[16], o[17]);                                                     \
      }                                                                                            \
      /* hand the next-side blocks to the partner, which owns the next cell of the ring */         \
      const double2 g3 = shfl_xor1(n3), g1 = shfl_xor1(n1), g8 = shfl_xor1(n8);                    \
      if (active) {                                                                                \
        double2* row = row0 + K * len;                                                             \
        double2 r;                                                                                 \
        r = fix_columns<5>(fx, o[10], o[11]); self[K].x += r.x; self[K].y += r.y;                  \
        /* carry from the previous cell of the ring: the partner's hand-over of THIS step (odd lane, weight mg) */ \
        /* or of the previous step (even lane, weight ms); weights are 0 / 1, at most one of them is 1 */ \
        r = fix_columns<4, true>(fx, o[8], o[9]);                                                        \
        r.x = fma(g3.x, mg, fma(saved[K][0].x, ms, r.x)); r.y = fma(g3.y, mg, fma(saved[K][0].y, ms, r.y)); \
        row[SFEM_SLOT(K, 4)] = r;                                                                  \
        r = fix_columns<2, true>(fx, o[4], o[5]);                                                        \
        r.x = fma(g1.x, mg, fma(saved[K][1].x, ms, r.x)); r.y = fma(g1.y, mg, fma(saved[K][1].y, ms, r.y)); \
        row[SFEM_SLOT(K, 2)] = r;                                                                  \
        r = fix_columns<9>(fx, o[18], o[19]);                                                      \
        r.x = fma(g8.x, mg, fma(saved[K][2].x, ms, r.x)); r.y = fma(g8.y, mg, fma(saved[K][2].y, ms, r.y)); \
        row[SFEM_SLOT(K, 9)] = r;                                                                  \
        row[SFEM_SLOT(K, 0)] = fix_columns<0, true>(fx, o[0], o[1]);                                     \
        row[SFEM_SLOT(K, 7)] = fix_columns<7>(fx, o[14], o[15]);                                   \
        row[SFEM_SLOT(K, 6)] = fix_columns<6>(fx, o[12], o[13]);                                   \
        if (!has_next) {                                                                           \
          double2* q3 = row + SFEM_SLOT(K, 3);                                                     \
          double2* q1 = row + SFEM_SLOT(K, 1);                                                     \
          double2* q8 = row + SFEM_SLOT(K, 8);                                                     \
          if (wrap) {                                                                              \
            const double2 t3 = *q3, t1 = *q1, t8 = *q8;                                            \
            n3.x += t3.x; n3.y += t3.y; n1.x += t1.x; n1.y += t1.y; n8.x += t8.x; n8.y += t8.y;    \
          }                                                                                        \
          *q3 = n3; *q1 = n1; *q8 = n8;                                                            \
        }                                                                                          \
      }                                                                                            \
      saved[K][0] = g3; saved[K][1] = g1; saved[K][2] = g8;                                        \
    }
    if (MASS) {
      SFEM_RING_ROW(0, canon_edge_mass_k0)
      SFEM_RING_ROW(1, canon_edge_mass_k1)
    } else {
      SFEM_RING_ROW(0, canon_edge_curl_k0)
      SFEM_RING_ROW(1, canon_edge_curl_k1)
    }
#undef SFEM_RING_ROW
#undef SFEM_SLOT
    rc1a = rc2a; rc1b = rc2b;
  }
  // the row edge's own block: both lanes' partial sums, stored by the even lane
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const double2 other = shfl_xor1(self[k]);
    self[k].x += other.x;
    self[k].y += other.y;
  }
  __syncwarp();
  if (valid && h == 0) {
    row0[self_slot] = self[0];
    row0[len + self_slot] = self[1];
    unsigned pr = pecraw;
    keep_opaque(pr);  // loaded at the start of the batch; first use here
    if (pr) pec_overwrite(decode_pecraw(pr), row0, len, self_slot);
  }
  tile_store_issue(tile, out_first, (unsigned)span * 16u, lane);
}

__global__ void __launch_bounds__(kThreadsEdges, SFEM_EDGE_CTAS) k_canon_edges(CanonArgs p) {
  extern __shared__ double2 smem[];
  __shared__ unsigned long long s_desc[16];
  __shared__ unsigned s_fsel[16];
  stage_lut(s_desc, s_fsel);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, h = lane & 1;
  double2* tile = smem + (size_t)w * p.tile_chunks;
  const int64_t nb = (p.r_end - p.r_begin + 15) / 16;
  const int64_t W = (int64_t)gridDim.x * (kThreadsEdges / 32);
  int64_t b = blockIdx.x * (int64_t)(kThreadsEdges / 32) + w;
  if (b >= nb) return;
  // prologue: this lane's first item and the record of its second one
  EdgeMeta m = load_edge_meta(p, b, lane);
  Item cur;
  uint4 rc1a = make_uint4(0, 0, 0, 0), rc1b = rc1a;
  {
    const int64_t i = m.i0 + h;
    if (i < m.i1) {
      const uint4 ra = p.rec[2 * i], rb = p.rec[2 * i + 1];
      if (i + 2 < m.i1) { rc1a = p.rec[2 * (i + 2)]; rc1b = p.rec[2 * (i + 2) + 1]; }
      gather_item<true>(p, ra, rb, cur);
    }
  }
#pragma unroll 1
  for (; b < nb; b += W) {
    EdgeMeta mn = load_edge_meta_raw(p, b + W, lane);  // beyond the end: empty; finished after the S pass
    const bool valid = m.i0 <= m.i1 && m.a1 > m.a;
    // the warp's span: chunks [2 a_first, 2 a_last_end) of the value arrays (lane pairs are consecutive entities)
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    const int last = 31 - __clz((int)vmask);
    const int64_t a_first = __shfl_sync(0xffffffffu, m.a, 0);
    const int64_t a_end = __shfl_sync(0xffffffffu, m.a1, last);
    const int span = (int)(2 * (a_end - a_first));
    const int off = (int)(2 * (m.a - a_first));
    int nsteps = valid ? (int)((m.i1 - m.i0 + 1) >> 1) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nsteps = max(nsteps, __shfl_xor_sync(0xffffffffu, nsteps, o));
    const unsigned pecraw = load_pecraw(p, p.ent_lo + p.r_begin + b * 16 + (lane >> 1), valid);
    edge_pass<false>(p, s_desc, s_fsel, tile, lane, m, pecraw, off, span, nsteps, cur, rc1a, rc1b, m.i0 + h, m.i1,
                     p.S + 4 * a_first);
    finish_edge_meta(p, mn);
    edge_pass<true>(p, s_desc, s_fsel, tile, lane, m, pecraw, off, span, nsteps, cur, rc1a, rc1b, mn.i0 + h, mn.i1,
                    p.T + 4 * a_first);
    m = mn;
  }
  tile_store_wait(lane);
}

// ---- edges, lane per item: ONE LANE PER (edge, cell) ITEM, no loop over the cells of an edge ---------------------
// A warp takes a BATCH of whole consecutive edges with at most 32 items and kE2TileChunks chunks of output
// (pattern.cu: k_e2_batches); lane j owns item j of the batch, items of an edge sit on consecutive lanes in ring order.
//   * what two neighbouring cells of a ring contribute to their three common column entities goes from the lane of
//     the previous cell to the lane of the next one by ONE indexed shuffle (source = lane - 1, or the last lane of
//     the ring for the first cell of a closed ring): every CSR slot is stored exactly once, closed rings included,
//     and nothing is carried from step to step -- there are no steps;
//   * the row edge's own block is a segmented suffix sum over the lanes of the edge (shuffles), stored by its first lane;
//   * the item is gathered ONCE for S and T; records, aux words, vertex coordinates and coefficients of the batches
//     ahead arrive through cp.async in per-warp shared-memory stages instead of registers (records two batches ahead,
//     coordinates one), so the prefetch costs no registers and no scoreboard stalls.
struct Edge2Args {
  const double* __restrict__ coords4;
  const double* __restrict__ eps;
  const double* __restrict__ mu;
  const double2* __restrict__ coef2;  // [C] (mu_r, eps_r) packed per call (one 16-byte gather per item)
  const uint4* __restrict__ rec;  // [nbatch][32][2]
  const uint4* __restrict__ aux;  // [nbatch][32]
  const uint8_t* __restrict__ pec;
  double* __restrict__ S;
  double* __restrict__ T;
  int64_t ent_lo;
  const int64_t* __restrict__ nbatch;  // number of batches (device: the pattern builder never reads it back)
};

#ifndef SFEM_E2_WARPS
#define SFEM_E2_WARPS 5
#endif
#ifndef SFEM_E2_CTAS
#define SFEM_E2_CTAS 2
#endif
// tiles per warp: 2 = S and T in separate tiles, so a pass never waits for the bulk copy of the pass before it
#ifndef SFEM_E2_TILES
#define SFEM_E2_TILES 2
#endif
constexpr int kE2Warps = SFEM_E2_WARPS;
// SFEM_E2_COORDS_SMEM = 1: vertex coordinates and coefficients of the next batch arrive through cp.async in a per-warp
// shared-memory stage (no registers, 12 warps per SM); 0: plain loads into registers one batch ahead (8-10 warps per SM).
#ifndef SFEM_E2_COORDS_SMEM
#define SFEM_E2_COORDS_SMEM 1
#endif
// per-warp shared memory: tile | [coordinates + coefficients stage [9][32] x 16 B] | 2 x (records [2][32] + aux [32]) x 16 B
constexpr int kE2StageCoords = SFEM_E2_COORDS_SMEM ? 9 * 32 : 0;  // uint4 units
constexpr int kE2StageRec = 3 * 32;                    // uint4 units per ping-pong half
constexpr int kE2WarpSmem = (SFEM_E2_TILES * kE2TileChunks + kE2StageCoords + 2 * kE2StageRec) * 16;

#ifdef SFEM_E2_NO_STREAM_HINT
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}
#endif
__device__ __forceinline__ void cp_async16_ca(void* sdst, const void* gsrc) {
#ifdef SFEM_E2_COORDS_CG
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
#else
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ double2 shfl_idx(double2 v, int src) {
  return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
__device__ __forceinline__ double2 shfl_down2(double2 v, int d) {
  return make_double2(__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d));
}

// records + aux words of batch b -> one ping-pong half of the warp's record stage (coalesced: 4 wavefronts per copy).
// They are read exactly once per step: L2 evict-first keeps the L2 for the gathered vertices and coefficients.
__device__ __forceinline__ void cp_async16_stream(void* sdst, const void* gsrc, unsigned long long policy) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)),
               "l"(gsrc), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void e2_issue_records(const Edge2Args& p, uint4* recst, int64_t b, int lane) {
  const uint4* r = p.rec + (b * 32 + lane) * 2;
#ifdef SFEM_E2_NO_STREAM_HINT
  cp_async16(recst + lane, r);
  cp_async16(recst + 32 + lane, r + 1);
  cp_async16(recst + 64 + lane, p.aux + b * 32 + lane);
#else
  unsigned long long policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  cp_async16_stream(recst + lane, r, policy);
  cp_async16_stream(recst + 32 + lane, r + 1, policy);
  cp_async16_stream(recst + 64 + lane, p.aux + b * 32 + lane, policy);
#endif
}
// Vertex data of the item whose record sits in `recst` -> the coordinate stage (cp.async).  A relabelled cell is
// (a, b, entry, exit) with (a, b) the row edge, and consecutive cells of a ring share a face: the exit vertex of one is
// the entry vertex of the next.  So an EDGE with m cells needs 2 + m (+1 for an open fan) vertices, not 4 m:
//   rows 0-3 of the stage: a, b (xy | z)   copied by the FIRST lane of the edge only, read by all its lanes
//   rows 4-5: entry vertex                 copied only where a chain starts without a predecessor (open fans)
//   rows 6-7: exit vertex                  every lane; the next cell of the ring reads it as its entry vertex
//   row  8:   (mu_r, eps_r) of the cell
// Every copy of a gathered address costs one shared-memory wavefront per returned sector (~15 per instruction for 32
// distinct vertices), which made the shared-memory pipe the limiter of this kernel; sharing cuts them by more than half.
__device__ __forceinline__ void e2_issue_coords(const Edge2Args& p, const uint4* recst, uint4* cst, int lane) {
  const uint4 ra = recst[lane];
  const uint4 rb = recst[32 + lane];
  const unsigned pk = recst[64 + lane].y;
  const int cell = (int)rb.x;
  if (cell >= 0) {
#if defined(SFEM_WHATIF_COORDS)
    const unsigned v[4] = {ra.x & 1023u, ra.y & 1023u, ra.z & 1023u, ra.w & 1023u};
#else
    const unsigned v[4] = {ra.x, ra.y, ra.z, ra.w};
#endif
    const bool first = ((pk >> 17) & 31u) == 0u;                       // first lane of its edge
    const bool chain_start = !((rb.y >> 8) & 1u) && (pk >> 27) == 0u;  // no predecessor in the ring
    if (first) {
      const double* xa = p.coords4 + (size_t)v[0] * 4;
      const double* xb = p.coords4 + (size_t)v[1] * 4;
      cp_async16_ca(cst + lane, xa);
      cp_async16_ca(cst + 32 + lane, xa + 2);
      cp_async16_ca(cst + 64 + lane, xb);
      cp_async16_ca(cst + 96 + lane, xb + 2);
    }
    if (chain_start) {
      const double* xe = p.coords4 + (size_t)v[2] * 4;
      cp_async16_ca(cst + 128 + lane, xe);
      cp_async16_ca(cst + 160 + lane, xe + 2);
    }
    const double* xx = p.coords4 + (size_t)v[3] * 4;
    cp_async16_ca(cst + 192 + lane, xx);
    cp_async16_ca(cst + 224 + lane, xx + 2);
    cp_async16_ca(cst + 256 + lane, p.coef2 + cell);
  }
}
// first: first lane of this lane's edge; src: lane of the previous cell of the ring (has_pred) or this lane
__device__ __forceinline__ void e2_read_coords(const uint4* cst, Item& it, int lane, int first, int src, bool has_pred) {
  const double2* c2 = reinterpret_cast<const double2*>(cst);
  const int e0 = has_pred ? 192 + src : 128 + lane;
  const double2 axy = c2[first], az = c2[32 + first], bxy = c2[64 + first], bz = c2[96 + first];
  const double2 exy = c2[e0], ez = c2[e0 + 32], xxy = c2[192 + lane], xz = c2[224 + lane];
  it.X[0] = axy.x; it.X[1] = axy.y; it.X[2] = az.x;
  it.X[3] = bxy.x; it.X[4] = bxy.y; it.X[5] = bz.x;
  it.X[6] = exy.x; it.X[7] = exy.y; it.X[8] = ez.x;
  it.X[9] = xxy.x; it.X[10] = xxy.y; it.X[11] = xz.x;
  const double2 c = c2[256 + lane];
  it.cmu = c.x; it.ceps = c.y;
}
// ... or -> registers (consumed one batch later)
#if !SFEM_E2_COORDS_SMEM
__device__ __forceinline__ void e2_gather_coords(const Edge2Args& p, const uint4* recst, Item& it, int lane) {
  const uint4 ra = recst[lane];
  const int cell = (int)recst[32 + lane].x;
  if (cell >= 0) {
    const unsigned v[4] = {ra.x, ra.y, ra.z, ra.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double2* x = reinterpret_cast<const double2*>(p.coords4 + (size_t)v[k] * 4);
      const double2 xy = x[0];
      const double z = p.coords4[(size_t)v[k] * 4 + 2];
      it.X[3 * k] = xy.x; it.X[3 * k + 1] = xy.y; it.X[3 * k + 2] = z;
    }
    it.cmu = p.mu[cell];
    it.ceps = p.eps[cell];
  }
}
#endif

template <bool MASS>
__device__ __forceinline__ void edge2_pass(double2* __restrict__ tile, int lane, bool active, const Slots& sl,
                                           unsigned pk, const double (&gm)[6], const Fix& fx, int src, double wprev,
                                           bool open_end, unsigned pecraw, int mmax, double* __restrict__ gdst,
                                           unsigned span) {
  const int off = (int)(pk & 511u), len = (int)((pk >> 9) & 255u), pos = (int)((pk >> 17) & 31u),
            rem = (int)((pk >> 22) & 31u);
  // the bulk store that last read this tile must have finished reading it before it is overwritten: with one tile
  // that is the copy of the pass before, with two tiles the one before that (at most one group may still be reading)
#if SFEM_E2_TILES == 2
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
  __syncwarp();
#else
  tile_store_wait(lane);
#endif
  double2* row0 = tile + off;
  double2 self[2];
  double o[20];
#if defined(SFEM_WHATIF_NOCONFLICT)
#define SFEM_E2_SLOT(U) (((U) * 32 + lane) - off)
#else
#define SFEM_E2_SLOT(U) slot_of<U>(sl)
#endif
#define SFEM_E2_ROW(K, FN)                                                                            \
  {                                                                                                   \
    double2 n3 = make_double2(0.0, 0.0), n1 = n3, n8 = n3;                                            \
    if (active) {                                                                                     \
      FN(gm, o);                                                                                      \
      n3 = fix_columns<3, true>(fx, o[6], o[7]);                                                      \
      n1 = fix_columns<1, true>(fx, o[2], o[3]);                                                      \
      n8 = fix_columns<8>(fx, o[16], o[17]);                                                          \
    }                                                                                                 \
    /* the next-side blocks of the previous cell of the ring (weight 0: none) */                      \
    const double2 g3 = shfl_idx(n3, src), g1 = shfl_idx(n1, src), g8 = shfl_idx(n8, src);             \
    self[K] = make_double2(0.0, 0.0);                                                                 \
    if (active) {                                                                                     \
      double2* row = row0 + K * len;                                                                  \
      double2 r;                                                                                      \
      self[K] = fix_columns<5>(fx, o[10], o[11]);                                                     \
      r = fix_columns<4, true>(fx, o[8], o[9]);                                                       \
      r.x = fma(g3.x, wprev, r.x); r.y = fma(g3.y, wprev, r.y);                                       \
      row[SFEM_E2_SLOT(4)] = r;                                                                        \
      r = fix_columns<2, true>(fx, o[4], o[5]);                                                       \
      r.x = fma(g1.x, wprev, r.x); r.y = fma(g1.y, wprev, r.y);                                       \
      row[SFEM_E2_SLOT(2)] = r;                                                                        \
      r = fix_columns<9>(fx, o[18], o[19]);                                                           \
      r.x = fma(g8.x, wprev, r.x); r.y = fma(g8.y, wprev, r.y);                                       \
      row[SFEM_E2_SLOT(9)] = r;                                                                        \
      row[SFEM_E2_SLOT(0)] = fix_columns<0, true>(fx, o[0], o[1]);                                     \
      row[SFEM_E2_SLOT(7)] = fix_columns<7>(fx, o[14], o[15]);                                         \
      row[SFEM_E2_SLOT(6)] = fix_columns<6>(fx, o[12], o[13]);                                         \
      if (open_end) {                                                                                 \
        row[SFEM_E2_SLOT(3)] = n3;                                                                     \
        row[SFEM_E2_SLOT(1)] = n1;                                                                     \
        row[SFEM_E2_SLOT(8)] = n8;                                                                     \
      }                                                                                               \
    }                                                                                                 \
  }
  if (MASS) {
    SFEM_E2_ROW(0, canon_edge_mass_k0)
    SFEM_E2_ROW(1, canon_edge_mass_k1)
  } else {
    SFEM_E2_ROW(0, canon_edge_curl_k0)
    SFEM_E2_ROW(1, canon_edge_curl_k1)
  }
#undef SFEM_E2_ROW
#undef SFEM_E2_SLOT
  // the row edge's own block: suffix sum over the lanes of the edge (fixed order: bit-reproducible)
#pragma unroll 1
  for (int d = 1; d < mmax; d <<= 1) {
    const double wd = d <= rem ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const double2 t = shfl_down2(self[k], d);
      self[k].x = fma(t.x, wd, self[k].x);
      self[k].y = fma(t.y, wd, self[k].y);
    }
  }
  const int self_slot = slot_of<5>(sl);
  if (active && pos == 0) {
    row0[self_slot] = self[0];
    row0[len + self_slot] = self[1];
  }
  unsigned pr = pecraw;
  keep_opaque(pr);  // loaded at the start of the batch; first use here
  if (__any_sync(0xffffffffu, pr != 0u)) {
    // row-only ident fused: the lanes of a constrained edge clear its rows, the first one sets the unit entries
    const unsigned pecbits = decode_pecraw(pr);
    __syncwarp();
    const int m = pos + rem + 1;
#pragma unroll
    for (int k = 0; k < 2; ++k)
      if ((pecbits >> k) & 1u)
        for (int s = pos; s < len; s += m) row0[k * len + s] = make_double2(0.0, 0.0);
    __syncwarp();
    if (pos == 0) {
      if (pecbits & 1u) row0[self_slot] = make_double2(1.0, 0.0);
      if (pecbits & 2u) row0[len + self_slot] = make_double2(0.0, 1.0);
    }
  }
  tile_store_issue(tile, gdst, span * 16u, lane);
}

#ifdef SFEM_E2_MAXREG
__global__ void __maxnreg__(SFEM_E2_MAXREG) k_canon_edges2(Edge2Args p) {
#else
__global__ void __launch_bounds__(kE2Warps * 32, SFEM_E2_CTAS) k_canon_edges2(Edge2Args p) {
#endif
  extern __shared__ uint4 smem4[];
  __shared__ unsigned long long s_desc[16];
  __shared__ unsigned s_fsel[16];
  stage_lut(s_desc, s_fsel);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint4* base = smem4 + (size_t)w * (kE2WarpSmem / 16);
  double2* tile = reinterpret_cast<double2*>(base);
  uint4* cst = base + SFEM_E2_TILES * kE2TileChunks;
  uint4* recst = cst + kE2StageCoords;
  const int64_t nb = *p.nbatch;
  const int64_t W = (int64_t)gridDim.x * kE2Warps;
  int64_t b = blockIdx.x * (int64_t)kE2Warps + w;
  if (b >= nb) return;
  // prologue: records of the first batch, then its coordinates together with the records of the second one
  Item cur;
  e2_issue_records(p, recst, b, lane);
  cp_async_commit();
  cp_async_wait_all();
#if SFEM_E2_COORDS_SMEM
  e2_issue_coords(p, recst, cst, lane);
#else
  e2_gather_coords(p, recst, cur, lane);
#endif
  if (b + W < nb) e2_issue_records(p, recst + kE2StageRec, b + W, lane);
  cp_async_commit();
  int ps = 0;
#pragma unroll 1
  for (; b < nb; b += W, ps ^= 1) {
#if SFEM_E2_COORDS_SMEM
    cp_async_wait_all();  // coordinates of this batch, records of the next one
#endif
    uint4* rcur = recst + ps * kE2StageRec;
    const uint4 rb = rcur[32 + lane];
    const uint4 ax = rcur[64 + lane];
    const bool active = (int)rb.x >= 0;
    Slots sl;
    sl.w1 = rb.y; sl.w2 = rb.z; sl.w3 = rb.w;
    const unsigned flags = active ? flags_of(sl) : 0u;
    const unsigned pk = ax.y;
    const int delta = (int)(pk >> 27);
    const bool has_prev = flags & 1u;
    const int src = has_prev ? lane - 1 : lane + delta;
    const bool has_pred = has_prev || delta;
    const double wprev = has_pred ? 1.0 : 0.0;
    const bool open_end = !(flags & 2u) && !(flags & 4u);
    // geometry of both forms from the coordinates gathered one batch ago
    double gS[6], gT[6];
    Fix fx;
    fx.desc = 0ull; fx.fsel = 0u;
    unsigned pecraw = 0u;
#if SFEM_E2_COORDS_SMEM
    __syncwarp();  // the stage is read across lanes
#endif
    if (active) {
#if SFEM_E2_COORDS_SMEM
      e2_read_coords(cst, cur, lane, lane - (int)((pk >> 17) & 31u), src, has_pred);
#endif
      canon_geometry<false>(cur, gS);
      canon_geometry<true>(cur, gT);
      fx.desc = s_desc[desc_index(sl)];
      fx.fsel = s_fsel[desc_index(sl)];
      if (p.pec) pecraw = *reinterpret_cast<const unsigned short*>(p.pec + 2 * (p.ent_lo + (int64_t)ax.x));
    }
    // the item buffer is free again: coordinates of the next batch (its records have landed), records of the one after
#if SFEM_E2_COORDS_SMEM
    __syncwarp();  // every lane has read the stage
    if (b + W < nb) e2_issue_coords(p, recst + (ps ^ 1) * kE2StageRec, cst, lane);
#else
    cp_async_wait_all();
    if (b + W < nb) e2_gather_coords(p, recst + (ps ^ 1) * kE2StageRec, cur, lane);
#endif
    if (b + 2 * W < nb) e2_issue_records(p, rcur, b + 2 * W, lane);
    cp_async_commit();
    const int mmax = __reduce_max_sync(0xffffffffu, active ? (int)(((pk >> 17) & 31u) + ((pk >> 22) & 31u) + 1u) : 0);
    const unsigned a_first = __shfl_sync(0xffffffffu, ax.z, 0), span = __shfl_sync(0xffffffffu, ax.w, 0);
    edge2_pass<false>(tile, lane, active, sl, pk, gS, fx, src, wprev, open_end, pecraw, mmax, p.S + 4 * (size_t)a_first, span);
    edge2_pass<true>(tile + (SFEM_E2_TILES - 1) * kE2TileChunks, lane, active, sl, pk, gT, fx, src, wprev, open_end, pecraw, mmax, p.T + 4 * (size_t)a_first, span);
  }
  tile_store_wait(lane);
}

// ---- faces: a PAIR of lanes per face, one lane per incident cell; the tile is store-only -------------------------
// Entities of the shared face (canonical 2,4,5 = its edges, 9 = itself) look alike from both cells (identity
// fix-up): the odd lane hands its values over by shuffle and the even lane stores the sums.
template <bool MASS>
__device__ __forceinline__ void face_pass(const CanonArgs& p, double2* __restrict__ tile, int lane, bool valid,
                                          bool active, unsigned pecbits, int len, int off, int span, const Slots& it,
                                          const double (&gm)[6], const Fix& fx, double* __restrict__ out_first) {
  const int h = lane & 1;
#if SFEM_FACE_TILES == 2
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
  __syncwarp();
#else
  tile_store_wait(lane);
#endif
  double2* row0 = tile + off;
  double o[20];
#define SFEM_FACE_ROW(K, FN)                                                                              \
  {                                                                                                       \
    double2 s2 = make_double2(0.0, 0.0), s4 = s2, s5 = s2, s9 = s2;                                       \
    if (active) {                                                                                         \
      FN(gm, o);                                                                                          \
      s2 = make_double2(o[4], o[5]); s4 = make_double2(o[8], o[9]);                                       \
      s5 = make_double2(o[10], o[11]); s9 = make_double2(o[18], o[19]);                                   \
    }                                                                                                     \
    const double2 t2 = shfl_xor1(s2), t4 = shfl_xor1(s4), t5 = shfl_xor1(s5), t9 = shfl_xor1(s9);         \
    if (active) {                                                                                         \
      double2* row = row0 + K * len;                                                                      \
      row[slot_of<0>(it)] = fix_columns<0, true>(fx, o[0], o[1]);                                               \
      row[slot_of<1>(it)] = fix_columns<1, true>(fx, o[2], o[3]);                                               \
      row[slot_of<3>(it)] = fix_columns<3, true>(fx, o[6], o[7]);                                               \
      row[slot_of<6>(it)] = fix_columns<6>(fx, o[12], o[13]);                                             \
      row[slot_of<7>(it)] = fix_columns<7>(fx, o[14], o[15]);                                             \
      row[slot_of<8>(it)] = fix_columns<8>(fx, o[16], o[17]);                                             \
      if (h == 0) {  /* the odd lane of a boundary face is inactive and handed over zeros */              \
        row[slot_of<2>(it)] = make_double2(s2.x + t2.x, s2.y + t2.y);                                     \
        row[slot_of<4>(it)] = make_double2(s4.x + t4.x, s4.y + t4.y);                                     \
        row[slot_of<5>(it)] = make_double2(s5.x + t5.x, s5.y + t5.y);                                     \
        row[slot_of<9>(it)] = make_double2(s9.x + t9.x, s9.y + t9.y);                                     \
      }                                                                                                   \
    }                                                                                                     \
  }
  if (MASS) {
    SFEM_FACE_ROW(0, canon_face_mass_k0)
    SFEM_FACE_ROW(1, canon_face_mass_k1)
  } else {
    SFEM_FACE_ROW(0, canon_face_curl_k0)
    SFEM_FACE_ROW(1, canon_face_curl_k1)
  }
#undef SFEM_FACE_ROW
  __syncwarp();
  if (valid && h == 0 && pecbits) pec_overwrite(pecbits, row0, len, slot_of<9>(it));  // canonical 9 = the row face
  tile_store_issue(tile, out_first, (unsigned)span * 16u, lane);
}

// incidence range of this lane's face in batch `batch`, RAW: an instruction that consumes a load stalls the in-order
// warp until it lands, so the range is only turned into a record index one iteration later (face_item_index)
struct FaceRange {
  int64_t ip0, ip1;
};
__device__ __forceinline__ FaceRange load_face_range(const CanonArgs& p, int64_t batch, int64_t nb, int lane) {
  FaceRange f;
  f.ip0 = 0; f.ip1 = 0;
  const int64_t r = p.r_begin + batch * 16 + (lane >> 1);
  if (batch < nb && r < p.r_end) {
    const int64_t g = p.ent_lo + r;
    f.ip0 = p.inc_ptr[g];
    f.ip1 = p.inc_ptr[g + 1];
  }
  return f;
}
// record index of this lane's cell (or -1), block-relative
__device__ __forceinline__ int64_t face_item_index(const CanonArgs& p, const FaceRange& f, int lane) {
  const int64_t i = f.ip0 + (lane & 1);
  return i < f.ip1 ? i - p.inc_base : -1;
}

// Persistent warps over batches of 16 faces; the gather chain (incidence pointer -> record -> coordinates) of the
// batches 3, 2 and 1 iterations ahead is in flight while the current batch is computed.
__global__ void __launch_bounds__(kThreadsFaces, SFEM_FACE_CTAS) k_canon_faces(CanonArgs p) {
  extern __shared__ double2 smem[];
  __shared__ unsigned long long s_desc[16];
  __shared__ unsigned s_fsel[16];
  stage_lut(s_desc, s_fsel);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double2* tile = smem + (size_t)w * SFEM_FACE_TILES * p.tile_chunks;
  const int64_t nb = (p.r_end - p.r_begin + 15) / 16;
  const int64_t W = (int64_t)gridDim.x * (kThreadsFaces / 32);
  int64_t b = blockIdx.x * (int64_t)(kThreadsFaces / 32) + w;
  if (b >= nb) return;
  Item cur;
  int64_t i0 = face_item_index(p, load_face_range(p, b, nb, lane), lane);
  int64_t i1 = face_item_index(p, load_face_range(p, b + W, nb, lane), lane);
  int64_t i2 = face_item_index(p, load_face_range(p, b + 2 * W, nb, lane), lane);
  FaceRange f3 = load_face_range(p, b + 3 * W, nb, lane);
  uint4 r1a = make_uint4(0, 0, 0, 0), r1b = r1a;
  if (i0 >= 0) gather_item(p, p.rec[2 * i0], p.rec[2 * i0 + 1], cur);
  if (i1 >= 0) { r1a = p.rec[2 * i1]; r1b = p.rec[2 * i1 + 1]; }
  // row pointers of this lane's face, one batch ahead as well
  int64_t a = 0, a1 = 0;
  {
    const int64_t r = p.r_begin + b * 16 + (lane >> 1);
    if (r < p.r_end) { a = p.ent_rowptr[r]; a1 = p.ent_rowptr[r + 1]; }
  }
#pragma unroll 1
  for (; b < nb; b += W) {
    // stage 3 / 2 / 1 of the batches ahead
    const int64_t i3 = face_item_index(p, f3, lane);        // loaded one iteration ago
    f3 = load_face_range(p, b + 4 * W, nb, lane);
    uint4 r2a = make_uint4(0, 0, 0, 0), r2b = r2a;
    if (i2 >= 0) { r2a = p.rec[2 * i2]; r2b = p.rec[2 * i2 + 1]; }
    // everything that is needed from the current item, then its buffer receives the next batch's
    const bool active = i0 >= 0;
    const Slots sl = slots_of(cur);
    double gS[6], gT[6];
    Fix fx;
    fx.desc = 0ull; fx.fsel = 0u;
    if (active) {
      canon_geometry<false>(cur, gS);
      canon_geometry<true>(cur, gT);
      fx.desc = s_desc[desc_index(sl)];
      fx.fsel = s_fsel[desc_index(sl)];
    }
    if (i1 >= 0) gather_item(p, r1a, r1b, cur);
    // current batch
    const int64_t r = p.r_begin + b * 16 + (lane >> 1);
    const bool valid = r < p.r_end;
    int64_t an = 0, a1n = 0;
    {
      const int64_t rn = r + W * 16;
      if (b + W < nb && rn < p.r_end) { an = p.ent_rowptr[rn]; a1n = p.ent_rowptr[rn + 1]; }
    }
    const unsigned pecbits = load_pecbits(p, p.ent_lo + r, valid);
    const int len = (int)(a1 - a);
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    const int last = 31 - __clz((int)vmask);
    const int64_t a_first = __shfl_sync(0xffffffffu, a, 0);
    const int64_t a_end = __shfl_sync(0xffffffffu, a1, last);
    const int span = (int)(2 * (a_end - a_first));
    const int off = (int)(2 * (a - a_first));
    face_pass<false>(p, tile, lane, valid, active, pecbits, len, off, span, sl, gS, fx, p.S + 4 * a_first);
    face_pass<true>(p, tile + (SFEM_FACE_TILES - 1) * p.tile_chunks, lane, valid, active, pecbits, len, off, span, sl, gT, fx,
                    p.T + 4 * a_first);
    r1a = r2a; r1b = r2b;
    i0 = i1; i1 = i2; i2 = i3;
    a = an; a1 = a1n;
  }
  tile_store_wait(lane);
}

// ---- order 1 (Whitney elements): one CSR row per edge, one function per column edge ----------------------------
// Same decomposition as the order-2 edge kernel (lane pair per edge, ring order, carries by shuffle, every slot stored
// once), but a row is short (1 + 3 m values), so S and T are assembled in ONE sweep into two small tiles and written
// out with coalesced 8-byte streaming stores.  A reversed column edge only changes the sign of its function.
__device__ __forceinline__ double flip_sign(const Fix& f, int u, double v) {
  const unsigned sign = ((unsigned)(f.desc >> 32) << (15 - u)) & 0x80000000u;  // descriptor bit 48 + u -> bit 31
  return __hiloint2double(__double2hiint(v) ^ (int)sign, __double2loint(v));
}

__device__ __forceinline__ void canon_geometry_both(const Item& it, double (&gT)[6], double (&gS)[6]) {
  double J[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int d = 0; d < 3; ++d) J[d][a] = it.X[3 * (a + 1) + d] - it.X[d];
  double A[3][3];  // adj(J) = det J^-1
  A[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
  A[1][0] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
  A[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
  A[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
  A[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
  A[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
  A[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
  A[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
  A[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
  const double det = J[0][0] * A[0][0] + J[0][1] * A[1][0] + J[0][2] * A[2][0];
  const double ad = fabs(det);
  const double st = it.ceps / ad, ss = 1.0 / (ad * it.cmu);
  const int pa[6] = {0, 1, 2, 0, 0, 1}, pb[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    const int a = pa[q], b = pb[q];
    gT[q] = (A[a][0] * A[b][0] + A[a][1] * A[b][1] + A[a][2] * A[b][2]) * st;
    gS[q] = (J[0][a] * J[0][b] + J[1][a] * J[1][b] + J[2][a] * J[2][b]) * ss;
  }
}

constexpr int kThreadsEdges1 = 128;

__global__ void __launch_bounds__(kThreadsEdges1, 3) k_canon1_edges(CanonArgs p) {
  extern __shared__ double2 smem[];
  __shared__ unsigned long long s_desc[16];
  __shared__ unsigned s_fsel[16];
  stage_lut(s_desc, s_fsel);
  (void)s_fsel;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, h = lane & 1;
  double* tileS = reinterpret_cast<double*>(smem) + (size_t)w * 2 * p.tile_chunks;  // tile_chunks = doubles per matrix tile
  double* tileT = tileS + p.tile_chunks;
  const int64_t nb = (p.r_end - p.r_begin + 15) / 16;
  const int64_t W = (int64_t)gridDim.x * (kThreadsEdges1 / 32);
  int64_t b = blockIdx.x * (int64_t)(kThreadsEdges1 / 32) + w;
  if (b >= nb) return;
  const uint4* __restrict__ rec = p.rec;
  EdgeMeta m = load_edge_meta(p, b, lane);
  Item cur;
  uint4 rc1a = make_uint4(0, 0, 0, 0), rc1b = rc1a;
  {
    const int64_t i = m.i0 + h;
    if (i < m.i1) {
      const uint4 ra = rec[2 * i], rb = rec[2 * i + 1];
      if (i + 2 < m.i1) { rc1a = rec[2 * (i + 2)]; rc1b = rec[2 * (i + 2) + 1]; }
      gather_item(p, ra, rb, cur);
    }
  }
#pragma unroll 1
  for (; b < nb; b += W) {
    const EdgeMeta mn = load_edge_meta(p, b + W, lane);  // beyond the end: empty
    const bool valid = m.i0 <= m.i1 && m.a1 > m.a;
    const int len = (int)(m.a1 - m.a);
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    const int last = 31 - __clz((int)vmask);
    const int64_t a_first = __shfl_sync(0xffffffffu, m.a, 0);
    const int64_t a_end = __shfl_sync(0xffffffffu, m.a1, last);
    const int span = (int)(a_end - a_first);
    const int off = (int)(m.a - a_first);
    int nsteps = valid ? (int)((m.i1 - m.i0 + 1) >> 1) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nsteps = max(nsteps, __shfl_xor_sync(0xffffffffu, nsteps, o));
    const int64_t g = p.ent_lo + p.r_begin + b * 16 + (lane >> 1);
    const bool pecrow = valid && p.pec && p.pec[g];
    const int64_t nf = mn.i0 + h, nf_end = mn.i1;
    uint4 rfa = make_uint4(0, 0, 0, 0), rfb = rfa;
    if (nf < nf_end) { rfa = rec[2 * nf]; rfb = rec[2 * nf + 1]; }
    double* rowS = tileS + off;
    double* rowT = tileT + off;
    double selfS = 0.0, selfT = 0.0;
    double savedS[2] = {0.0, 0.0}, savedT[2] = {0.0, 0.0};
    int self_slot = 0;
    int64_t i = m.i0 + h;
    __syncwarp();  // the previous batch's write-out has finished reading the tiles
#pragma unroll 1
    for (int s = 0; s < nsteps; ++s, i += 2) {
      const bool active = valid && i < m.i1;
      const Slots sl = slots_of(cur);
      const unsigned flags = active ? flags_of(sl) : 0u;
      const bool has_prev = flags & 1u, has_next = flags & 2u, wrap = flags & 4u;
      double gT[6], gS[6];
      Fix fx;
      fx.desc = 0ull; fx.fsel = 0u;
      if (active) {
        canon_geometry_both(cur, gT, gS);
        fx.desc = s_desc[desc_index(sl)];
        if (s == 0 && h == 0) self_slot = slot_of<5>(sl);
      }
      uint4 rc2a = make_uint4(0, 0, 0, 0), rc2b = rc2a;
      if (s + 1 < nsteps) {
        if (i + 2 < m.i1) {
          if (i + 4 < m.i1) { rc2a = rec[2 * (i + 4)]; rc2b = rec[2 * (i + 4) + 1]; }
          gather_item(p, rc1a, rc1b, cur);
        }
      } else if (nf < nf_end) {
        if (nf + 2 < nf_end) { rc2a = rec[2 * (nf + 2)]; rc2b = rec[2 * (nf + 2) + 1]; }
        gather_item(p, rfa, rfb, cur);
      }
      double oS[6], oT[6];
#pragma unroll
      for (int u = 0; u < 6; ++u) { oS[u] = 0.0; oT[u] = 0.0; }
      if (active) {
        canon1_edge_curl(gS, oS);
        canon1_edge_mass(gT, oT);
#pragma unroll
        for (int u = 0; u < 5; ++u) { oS[u] = flip_sign(fx, u, oS[u]); oT[u] = flip_sign(fx, u, oT[u]); }
      }
      // next-side edges (3 = (0,3), 1 = (1,3)) go to the partner lane, which owns the next cell of the ring
      const double gS3 = __shfl_xor_sync(0xffffffffu, oS[3], 1), gS1 = __shfl_xor_sync(0xffffffffu, oS[1], 1);
      const double gT3 = __shfl_xor_sync(0xffffffffu, oT[3], 1), gT1 = __shfl_xor_sync(0xffffffffu, oT[1], 1);
      const double cS4 = sel_flag(h, gS3, savedS[0]), cS2 = sel_flag(h, gS1, savedS[1]);
      const double cT4 = sel_flag(h, gT3, savedT[0]), cT2 = sel_flag(h, gT1, savedT[1]);
      savedS[0] = gS3; savedS[1] = gS1; savedT[0] = gT3; savedT[1] = gT1;
      __syncwarp();
      if (active) {
        selfS += oS[5];
        selfT += oT[5];
        const int s4 = slot_of<4>(sl), s2 = slot_of<2>(sl), s0 = slot_of<0>(sl);
        rowS[s4] = oS[4] + sel_flag(has_prev, cS4, 0.0);
        rowT[s4] = oT[4] + sel_flag(has_prev, cT4, 0.0);
        rowS[s2] = oS[2] + sel_flag(has_prev, cS2, 0.0);
        rowT[s2] = oT[2] + sel_flag(has_prev, cT2, 0.0);
        rowS[s0] = oS[0];
        rowT[s0] = oT[0];
        if (!has_next) {
          const int s3 = slot_of<3>(sl), s1 = slot_of<1>(sl);
          double a3 = oS[3], a1 = oS[1], b3 = oT[3], b1 = oT[1];
          if (wrap) { a3 += rowS[s3]; a1 += rowS[s1]; b3 += rowT[s3]; b1 += rowT[s1]; }
          rowS[s3] = a3; rowS[s1] = a1; rowT[s3] = b3; rowT[s1] = b1;
        }
      }
      rc1a = rc2a; rc1b = rc2b;
    }
    selfS += __shfl_xor_sync(0xffffffffu, selfS, 1);
    selfT += __shfl_xor_sync(0xffffffffu, selfT, 1);
    __syncwarp();
    if (valid && h == 0) {
      rowS[self_slot] = selfS;
      rowT[self_slot] = selfT;
      if (pecrow) {  // row-only ident fused: unit row
        for (int k = 0; k < len; ++k) { rowS[k] = 0.0; rowT[k] = 0.0; }
        rowS[self_slot] = 1.0;
        rowT[self_slot] = 1.0;
      }
    }
    __syncwarp();
    for (int j = lane; j < span; j += 32) {
      __stcs(p.S + a_first + j, tileS[j]);
      __stcs(p.T + a_first + j, tileT[j]);
    }
    m = mn;
  }
}

}  // namespace

static bool canonical1_supported(const sfem_pattern_s* pat) {
  return pat->space->mesh->gdim == 3 && pat->space->order == 1 && pat->rec != nullptr && pat->ring_ok &&
         (size_t)(kThreadsEdges1 / 32) * 2 * 16 * pat->max_row_len_edges * sizeof(double) <= 64 * 1024;
}

bool canonical_supported(const sfem_pattern_s* pat) {
  if (pat->space->order == 1) return canonical1_supported(pat);
  // order-2 tets, packed records built (rows <= 255 entities), edges in ring order (manifold, <= 64 cells), faces
  // between at most two cells, and the edge kernel's tiles (16 edges x 2 rows per warp) must fit in shared memory
  const size_t edge_smem = pat->rec_e2 ? 0 : (size_t)(kThreadsEdges / 32) * 16 * 2 * pat->max_row_len_edges * sizeof(double2);
  return pat->space->mesh->gdim == 3 && pat->space->order == 2 && pat->rec != nullptr && pat->ring_ok &&
         pat->max_row_len_other <= 16 && edge_smem <= 200 * 1024;
}

int canonical_launches(const sfem_pattern_s* pat) {
  if (pat->space->order == 1) return 1;
  const int64_t E = pat->space->E;
  // edge rows: k_pack_coef + k_canon_edges2 (or the lane-pair kernel alone); face rows: k_canon_faces
  // + cell geometry and row gather of the generic kernel for the rows of the per-entity fallback
  return (pat->ent_lo < E ? (pat->rec_e2 ? 2 : 1) : 0) + (pat->ent_hi > E ? 1 : 0) + (pat->n_gen_rows > 0 ? 2 : 0);
}

// (mu_r, eps_r) of every cell side by side: one 16-byte gather per item instead of two 8-byte ones
__global__ void k_pack_coef(const double* __restrict__ mu, const double* __restrict__ eps, double2* __restrict__ out, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_double2(mu ? mu[i] : 1.0, eps ? eps[i] : 1.0);
}

__global__ void k_fill_ones(double* a, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) a[i] = 1.0;
}

int assemble_canonical_ST(const sfem_pattern_s* pat, const double* eps, const double* mu, const uint8_t* pec,
                          double* S, double* T, cudaStream_t st) {
  sfem_space_s* sp = pat->space;
  const sfem_mesh_s* m = sp->mesh;
  const double* eps_in = eps;
  const double* mu_in = mu;
  if (!eps || !mu) {  // default coefficients: the kernels always read a per-cell array
    if (!sp->ones) {
      SFEM_CUDA(dev_malloc(&sp->ones, m->C * sizeof(double), st));
      k_fill_ones<<<(unsigned)((m->C + 255) / 256), 256, 0, st>>>(sp->ones, m->C);
    }
    if (!eps) eps = sp->ones;
    if (!mu) mu = sp->ones;
  }
  CanonArgs p;
  p.coords4 = sp->coords4; p.cells = m->cells; p.eps = eps; p.mu = mu;
  p.inc_ptr = sp->inc_ptr; p.ent_rowptr = pat->ent_rowptr; p.rec = pat->rec; p.pec = pec; p.S = S; p.T = T;
  p.ent_lo = pat->ent_lo; p.inc_base = pat->inc_base;
  const int64_t E = sp->E;
  const int64_t n_edge_rows = std::max<int64_t>(0, std::min(pat->ent_hi, E) - pat->ent_lo);
  const int64_t n_rows = pat->n_rows_ent;
  int num_sms = 148;
  SFEM_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, sp->device));
  if (sp->order == 1) {
    if (n_rows == 0) return 0;
    p.r_begin = 0; p.r_end = n_rows;
    p.tile_chunks = 16 * pat->max_row_len_edges;  // doubles per matrix tile
    const size_t smem = (size_t)(kThreadsEdges1 / 32) * 2 * p.tile_chunks * sizeof(double);
    SFEM_CUDA(cudaFuncSetAttribute(k_canon1_edges, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t nblocks = (n_rows + kThreadsEdges1 / 2 - 1) / (kThreadsEdges1 / 2);
    const unsigned grid = (unsigned)std::min<int64_t>(nblocks, (int64_t)num_sms * 3);
    k_canon1_edges<<<grid, kThreadsEdges1, smem, st>>>(p);
    SFEM_CUDA(cudaGetLastError());
    return 0;
  }
  if (n_edge_rows > 0 && pat->rec_e2) {
    if (!sp->coef2) SFEM_CUDA(dev_malloc(&sp->coef2, m->C * sizeof(double2), st));
    k_pack_coef<<<(unsigned)((m->C + 255) / 256), 256, 0, st>>>(mu_in, eps_in, sp->coef2, m->C);
    Edge2Args q;
    q.coords4 = sp->coords4; q.eps = eps; q.mu = mu; q.coef2 = sp->coef2; q.rec = pat->rec_e2; q.aux = pat->aux_e2; q.pec = pec;
    q.S = S; q.T = T; q.ent_lo = pat->ent_lo; q.nbatch = pat->nbatch_e2_dev;
    const size_t smem = (size_t)kE2Warps * kE2WarpSmem;
    SFEM_CUDA(cudaFuncSetAttribute(k_canon_edges2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t nblocks = (pat->n_batch_e2 + kE2Warps - 1) / kE2Warps;
    const unsigned grid = (unsigned)std::min<int64_t>(nblocks, (int64_t)num_sms * SFEM_E2_CTAS);  // persistent
    k_canon_edges2<<<grid, kE2Warps * 32, smem, st>>>(q);
  } else if (n_edge_rows > 0) {
    p.r_begin = 0; p.r_end = n_edge_rows;
    p.tile_chunks = 16 * 2 * pat->max_row_len_edges;
    const size_t smem = (size_t)(kThreadsEdges / 32) * p.tile_chunks * sizeof(double2);
    SFEM_CHECK(smem <= 227 * 1024, "edge valence too high for the canonical kernel's shared-memory tile; use the generic kernel");
    SFEM_CUDA(cudaFuncSetAttribute(k_canon_edges, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t nblocks = (n_edge_rows + kThreadsEdges / 2 - 1) / (kThreadsEdges / 2);
    const unsigned grid = (unsigned)std::min<int64_t>(nblocks, (int64_t)num_sms * SFEM_EDGE_CTAS);  // persistent
    k_canon_edges<<<grid, kThreadsEdges, smem, st>>>(p);
  }
  if (n_rows > n_edge_rows) {
    // a face shared by more than two cells is not a valid tetrahedral mesh
    SFEM_CHECK(pat->max_row_len_other <= 16, "a face is shared by more than two cells: not a valid tetrahedral mesh");
    p.r_begin = n_edge_rows; p.r_end = n_rows;
    p.tile_chunks = 16 * 2 * 16;
    const size_t smem = (size_t)(kThreadsFaces / 32) * SFEM_FACE_TILES * p.tile_chunks * sizeof(double2);
    SFEM_CUDA(cudaFuncSetAttribute(k_canon_faces, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t nblocks = (n_rows - n_edge_rows + kThreadsFaces / 2 - 1) / (kThreadsFaces / 2);
    const unsigned grid = (unsigned)std::min<int64_t>(nblocks, (int64_t)num_sms * SFEM_FACE_CTAS);
    k_canon_faces<<<grid, kThreadsFaces, smem, st>>>(p);
  }
  SFEM_CUDA(cudaGetLastError());
  // per-entity fallback: the few edge rows the batched layout does not take (an edge with more cells than a warp has lanes,
  // a row longer than the tile) are assembled by the generic kernel into the same arrays
  if (pat->n_gen_rows > 0)
    SFEM_TRY(assemble_generic(pat, 3, eps_in, mu_in, nullptr, pec, S, T, st, false, true, pat->gen_rows, pat->n_gen_rows));
  return 0;
}

}  // namespace sfem
