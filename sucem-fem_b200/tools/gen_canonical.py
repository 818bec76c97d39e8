#!/usr/bin/env python
"""Generate csrc/generated/canonical_rows.inc: straight-line code for the element-matrix rows of the CANONICAL
row entity -- edge e5=(v0,v1) (local dofs 10,11) and face f3=(v0,v1,v2) (local dofs 18,19) of the order-2 UFC
reference tetrahedron -- and the look-up tables that map a cell relabelled into that frame back to its true
(vertex-sorted) frame.

Why: the value kernels give one thread one row entity.  Relabelling the cell's vertices so that the row entity
always sits at the canonical position makes the reference-tensor coefficients compile-time constants for every
thread of a warp (operands of the DFMAs come from the constant bank, structural zeros vanish), instead of
per-lane gathers from shared memory that cap the fp64 pipe at a quarter of its rate.  The price is a fix-up per
column entity, because a relabelled cell sees some entities with the opposite orientation:
    edge seen high->low      :  (phi_0, phi_1) = (-phi'_1, -phi'_0)
    face seen in another order:  phi = Q phi' with Q in {I, swap, ...} (entries 0, +-1), see FACE_Q below
Both follow from the dual-basis definition of the element: the functionals of an entity
transform among themselves, l = R l', hence the nodal functions with Q = R^-T.

Reads csrc/generated/n1curl_tables.inc (written by gen_tables.py); verified against the full-tensor path by
tests/test_canonical_tables.py and, on the GPU, by the generic kernel.
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GEN = os.path.join(HERE, '..', 'csrc', 'generated')
TET_EDGES = ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))
TET_FACES = ((1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2))


def load_table(name):
    txt = open(os.path.join(GEN, 'n1curl_tables.inc')).read()
    m = re.search(r'SFEM_TABLE double %s((?:\[\d+\])+) = \{(.*?)\};' % name, txt, re.S)
    dims = [int(x) for x in re.findall(r'\[(\d+)\]', m.group(1))]
    vals = [float(x) for x in m.group(2).replace('\n', ' ').split(',') if x.strip()]
    return np.array(vals).reshape(dims)


def relabel_lut(rowtype, t, swap=0):
    """sigma (relabelled vertex -> true local vertex), umap (canonical entity -> true local entity),
    Q (2x2 per canonical entity: phi_true = Q phi_canonical)."""
    eidx = {frozenset(e): i for i, e in enumerate(TET_EDGES)}
    fidx = {frozenset(f): i for i, f in enumerate(TET_FACES)}
    if rowtype == 'edge':
        i = TET_EDGES[t]
        others = [v for v in range(4) if v not in i]
        sigma = [i[0], i[1], others[swap], others[1 - swap]]
    else:
        f = TET_FACES[t]
        sigma = [f[0], f[1], f[2], t]
    umap, Q = [], []
    for (p, q) in TET_EDGES:
        a, b = sigma[p], sigma[q]
        umap.append(eidx[frozenset((a, b))])
        Q.append(np.eye(2) if a < b else np.array([[0, -1.], [-1, 0]]))
    E = np.eye(4)
    for (p, q, r) in TET_FACES:
        vs = [sigma[p], sigma[q], sigma[r]]
        umap.append(6 + fidx[frozenset(vs)])
        s = sorted(vs)
        tc = np.stack([E[vs[1]] - E[vs[0]], E[vs[2]] - E[vs[0]]])
        ta = np.stack([E[s[1]] - E[s[0]], E[s[2]] - E[s[0]]])
        Rm = np.linalg.lstsq(tc.T, ta.T, rcond=None)[0].T
        Q.append(np.round(np.linalg.inv(Rm).T))
    return sigma, umap, Q


# a face fix-up selects an ordered pair out of (v0, v1, w = -(v0+v1)).  Rows of Q -> codes 0,1,2; the kernel gets
# three bits per face: swap (u = (v1, v0) instead of (v0, v1)), w0 (out0 = w instead of u0), w1 (out1 = w instead of u1)
SEL = {(1, 0): 0, (0, 1): 1, (-1, -1): 2}


def face_bits(s0, s1):
    swap = int(s0 == 1 or s1 == 0)
    w0, w1 = int(s0 == 2), int(s1 == 2)
    u = (1, 0) if swap else (0, 1)
    assert (2 if w0 else u[0], 2 if w1 else u[1]) == (s0, s1)
    return swap | (w0 << 1) | (w1 << 2)


def pack_descriptor(rowtype, t, swap=0):
    """64-bit descriptor: bits 0-7 sigma (2 bits each), 8-47 umap (4 bits each), 48-53 edge flips,
    separate 16-bit word of face selectors: per canonical face f a nibble (swap, w0, w1), see face_bits."""
    sigma, umap, Q = relabel_lut(rowtype, t, swap)
    d = 0
    for i, s in enumerate(sigma):
        d |= s << (2 * i)
    for i, u in enumerate(umap):
        d |= u << (8 + 4 * i)
    for i in range(6):
        if Q[i][0, 0] == 0:
            d |= 1 << (48 + i)
    hi = 0
    for f in range(4):
        q = Q[6 + f]
        s0 = SEL[(int(q[0, 0]), int(q[0, 1]))]
        s1 = SEL[(int(q[1, 0]), int(q[1, 1]))]
        hi |= face_bits(s0, s1) << (4 * f)
    return d, hi


# Column EDGE entities that a relabelled cell can see reversed are emitted in the (d, m) = ((v0 - v1)/2, (v0 + v1)/2)
# basis: reversing an edge maps (v0, v1) -> (-v1, -v0), i.e. d -> d, m -> -m, so the kernel's fix-up is one sign flip
# and two adds (v0 = m + d, v1 = m - d) instead of a swap-and-negate of two doubles.
DM_ENTITIES = {'edge': (0, 1, 2, 3, 4), 'face': (0, 1, 3)}


def emit_rows(name, table, row, coefs, dm=(), ncols=20):
    """o[j] = sum_p table[row][j][p] * g[p]; coefficients referenced from the __constant__ pool kCanonCoef so that
    they become constant-bank operands of the DFMAs (literals would be materialised through UMOV pairs)."""
    out = '__device__ __forceinline__ void %s(const double (&g)[6], double (&o)[%d]) {\n' % (name, ncols)
    nterms = 0
    tab = np.array(table[row], dtype=np.float64)
    for u in dm:
        v0, v1 = tab[2 * u].copy(), tab[2 * u + 1].copy()
        tab[2 * u], tab[2 * u + 1] = 0.5 * (v0 - v1), 0.5 * (v0 + v1)
    for j in range(ncols):
        terms = [(p, tab[j, p]) for p in range(6) if tab[j, p] != 0.0]
        nterms += len(terms)
        if not terms:
            out += '  o[%d] = 0.0;\n' % j
            continue
        parts = []
        for p, c in terms:
            c = float(c)
            if c not in coefs:
                coefs[c] = len(coefs)
            parts.append('kCanonCoef[%d] * g[%d]' % (coefs[c], p))
        out += '  o[%d] = %s;\n' % (j, ' + '.join(parts))
    out += '}\n'
    return out, nterms


def main():
    mass, curl = load_table('kMassTet2'), load_table('kCurlTet2')
    src = '// GENERATED by tools/gen_canonical.py -- do not edit.\n'
    stats = {}
    coefs, body = {}, ''
    for nm, tab in (('mass', mass), ('curl', curl)):
        for rt, rows in (('edge', (10, 11)), ('face', (18, 19))):
            for k, row in enumerate(rows):
                code, n = emit_rows('canon_%s_%s_k%d' % (rt, nm, k), tab, row, coefs, DM_ENTITIES[rt])
                body += code
                stats['%s_%s_k%d' % (rt, nm, k)] = n
    # order 1 (Whitney): one function per edge, one row (the canonical edge e5 = local dof 5), 6 columns; a reversed
    # column edge only changes sign, so no change of basis is needed
    mass1, curl1 = load_table('kMassTet1'), load_table('kCurlTet1')
    for nm, tab in (('mass', mass1), ('curl', curl1)):
        code, n = emit_rows('canon1_edge_%s' % nm, tab, 5, coefs, ncols=6)
        body += code
        stats['order1_edge_%s' % nm] = n
    pool = sorted(coefs, key=coefs.get)
    src += '__constant__ double kCanonCoef[%d] = {%s};\n' % (len(pool), ', '.join(repr(c) for c in pool))
    src += body
    with open(os.path.join(GEN, 'canonical_rows.inc'), 'w') as f:
        f.write(src)
    # descriptors: low 64 bits (sigma, umap, edge flips), separate 16 bits of face selectors.
    #   kCanonDesc[t], t = 0..9: true local entity of the row entity in its cell (0-5 edges, 6-9 faces); the cell's
    #     other vertices follow in ascending order.
    #   kRingDesc[2 t + swap], t = 0..5, and [12 + f], f = 0..3: edge rows visited in RING order around the edge:
    #     relabelled vertices (a, b, entry, exit) with (entry, exit) = the cell's two other vertices in ascending
    #     (swap = 0) or descending (swap = 1) local order; face rows as above.
    lut = '// GENERATED by tools/gen_canonical.py -- do not edit.\n'
    lo, hi = zip(*([pack_descriptor('edge', t) for t in range(6)] + [pack_descriptor('face', t) for t in range(4)]))
    lut += 'SFEM_TABLE unsigned long long kCanonDesc[10] = {%s};\n' % ', '.join('0x%xull' % x for x in lo)
    lut += 'SFEM_TABLE unsigned int kCanonFaceSel[10] = {%s};\n' % ', '.join('0x%xu' % x for x in hi)
    lo, hi = zip(*([pack_descriptor('edge', t, sw) for t in range(6) for sw in range(2)] +
                   [pack_descriptor('face', t) for t in range(4)]))
    lut += 'SFEM_TABLE unsigned long long kRingDesc[16] = {%s};\n' % ', '.join('0x%xull' % x for x in lo)
    lut += 'SFEM_TABLE unsigned int kRingFaceSel[16] = {%s};\n' % ', '.join('0x%xu' % x for x in hi)
    with open(os.path.join(GEN, 'canonical_lut.inc'), 'w') as f:
        f.write(lut)
    sys.stderr.write('multiply-add terms: %r\n' % stats)


if __name__ == '__main__':
    main()
