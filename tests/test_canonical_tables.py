"""Host-side check of the generated canonical-frame code (no GPU): parse csrc/generated/canonical_rows.inc and
canonical_lut.inc exactly as the CUDA kernels consume them (descriptor bit layout included) and compare the rows they
produce, for every descriptor, with the oracle's full element matrices on random tetrahedra.  Also the generic tables
(tools/gen_tables.py, sympy) against the oracle's exact-rational ones (oracle/refelem.py)."""
import os
import re

import numpy as np

from oracle import fem, refelem
from conftest import ROOT

GEN = os.path.join(ROOT, 'sucem-fem_b200', 'csrc', 'generated')
TE, TF = refelem.TET_EDGES, refelem.TET_FACES


def _parse_rows():
    txt = open(os.path.join(GEN, 'canonical_rows.inc')).read()
    coef = np.array([float(x) for x in re.search(r'kCanonCoef\[\d+\] = \{(.*?)\};', txt, re.S).group(1).split(',')])
    fns = {}
    for name, body in re.findall(r'void (canon_\w+)\(const double \(&g\)\[6\], double \(&o\)\[20\]\) \{(.*?)\n\}', txt, re.S):
        M = np.zeros((20, 6))
        for j, expr in re.findall(r'o\[(\d+)\] = (.*?);', body):
            for ci, p in re.findall(r'kCanonCoef\[(\d+)\] \* g\[(\d)\]', expr):
                M[int(j), int(p)] += coef[int(ci)]
        fns[name] = M
    return fns


def _parse_lut(name):
    txt = open(os.path.join(GEN, 'canonical_lut.inc')).read()
    body = re.search(r'%s\[\d+\] = \{(.*?)\};' % name, txt).group(1)
    return [int(x.strip().rstrip('ul'), 16) for x in body.split(',')]


def _sym6(G):
    return np.array([G[0, 0], G[1, 1], G[2, 2], G[0, 1], G[0, 2], G[1, 2]])


DM = {'edge': (0, 1, 2, 3, 4), 'face': (0, 1, 3)}     # tools/gen_canonical.py DM_ENTITIES


def _fix(u, desc, fsel, v0, v1, kind):
    """the kernels' fix_columns<U, DM>"""
    if u < 6 and u in DM[kind]:
        d, m = v0, v1
        sm = -m if (desc >> (48 + u)) & 1 else m
        return sm + d, sm - d
    if u < 6:
        assert not (desc >> (48 + u)) & 1           # never seen reversed
        return v0, v1
    nib = (fsel >> (4 * (u - 6))) & 15
    w = -(v0 + v1)
    u0, u1 = (v1, v0) if nib & 1 else (v0, v1)
    return (w if nib & 2 else u0), (w if nib & 4 else u1)


def test_canonical_rows_and_descriptors_reproduce_true_rows():
    fns = _parse_rows()
    desc, fsel = _parse_lut('kRingDesc'), _parse_lut('kRingFaceSel')
    assert desc[:12:2] == _parse_lut('kCanonDesc')[:6] and desc[12:] == _parse_lut('kCanonDesc')[6:]
    rng = np.random.default_rng(11)
    cells = np.array([[0, 1, 2, 3]])
    for trial in range(3):
        co = rng.standard_normal((4, 3))
        eps, mu = rng.uniform(1, 16), rng.uniform(1, 2)
        Mfull = fem.element_mass(co, cells, 2, [eps])[0]
        Sfull = fem.element_stiffness(co, cells, 2, [mu])[0]
        for di in range(16):
            d, f = desc[di], fsel[di]
            sigma = [(d >> (2 * k)) & 3 for k in range(4)]
            umap = [(d >> (8 + 4 * u)) & 15 for u in range(10)]
            if di < 12:          # edge row: true local edge t = di // 2, relabelled (a, b, entry, exit)
                t, rows, kind = di // 2, (2 * (di // 2), 2 * (di // 2) + 1), 'edge'
                assert (sigma[0], sigma[1]) == TE[t]
            else:                # face row: true local face
                t, kind = di - 12, 'face'
                rows = (12 + 2 * t, 12 + 2 * t + 1)
                assert tuple(sigma[:3]) == TF[t] and sigma[3] == t
            x = co[sigma]
            J = (x[1:] - x[0]).T
            det = np.linalg.det(J)
            Ji = np.linalg.inv(J)
            gT = _sym6(Ji @ Ji.T) * abs(det) * eps
            gS = _sym6(J.T @ J) / (abs(det) * mu)
            for A, g, nm in ((Mfull, gT, 'mass'), (Sfull, gS, 'curl')):
                for k in range(2):
                    o = fns['canon_%s_%s_k%d' % (kind, nm, k)] @ g
                    got = np.zeros(20)
                    for u in range(10):
                        got[2 * umap[u]], got[2 * umap[u] + 1] = _fix(u, d, f, o[2 * u], o[2 * u + 1], kind)
                    assert np.allclose(got, A[rows[k]], rtol=1e-11, atol=1e-13 * np.abs(A).max()), (di, nm, k)


def test_generic_tables_match_oracle_tensors():
    txt = open(os.path.join(GEN, 'n1curl_tables.inc')).read()

    def table(name):
        m = re.search(r'SFEM_TABLE double %s((?:\[\d+\])+) = \{(.*?)\};' % name, txt, re.S)
        dims = [int(x) for x in re.findall(r'\[(\d+)\]', m.group(1))]
        return np.array([float(x) for x in m.group(2).replace('\n', ' ').split(',') if x.strip()]).reshape(dims)

    pairs3, pairs2 = [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)], [(0, 0), (1, 1), (0, 1)]
    for dim, order, tag in ((3, 1, 'Tet1'), (3, 2, 'Tet2'), (2, 1, 'Tri1'), (2, 2, 'Tri2')):
        pairs = pairs3 if dim == 3 else pairs2
        Mh, Ch, Fh = refelem.mass_tensor(dim, order), refelem.curl_tensor(dim, order), refelem.facet_tensor(dim, order)
        sym = lambda T: np.stack([T[a, b] + (T[b, a] if a != b else 0) for a, b in pairs], axis=-1)
        assert np.allclose(table('kMass' + tag), sym(Mh), rtol=1e-15, atol=1e-18)
        if dim == 3:
            assert np.allclose(table('kCurl' + tag), sym(Ch), rtol=1e-15, atol=1e-18)
        else:
            assert np.allclose(table('kCurl' + tag), Ch[0, 0], rtol=1e-15, atol=1e-18)
        symf = np.stack([np.stack([Fh[lf, a, b] + (Fh[lf, b, a] if a != b else 0) for a, b in pairs], axis=-1)
                         for lf in range(Fh.shape[0])])
        assert np.allclose(table('kFacet' + tag), symf, rtol=1e-15, atol=1e-18)
