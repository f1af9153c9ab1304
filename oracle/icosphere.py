"""Icosphere generator in meshzoo's vertex/cell order -- ORACLE, test infrastructure.

PARITY UNPINNED.  The reference obtains its mesh from the third-party package
``meshzoo`` (`util.py:6,43`: ``mz.icosa_sphere(divisions)``; ``requirements.yaml:9``
says ``meshzoo>=0.7.3``, no lock file).  meshzoo is not vendored in
/root/reference, not installed in this image and there is no network, so nothing
can pin this generator bit-for-bit.  It restates the published meshzoo 0.7.x
layout and is checked against every invariant the
reference relies on (tests/test_mesh.py): V = 10k^2+2, T = 20k^2, the valence-5
vertices are exactly 0..11 (`util.py:640-650`), consistent winding
(`util.py:608-610`), pole vertices for even k (`util.py:23`), triangles
[0,11,5] / [0,10,11] at k=1 (`util.py:619`).

Layout: [12 corners | 30 edges x (k-1) | 20 faces x (k-1)(k-2)/2 interiors].
This numpy version is vectorised per face; the CUDA generator
(nixis_b200/csrc/nxb_mesh.cu) is closed-form per vertex / per triangle and is
tested for BIT equality against this one, so the arithmetic is spelled out:
  edge point   p = (1 - j/k) * c[i0] + (j/k) * c[i1]                 j = 1..k-1
  face point   p = ((1 - i/k - j/k) * c0 + (j/k) * c1) + (i/k) * c2   i = 1..k-1, j = 1..k-i-1
  all points   p / sqrt((x*x + y*y) + z*z)
every product and sum rounded once to float64, left to right, no FMA.
"""
import numpy as np

_T = (1.0 + np.sqrt(5.0)) / 2.0
CORNERS = np.array(
    [[-1, _T, 0], [1, _T, 0], [-1, -_T, 0], [1, -_T, 0],
     [0, -1, _T], [0, 1, _T], [0, -1, -_T], [0, 1, -_T],
     [_T, 0, -1], [_T, 0, 1], [-_T, 0, -1], [-_T, 0, 1]], dtype=np.float64)

FACES = np.array(
    [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11),
     (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
     (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9),
     (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)], dtype=np.int64)

# CPython iteration order of the set of sorted corner pairs
EDGES = np.array(
    [(3, 4), (4, 9), (8, 9), (0, 5), (2, 11), (1, 9), (0, 11), (7, 10), (6, 8), (4, 5),
     (3, 9), (3, 6), (5, 9), (4, 11), (0, 1), (0, 7), (2, 4), (10, 11), (0, 10), (1, 5),
     (2, 10), (1, 8), (6, 7), (6, 10), (3, 8), (5, 11), (2, 3), (1, 7), (2, 6), (7, 8)],
    dtype=np.int64)

EDGE_ID = {(int(a), int(b)): e for e, (a, b) in enumerate(EDGES)}


def face_edge_table():
    """For each face and side s (0: c0->c1, 1: c1->c2, 2: c2->c0): (edge id, reversed?)."""
    tab = np.zeros((20, 3, 2), dtype=np.int64)
    for f, (a, b, c) in enumerate(FACES):
        for s, (p, q) in enumerate(((a, b), (b, c), (c, a))):
            rev = p > q
            tab[f, s] = (EDGE_ID[(min(p, q), max(p, q))], int(rev))
    return tab


def counts(k):
    return 10 * k * k + 2, 20 * k * k


def _local_to_global(k, f, ftab):
    """Translation table for face f: local node (row-major rows of shrinking
    length, row r has k-r+1 nodes) -> global vertex id."""
    n = k
    nn = (n + 1) * (n + 2) // 2
    tt = np.empty(nn, dtype=np.int64)
    c0, c1, c2 = FACES[f]
    row_start = np.concatenate([[0], np.cumsum(np.arange(n + 1, 0, -1))])[: n + 1]
    tt[0], tt[n], tt[nn - 1] = c0, c1, c2
    edge_base = lambda e: 12 + e * (n - 1)
    j = np.arange(n - 1)
    # side 0 along the first row
    e, rev = ftab[f, 0]
    tt[1:n] = edge_base(e) + (j[::-1] if rev else j)
    # side 1 up the right side: row r=1..n-1, last node of the row
    e, rev = ftab[f, 1]
    r = np.arange(1, n)
    tt[row_start[r] + (n - r)] = edge_base(e) + ((n - 2 - j) if rev else j)
    # side 2 (c2->c0) runs down the left side, so going up it is reversed
    e, rev = ftab[f, 2]
    tt[row_start[r]] = edge_base(e) + (j if rev else (n - 2 - j))
    # interiors
    base = 12 + 30 * (n - 1) + f * ((n - 1) * (n - 2) // 2)
    off = 0
    for r in range(1, n - 1):
        cnt = n - r - 1
        tt[row_start[r] + 1: row_start[r] + 1 + cnt] = base + off + np.arange(cnt)
        off += cnt
    return tt


def _local_cells(k):
    n = k
    out = np.empty((n * n, 3), dtype=np.int64)
    pos = 0
    start = 0
    for i in range(n):
        j = np.arange(n - i)
        up = np.stack([start + j, start + j + 1, start + n - i + j + 1], axis=1)
        out[pos:pos + len(up)] = up
        pos += len(up)
        j = np.arange(n - i - 1)
        dn = np.stack([start + j + 1, start + n - i + j + 2, start + n - i + j + 1], axis=1)
        out[pos:pos + len(dn)] = dn
        pos += len(dn)
        start += n - i + 1
    return out


def icosa_sphere(k):
    """points f64[V,3] on the unit sphere, cells int64[T,3]."""
    n = int(k)
    assert n >= 1
    V, T = counts(n)
    pts = np.empty((V, 3), dtype=np.float64)
    pts[:12] = CORNERS
    t = np.arange(1, n, dtype=np.float64) / n
    for e, (i0, i1) in enumerate(EDGES):
        pts[12 + e * (n - 1): 12 + (e + 1) * (n - 1)] = (
            np.outer(1 - t, CORNERS[i0]) + np.outer(t, CORNERS[i1]))
    ftab = face_edge_table()
    nint = (n - 1) * (n - 2) // 2
    if nint:
        ii = np.concatenate([np.full(n - i - 1, i) for i in range(1, n)]).astype(np.float64) / n
        jj = np.concatenate([np.arange(1, n - i) for i in range(1, n)]).astype(np.float64) / n
        bary = np.stack([1.0 - ii - jj, jj, ii])
    lc = _local_cells(n)
    cells = np.empty((T, 3), dtype=np.int64)
    for f in range(20):
        if nint:
            base = 12 + 30 * (n - 1) + f * nint
            c0, c1, c2 = CORNERS[FACES[f]]
            # explicit left-to-right products and sums (no BLAS: the CUDA generator
            # reproduces these roundings one by one)
            pts[base: base + nint] = (np.outer(bary[0], c0) + np.outer(bary[1], c1)) + np.outer(bary[2], c2)
        tt = _local_to_global(n, f, ftab)
        cells[f * n * n: (f + 1) * n * n] = tt[lc]
    norms = np.sqrt((pts[:, 0] * pts[:, 0] + pts[:, 1] * pts[:, 1]) + pts[:, 2] * pts[:, 2])
    pts = pts / norms[:, None]
    return np.ascontiguousarray(pts), cells
