"""Invariants the reference relies on for the meshzoo icosphere.
The generator itself is parity-unpinned (meshzoo is absent), so these are the pins there are."""
import numpy as np
import pytest

from oracle import icosphere


@pytest.mark.parametrize("k", [1, 2, 3, 4, 7, 8, 16, 33])
def test_icosphere_invariants(k, oracle):
    pts, cells = icosphere.icosa_sphere(k)
    V, T = 10 * k * k + 2, 20 * k * k
    assert pts.shape == (V, 3) and cells.shape == (T, 3)                    # util.py:27-30
    assert np.abs(np.linalg.norm(pts, axis=1) - 1).max() < 1e-15
    assert cells.min() == 0 and cells.max() == V - 1
    # consistent winding: every directed edge exactly once (util.py:608-610 needs it)
    e = np.concatenate([cells[:, [0, 1]], cells[:, [1, 2]], cells[:, [2, 0]]])
    key = e[:, 0] * V + e[:, 1]
    assert len(np.unique(key)) == len(key) == 3 * T
    rev = e[:, 1] * V + e[:, 0]
    assert np.array_equal(np.sort(key), np.sort(rev))                       # closed surface
    # valence-5 vertices are exactly 0..11 (util.py:640-650)
    val = np.bincount(e[:, 0], minlength=V)
    assert (val[:12] == 5).all() and (val[12:] == 6).all()
    # outward orientation (counter-clockwise seen from outside)
    a, b, c = pts[cells[:, 0]], pts[cells[:, 1]], pts[cells[:, 2]]
    assert (np.einsum("ij,ij->i", np.cross(b - a, c - a), a + b + c) > 0).all()
    if k % 2 == 0:                                                          # util.py:23
        for pole in ([0, 0, 1.0], [0, 0, -1.0]):
            assert np.abs(pts - pole).sum(1).min() < 1e-12
    if k == 1:
        tri = {tuple(t) for t in cells.tolist()}
        assert (0, 11, 5) in tri and (0, 10, 11) in tri                    # util.py:619


def test_spot_values_k8():
    pts, cells = icosphere.icosa_sphere(8)
    assert np.allclose(pts[12], [0.49063444264892453, -0.863953838921453, 0.11340902917961727], atol=1e-15)
    assert cells[:3].tolist() == [[0, 54, 33], [54, 55, 222], [55, 56, 223]]
    assert cells[-1].tolist() == [47, 159, 1]
