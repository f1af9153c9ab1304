"""Drop-in for the reference's `opensimplex` module (opensimplex.py) on B200.

Same names, argument order and defaults: `init`, `noise2d/3d/4d`, `noisearr2d/3d/4d`.
`init` runs on the host (256 iterations, opensimplex.py:90-112, including the int32
truncation of `over`); every noise evaluation runs in the sm_100a kernels.  Scalar calls
launch a 1-element kernel: they exist for API parity, arrays are the fast path.
Results are FP32-accurate values returned as float64, like the reference's dtype.
"""
import numpy as np
import torch

from . import runtime as rt

DEFAULT_SEED = 0          # opensimplex.py:35


def init(seed=DEFAULT_SEED):
    """-> (perm, perm_grad_index_3D), two int32[256] arrays, bit-identical to the reference."""
    return rt.init_perm(seed)


def _arr(tables, coords):
    dev = [rt.upload_f32(np.atleast_1d(np.asarray(c, dtype=np.float64)).ravel()) for c in coords]
    return rt.download_f64(rt.noise_array(tables, dev))


def noisearr2d(x, y, perm):
    return _arr(rt.tables_for(perm, None), (x, y))


def noisearr3d(x, y, z, perm, perm_grad_index_3D):
    return _arr(rt.tables_for(perm, perm_grad_index_3D), (x, y, z))


def noisearr4d(x, y, z, w, perm):
    return _arr(rt.tables_for(perm, None), (x, y, z, w))


def noise2d(x, y, perm):
    return float(noisearr2d([x], [y], perm)[0])


def noise3d(x, y, z, perm, perm_grad_index_3D):
    return float(noisearr3d([x], [y], [z], perm, perm_grad_index_3D)[0])


def noise4d(x, y, z, w, perm):
    return float(noisearr4d([x], [y], [z], [w], perm)[0])
