"""ctypes binding of libnixis_b200.so (the C-ABI in include/nixis_b200.h).

There is NO CPU fallback: if the shared library is missing or a call fails, this
module raises.  Build with `python -m nixis_b200.build` (nvcc, sm_100a).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("NXB_SO") or os.path.join(_HERE, "libnixis_b200.so")     # NXB_SO: A/B runs of kernel variants (tools/ab_probe.py)

_p = C.c_void_p
_i64 = C.c_int64
_i = C.c_int
_f = C.c_float
_d = C.c_double

# name -> (restype, argtypes); every symbol include/nixis_b200.h declares
SIGNATURES = {
    "nxb_version": (_i, []),
    "nxb_last_error": (_i, [C.c_char_p, _i]),
    "nxb_device_info": (_i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i64), C.POINTER(_i), C.POINTER(_i)]),
    "nxb_ffma_peak": (_i, [_i, C.POINTER(_d)]),
    "nxb_init_perm": (_i, [_i64, _p, _p]),
    "nxb_tables_create": (_i, [_p, _p, C.POINTER(_p)]),
    "nxb_tables_destroy": (_i, [_p]),
    "nxb_noise3_f32": (_i, [_p, _p, _p, _p, _i64, _p, _p]),
    "nxb_noise2_f32": (_i, [_p, _p, _p, _i64, _p, _p]),
    "nxb_noise4_f32": (_i, [_p, _p, _p, _p, _p, _i64, _p, _p]),
    "nxb_fbm3_f32": (_i, [_p, _p, _i64, _i, _p, _p, _p, _p, _p, _p]),
    "nxb_fbm3_pos64_f32": (_i, [_p, _p, _d, _i64, _i, _p, _p, _p, _p, _p, _p]),
    "nxb_fbm3_f64": (_i, [_p, _p, _i64, _i, _p, _p, _d, _d, _p, _p, _p]),
    "nxb_fbm4_f32": (_i, [_p, _p, _i64, _i, _p, _p, _p, _p, _p, _p, _p]),
    "nxb_mask_le_f32": (_i, [_p, _i64, _f, _p, _p]),
    "nxb_mesh_icosa_points": (_i, [_i, _i64, _i64, _p, _p, _p]),
    "nxb_mesh_icosa_cells": (_i, [_i, _i64, _i64, _p, _p]),
    "nxb_mesh_icosa_adj_rows_workspace": (_i64, [_i64]),
    "nxb_mesh_icosa_adj_rows": (_i, [_i, _i64, _i64, _p, _p, _p, _p]),
    "nxb_xyz_f64_to_f32": (_i, [_p, _i64, _d, _p, _p]),
    "nxb_ll_grid_f64": (_i, [_i, _i, _d, _p, _p]),
    "nxb_ico_nearest3_f64": (_i, [_i, _d, _p, _i64, _p, _p, _p]),
    "nxb_idw_gray_f64": (_i, [_p, _p, _p, _i64, _p, _p]),
    "nxb_idw_map": (_i, [_p, _p, _p, _i, _i64, _d, _d, _d, _d, _d, _i, _i, _p, _p]),
    "nxb_climate_surface_temp_f32": (_i, [_p, _i64, _d, _d, _p, _p]),
    "nxb_climate_insolation_f32": (_i, [_p, _i64, _d, _p, _i, _p, _i, _p, _p, _p]),
    "nxb_climate_slice_verts": (_i, [_d, _p]),
    "nxb_climate_seasonal_tilt": (_d, [_d, _d]),
    "nxb_climate_interpolate_f32": (_i, [_p, _i64, _d, _p, _i, _p, _p]),
    "nxb_adj_build_workspace": (_i64, [_i64]),
    "nxb_adj_build": (_i, [_p, _i64, _i64, _p, _p, _p]),
    "nxb_adj_sort": (_i, [_p, _p, _i64, _p]),
    "nxb_minmax_reset": (_i, [_p, _p]),
    "nxb_minmax_f32": (_i, [_p, _i64, _p, _p]),
    "nxb_rescale_f32": (_i, [_p, _i64, _f, _f, _f, _f, _i, _f, _i, _p, _p]),
    "nxb_power_summary_f32": (_i, [_p, _p, _i64, _i, _p, _p]),
    "nxb_power_apply_f32": (_i, [_p, _p, _i64, _i, _f, _f, _f, _f, _p, _p]),
    "nxb_edge_lengths_f64": (_i, [_p, _p, _i64, _p, _p]),
    "nxb_mesh_icosa_edge_lengths": (_i, [_i, _p, _i64, _i64, _d, _p, _p]),
    "nxb_erode_plan_bytes": (_i64, [_i64]),
    "nxb_erode_plan_build": (_i, [_p, _i64, _i64, _p, _p, _p]),
    "nxb_erode_dist3_floats": (_i64, [_i64]),
    "nxb_erode_dist3_build": (_i, [_p, _p, _i64, _p, _p]),
    "nxb_erode3_plan_step_f32": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _f, _p]),
    "nxb_erode3_run_f32": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _f, _i64, _p]),
    "nxb_erode3_run_comm_f32": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _f, _i64,
                                     _p, _i, _p, _p, _p, _p, _p, _i, C.c_uint32, _p, _p]),
    "nxb_erode3_step_f64": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _d, _p]),
    "nxb_erode1_step_f32": (_i, [_p, _p, _p, _i64, _i64, _p]),
    "nxb_halo_put_f32": (_i, [_p, _p, _i, _p, _p, _p, _p, _p, C.c_uint32, _p, _p]),
    "nxb_halo_wait": (_i, [_p, _p, _i, C.c_uint32, _p]),
    "nxb_halo_wait_stream": (_i, [_p, _p, _i, C.c_uint32, _p]),
    "nxb_peer_alloc": (_i, [_i64, C.POINTER(_p), _p]),
    "nxb_peer_open": (_i, [_p, C.POINTER(_p)]),
    "nxb_peer_close": (_i, [_p]),
    "nxb_peer_free": (_i, [_p]),
    "nxb_gather_f32": (_i, [_p, _p, _i64, _p, _p]),
    "nxb_scatter_f32": (_i, [_p, _p, _i64, _p, _p]),
    "nxb_f32_to_f64": (_i, [_p, _i64, _p, _p]),
    "nxb_f64_to_f32": (_i, [_p, _i64, _p, _p]),
}

_lib = None


class NxbError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises ImportError if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: the CUDA extension was not built "
            "(run `python -m nixis_b200.build`); nixis_b200 has no CPU fallback")
    lib = C.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    buf = C.create_string_buffer(512)
    load().nxb_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(rc, what=""):
    if rc != 0:
        raise NxbError(f"{what or 'nxb call'} failed with status {rc}: {last_error()}")


# kernels launched per successful call (for bench.py's gpu_launches claim)
KERNELS_PER_CALL = {"nxb_adj_build": 2, "nxb_mesh_icosa_adj_rows": 2, "nxb_ffma_peak": 5, "nxb_init_perm": 0, "nxb_tables_create": 0,
                    "nxb_tables_destroy": 0, "nxb_climate_slice_verts": 0, "nxb_halo_wait_stream": 0,
                    "nxb_peer_alloc": 0, "nxb_peer_open": 0, "nxb_peer_close": 0, "nxb_peer_free": 0}
launch_count = 0


def call(name, *args, launches=None):
    """Call an int-returning entry point and raise on a non-zero status.  `launches`: kernels the
    call launches when that depends on its arguments (the C-side sweep loops)."""
    global launch_count
    check(getattr(load(), name)(*args), name)
    launch_count += KERNELS_PER_CALL.get(name, 1) if launches is None else launches
