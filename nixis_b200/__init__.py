"""nixis_b200 -- the terrain hot path of MightyBOBcnc/nixis on B200 (sm_100a).

Modules mirror the reference's module names so nixis.py can import them in place of its own:
    nixis_b200.opensimplex   init, noise2d/3d/4d, noisearr2d/3d/4d
    nixis_b200.terrain       sample_noise, sample_octaves, make_bool_elevation_mask
    nixis_b200.util          create_mesh, rescale, power_rescale, find_percent_val,
                             build_adjacency, sort_adjacency
    nixis_b200.erosion       erode_terrain3, erosion_iteration3, erode_terrain1, erosion_iteration1
    nixis_b200.pipeline      device-resident / multi-GPU driver of the same kernels

All compute runs in libnixis_b200.so (hand-written CUDA, C-ABI in include/nixis_b200.h).
There is no CPU fallback: without the built library or without a GPU, calls raise.
"""
__version__ = "0.1.0"
