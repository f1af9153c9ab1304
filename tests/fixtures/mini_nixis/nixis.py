"""A stand-in for the reference's CLI script, for boxes that do not hold a checkout of it.

tools/run_nixis.py swaps the hot path of the reference's `nixis.py` onto nixis_b200 by pre-seeding
sys.modules.  /root/reference does not exist on the GPU box, so this file gives the runner a script to
drive there: it imports the hot-path functions THE WAY nixis.py does (module names `opensimplex`, `util`,
`terrain`, `erosion`, `cfg`; nixis.py:12-18) and replays the call order of nixis.py:247-417 -- mesh,
optional image query, seed tables, fBm, the height-assembly chain, adjacency, erosion, image export --
with the reference's constants (nixis.py:312-320, 410).  It is written for this test, not copied: no
prompts, no viewer, no timing chatter; results are saved to --out for the test to check.
"""
import argparse
import os
import sys

import numpy as np

import cfg
import opensimplex as osi
from util import *                                         # noqa: F401,F403  (nixis.py:14)
from terrain import sample_octaves, make_bool_elevation_mask
from erosion import *                                      # noqa: F401,F403  (nixis.py:17)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-d", "--divisions", type=int, default=320)
    ap.add_argument("-s", "--seed", type=int, default=0)
    ap.add_argument("-r", "--radius", type=float, default=1.0)
    ap.add_argument("--novis", action="store_true")
    ap.add_argument("--save_img", action="store_true")
    ap.add_argument("--img_width", type=int, default=256)
    ap.add_argument("--erode_iters", type=int, default=11)      # nixis.py:410 num_iter=11
    ap.add_argument("--snapshot", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()

    do_erode = False                                            # nixis.py:131 (tools/run_nixis.py --erode flips it)
    divisions = args.divisions + (args.divisions % 2)           # nixis.py:142-146: odd -> even
    world_radius = args.radius
    cfg.SAVE_DIR = cfg.SNAP_DIR = os.path.dirname(os.path.abspath(args.out)) if args.out else os.getcwd()

    points, cells = create_mesh(divisions)                      # nixis.py:247
    if world_radius != 1:
        points *= world_radius                                  # nixis.py:249
    export_list = {}
    if args.save_img or args.snapshot:                          # nixis.py:270-283
        width, height_px = args.img_width, args.img_width // 2
        ll = make_ll_arr(width, height_px, world_radius)
        build_KDTree(points)
        cfg.IMG_QUERY_DATA = cfg.KDT.query(ll, k=3)

    perm, pgi = osi.init(args.seed)                             # nixis.py:308
    min_alt, max_alt, ocean_percent = -4000, 8850, 55.0         # nixis.py:312-320
    height = sample_octaves(points, None, perm, pgi, 7, 1.5, 0.4, 2.5, 0.5, world_radius)
    height = rescale(height, min_alt, max_alt)
    minval, maxval = np.amin(height), np.amax(height)
    ocean_level = find_percent_val(minval, maxval, ocean_percent)
    ocean = make_bool_elevation_mask(height, ocean_level)
    if args.save_img:
        export_list["ocean"] = [ocean, 'gray']
    height = power_rescale(height, mask=ocean, mode=1, power=0.5)
    height = power_rescale(height, mask=ocean, mode=0, power=2.0)
    height -= ocean_level
    height = rescale(height, min_alt, max_alt, mid=0)
    assembled = height.copy()
    if args.save_img and not do_erode:                          # nixis.py:382-389
        export_list["height_absolute"] = [(rescale(height, -4000, 8850) + (32768 - find_percent_val(-4000, 8850, ocean_percent))).astype('uint16'), 'gray']
        export_list["height_relative"] = [(rescale(height, 0, 65535)).astype('uint16'), 'gray']
    neighbors = None
    if do_erode:                                                # nixis.py:397-417
        neighbors = build_adjacency(cells)
        sort_adjacency(neighbors)
        erode_terrain3(points, neighbors, height, num_iter=args.erode_iters, snapshot=args.snapshot)
        if args.save_img:
            export_list["height"] = [height, 'gray']
    images = {}
    if args.save_img:                                           # nixis.py:516-540
        images = build_image_data(export_list)
        save_image(images, cfg.SAVE_DIR, "mini")
    if args.out:
        np.savez(args.out, points=points, cells=cells, perm=perm, pgi=pgi, assembled=assembled, height=height,
                 ocean=ocean, ocean_level=ocean_level, eroded=bool(do_erode),
                 neighbors=neighbors if neighbors is not None else np.zeros((0, 6), np.int32),
                 **{f"img_{k}": v for k, v in images.items()})
    print("mini_nixis finished: V =", len(points), "eroded =", bool(do_erode))


main()
