"""Run the UNMODIFIED reference CLI (nixis.py) with its hot path swapped onto nixis_b200.

    python tools/run_nixis.py [--ref /path/to/nixis] [--erode] [--climate] -- -d 320 -s 12345 --novis --save_img

The reference imports its hot-path functions by module name (nixis.py:12-18), and its own directory
comes first on sys.path, so the swap is done by pre-seeding sys.modules (SURVEY 8b "swap-in
mechanism"):
    opensimplex, terrain, erosion, climate -> the nixis_b200 modules of the same name
    util    -> the reference's util with every function nixis_b200.util provides laid over it
    cfg     -> ONE shared module (nixis.py and the export helpers must see the same KDT / query data)
    gui     -> a stub `visualize` (pyvista is a viewer, not part of the path)
    meshzoo, meshio -> stubs, only so that the reference's util imports (they are never called)
`do_erode` / `do_climate` are local literals in nixis.py (:131-132); --erode / --climate flip them in
an in-memory copy of the source -- nothing in the reference tree is written.
"""
import argparse
import importlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def install(ref):
    sys.path.insert(0, ROOT)
    import nixis_b200.opensimplex, nixis_b200.terrain, nixis_b200.erosion, nixis_b200.climate, nixis_b200.util
    import nixis_b200.cfg as shared_cfg
    for name in ("meshzoo", "meshio"):
        sys.modules.setdefault(name, types.ModuleType(name))
    gui = types.ModuleType("gui")
    gui.visualize = lambda *a, **k: print("(viewer disabled by tools/run_nixis.py)")
    sys.modules["gui"] = gui
    sys.modules["cfg"] = shared_cfg
    nixis_b200.util.cfg = shared_cfg
    # the reference's util as the base layer (file I/O helpers, pretty printers), ours on top
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nixis_numba_cache")
    sys.path.insert(0, ref)
    spec = importlib.util.spec_from_file_location("util", os.path.join(ref, "util.py"))
    util = importlib.util.module_from_spec(spec)
    sys.modules["util"] = util
    try:
        spec.loader.exec_module(util)
    except Exception as exc:                      # numba / scipy / PIL missing: our functions alone
        print(f"(reference util not importable: {exc}; using nixis_b200.util only)")
        util = types.ModuleType("util")
        sys.modules["util"] = util
    swapped = []
    for name in dir(nixis_b200.util):
        obj = getattr(nixis_b200.util, name)
        if not name.startswith("_") and callable(obj) and getattr(obj, "__module__", "") == "nixis_b200.util":
            setattr(util, name, obj)
            swapped.append(name)
    util.cfg = shared_cfg
    sys.modules["opensimplex"] = nixis_b200.opensimplex
    sys.modules["terrain"] = nixis_b200.terrain
    sys.modules["erosion"] = nixis_b200.erosion
    sys.modules["climate"] = nixis_b200.climate
    return swapped


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--ref", default=os.environ.get("NIXIS_REF", "/root/reference"))
    ap.add_argument("--erode", action="store_true", help="set do_erode = True (nixis.py:131)")
    ap.add_argument("--climate", action="store_true", help="set do_climate = True (nixis.py:132)")
    ap.add_argument("rest", nargs=argparse.REMAINDER, help="arguments for nixis.py (after --)")
    args = ap.parse_args()
    ref = os.path.abspath(args.ref)
    src_path = os.path.join(ref, "nixis.py")
    if not os.path.exists(src_path):
        raise SystemExit(f"{src_path} not found: point --ref at a checkout of MightyBOBcnc/nixis")
    swapped = install(ref)
    print(f"hot path swapped onto nixis_b200: util.{{{', '.join(sorted(swapped))}}} + opensimplex, terrain, erosion, climate")
    src = open(src_path).read()
    for flag, on in (("do_erode", args.erode), ("do_climate", args.climate)):
        if on:
            assert f"{flag} = False" in src
            src = src.replace(f"{flag} = False", f"{flag} = True", 1)
    rest = [a for a in args.rest if a != "--"]
    sys.argv = [src_path] + rest
    os.chdir(ref)                                  # options.json is opened cwd-relative (util.py:381)
    glb = {"__name__": "__main__", "__file__": src_path}
    exec(compile(src, src_path, "exec"), glb)


if __name__ == "__main__":
    main()
