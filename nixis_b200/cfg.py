"""Shared globals of the export path, same names as the reference's cfg.py (cfg.py:3-14).
When nixis_b200.util is imported in place of the reference's util INSIDE the reference tree, the
reference's own `cfg` module is used instead (so nixis.py and these functions see the same state)."""
WORK_DIR = None
SAVE_DIR = None
SNAP_DIR = None
WORLD_CONFIG = {}
KDT = None
IMG_QUERY_DATA = None
