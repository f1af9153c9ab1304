"""Shared state of the export path.  The attribute NAMES are the reference's (cfg.py:3-14) because
nixis.py and the export helpers address them as `cfg.<NAME>`; inside the reference tree
nixis_b200.util uses the reference's own `cfg` module instead, so both sides see one state."""

_STATE = {
    "WORK_DIR": None,           # set by the CLI
    "SAVE_DIR": None,
    "SNAP_DIR": None,
    "WORLD_CONFIG": {},         # settings that recreate a planet
    "KDT": None,                # nearest-vertex searcher (nixis_b200.util.IcoNearest)
    "IMG_QUERY_DATA": None,     # (distances, vertex ids) of the 3 nearest vertices of every pixel
}
globals().update(_STATE)
