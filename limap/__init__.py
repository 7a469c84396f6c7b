"""Import alias so that code written against the reference package (`import limap.triangulation`,
`limap.optimize`, `limap.base`, `limap.vplib`) resolves to the B200 engine's operator surface.
Only the hot-path packages exist; everything else of cvg/limap is out of scope."""
import importlib
import sys

for _name in ("base", "triangulation", "optimize", "vplib", "merging", "util", "util.io", "runners", "visualize", "pointsfm"):
    try:
        _m = importlib.import_module(f"limap_b200.{_name}")
    except ModuleNotFoundError:
        continue
    sys.modules[f"limap.{_name}"] = _m
    if "." not in _name:
        globals()[_name] = _m
