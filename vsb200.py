"""Import shim: the package directory is `video-stitcher_b200/` (not an importable name), so load its modules by path."""
import importlib.util
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "video-stitcher_b200")


def _load(name):
    full = "vsb200_" + name
    if full in sys.modules:
        return sys.modules[full]
    spec = importlib.util.spec_from_file_location(full, os.path.join(_PKG, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    return mod


binding = _load("binding")
synth = _load("synth")
dist = _load("dist")
