"""TEST INFRASTRUCTURE ONLY.  Loads the reference's pybind modules built by oracle/build_ref.sh
(_raymarching, _gridencoder, _shencoder: the reference's L0 FFI) so GPU tests can call
the reference kernels directly.  /root/reference is NOT needed at run time (the .so files travel)."""
import glob
import importlib.util
import os
from types import SimpleNamespace

_HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name):
    hits = glob.glob(os.path.join(_HERE, "_ref", name + "*.so"))
    if not hits:
        return None
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location(name, hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load():
    mods = {n: _load(n) for n in ("_raymarching", "_gridencoder", "_shencoder")}
    if any(v is None for v in mods.values()):
        return None
    return SimpleNamespace(raymarching=mods["_raymarching"], gridencoder=mods["_gridencoder"], shencoder=mods["_shencoder"])
