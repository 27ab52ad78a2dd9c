"""Load the *unmodified* reference ``sspslam`` sources.

``import sspslam`` fails in this environment because its package ``__init__`` pulls in
matplotlib/LaTeX and ``nengo_loihi``.  This helper registers empty package shells whose
``__path__`` points at the reference tree, executes only the numerics modules, and lets
the ``nengo`` imports resolve to the declaration layer in :mod:`nengo_shim`.  No
reference file is edited or copied.  It is used by the parity tests and the golden
vector generator here; nothing on the GPU box needs it.
"""
import importlib
import importlib.util
import os
import sys
import types

from . import nengo_shim

DEFAULT_ROOT = "/root/reference"


def reference_available(root=DEFAULT_ROOT):
    return os.path.isfile(os.path.join(root, "sspslam", "sspspace.py"))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference(root=DEFAULT_ROOT):
    """Return the reference ``sspslam`` package (sspspace + networks + utils.utils)."""
    if "sspslam" in sys.modules and getattr(sys.modules["sspslam"], "_b200_loaded", False):
        return sys.modules["sspslam"]
    if not reference_available(root):
        raise FileNotFoundError(f"reference checkout not found under {root}")
    nengo_shim.install()
    base = os.path.join(root, "sspslam")
    pkg = types.ModuleType("sspslam")
    pkg.__path__ = [base]
    pkg._b200_loaded = True
    sys.modules["sspslam"] = pkg
    upkg = types.ModuleType("sspslam.utils")
    upkg.__path__ = [os.path.join(base, "utils")]
    sys.modules["sspslam.utils"] = upkg
    uu = _load("sspslam.utils.utils", os.path.join(base, "utils", "utils.py"))
    for name in ("Rd_sampling", "sparsity_to_x_intercept", "get_mean_and_ci"):
        setattr(upkg, name, getattr(uu, name))
    pkg.utils = upkg
    ssp = _load("sspslam.sspspace", os.path.join(base, "sspspace.py"))
    for name in ("SPSpace", "SSPSpace", "RandomSSPSpace", "HexagonalSSPSpace"):
        setattr(pkg, name, getattr(ssp, name))
    pkg.sspspace = ssp
    pkg.networks = importlib.import_module("sspslam.networks")
    return pkg
