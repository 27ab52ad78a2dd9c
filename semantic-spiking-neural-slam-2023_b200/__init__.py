"""B200-native simulator backend for the SSP-SLAM per-timestep hot path.

Public surface (mirrors what the reference drivers touch):

* ``Simulator(network, dt)`` — drop-in for ``nengo.Simulator`` on ``PathIntegration`` /
  ``SLAMNetwork`` / ``SLAMViewNetwork`` graphs, batched over trials, CUDA only.
* ``SPSpace`` / ``SSPSpace`` / ``HexagonalSSPSpace`` / ``RandomSSPSpace`` with
  device encode / grid-decode.
* ``networks`` — from-scratch declarations of the reference topologies (used where the
  reference checkout is not present, e.g. on the GPU box).
* ``results`` — the drivers' post-processing and ``.npz`` schema (``run_slam.py:236-293``, ``run_pathint.py:168-203``).
* ``nengo_shim`` — nengo-compatible declaration layer (``install()`` registers it as
  ``nengo`` when the real package is missing).
"""
from . import nengo_shim  # noqa: F401

__version__ = "0.1.0"

_LAZY = {
    "Simulator": ("simulator", "Simulator"),
    "SPSpace": ("sspspace", "SPSpace"),
    "SSPSpace": ("sspspace", "SSPSpace"),
    "HexagonalSSPSpace": ("sspspace", "HexagonalSSPSpace"),
    "RandomSSPSpace": ("sspspace", "RandomSSPSpace"),
    "build_model": ("builder", "build_model"),
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod, attr = _LAZY[name]
        return getattr(importlib.import_module(f"{__name__}.{mod}"), attr)
    if name in ("networks", "sspspace", "builder", "lowering", "simulator", "inputs", "cabi", "refload", "results",
                "scenarios", "sharding"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
