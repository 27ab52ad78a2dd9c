"""nengo-compatible declaration layer; ``install()`` registers it as ``nengo`` when the
real package is absent."""
import importlib
import sys

from .params import Default
from .core import Network, Ensemble, Node, Connection, Probe, Neurons, LearningRule, ObjView, Config
from .neurons import LIF, LIFRate, RectifiedLinear, Direct, NeuronType
from .synapses import Lowpass, Synapse
from .learning_rules import PES, Voja
from .rc import rc
from . import dists, solvers, processes, exceptions, networks, utils, builder, core, neurons, synapses
from . import learning_rules

# nengo exposes these as sub-modules; the reference imports from them directly
# (``binding.py:5-9``: nengo.connection / nengo.network / nengo.node).
connection = core
network = core
node = core
ensemble = core
probe = core

_ALIASES = {
    "": None, ".dists": dists, ".solvers": solvers, ".processes": processes,
    ".exceptions": exceptions, ".networks": networks,
    ".networks.ensemblearray": networks.ensemblearray, ".utils": utils,
    ".utils.numpy": utils.numpy, ".builder": builder, ".builder.ensemble": builder.ensemble,
    ".connection": core, ".network": core, ".node": core, ".ensemble": core, ".probe": core,
    ".neurons": neurons, ".synapses": synapses, ".learning_rules": learning_rules,
}


def real_nengo_available():
    if "nengo" in sys.modules:
        return getattr(sys.modules["nengo"], "__name__", "") == "nengo" and \
            not getattr(sys.modules["nengo"], "_IS_B200_SHIM", False)
    try:
        return importlib.util.find_spec("nengo") is not None
    except (ImportError, ValueError):
        return False


_IS_B200_SHIM = True


def install(force=False):
    """Make ``import nengo`` resolve to this layer (only if real nengo is missing)."""
    if not force and real_nengo_available():
        return sys.modules.get("nengo") or importlib.import_module("nengo")
    me = sys.modules[__name__]
    for suffix, mod in _ALIASES.items():
        sys.modules["nengo" + suffix] = me if mod is None else mod
    return me
