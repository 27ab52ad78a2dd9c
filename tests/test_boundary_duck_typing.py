"""The boundary accepts object graphs that are NOT instances of this repo's declaration classes (a
network built with the real ``nengo`` package).  Real nengo cannot be installed here, so the look-alike is a SECOND,
independently imported copy of the declaration layer (every class is a distinct type object with nengo's class names),
with nengo >= 3's transform objects (``Dense`` / ``NoTransform`` carrying ``.init``) substituted for the plain arrays.
``build_model`` + ``lower`` must produce the same built parameters and the same device plan, bit for bit."""
import importlib.util
import os
import sys

import numpy as np
import pytest

from sspslam_b200 import builder, compat, lowering, scenarios
from sspslam_b200 import nengo_shim as ns

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def clone():
    """A second copy of the whole package under another name: ``clone.nengo_shim.Ensemble is not ns.Ensemble``."""
    real = os.path.join(ROOT, "semantic-spiking-neural-slam-2023_b200")
    spec = importlib.util.spec_from_file_location("sspslam_lookalike", os.path.join(real, "__init__.py"),
                                                  submodule_search_locations=[real])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["sspslam_lookalike"] = mod
    spec.loader.exec_module(mod)
    yield mod
    for name in [n for n in sys.modules if n.startswith("sspslam_lookalike")]:
        del sys.modules[name]


class Dense:                       # stands in for nengo.transforms.Dense: the array lives in ``.init``
    def __init__(self, init):
        self.init = init


class NoTransform:                 # stands in for nengo.transforms.NoTransform
    pass


def _wrap_transforms(net):
    for conn in net.all_connections:
        conn.transform = NoTransform() if conn.transform is None else Dense(conn.transform)


KW = dict(n_trials=1, n_steps=40, ssp_dim=19, pi_n_neurons=30, mem_n_neurons=64, circonv_n_neurons=16, n_landmarks=6,
          T=20.0, view_rad=0.6)


@pytest.mark.parametrize("neuron_type,view", [("lif", False), ("lifrate", True)])
def test_lookalike_graph_builds_and_lowers_to_the_same_plan(clone, neuron_type, view):
    sc_a = scenarios.make_slam(neuron_type=neuron_type, view=view, **KW)
    sc_b = clone.scenarios.make_slam(neuron_type=neuron_type, view=view, **KW)
    net_b = sc_b.network
    assert not isinstance(net_b, ns.Network) and not isinstance(net_b.all_ensembles[0], ns.Ensemble)
    assert not isinstance(net_b.all_ensembles[0].neuron_type, (ns.LIF, ns.LIFRate))
    _wrap_transforms(net_b)
    ma, mb = builder.build_model(sc_a.network, dt=0.001), builder.build_model(net_b, dt=0.001)
    for ea, eb in zip(sc_a.network.all_ensembles, net_b.all_ensembles):
        assert np.array_equal(ma.params[ea].scaled_encoders, mb.params[eb].scaled_encoders)
        assert np.array_equal(ma.params[ea].bias, mb.params[eb].bias)
    pa, pb = lowering.lower(sc_a.network, ma, chunk_cap=40), lowering.lower(net_b, mb, chunk_cap=40)
    assert pa.scalars == pb.scalars and set(pa.arrays) == set(pb.arrays)
    for name in pa.arrays:
        assert np.array_equal(pa.arrays[name], pb.arrays[name]), name
    assert [p.kind for p in pa.probes] == [p.kind for p in pb.probes]
    assert pa.stats["n_learned"] > 0 and len(pa.arrays["cleanup"]) == 1 and len(pa.arrays["gate"]) == 1


def test_lookalike_pathint_with_relu(clone):
    a = scenarios.make_pathint(n_trials=1, n_steps=30, ssp_dim=19, pi_n_neurons=40, neuron_type="relu")
    b = clone.scenarios.make_pathint(n_trials=1, n_steps=30, ssp_dim=19, pi_n_neurons=40, neuron_type="relu")
    _wrap_transforms(b.network)
    pa = lowering.lower(a.network, builder.build_model(a.network), chunk_cap=30)
    pb = lowering.lower(b.network, builder.build_model(b.network), chunk_cap=30)
    for name in pa.arrays:
        assert np.array_equal(pa.arrays[name], pb.arrays[name]), name


def test_compat_classification_is_by_name_and_refuses_lookalike_subclasses():
    class SpikingRectifiedLinear(ns.RectifiedLinear):
        pass

    class Alpha:
        tau = 0.01

    class Lowpass:
        tau = 0.02

    assert compat.neuron_kind(ns.LIF()) == "lif" and compat.neuron_kind(ns.LIFRate()) == "lifrate"
    assert compat.neuron_kind(SpikingRectifiedLinear()) is None          # steps differently: not its base class
    assert compat.synapse_tau(Lowpass()) == 0.02 and compat.synapse_tau(None) is None
    with pytest.raises(NotImplementedError):
        compat.synapse_tau(Alpha())

    class Conn:
        def __init__(self, t):
            self.transform = t
    assert compat.transform_of(Conn(NoTransform())) is None and compat.transform_of(Conn(None)) is None
    assert compat.transform_of(Conn(Dense(2.5))) == 2.5
    assert np.array_equal(compat.transform_of(Conn(Dense(np.eye(3)))), np.eye(3))
    assert np.array_equal(compat.transform_of(Conn([[1, 2]])), [[1.0, 2.0]])


def test_unsupported_neuron_type_is_refused_loudly():
    class AdaptiveLIF(ns.LIF):
        pass
    with ns.Network(seed=1) as net:
        net.config[ns.Ensemble].neuron_type = AdaptiveLIF()
        a = ns.Node(lambda t: [0.5])
        e = ns.Ensemble(20, 1)
        ns.Connection(a, e, synapse=None)
        ns.Probe(e, synapse=0.01)
    with pytest.raises(NotImplementedError):
        lowering.lower(net, builder.build_model(net))
