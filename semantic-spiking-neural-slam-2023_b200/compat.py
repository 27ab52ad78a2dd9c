"""Duck-typed view of a nengo-style object graph.

The drop-in boundary is ``Simulator(network)`` where ``network`` was built either with this repo's declaration layer
(``nengo_shim``) or with the real ``nengo`` package on a machine that has it (``experiments/run_slam.py:151-199``).  The
builder and the lowering therefore never ask ``isinstance(x, nengo_shim.Something)``: they classify objects by the class
NAMES in their MRO and read the documented nengo attributes (same names in both packages), with the two places where
nengo >= 3 differs from a plain array handled here:

* ``conn.transform`` is a ``nengo.transforms.Dense`` / ``NoTransform`` object (``.init`` holds the array / scalar);
* solvers / distributions / synapses are recognised by name and used through their public attributes only
  (``solver.reg``, ``solver.weights``, ``dist.sample(n, d, rng=)``, ``synapse.tau``).

Anything that is not one of the constructs of the SSP-SLAM graphs is refused loudly by the callers.
"""
from __future__ import annotations

import numpy as np

_OBJECT_KINDS = ("Ensemble", "Neurons", "Node", "Connection", "Probe", "LearningRule", "Network")


def _mro_names(obj):
    return [c.__name__ for c in type(obj).__mro__]


def kind(obj):
    """'ensemble' | 'neurons' | 'node' | 'connection' | 'probe' | 'learning_rule' | 'network' | None."""
    names = _mro_names(obj)
    for k in _OBJECT_KINDS:
        if k in names:
            return {"LearningRule": "learning_rule"}.get(k, k.lower())
    return None


def is_ensemble(obj):
    return kind(obj) == "ensemble"


def is_neurons(obj):
    return kind(obj) == "neurons"


def is_node(obj):
    return kind(obj) == "node"


def is_connection(obj):
    return kind(obj) == "connection"


def is_learning_rule(obj):
    return kind(obj) == "learning_rule"


def neuron_kind(nt):
    """'lif' | 'lifrate' | 'relu' | 'direct' | None — by the EXACT class name (a subclass such as
    ``SpikingRectifiedLinear`` or ``AdaptiveLIF`` steps differently and must not be mistaken for its base)."""
    return {"LIF": "lif", "LIFRate": "lifrate", "RectifiedLinear": "relu", "Direct": "direct"}.get(type(nt).__name__)


def rule_kind(lrt):
    """'pes' | 'voja' | None — exact class name."""
    return {"PES": "pes", "Voja": "voja"}.get(type(lrt).__name__)


def is_distribution(x):
    return "Distribution" in _mro_names(x) and hasattr(x, "sample")


def get_samples(dist_or_samples, n, d=None, rng=None):
    """``nengo.dists.get_samples``: sample a distribution, pass an array through."""
    if is_distribution(dist_or_samples):
        return np.asarray(dist_or_samples.sample(n, d, rng=rng) if d is not None else dist_or_samples.sample(n, rng=rng),
                          dtype=np.float64)
    return np.array(dist_or_samples, dtype=np.float64)


def is_lstsq_l2(solver):
    return type(solver).__name__ == "LstsqL2" and hasattr(solver, "reg")


def synapse_tau(syn):
    """Time constant of a ``Lowpass`` synapse; other synapse models are outside the hot path."""
    if syn is None:
        return None
    if type(syn).__name__ != "Lowpass" or not hasattr(syn, "tau"):
        raise NotImplementedError(f"synapse {syn!r}: only Lowpass synapses are on the hot path")
    return float(syn.tau)


def transform_of(conn):
    """The connection's transform as ``None`` / scalar / 1-D / 2-D float array.  nengo >= 3 wraps it in a
    ``nengo.transforms`` object (``NoTransform``; ``Dense`` with ``.init``); older versions and the shim give the array."""
    t = conn.transform
    if t is None:
        return None
    name = type(t).__name__
    if name == "NoTransform":
        return None
    if hasattr(t, "init") and not isinstance(t, np.ndarray):
        if name not in ("Dense",):
            raise NotImplementedError(f"transform {name} is outside the hot path (only dense / scalar transforms)")
        init = t.init
        if is_distribution(init):
            raise NotImplementedError("randomly initialised transforms are outside the hot path")
        t = init
    return np.asarray(t, dtype=np.float64)


def lstsq_l2_factor(A, reg):
    """``LstsqL2`` with nengo's Cholesky sub-solver: factor of ``A^T A + m sigma^2 I`` (or of
    ``A A^T + ...`` when there are fewer evaluation points than neurons), ``sigma = reg * max(A)``."""
    import scipy.linalg
    m, n = A.shape
    sigma = reg * A.max()
    transpose = m < n
    G = A @ A.T if transpose else A.T @ A
    G[np.diag_indices_from(G)] += m * sigma ** 2
    return scipy.linalg.cho_factor(G, overwrite_a=True), transpose


def lstsq_l2_solve(A, Y, factor):
    import scipy.linalg
    chol, transpose = factor
    b = Y if transpose else A.T @ Y
    x = scipy.linalg.cho_solve(chol, b)
    return A.T @ x if transpose else x
