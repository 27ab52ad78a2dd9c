"""ORACLE (test infrastructure, never shipped or measured as the product).

Independent CPU restatement of the BUILD half of ``nengo.Simulator.__init__`` for the SSP-SLAM graphs — what the
reference obtains at ``experiments/run_slam.py:198-199`` / ``run_pathint.py:147-148`` before the first step: seed
assignment, evaluation points, encoders, gain / bias, tuning curves, ``LstsqL2`` decoders, transform folding.

It is written against SURVEY.md Appendix A (A.2 seeds, A.3 ensemble build, A.6 connections, A.7 solver) and deliberately
shares NO code with the product's builder (``sspslam_b200/builder.py``, ``compat.py``): different traversal, its own
gain / bias and rate formulas, and the regularised least-squares problem is solved as an augmented ``lstsq`` system
(QR / SVD) rather than by the product's Cholesky factorisation of the normal equations.  ``tests/test_oracle_build.py``
asserts that the two builds agree (seeds exactly, sampled quantities to 1e-12, decoders to 1e-9 of their scale), so a
bug in either shows up; ``RefSimulator`` uses this build when it is not handed a model.

**Parity unpinned** (as for ``nengo_ref_sim.py``): nengo itself is absent from the image; the analytic known-answer
tests for A.2 / A.3 / A.7 are in ``tests/test_oracle_build.py``.  What stays unverifiable without a nengo checkout is
listed in DESIGN.md §2.

The only graph-side code it calls are the sampling methods of the distribution objects attached to the declared
ensembles (``ens.encoders.sample(n, d, rng)``): they are part of the network declaration, not of a builder.
"""
from __future__ import annotations

import numpy as np

MAXINT = np.iinfo(np.int32).max          # nengo.utils.numpy.maxint


def _names(obj):
    return {c.__name__ for c in type(obj).__mro__}


class RefBuiltEnsemble:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class RefBuiltConnection:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class RefModel:
    """``params[obj]`` / ``seeds[obj]`` / ``probe_conns[probe]`` like the product's built model."""

    def __init__(self, network, dt):
        self.toplevel, self.dt = network, float(dt)
        self.seeds, self.params, self.probe_conns = {}, {}, {}

    def initial_voltage(self, ens, trial_seed=None):
        """App. A.2 / A.4: LIF voltages start at ``RandomState(seed + 1).uniform(0, 1)``, everything else at 0.  An integer
        ``trial_seed`` selects the batching extension's counter-based start state (a splitmix64 hash of ensemble seed,
        trial seed and neuron index — the definition, restated, of what ``Simulator(trial_seeds=...)`` documents)."""
        n = ens.n_neurons
        if type(ens.neuron_type).__name__ != "LIF":
            return np.zeros(n)
        if trial_seed is None:
            return np.random.RandomState(self.seeds[ens] + 1).uniform(0.0, 1.0, size=n)
        mask = (1 << 64) - 1
        out = np.empty(n)
        for i in range(n):
            x = (self.seeds[ens] + 1 + int(trial_seed) * 0x9E3779B97F4A7C15 + (i + 1) * 0xD1B54A32D192ED03) & mask
            x ^= x >> 30
            x = (x * 0xBF58476D1CE4E5B9) & mask
            x ^= x >> 27
            x = (x * 0x94D049BB133111EB) & mask
            x ^= x >> 31
            out[i] = (x >> 11) / float(1 << 53)
        return out


# ----------------------------------------------------------------------------- A.2 seeds
def assign_seeds(network, top_seed):
    """One ``randint(maxint)`` per child object from ``RandomState(network seed)``, children visited by TYPE in the order
    Connection, Ensemble, Network, Node, Probe and, inside a type, in creation order; the draw happens even when the
    object carries its own seed (which then wins).  Sub-networks recurse with the seed they were given."""
    seeds = {network: int(top_seed)}
    stack = [network]
    while stack:
        net = stack.pop(0)
        rng = np.random.RandomState(seeds[net])
        for group in (net.connections, net.ensembles, net.networks, net.nodes, net.probes):
            for obj in group:
                drawn = rng.randint(MAXINT)
                own = getattr(obj, "seed", None)
                seeds[obj] = int(drawn if own is None else own)
        stack.extend(net.networks)
    return seeds


# ----------------------------------------------------------------------------- A.3 / A.4 neuron statics
def gain_bias(neuron_type, max_rates, intercepts):
    kind = type(neuron_type).__name__
    max_rates, intercepts = np.asarray(max_rates, float), np.asarray(intercepts, float)
    if kind in ("LIF", "LIFRate"):
        # rate(J) = 1 / (tau_ref + tau_rc ln(1 + 1/(J-1))) = max_rate at J = J_max:  J_max = 1 + 1/(exp((1/max_rate - tau_ref)/tau_rc) - 1)
        j_max = 1.0 + 1.0 / np.expm1((1.0 / max_rates - neuron_type.tau_ref) / neuron_type.tau_rc)
        gain = (j_max - 1.0) / (1.0 - intercepts)          # J(x=1) = J_max and J(x=intercept) = 1
        return gain, 1.0 - gain * intercepts
    if kind == "RectifiedLinear":
        gain = max_rates / (1.0 - intercepts)
        return gain, -intercepts * gain
    raise NotImplementedError(kind)


def rates(neuron_type, J):
    kind = type(neuron_type).__name__
    J = np.asarray(J, float)
    amp = getattr(neuron_type, "amplitude", 1.0)
    if kind in ("LIF", "LIFRate"):
        out = np.zeros_like(J)
        on = J > 1.0
        out[on] = amp / (neuron_type.tau_ref + neuron_type.tau_rc * np.log1p(1.0 / (J[on] - 1.0)))
        return out
    if kind == "RectifiedLinear":
        return amp * np.clip(J, 0.0, None)
    raise NotImplementedError(kind)


def n_eval_points(n_neurons, dims):
    return int(max(min(max(500 * dims, 750), 2500), 2 * n_neurons))


def _is_dist(x):
    return "Distribution" in _names(x)


def build_ensemble(ens, seed):
    rng = np.random.RandomState(seed)
    # draw order: eval points, encoders, max_rates, intercepts
    if _is_dist(ens.eval_points):
        n_pts = ens.n_eval_points if ens.n_eval_points else n_eval_points(ens.n_neurons, ens.dimensions)
        pts = np.asarray(ens.eval_points.sample(n_pts, ens.dimensions, rng=rng), float)
    else:
        pts = np.array(ens.eval_points, float)
    pts = pts * ens.radius
    if _is_dist(ens.encoders):
        enc = np.asarray(ens.encoders.sample(ens.n_neurons, ens.dimensions, rng=rng), float)
    else:
        enc = np.array(ens.encoders, float)
    if ens.normalize_encoders:
        enc = enc / np.sqrt(np.sum(enc * enc, axis=1))[:, None]
    if ens.gain is not None and ens.bias is not None:
        gain, bias, max_rates, intercepts = np.array(ens.gain, float), np.array(ens.bias, float), None, None
    else:
        max_rates = (np.asarray(ens.max_rates.sample(ens.n_neurons, rng=rng), float) if _is_dist(ens.max_rates)
                     else np.array(ens.max_rates, float))
        intercepts = (np.asarray(ens.intercepts.sample(ens.n_neurons, rng=rng), float) if _is_dist(ens.intercepts)
                      else np.array(ens.intercepts, float))
        gain, bias = gain_bias(ens.neuron_type, max_rates, intercepts)
    return RefBuiltEnsemble(eval_points=pts, encoders=enc, intercepts=intercepts, max_rates=max_rates,
                            scaled_encoders=enc * (gain / ens.radius)[:, None], gain=gain, bias=bias)


# ----------------------------------------------------------------------------- A.7 solver
def lstsq_l2(A, Y, reg):
    """``LstsqL2(reg)``: minimise ``|A X - Y|^2 + m sigma^2 |X|^2`` with ``sigma = reg * max(A)``, ``m`` = number of
    evaluation points — solved here as the equivalent augmented least-squares system ``[A; sqrt(m) sigma I] X = [Y; 0]``."""
    m, n = A.shape
    sigma = reg * np.max(A)
    aug_a = np.vstack([A, np.sqrt(m) * sigma * np.eye(n)])
    aug_y = np.vstack([Y, np.zeros((n, Y.shape[1]))])
    X, *_ = np.linalg.lstsq(aug_a, aug_y, rcond=None)
    return X


def _transform(conn):
    t = conn.transform
    if t is None or type(t).__name__ == "NoTransform":
        return None
    if hasattr(t, "init") and not isinstance(t, np.ndarray):
        t = t.init
    return np.asarray(t, float)


def _fold(transform, decoders):
    if transform is None:
        return decoders
    if transform.ndim == 0:
        return float(transform) * decoders
    if transform.ndim == 1:
        return transform[:, None] * decoders
    return transform @ decoders


def build(network, dt=0.001, seed=None):
    """Build every ensemble, connection and probe of ``network`` (App. A.2 - A.7)."""
    model = RefModel(network, dt)
    top = getattr(network, "seed", None)
    if top is None:
        top = seed if seed is not None else np.random.randint(MAXINT)
    model.seeds = assign_seeds(network, top)
    for ens in network.all_ensembles:
        model.params[ens] = build_ensemble(ens, model.seeds[ens])

    acts = {}

    def activities(ens):
        if ens not in acts:
            p = model.params[ens]
            x = (p.eval_points @ p.encoders.T) / ens.radius
            acts[ens] = rates(ens.neuron_type, p.gain[None, :] * x + p.bias[None, :])
        return acts[ens]

    def decoders_for(ens, solver, targets):
        if type(solver).__name__ != "LstsqL2" or getattr(solver, "weights", False):
            raise NotImplementedError("only LstsqL2(weights=False) is on the hot path")
        return lstsq_l2(activities(ens), targets, float(solver.reg)).T                 # (size_mid, n_neurons)

    for conn in network.all_connections:
        pre = conn.pre_obj
        tr = _transform(conn)
        if "Ensemble" in _names(pre):
            pts = model.params[pre].eval_points
            sel = pts if conn.pre_slice == slice(None) else pts[:, conn.pre_slice]
            if sel.ndim == 1:
                sel = sel[:, None]
            if conn.function is None:
                targets = sel
            elif isinstance(conn.function, np.ndarray):
                targets = conn.function
            else:
                targets = np.stack([np.asarray(conn.function(p), float).reshape(-1) for p in sel])
            dec = decoders_for(pre, conn.solver, targets)
            model.params[conn] = RefBuiltConnection(eval_points=pts, transform=tr, weights=_fold(tr, dec), decoders=dec)
        elif "Neurons" in _names(pre):
            raise NotImplementedError("connections from ens.neurons are outside the hot path")
        else:
            if conn.function is not None:
                raise NotImplementedError("functions on Node -> X connections are outside the hot path")
            model.params[conn] = RefBuiltConnection(eval_points=None, transform=tr, weights=tr, decoders=None)
    for probe in network.all_probes:
        obj = probe.obj
        if "Ensemble" in _names(obj) and probe.attr == "decoded_output":
            dec = decoders_for(obj, probe.solver, model.params[obj].eval_points)
            model.probe_conns[probe] = dec[np.arange(obj.dimensions)[probe.slice]]
        model.params[probe] = None
    return model
