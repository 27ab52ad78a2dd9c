"""Host-side *builder*: nengo-style network declaration -> built parameters.

This is the part of ``nengo.Simulator.__init__`` that the reference relies on at
``experiments/run_slam.py:198-199`` / ``run_pathint.py:147-148`` (SURVEY.md §3.4 and
App. A.2-A.8): seed assignment, eval-point / encoder / gain / bias sampling, tuning
curves, regularised least-squares decoders, ``transform @ decoders`` folding.  It is
pure NumPy/SciPy host logic and is shared by the CUDA lowering (product) and the
operator-level oracle (checker) so that both step *the same built model*.
"""
from __future__ import annotations

import dataclasses
import os

import numpy as np

from . import compat
from .nengo_shim.exceptions import BuildError
from .nengo_shim.utils.numpy import maxint

_TYPE_ORDER = ("connections", "ensembles", "networks", "nodes", "probes")  # App. A.2


@dataclasses.dataclass
class BuiltEnsemble:
    eval_points: np.ndarray
    encoders: np.ndarray
    intercepts: np.ndarray
    max_rates: np.ndarray
    scaled_encoders: np.ndarray
    gain: np.ndarray
    bias: np.ndarray


@dataclasses.dataclass
class BuiltConnection:
    eval_points: np.ndarray | None
    solver_info: dict | None
    transform: np.ndarray | None
    weights: np.ndarray | float | None  # folded ``transform @ decoders`` (size_out x n) for decoded conns
    decoders: np.ndarray | None = None  # unfolded decoders (size_mid x n)


class BuiltModel:
    """``sim.model``-like container: ``params[obj]``, ``seeds[obj]``, ``dt``."""

    def __init__(self, network, dt):
        self.toplevel = network
        self.dt = float(dt)
        self.seeds: dict = {}
        self.params: dict = {}
        self.probe_conns: dict = {}  # probe -> implicit decoded connection weights

    def initial_voltage(self, ens, trial_seed=None):
        """LIF start voltages ~ U(0,1) (App. A.2/A.4).

        ``trial_seed=None`` is nengo's own draw from ``RandomState(seed+1)``.  An integer
        ``trial_seed`` is this backend's batching extension: a counter-based hash of
        (ensemble seed, trial seed, neuron index), so a trial's start state does not depend
        on which other trials share the batch (and thousands of trials cost no RNG set-up)."""
        if compat.neuron_kind(ens.neuron_type) != "lif":
            return np.zeros(ens.n_neurons)
        if trial_seed is None:
            return np.random.RandomState(self.seeds[ens] + 1).uniform(0.0, 1.0, size=ens.n_neurons)
        return self.initial_voltages(ens, [trial_seed])[0]

    def initial_voltages(self, ens, trial_seeds):
        """[n_trials, n_neurons] start voltages; entries of ``trial_seeds`` may be ``None``."""
        n = ens.n_neurons
        out = np.zeros((len(trial_seeds), n))
        if compat.neuron_kind(ens.neuron_type) != "lif":
            return out
        ints = np.array([-1 if s is None else int(s) for s in trial_seeds], dtype=np.int64)
        hashed = ints >= 0
        if hashed.any():
            with np.errstate(over="ignore"):
                x = (np.uint64(self.seeds[ens] + 1)
                     + ints[hashed].astype(np.uint64)[:, None] * np.uint64(0x9E3779B97F4A7C15)
                     + (np.arange(n, dtype=np.uint64)[None, :] + np.uint64(1)) * np.uint64(0xD1B54A32D192ED03))
                x ^= x >> np.uint64(30)
                x *= np.uint64(0xBF58476D1CE4E5B9)
                x ^= x >> np.uint64(27)
                x *= np.uint64(0x94D049BB133111EB)
                x ^= x >> np.uint64(31)
            out[hashed] = (x >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
        if (~hashed).any():
            out[~hashed] = self.initial_voltage(ens, None)
        return out


# ----------------------------------------------------------------------------- seeds
def _assign_seeds(net, seeds):
    rng = np.random.RandomState(seeds[net])
    for attr in _TYPE_ORDER:
        for obj in getattr(net, attr):
            drawn = rng.randint(maxint)  # drawn even when the object has its own seed
            own = getattr(obj, "seed", None)
            seeds[obj] = drawn if own is None else own
    for sub in net.networks:
        _assign_seeds(sub, seeds)


def n_eval_points_default(n_neurons, dimensions):
    return int(max(np.clip(500 * dimensions, 750, 2500), 2 * n_neurons))


# ----------------------------------------------------------------------------- ensembles
def _build_ensemble(model, ens):
    rng = np.random.RandomState(model.seeds[ens])
    # draw order: eval points -> encoders -> max_rates -> intercepts (App. A.3)
    if compat.is_distribution(ens.eval_points):
        n_pts = ens.n_eval_points or n_eval_points_default(ens.n_neurons, ens.dimensions)
        eval_points = ens.eval_points.sample(n_pts, ens.dimensions, rng=rng)
    else:
        eval_points = np.array(ens.eval_points, dtype=np.float64)
    eval_points = eval_points * ens.radius

    if compat.is_distribution(ens.encoders):
        encoders = np.asarray(ens.encoders.sample(ens.n_neurons, ens.dimensions, rng=rng), dtype=np.float64)
    else:
        encoders = np.array(ens.encoders, dtype=np.float64)
    if ens.normalize_encoders:
        encoders = encoders / np.linalg.norm(encoders, axis=1, keepdims=True)
    if not np.all(np.isfinite(encoders)):
        raise BuildError(f"non-finite encoders in {ens!r}")

    if ens.gain is not None and ens.bias is not None:
        gain = np.array(ens.gain, dtype=np.float64)
        bias = np.array(ens.bias, dtype=np.float64)
        max_rates = intercepts = None
    else:
        max_rates = compat.get_samples(ens.max_rates, ens.n_neurons, rng=rng)
        intercepts = compat.get_samples(ens.intercepts, ens.n_neurons, rng=rng)
        gain, bias = ens.neuron_type.gain_bias(max_rates, intercepts)
    if not (np.all(np.isfinite(gain)) and np.all(np.isfinite(bias))):
        raise BuildError(f"non-finite gain/bias in {ens!r}")

    scaled = encoders * (gain / ens.radius)[:, None]
    model.params[ens] = BuiltEnsemble(eval_points, encoders, intercepts, max_rates, scaled, gain, bias)


# ----------------------------------------------------------------------------- decoders
class _DecoderCache:
    """Tuning curves and the regularised Gram factor are shared by every decoded
    connection leaving one ensemble with the same solver (509 solves)."""

    def __init__(self, model):
        self.model = model
        self._acts = {}
        self._factor = {}
        self._presolved = {}

    def activities(self, ens):
        if ens not in self._acts:
            p = self.model.params[ens]
            x = p.eval_points @ (p.encoders.T / ens.radius)
            self._acts[ens] = ens.neuron_type.rates(x, p.gain, p.bias)
            if np.count_nonzero(self._acts[ens]) == 0:
                raise BuildError(f"all tuning curves of {ens!r} are zero")
        return self._acts[ens]

    def presolve_on_device(self, network, device=0):
        """every LstsqL2 system of the model on the GPU (``libssb_builder.so``), batched over the ensembles
        of one shape; ``solve`` then only looks the decoders up.  Same arithmetic as the host path (float64 normal
        equations + Cholesky), summation order aside."""
        from . import cabi
        jobs = {}                                     # (ens, reg) -> [(key object, targets)]
        for conn in network.all_connections:
            pre = conn.pre_obj
            if compat.is_ensemble(pre) and compat.is_lstsq_l2(conn.solver) and not conn.solver.weights \
                    and conn.eval_points is None and compat.neuron_kind(pre.neuron_type) != "direct":
                jobs.setdefault((pre, float(conn.solver.reg)), []).append(
                    (conn, _targets(conn, self.model.params[pre].eval_points)))
        for probe in network.all_probes:
            obj = probe.obj
            if compat.is_ensemble(obj) and probe.attr == "decoded_output" and compat.is_lstsq_l2(probe.solver):
                jobs.setdefault((obj, float(probe.solver.reg)), []).append((probe, self.model.params[obj].eval_points))
        groups = {}                                   # (m, n, reg) -> [(ens, [(key, targets)])]
        for (ens, reg), lst in jobs.items():
            A = self.activities(ens)
            if A.shape[0] >= A.shape[1]:              # (fewer evaluation points than neurons: host path)
                groups.setdefault((A.shape[0], A.shape[1], reg), []).append((ens, lst))
        for (m, n, reg), members in groups.items():
            kmax = max(sum(t.shape[1] for _, t in lst) for _, lst in members)
            A = np.stack([self.activities(ens) for ens, _ in members])
            Y = np.zeros((len(members), m, kmax))
            for s, (_, lst) in enumerate(members):
                col = 0
                for _, t in lst:
                    Y[s, :, col:col + t.shape[1]] = t
                    col += t.shape[1]
            X = cabi.solve_decoders(A, Y, reg, device)
            for s, (ens, lst) in enumerate(members):
                col = 0
                for key, t in lst:
                    self._presolved[(ens, reg, key)] = X[s, :, col:col + t.shape[1]].copy()
                    col += t.shape[1]

    def solve(self, ens, solver, targets, key=None):
        if compat.is_lstsq_l2(solver) and (ens, float(solver.reg), key) in self._presolved:
            return self._presolved[(ens, float(solver.reg), key)]
        A = self.activities(ens)
        if not compat.is_lstsq_l2(solver):
            X, _ = solver(A, targets)
            return X
        key = (ens, float(solver.reg))
        if key not in self._factor:
            self._factor[key] = compat.lstsq_l2_factor(A, float(solver.reg))
        return compat.lstsq_l2_solve(A, targets, self._factor[key])


def _targets(conn, eval_points):
    pts = eval_points[:, conn.pre_slice] if conn.pre_slice != slice(None) else eval_points
    if pts.ndim == 1:
        pts = pts[:, None]
    if conn.function is None:
        return pts
    if isinstance(conn.function, np.ndarray):
        return conn.function
    out = np.zeros((len(pts), conn.size_mid))
    for i, p in enumerate(pts):
        out[i] = np.asarray(conn.function(p), dtype=np.float64).reshape(-1)
    return out


def fold_transform(transform, mat):
    """``transform * decoders`` with nengo's scalar / diagonal / dense cases (App. A.6)."""
    if transform is None:
        return mat
    t = np.asarray(transform, dtype=np.float64)
    if t.ndim == 0:
        return t * mat
    if t.ndim == 1:
        return t[:, None] * mat
    return t @ mat


def _build_connection(model, conn, cache):
    pre = conn.pre_obj
    transform = compat.transform_of(conn)
    if compat.is_ensemble(pre):
        if conn.solver.weights:
            raise NotImplementedError("weight solvers (solver.weights=True) are outside the hot path")
        if compat.neuron_kind(pre.neuron_type) == "direct":
            raise NotImplementedError("Direct-mode ensembles are not supported")
        eval_points = model.params[pre].eval_points if conn.eval_points is None \
            else np.array(conn.eval_points, dtype=np.float64)
        if conn.eval_points is not None:
            raise NotImplementedError("per-connection eval_points are outside the hot path")
        decoders = cache.solve(pre, conn.solver, _targets(conn, eval_points), key=conn).T  # size_mid x n
        weights = fold_transform(transform, decoders)
        model.params[conn] = BuiltConnection(eval_points, {}, transform, weights, decoders)
    elif compat.is_neurons(pre):
        raise NotImplementedError("connections *from* ens.neurons are outside the hot path")
    else:
        if conn.function is not None:
            raise NotImplementedError("functions on Node->X connections are outside the hot path")
        model.params[conn] = BuiltConnection(None, None, transform, transform)


def _build_probe(model, probe, cache):
    obj = probe.obj
    if compat.is_ensemble(obj) and probe.attr == "decoded_output":
        dec = cache.solve(obj, probe.solver, model.params[obj].eval_points, key=probe).T
        model.probe_conns[probe] = dec[np.arange(obj.dimensions)[probe.slice]]
    model.params[probe] = None


def build_model(network, dt=0.001, seed=None, seed_override=None, device_solver=None):
    """Build every ensemble / connection / probe of ``network`` (and sub-networks).

    ``seed_override`` replaces the network's own seed: the same declared graph built as if the driver had been started
    with another ``--seed`` (``nengo.Network(seed=args.seed)``, run_slam.py:151) - one built model per trial of a batch
    whose trials have their own network seeds.

    ``device_solver``: GPU index on which the decoder systems are solved (``libssb_builder.so``; default: environment
    variable ``SSB_DEVICE_BUILDER``, else the host solver)."""
    model = BuiltModel(network, dt)
    top_seed = getattr(network, "seed", None) if seed_override is None else int(seed_override)
    if top_seed is None:
        top_seed = seed if seed is not None else np.random.randint(maxint)
    model.seeds[network] = int(top_seed)
    _assign_seeds(network, model.seeds)

    with _blas_threads():
        return _build_all(model, network, device_solver)


class _blas_threads:
    """The builder's matrices are small (<= 2 500 x 1 000): multi-threaded BLAS / LAPACK only thrashes on them (measured:
    ``cho_factor`` of a 500 x 500 Gram matrix 121 ms with 8 threads, 2.2 ms with one).  ``SSB_BUILDER_THREADS`` overrides."""

    def __enter__(self):
        self._ctx = None
        try:
            from threadpoolctl import threadpool_limits
            self._ctx = threadpool_limits(limits=int(os.environ.get("SSB_BUILDER_THREADS", "1")))
        except Exception:
            pass
        return self

    def __exit__(self, *exc):
        if self._ctx is not None:
            self._ctx.__exit__(*exc)


def _builder_procs(n_ensembles):
    """Worker processes of the host builder: ``SSB_BUILDER_PROCS`` (0 / 1 = in-process), else the cores this process may use
    when the network is large enough to pay for the fork (and this process is not itself a pool worker)."""
    import multiprocessing as mp
    if mp.current_process().daemon:
        return 1
    env = os.environ.get("SSB_BUILDER_PROCS")
    if env is not None:
        return max(1, int(env))
    if n_ensembles < 128 or not hasattr(os, "fork"):
        return 1
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cores //= max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))        # one process per GPU shares the host
    return max(1, min(cores, 16, n_ensembles // 16))


_PAR_CTX = None      # (model, ensembles, connections / probes grouped by their pre-ensemble) for the forked workers


def _par_job(i):
    model, ensembles, by_ens = _PAR_CTX
    ens = ensembles[i]
    _build_ensemble(model, ens)
    cache = _DecoderCache(model)
    conns, probes = by_ens.get(ens, ((), ()))
    out_c, out_p = [], []
    for ci, conn in conns:
        _build_connection(model, conn, cache)
        out_c.append((ci, model.params[conn]))
    for pi, probe in probes:
        _build_probe(model, probe, cache)
        out_p.append((pi, model.probe_conns.get(probe)))
    return model.params[ens], out_c, out_p


def _build_parallel(model, network, procs):
    """Every ensemble's sampling, tuning curves and decoder solves depend on its own seed only, so ensembles (with the
    decoded connections and probes leaving them) are built by a fork pool; the parent stores the results in the order of
    the in-process build, so the two are indistinguishable (tests/test_builder_parallel.py)."""
    import multiprocessing as mp
    global _PAR_CTX
    ensembles, conns, probes = list(network.all_ensembles), list(network.all_connections), list(network.all_probes)
    by_ens = {}
    for ci, c in enumerate(conns):
        if compat.is_ensemble(c.pre_obj):
            by_ens.setdefault(c.pre_obj, ([], []))[0].append((ci, c))
    for pi, p in enumerate(probes):
        if compat.is_ensemble(p.obj) and p.attr == "decoded_output":
            by_ens.setdefault(p.obj, ([], []))[1].append((pi, p))
    _PAR_CTX = (model, ensembles, by_ens)
    try:
        with mp.get_context("fork").Pool(procs) as pool:
            results = pool.map(_par_job, range(len(ensembles)), chunksize=max(1, len(ensembles) // (8 * procs)))
    finally:
        _PAR_CTX = None
    built_c, built_p = {}, {}
    for ens, (pe, out_c, out_p) in zip(ensembles, results):
        model.params[ens] = pe
        built_c.update(out_c)
        built_p.update(out_p)
    cache = _DecoderCache(model)
    for ci, conn in enumerate(conns):
        if ci in built_c:
            model.params[conn] = built_c[ci]
        else:
            _build_connection(model, conn, cache)
    for pi, probe in enumerate(probes):
        if pi in built_p:
            model.probe_conns[probe] = built_p[pi]
            model.params[probe] = None
        else:
            _build_probe(model, probe, cache)
    return model


def _build_all(model, network, device_solver):
    if device_solver is None and os.environ.get("SSB_DEVICE_BUILDER") not in (None, "", "off"):
        device_solver = int(os.environ["SSB_DEVICE_BUILDER"])
    procs = _builder_procs(len(network.all_ensembles)) if device_solver is None else 1
    if procs > 1:
        return _build_parallel(model, network, procs)
    cache = _DecoderCache(model)
    for ens in network.all_ensembles:
        _build_ensemble(model, ens)
    if device_solver is not None:
        cache.presolve_on_device(network, int(device_solver))
    for conn in network.all_connections:
        _build_connection(model, conn, cache)
    for probe in network.all_probes:
        _build_probe(model, probe, cache)
    return model
