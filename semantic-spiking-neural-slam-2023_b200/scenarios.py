"""Synthetic workloads that follow the reference drivers (``experiments/run_pathint.py``,
``run_slam.py``): random 2-D paths, R_d landmarks, hexagonal SSP space,
the network declaration and per-trial input tables.  Used by ``bench.py``, ``smoke()``
and the GPU tests (the GPU box has no reference checkout)."""
from __future__ import annotations

import dataclasses
import numpy as np

from . import nengo_shim as nengo
from . import networks, inputs
from .sspspace import HexagonalSSPSpace, SPSpace


@dataclasses.dataclass
class Scenario:
    network: object
    probe: object
    trial_inputs: dict
    ssp_space: object
    paths: np.ndarray          # [n_trials, T, dim]
    real_ssp: np.ndarray       # [n_trials, T, d]
    extra: dict
    dt: float = 0.001


def _neuron_type(name):
    return {"lif": nengo.LIF, "lifrate": nengo.LIFRate, "relu": nengo.RectifiedLinear}[name]()


def make_space(domain_dim=2, ssp_dim=55, length_scale=0.2, radius=1.0, backend="host"):
    bounds = radius * np.tile([-1.0, 1.0], (domain_dim, 1))
    return HexagonalSSPSpace(domain_dim, ssp_dim=ssp_dim, domain_bounds=bounds, length_scale=length_scale,
                             rng=np.random.default_rng(0), backend=backend)


class _PathintJob:
    """Inputs of one path-integration trial (picklable, see ``_map_trials``)."""

    def __init__(self, space, scale, T, dt, limit, seed, n_steps):
        self.__dict__.update(locals())

    def __call__(self, i):
        path = inputs.random_path(self.T, self.dt, self.limit, self.seed + 1000 * i, 2)
        vels = inputs.velocities(path, self.dt) * self.scale
        n_keep = max(self.n_steps + 2, 4)
        real = self.space.encode_host(path[:n_keep])
        tb = inputs.pathint_tables(real, vels[:n_keep], self.n_steps, self.dt)
        return tb, path[:n_keep], real[:self.n_steps], vels[:n_keep]


def make_pathint(n_trials=1, n_steps=1000, ssp_dim=55, pi_n_neurons=500, T=20.0, limit=0.1, seed=0,
                 neuron_type="lif", dt=0.001, tau=0.05, with_gcs=False, n_gcs=1000, distinct_tables=None, workers=None):
    """``run_pathint.py`` workload; trial i uses path seed ``seed + 1000*i``."""
    space = make_space(2, ssp_dim)
    d = space.ssp_dim
    scale = inputs.velocity_scale(space.phase_matrix, inputs.velocities(inputs.random_path(T, dt, limit, seed, 2), dt))
    n_distinct = n_trials if distinct_tables is None else min(n_trials, distinct_tables)
    results = _map_trials(_PathintJob(space, scale, T, dt, limit, seed, n_steps), n_distinct, workers)
    reps = -(-n_trials // n_distinct)
    results = (results * reps)[:n_trials]
    tabs = [r[0] for r in results]
    paths = [r[1][:n_steps] for r in results]
    ssps = [r[2] for r in results]
    full_paths = [r[1] for r in results]
    full_vels = [r[3] for r in results]
    vel0, init0 = tabs[0]["vel"], tabs[0]["init"]
    model = nengo.Network(seed=seed)
    model.config[nengo.Ensemble].neuron_type = _neuron_type(neuron_type)
    with model:
        vel_in = nengo.Node(lambda t: vel0[int(round(t / dt)) - 1], label="vel_input")
        init = nengo.Node(lambda t: init0[int(round(t / dt)) - 1], label="init_state")
        pi = networks.PathIntegration(space, pi_n_neurons, tau, scaling_factor=scale, stable=True,
                                      solver_weights=False, with_gcs=with_gcs, n_gcs=n_gcs)
        nengo.Connection(vel_in, pi.velocity_input, synapse=None)
        nengo.Connection(init, pi.input, synapse=None)
        probe = nengo.Probe(pi.output, synapse=0.05)
    trial_inputs = {vel_in: np.stack([t["vel"] for t in tabs]), init: np.stack([t["init"] for t in tabs])}
    synth = dict(nodes={"vel": vel_in, "init": init}, ssp_space=space, path=np.stack(full_paths),
                 vels_scaled=np.stack(full_vels))
    return Scenario(model, probe, trial_inputs, space, np.stack(paths), np.stack(ssps),
                    dict(pathint=pi, vel_scale=scale, input_synthesis=synth), dt)


class _TrialJob:
    """Inputs of one trial (path, scaled velocities, landmark set, tables for the first ``table_steps`` steps); a
    picklable callable so that a batch of distinct trials can be synthesised by a process pool."""

    def __init__(self, space, lm_vectors, scale, recorded, T, dt, limit, seed, domain_dim, n_landmarks, view_rad,
                 n_steps, table_steps, view, view_bound, table_dtype, trial0):
        self.__dict__.update(locals())

    def __call__(self, i):
        s_i = self.seed + 1000 * (self.trial0 + i)
        n_keep = max(self.n_steps + 2, 4)
        if self.recorded is not None:
            path = self.recorded[:n_keep]
        else:
            path = inputs.random_path(self.T, self.dt, self.limit, s_i, self.domain_dim)[:n_keep]
        vels = inputs.velocities(path, self.dt) * self.scale
        obj_locs = 1.8 * (inputs.rd_sampling(self.n_landmarks, self.domain_dim, seed=s_i) - 0.5)
        n_tab = self.table_steps
        head = path[:max(n_tab + 2, 4)]            # the table index rules only look at rows < n_tab + 2
        vec_to_lm = obj_locs[None, :, :] - head[:, None, :]
        space, lmv, dt = self.space, self.lm_vectors, self.dt
        real = space.encode_host(head)
        kw = dict(real_ssp=real)     # (the ``pathlen - 2`` clip of the index rule never bites for steps <= n_tab)
        if self.view and self.view_bound:
            vt = inputs.slamview_tables(space, lmv, vels[:len(head)], vec_to_lm, self.view_rad, n_tab, dt)
            tb = inputs.slam_tables(space.encode_host, lmv, vels[:len(head)], vec_to_lm, self.view_rad, n_tab, dt,
                                    none_in_view_value=1.0, **kw)
            tb.update(vel=vt["vel"], lm_sp=vt["view"], nolm=vt["nolm"])
        elif self.view:
            # default: the superposed landmark SPs as the view key (what the on-device input synthesis evaluates)
            tb = inputs.slam_tables(space.encode_host, lmv, vels[:len(head)], vec_to_lm, self.view_rad, n_tab, dt,
                                    none_in_view_value=1.0, **kw)
        else:
            tb = inputs.slam_tables(space.encode_host, lmv, vels[:len(head)], vec_to_lm, self.view_rad, n_tab, dt, **kw)
        tb = {k: np.asarray(v, dtype=self.table_dtype) for k, v in tb.items()}
        return tb, path, real[:n_tab].astype(self.table_dtype), vels, obj_locs


def _map_trials(job, n, workers=None):
    """Evaluate ``job(i)`` for i < n, on a fork pool when there are enough trials to pay for it."""
    import os
    if workers is None:
        workers = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    workers = int(max(1, min(workers, n // 8)))
    if workers <= 1:
        return [job(i) for i in range(n)]
    import multiprocessing as mp
    with mp.get_context("fork").Pool(workers) as pool:
        return pool.map(job, range(n), chunksize=max(1, n // (4 * workers)))


def make_slam(n_trials=1, n_steps=1000, ssp_dim=55, pi_n_neurons=500, mem_n_neurons=970, circonv_n_neurons=100,
              n_landmarks=50, view_rad=0.2, T=200.0, limit=0.1, seed=0, dt=0.001, length_scale=0.2,
              shift_rate=0.2, update_thres=0.2, neuron_type="lif", weights_probe=False, view=False,
              distinct_tables=None, domain_dim=2, grid_points_per_dim=100, gc_n_neurons=0, approx_vel=False,
              vel_n_neurons=500, loihi=False, dotprod_n_neurons=50, inverse_memory=False, voja=True, view_bound=False,
              table_steps=None, path_data=None, data_dt=0.001, workers=None, table_dtype=np.float64, trial0=0):
    """``run_slam.py`` (or ``run_slamview.py`` when ``view``) workload, batched over trials.

    ``distinct_tables``: synthesise only that many distinct trials' tables and tile them
    over the batch (bench warm-up economy); state/voltages still differ per trial.
    ``loihi``: the all-neural ``SLAMLoihiNetwork`` with the driver's arguments of ``run_slam.py:171-176``.
    ``inverse_memory``: the topology of ``experiments/slam_map_new.py:207-263`` — a second, uncorrected path integrator and
    a second Voja + PES memory mapping landmark locations back to landmark SPs, with that script's probes.
    ``view_bound`` (with ``view``): feed the local view of ``run_slamview.py:103,123-130`` — the normalised sum of
    ``SP_l (*) encode(displacement_l)`` with that driver's index rules (``inputs.slamview_tables``) — instead of the SP sum;
    tables only (the on-device input synthesis evaluates the SP-sum form).
    ``table_steps``: build host tables / ``real_ssp`` for the first ``table_steps`` steps only, while the per-trial paths
    handed to the on-device input synthesis cover ``n_steps`` (long sustained runs need no host tables).
    ``path_data`` / ``data_dt``: a recorded path (array or ``.npy`` file) with the ``--path-data`` rules of
    ``run_slam.py:100-112`` (``inputs.load_path``); every trial then walks the same path and trial ``i`` differs in its
    landmark set (``Rd_sampling(seed + 1000 i)``) and start state — BASELINE configs[2] / [3].
    ``trial0``: global id of the first trial (sharded runs: rank r passes ``r * n_trials``; trial ``i`` uses path / landmark
    seed ``seed + 1000 * (trial0 + i)`` while the network seed — the shared static weights — stays ``seed``).
    ``table_dtype``: dtype of the host tables (the device arena is float32; large batches pass ``np.float32``).
    ``workers``: processes used to synthesise the distinct trials (default: the host cores this process may use)."""
    recorded = None
    if path_data is not None:
        recorded = inputs.load_path(path_data, data_dt, dt)
        domain_dim = recorded.shape[1]
    space = make_space(domain_dim, ssp_dim, length_scale)
    d = space.ssp_dim
    lm_space = SPSpace(n_landmarks, d, seed=seed)
    n_distinct = n_trials if distinct_tables is None else min(n_trials, distinct_tables)
    table_steps = n_steps if table_steps is None else min(int(table_steps), n_steps)
    if recorded is not None:
        scale = inputs.velocity_scale(space.phase_matrix, inputs.velocities(recorded, dt))
    else:
        scale = inputs.velocity_scale(space.phase_matrix,
                                      inputs.velocities(inputs.random_path(T, dt, limit, seed, domain_dim), dt))
    job = _TrialJob(space, lm_space.vectors, scale, recorded, T, dt, limit, seed, domain_dim, n_landmarks, view_rad,
                    n_steps, table_steps, view, view_bound, table_dtype, trial0)
    results = _map_trials(job, n_distinct, workers)
    tabs = [r[0] for r in results]
    paths = [r[1][:table_steps] for r in results]
    ssps = [r[2] for r in results]
    syn_paths = [r[1] for r in results]
    syn_vels = [r[3] for r in results]
    syn_lms = [r[4] for r in results]
    t0 = tabs[0]

    def tab_fn(name):
        arr = t0[name]
        return lambda t: arr[int(round(t / dt)) - 1]

    np.random.seed(seed)  # OVC encoders come from the global stream (slam.py:206)
    model = nengo.Network(seed=seed)
    model.config[nengo.Ensemble].neuron_type = _neuron_type(neuron_type)
    with model:
        vel_in = nengo.Node(tab_fn("vel"), label="vel_input")
        init = nengo.Node(tab_fn("init"), label="init_state")
        lm_id = nengo.Node(tab_fn("lm_sp"), label="lm_sp_input")
        is_lm = nengo.Node(tab_fn("nolm"), label="lm_in_view_input")
        if loihi:
            lm_vec = nengo.Node(tab_fn("lmvec_ssp"), label="lm_vecssp_input")
            slam = networks.SLAMLoihiNetwork(space, lm_space, view_rad, n_landmarks, pi_n_neurons, mem_n_neurons,
                                             circonv_n_neurons, dotprod_n_neurons, vel_in, lm_vec, lm_id, is_lm,
                                             tau_pi=0.05, update_thres=update_thres, vel_scaling_factor=scale,
                                             shift_rate=0.1, pes_learning_rate=1e-3, encoders=None, seed=seed)
            table_nodes = {"vel": vel_in, "init": init, "lm_sp": lm_id, "nolm": is_lm, "lmvec_ssp": lm_vec}
        elif view:
            slam = networks.SLAMViewNetwork(space, lm_space, view_rad, n_landmarks, pi_n_neurons, mem_n_neurons,
                                            circonv_n_neurons, tau_pi=0.05, update_thres=update_thres,
                                            vel_scaling_factor=scale, shift_rate=0.02, voja_learning_rate=5e-4,
                                            pes_learning_rate=1e-3, grid_points_per_dim=grid_points_per_dim)
            nengo.Connection(lm_id, slam.view_input, synapse=None)
            table_nodes = {"vel": vel_in, "init": init, "lm_sp": lm_id, "nolm": is_lm}
        else:
            lm_vec = nengo.Node(tab_fn("lmvec_ssp"), label="lm_vecssp_input")
            slam = networks.SLAMNetwork(space, lm_space, view_rad, n_landmarks, pi_n_neurons, mem_n_neurons,
                                        circonv_n_neurons, tau_pi=0.05, update_thres=update_thres,
                                        vel_scaling_factor=scale, shift_rate=shift_rate, voja_learning_rate=1e-4,
                                        pes_learning_rate=5e-3, intercept=0.1, seed=seed,
                                        grid_points_per_dim=grid_points_per_dim, gc_n_neurons=gc_n_neurons, voja=voja)
            nengo.Connection(lm_vec, slam.landmark_vec_ssp, synapse=None)
            nengo.Connection(lm_id, slam.landmark_id_input, synapse=None)
            table_nodes = {"vel": vel_in, "init": init, "lm_sp": lm_id, "nolm": is_lm, "lmvec_ssp": lm_vec}
        if not loihi:
            nengo.Connection(is_lm, slam.no_landmark_in_view, synapse=None)
        if loihi:
            pass                     # the driver's nodes are the network's inputs
        elif approx_vel:   # run_slam.py:154-160: the velocity passes through a spiking ensemble (vel_syn = 0.01)
            vel_ens = nengo.Ensemble(vel_n_neurons, domain_dim)
            nengo.Connection(vel_in, vel_ens, synapse=None)
            nengo.Connection(vel_ens, slam.velocity_input, synapse=0.01)
        else:
            nengo.Connection(vel_in, slam.velocity_input, synapse=None)
        nengo.Connection(init, slam.pathintegrator.input, synapse=None)
        probe = nengo.Probe(slam.pathintegrator.output, synapse=0.05)
        more = {}
        if inverse_memory:
            pi2 = networks.PathIntegration(space, pi_n_neurons, 0.05, scaling_factor=scale, stable=True,
                                           solver_weights=False)
            nengo.Connection(vel_in, pi2.velocity_input, synapse=None)
            nengo.Connection(init, pi2.input, synapse=None)
            inv = networks.AssociativeMemory(mem_n_neurons, d, d, 0.1, voja_learning_rate=5e-4, pes_learning_rate=1e-2,
                                             voja=True, encoders=space.sample_grid_encoders(mem_n_neurons), radius=1.3)
            nengo.Connection(slam.landmark_ssp_ens.output, inv.key_input, synapse=0.05)
            nengo.Connection(lm_id, inv.value_input, synapse=None)
            nengo.Connection(is_lm, inv.learning, synapse=None)
            every = n_steps * dt
            more = dict(pathintegrator2=pi2, invassomemory=inv,
                        ssp_pi_p=nengo.Probe(pi2.output, synapse=0.05),
                        newpos_p=nengo.Probe(slam.position_estimate.output, synapse=0.05),
                        objssp_p=nengo.Probe(slam.landmark_ssp_ens.output, synapse=0.05),
                        recall_p=nengo.Probe(slam.assomemory.recall, synapse=0.05),
                        isitem_p=nengo.Probe(is_lm, synapse=None),
                        mem_weights=nengo.Probe(slam.assomemory.conn_out, "weights", sample_every=every / 2),
                        meminv_weights=nengo.Probe(inv.conn_out, "weights", sample_every=every),
                        mem_encoders=nengo.Probe(slam.assomemory.conn_in.learning_rule, "scaled_encoders",
                                                 sample_every=every / 2),
                        meminv_encoders=nengo.Probe(inv.conn_in.learning_rule, "scaled_encoders", sample_every=every))
        wprobe = None
        if weights_probe:
            wprobe = nengo.Probe(slam.assomemory.conn_out, "weights", sample_every=n_steps * dt)
    reps = -(-n_trials // n_distinct)
    trial_inputs = {}
    for name, node in table_nodes.items():
        stacked = np.stack([t[name] for t in tabs])
        trial_inputs[node] = np.tile(stacked, (reps, 1, 1))[:n_trials]
    paths = np.tile(np.stack(paths), (reps, 1, 1))[:n_trials]
    ssps = np.tile(np.stack(ssps), (reps, 1, 1))[:n_trials]
    tile = lambda lst: np.tile(np.stack(lst), (reps, 1, 1))[:n_trials]
    synth = dict(nodes={k: table_nodes.get(k) for k in ("vel", "init", "lmvec_ssp", "lm_sp", "nolm")}, ssp_space=space,
                 path=tile(syn_paths), vels_scaled=tile(syn_vels), landmarks=tile(syn_lms), lm_vectors=lm_space.vectors,
                 view_rad=view_rad, none_in_view_value=1.0 if view else 10.0)
    if view and view_bound:
        synth = None
    return Scenario(model, probe, trial_inputs, space, paths, ssps,
                    dict(slam=slam, vel_scale=scale, weights_probe=wprobe, lm_space=lm_space, input_synthesis=synth,
                         **more), dt)
