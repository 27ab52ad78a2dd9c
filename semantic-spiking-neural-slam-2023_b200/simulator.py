"""``Simulator(network, dt)`` — drop-in for ``nengo.Simulator`` on the SSP-SLAM graphs.

Mirrors the surface the reference drivers use:
``experiments/run_slam.py:198-199`` (construction), ``:232-233`` (``with sim: sim.run(T)``),
``:243,250`` (``sim.trange()``, ``sim.data[probe]`` read *after* the ``with`` block),
``:265-266`` (``sim.data[Probe(conn,'weights')]``, ``sim.data[ensemble]``), plus
``run_steps`` / ``step`` / ``reset`` / ``n_steps`` / ``time`` / ``dt`` from nengo's convention.

Batching extension (not in the reference, which has no batch axis):
``Simulator(network, n_trials=B, trial_inputs={node: array[B, T, size]}, trial_seeds=[...])``
runs B independent trials that share the static weights; learned PES/Voja matrices and
all state are per trial.  ``sim.data[probe]`` then has a leading trial axis.
``trial_seeds`` select the start voltages: ``None`` = nengo's own draw for the built model, an int = a hash of (ensemble
seed, that int, neuron).  The default is ``[None, 0, 1, ...]`` so that trial 0 of a batch IS the unbatched nengo run; a
sharded job passes ``sharding.trial_seeds`` (global trial ids) explicitly, so a trial's start state does not depend on
how the batch is split over GPUs.  ``trial_network_seeds`` gives every trial its own NETWORK seed instead (one built model
per distinct seed, per-trial static weights).

``input_synthesis=dict(...)`` (optional) evaluates the drivers' per-step input closures on the device from
per-trial paths / landmarks (``slam.py:442-497``) instead of ``trial_inputs`` tables.

Everything that is stepped runs in the CUDA library behind ``include/sspslam_b200.h``;
if the library or a GPU is missing, construction raises.
"""
from __future__ import annotations

import numpy as np

from . import cabi, lowering
from . import nengo_shim as ns
from .builder import build_model, BuiltModel


class _SimData:
    def __init__(self, sim):
        self._sim = sim

    def __getitem__(self, key):
        sim = self._sim
        if key in sim._probe_infos:
            return sim._probe_array(key)
        if key in sim.model.params:
            return sim.model.params[key]
        raise KeyError(key)

    def __contains__(self, key):
        return key in self._sim._probe_infos or key in self._sim.model.params

    def keys(self):
        return list(self._sim._probe_infos) + list(self._sim.model.params)


def _build_one(args):
    network, dt, s = args
    return build_model(network, dt=dt, seed_override=s)


def _build_models(network, dt, seeds, workers=None):
    """One built model per network seed (identical seeds are built once), on a fork pool when there are many."""
    import os
    distinct = sorted(set(int(s) for s in seeds))
    if workers is None:
        workers = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else 1
    workers = max(1, min(int(workers), len(distinct) // 2))
    if workers > 1:
        import multiprocessing as mp
        global _POOL_NET
        _POOL_NET = (network, dt)
        with mp.get_context("fork").Pool(workers) as pool:
            built = [_unpack_model(network, dt, p) for p in pool.map(_build_pool_job, distinct, chunksize=1)]
    else:
        built = [build_model(network, dt=dt, seed_override=s) for s in distinct]
    by_seed = dict(zip(distinct, built))
    return [by_seed[int(s)] for s in seeds]


_POOL_NET = None


def _build_pool_job(seed):
    network, dt = _POOL_NET
    m = build_model(network, dt=dt, seed_override=seed)
    # objects of the (forked) network are keys of the model's dicts: send back plain lists in traversal order
    return _pack_model(network, m)


def _model_objects(network):
    objs = list(network.all_ensembles) + list(network.all_connections) + list(network.all_probes)
    return objs, [network] + list(network.all_networks) + objs + list(network.all_nodes)


def _unpack_model(network, dt, packed):
    """Re-key a model built in a forked worker (whose keys are the worker's copies of the objects) by this process's objects."""
    from .builder import BuiltModel
    objs, seed_objs = _model_objects(network)
    m = BuiltModel(network, dt)
    m.seeds = {o: s for o, s in zip(seed_objs, packed["seeds"]) if s is not None}
    m.params = dict(zip(objs, packed["params"]))
    m.probe_conns = {p: v for p, v in zip(network.all_probes, packed["probe_conns"]) if v is not None}
    return m


def _pack_model(network, m):
    objs, seed_objs = _model_objects(network)
    return dict(seeds=[m.seeds.get(o) for o in seed_objs], params=[m.params.get(o) for o in objs],
                probe_conns=[m.probe_conns.get(p) for p in network.all_probes])


class Simulator:
    def __init__(self, network, dt=0.001, seed=None, model: BuiltModel | None = None, progress_bar=True,
                 optimize=True, n_trials=None, trial_inputs=None, trial_seeds=None, device=0, chunk_steps=256,
                 input_synthesis=None, keep_probe_history=True, trial_network_seeds=None, build_workers=None):
        self.network = network
        self.dt = float(dt)
        self.closed = False
        self._lib = cabi.load()  # raises if the CUDA library has not been built
        self._batched = n_trials is not None
        self.n_trials = int(n_trials) if self._batched else 1
        self.chunk_steps = int(chunk_steps)
        self.models = None
        if trial_network_seeds is not None:
            # every trial is the driver started with its own --seed (run_slam.py:151): one built model per trial, per-trial
            # static weights on the device (narrow ensembles: per-trial weight arena; wide ensembles: encoders in the lenc
            # arena, decoders in the ldec arena)
            if len(trial_network_seeds) != self.n_trials:
                raise ValueError("trial_network_seeds must have one entry per trial")
            self.models = _build_models(network, self.dt, list(trial_network_seeds), build_workers)
            model = self.models[0]
        self.model = model if model is not None else build_model(network, dt=self.dt, seed=seed)
        self.plan = lowering.lower(network, self.model, chunk_cap=self.chunk_steps, n_trials=self.n_trials,
                                   per_trial=self.models is not None)
        if self.models is not None:
            if trial_seeds is None:
                trial_seeds = [None] * self.n_trials          # nengo's own start-voltage draw of every trial's model
        self._trial_inputs = dict(trial_inputs or {})
        if trial_seeds is None:
            trial_seeds = [None] + list(range(self.n_trials - 1)) if self._batched else [None]
        if len(trial_seeds) != self.n_trials:
            raise ValueError("trial_seeds must have one entry per trial")
        self.trial_seeds = list(trial_seeds)
        for node, arr in self._trial_inputs.items():
            arr = np.asarray(arr)
            if arr.ndim != 3 or arr.shape[0] != self.n_trials or arr.shape[2] != node.size_out:
                raise ValueError(f"trial_inputs[{node!r}] must have shape (n_trials, n_steps, {node.size_out})")

        import ctypes as C
        handle = C.c_void_p()
        cabi.check(self._lib.ssb_create(int(device), self.n_trials, C.byref(handle)), "ssb_create")
        self._h = handle
        self._keep = []
        for name, arr in self.plan.arrays.items():
            a = np.ascontiguousarray(arr)
            self._keep.append(a)
            cabi.check(self._lib.ssb_set_array(self._h, name.encode(), cabi._ptr(a), a.nbytes), f"set_array {name}")
        for name, val in self.plan.scalars.items():
            cabi.check(self._lib.ssb_set_scalar(self._h, name.encode(), float(val)), f"set_scalar {name}")
        cabi.check(self._lib.ssb_finalize(self._h), "ssb_finalize")
        self.B = int(self._lib.ssb_n_trials_padded(self._h))
        self._nt = int(self.plan.scalars["nt"])
        self._np = int(self.plan.scalars["n_probe"])
        self._tab_buf_ = None        # page-locked table staging of one chunk, allocated on first use
        self.keep_probe_history = bool(keep_probe_history)
        # two page-locked probe buffers: the host-side copy-out of chunk k overlaps the device work of chunk k + 1
        self._probe_bufs = [cabi.PinnedBuffer(max(1, self.chunk_steps * self._np * self.B)) for _ in range(2)]
        self._probe_cur = 0
        self._probe_pending = None   # (buffer index, n_steps) whose rows have not been copied out yet
        self._probe_infos = {info.probe: info for info in self.plan.probes}
        self._n_steps = 0
        self._synth = None
        if input_synthesis is not None:
            self._setup_input_synthesis(input_synthesis)
        self._init_state()

    @property
    def _tab_buf(self):
        if self._tab_buf_ is None:
            self._tab_buf_ = cabi.PinnedBuffer(max(1, self.chunk_steps * self._nt * self.B))
        return self._tab_buf_

    # ------------------------------------------------------------------ on-device input synthesis
    def _setup_input_synthesis(self, spec):
        """Evaluate the drivers' input closures on the device (``ssb_synth_setup``) instead of
        feeding per-step tables.  ``spec``: ``nodes`` {'vel','init','lmvec_ssp','lm_sp','nolm' -> Node}, ``ssp_space``,
        per-trial ``path`` / ``vels_scaled`` ``[n_trials, T, dim]``, ``landmarks`` ``[n_trials, n_lm, dim]``,
        ``lm_vectors`` ``[n_lm, d]``, ``view_rad``, ``none_in_view_value``, ``init_time`` — the arguments of
        ``get_slam_input_functions2`` (``slam.py:442-497``) / the ``run_pathint.py:134-136`` lambdas."""
        import ctypes as C
        space = spec["ssp_space"]
        path = np.asarray(spec["path"], dtype=np.float64)
        vels = np.asarray(spec["vels_scaled"], dtype=np.float64)
        if path.ndim != 3 or path.shape[0] != self.n_trials or vels.shape != path.shape:
            raise ValueError("input_synthesis: path / vels_scaled must be [n_trials, T, dim]")
        T, dim = path.shape[1], path.shape[2]
        d = space.ssp_dim
        col_of = {node: col0 for node, col0, size in self.plan.tables}
        nodes = spec["nodes"]
        missing = [n for n in col_of if n not in nodes.values()]
        if missing:
            raise ValueError(f"input_synthesis: table nodes without a synthesised signal: {missing}")
        cols = [col_of[nodes[k]] if nodes.get(k) is not None else -1 for k in ("vel", "init", "lmvec_ssp", "lm_sp", "nolm")]
        lms = spec.get("landmarks")
        n_lm = 0 if lms is None else np.asarray(lms).shape[1]

        def rows(a):                                      # [n_trials, R] -> float32 [R, B]
            out = np.zeros((a.shape[1], self.B), dtype=np.float32)
            out[:, :self.n_trials] = a.T
            return np.ascontiguousarray(out)
        path_rows, vel_rows = rows(path.reshape(self.n_trials, T * dim)), rows(vels.reshape(self.n_trials, T * dim))
        lm_rows = rows(np.asarray(lms, dtype=np.float64).reshape(self.n_trials, n_lm * dim)) if n_lm else None
        lm_sp = np.ascontiguousarray(spec["lm_vectors"], dtype=np.float64) if n_lm else None
        phases = np.ascontiguousarray(space._scaled_phases(), dtype=np.float64)            # [d, dim] = A / length_scale
        cfg = np.asarray([dim, d, n_lm, T] + cols, dtype=np.int32)
        fpar = np.asarray([spec.get("view_rad", 0.0), spec.get("none_in_view_value", 10.0)], dtype=np.float32)
        ptr = lambda a: cabi._ptr(a) if a is not None else None
        cabi.check(self._lib.ssb_synth_setup(self._h, ptr(cfg), ptr(fpar), ptr(phases), ptr(lm_sp), ptr(path_rows),
                                             ptr(vel_rows), ptr(lm_rows)), "ssb_synth_setup")
        self._synth = dict(T=T, init_time=float(spec.get("init_time", 0.05)))

    def _synth_steps(self, step0, n):
        from . import inputs
        t, i_prev, i_cur = inputs.step_indices(n, self.dt, self._synth["T"], step0)
        idx = np.ascontiguousarray(np.stack([i_prev, i_cur, (t < self._synth["init_time"]).astype(np.int64)], axis=1),
                                   dtype=np.int32)
        cabi.check(self._lib.ssb_synth_steps(self._h, cabi._ptr(idx), int(step0), int(n)), "ssb_synth_steps")

    # ------------------------------------------------------------------ state
    def _rows(self, per_trial):
        """[n_trials, rows] (or [rows] broadcast) -> float32 [rows, B] with zero padding."""
        a = np.asarray(per_trial, dtype=np.float32)
        if a.ndim == 1:
            a = np.broadcast_to(a[None, :], (self.n_trials, a.shape[0]))
        out = np.zeros((a.shape[1], self.B), dtype=np.float32)
        out[:, :self.n_trials] = a.T
        return out

    def _upload(self, arena, row0, rows):
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        cabi.check(self._lib.ssb_upload(self._h, arena.encode(), int(row0), int(rows.shape[0]), cabi._ptr(rows)),
                   f"upload {arena}")

    def _download(self, arena, row0, n_rows):
        out = np.empty((int(n_rows), self.B), dtype=np.float32)
        cabi.check(self._lib.ssb_download(self._h, arena.encode(), int(row0), int(n_rows), cabi._ptr(out)),
                   f"download {arena}")
        return out

    # learned decoders: a trial group's block is [neuron][trial][JP] floats (csrc/ssb_pes.cuh), JP = size_out rounded up to 4
    def _upload_decoders(self, conn, per_trial):
        """``per_trial`` [n_trials, size_out, n] -> device order [group][n * JP rows][32]."""
        row0, size_out, n = conn if isinstance(conn, tuple) else self.plan.learned_dec[conn]
        jp = -(-size_out // 4) * 4
        G = self.B // 32
        full = np.zeros((self.B, jp, n), dtype=np.float32)
        full[:self.n_trials, :size_out] = per_trial
        dev = np.ascontiguousarray(full.reshape(G, 32, jp, n).transpose(0, 3, 1, 2))       # [G][n][32][jp]
        dev = dev.reshape(G, n * jp, 32)
        cabi.check(self._lib.ssb_upload(self._h, b"ldec", int(row0), int(n * jp), cabi._ptr(dev)), "upload ldec")

    def _download_decoders(self, conn):
        """Device decoder block -> [n_trials, size_out, n] float32."""
        row0, size_out, n = self.plan.learned_dec[conn]
        jp = -(-size_out // 4) * 4
        G = self.B // 32
        dev = np.empty((G, n * jp, 32), dtype=np.float32)
        cabi.check(self._lib.ssb_download(self._h, b"ldec", int(row0), int(n * jp), cabi._ptr(dev)), "download ldec")
        full = dev.reshape(G, n, 32, jp).transpose(0, 2, 3, 1).reshape(self.B, jp, n)
        return np.ascontiguousarray(full[:self.n_trials, :size_out])

    def _init_state(self):
        m, plan = self.model, self.plan
        nn = int(plan.scalars["nn"])
        if self.models is not None:
            # per-trial static weights: the weight arena / wide encoders / wide decoders of every trial's model
            distinct, which = {}, []
            for mt in self.models:
                if id(mt) not in distinct:
                    wt, enc, dec = lowering.trial_weights(self.network, mt, self.n_trials)
                    if wt.size != plan.arrays["weights_pt"].size:
                        raise RuntimeError("per-trial models lower to different weight layouts")
                    distinct[id(mt)] = (len(distinct), wt, enc, dec)
                which.append(distinct[id(mt)][0])
            which = np.asarray(which)
            per_model = sorted(distinct.values(), key=lambda p: p[0])
            stacked = np.stack([p[1] for p in per_model])                                           # [n_models, n_w]
            w = np.zeros((stacked.shape[1], self.B), dtype=np.float32)
            w[:, :self.n_trials] = stacked.T[:, which]
            self._upload("wpt", 0, w)
            for ens, (row0, n, dims) in plan.pt_enc.items():
                e = np.stack([p[2][ens].reshape(-1) for p in per_model])                            # [n_models, n * dims]
                self._upload("lenc", row0, self._rows(e[which]))
            for conn, (row0, size_out, n) in plan.pt_dec.items():
                d = np.stack([p[3][conn] for p in per_model])                                       # [n_models, size_out, n]
                self._upload_decoders((row0, size_out, n), d[which])
        if nn:
            v0 = np.zeros((nn, self.B), dtype=np.float32)
            for ens, (row0, n) in plan.ens_state.items():
                if self.models is not None:
                    memo = {}
                    cols = []
                    for mt, ts in zip(self.models, self.trial_seeds):
                        if (id(mt), ts) not in memo:
                            memo[(id(mt), ts)] = mt.initial_voltage(ens, ts)
                        cols.append(memo[(id(mt), ts)])
                    v0[row0:row0 + n, :self.n_trials] = np.stack(cols).T
                else:
                    v0[row0:row0 + n, :self.n_trials] = m.initial_voltages(ens, self.trial_seeds).T
                if self.B > self.n_trials:
                    v0[row0:row0 + n, self.n_trials:] = v0[row0:row0 + n, :1]
            self._upload("st", 0, v0)   # packed LIF state: s >= 0 is the voltage of a non-refractory neuron
        for ens, (row0, n, dims) in plan.learned_enc.items():
            if self.models is not None:
                e = np.stack([mt.params[ens].scaled_encoders.reshape(-1) for mt in self.models])
                self._upload("lenc", row0, self._rows(e))
            else:
                self._upload("lenc", row0, self._rows(m.params[ens].scaled_encoders.reshape(-1)))
        for conn in plan.learned_dec:
            if self.models is not None:
                self._upload_decoders(conn, np.stack([np.asarray(mt.params[conn].weights, dtype=np.float32)
                                                      for mt in self.models]))
            else:
                w = np.asarray(m.params[conn].weights, dtype=np.float32)
                self._upload_decoders(conn, np.broadcast_to(w[None], (self.n_trials,) + w.shape))
        # per "rows" probe: list of [samples, size, n_trials] float32 chunks, already decimated by the probe's period
        self._probe_rows = {info.probe: [] for info in plan.probes if info.kind == "rows"}
        self._probe_steps_flushed = 0
        self._snap = {info.probe: [] for info in plan.probes if info.kind != "rows"}
        self._table_cache = {}

    # ------------------------------------------------------------------ nengo surface
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        """Release device memory; probe data stays readable (run_slam.py:242-252)."""
        if not self.closed:
            self.closed = True
            if self._h:
                self._launches_at_close = int(self._lib.ssb_total_launches(self._h))
                self._lib.ssb_sync(self._h)
                self._lib.ssb_destroy(self._h)
                self._h = None
            self._flush_probes()
            if self._tab_buf_ is not None:
                self._tab_buf_.free()
            for b in self._probe_bufs:
                b.free()
            if getattr(self, "_staged", None) is not None:
                self._staged.free()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_steps(self):
        return self._n_steps

    @property
    def time(self):
        return self._n_steps * self.dt

    @property
    def data(self):
        return _SimData(self)

    def trange(self, sample_every=None, dt=None):
        period = 1 if sample_every is None else int(round(sample_every / self.dt))
        n = self._n_steps // period
        return self.dt * period * np.arange(1, n + 1)

    def run(self, time_in_seconds, progress_bar=None):
        if time_in_seconds < 0:
            raise ValueError("run time must be positive")
        self.run_steps(int(np.round(float(time_in_seconds) / self.dt)))

    def step(self):
        self.run_steps(1)

    def reset(self, seed=None):
        """Back to step 0 with the same built model, start voltages and ``trial_seeds`` (nengo's ``reset(seed)`` would
        rebuild the model with another seed; that is refused rather than ignored)."""
        self._check_open()
        if seed is not None:
            raise NotImplementedError("reset(seed=...) would rebuild the model; construct a new Simulator instead")
        cabi.check(self._lib.ssb_reset(self._h), "ssb_reset")
        self._probe_pending = None
        self._n_steps = 0
        self._init_state()

    def _check_open(self):
        if self.closed:
            raise ns.exceptions.SimulatorClosed("Simulator is closed")

    # ------------------------------------------------------------------ input tables
    def _node_table(self, node, step0, n):
        """Values of a ``t``-only node for steps step0+1 .. step0+n -> [n_trials|1, n, size]."""
        if node in self._trial_inputs:
            arr = np.asarray(self._trial_inputs[node])
            if step0 + n > arr.shape[1]:
                raise ValueError(f"trial_inputs[{node!r}] holds {arr.shape[1]} steps, need {step0 + n}")
            return arr[:, step0:step0 + n, :]
        out = np.empty((1, n, node.size_out), dtype=np.float64)
        fn = node.output
        for i in range(n):
            t = (step0 + i + 1) * self.dt   # nengo: time = step * dt after the increment (App. A.1)
            out[0, i] = np.asarray(fn(t), dtype=np.float64).reshape(-1)
        if not np.all(np.isfinite(out)):
            raise ns.exceptions.SimulationError(f"{node!r} returned a non-finite value")
        return out

    def stage_inputs(self, step0, n_steps):
        """Pre-pack the input tables of steps [step0, step0+n_steps) into page-locked host memory in
        the device layout ``[step][row][trial]``.  ``run_steps`` then copies straight from this
        staging area (host -> device per chunk) instead of transposing NumPy tables chunk by chunk."""
        self._check_open()
        if self._nt == 0:
            return
        n_steps = int(n_steps)
        if getattr(self, "_staged", None) is not None:
            self._staged.free()
        self._staged = cabi.PinnedBuffer(n_steps * self._nt * self.B)
        self._staged_step0 = int(step0)
        self._staged_steps = n_steps
        view = self._staged.array.reshape(n_steps, self._nt, self.B)
        blk = 256
        for s0 in range(0, n_steps, blk):
            n = min(blk, n_steps - s0)
            self._fill_tables(step0 + s0, n, view[s0:s0 + n])

    def _staged_ptr(self, step0, n):
        st = getattr(self, "_staged", None)
        if st is None or st.ptr is None:
            return None
        if step0 < self._staged_step0 or step0 + n > self._staged_step0 + self._staged_steps:
            return None
        return st.ptr + (step0 - self._staged_step0) * self._nt * self.B * 4

    def _fill_tables(self, step0, n, buf=None):
        if self._nt == 0:
            return
        if buf is None:
            buf = self._tab_buf.array[:n * self._nt * self.B].reshape(n, self._nt, self.B)
        for node, col0, size in self.plan.tables:
            vals = self._node_table(node, step0, n)              # [T|1, n, size]
            block = np.transpose(vals, (1, 2, 0))                 # [n, size, T|1]
            buf[:, col0:col0 + size, :self.n_trials] = block
            if self.B > self.n_trials:
                buf[:, col0:col0 + size, self.n_trials:] = 0.0

    # ------------------------------------------------------------------ stepping
    def run_steps(self, n_steps, progress_bar=None):
        self._check_open()
        n_steps = int(n_steps)
        lib = self._lib
        periods = [info.period for info in self.plan.probes if info.kind != "rows"]
        done = 0
        while done < n_steps:
            n = min(self.chunk_steps, n_steps - done)
            for p in periods:  # stop exactly on snapshot steps of weight / encoder probes
                to_next = p - (self._n_steps % p)
                n = min(n, to_next)
            if self._synth is not None:                    # inputs are evaluated on the device: 12 bytes per step
                self._synth_steps(self._n_steps, n)
                src = None
            else:
                src = self._staged_ptr(self._n_steps, n)
                if src is None:
                    self._fill_tables(self._n_steps, n)
                    src = self._tab_buf.ptr
            # one pipelined call: tables host -> device, steps, probe rows device -> host (16-step sub-chunks overlap)
            pbuf = self._probe_bufs[self._probe_cur]
            if self._probe_pending is not None and self._probe_pending[0] == self._probe_cur:
                self._flush_probes()
            rc = lib.ssb_run_steps_io(self._h, src, n, pbuf.ptr if self._np else None)
            cabi.check(rc, "ssb_run_steps_io")
            if self._np:
                self._flush_probes()                       # the previous chunk (other buffer), while the device works
            cabi.check(lib.ssb_io_wait(self._h), "ssb_io_wait")
            if self._np:
                self._probe_pending = (self._probe_cur, n)
                self._probe_cur ^= 1
            self._n_steps += n
            done += n
            for info in self.plan.probes:
                if info.kind != "rows" and self._n_steps % info.period == 0:
                    self._snap[info.probe].append(self._snapshot(info))

    def _flush_probes(self):
        """Copy the rows of the last finished chunk out of its page-locked buffer."""
        if self._probe_pending is None:
            return
        idx, n = self._probe_pending
        self._probe_pending = None
        buf = self._probe_bufs[idx]
        if buf.array is None:
            return
        chunk = buf.array[:n * self._np * self.B].reshape(n, self._np, self.B)
        step0 = self._probe_steps_flushed                  # steps already copied out: this chunk holds step0+1 ..
        for probe, rows in self._probe_rows.items():
            info = self._probe_infos[probe]
            first = (-(step0 + 1)) % info.period            # offset of the first step with (step % period) == 0
            part = chunk[first::info.period, info.row0:info.row0 + info.size, :self.n_trials]
            if not self.keep_probe_history:
                rows.clear()
            if part.shape[0]:
                rows.append(part.copy())
        self._probe_steps_flushed += n

    def _snapshot(self, info):
        if info.kind == "weights":
            return self._download_decoders(info.conn).astype(np.float64)
        row0, n, dims = self.plan.learned_enc[info.ens]
        e = self._download("lenc", row0, n * dims)[:, :self.n_trials]
        return e.T.reshape(self.n_trials, n, dims).astype(np.float64)

    def _probe_array(self, probe):
        info = self._probe_infos[probe]
        if info.kind == "rows":
            self._flush_probes()
            rows = self._probe_rows[probe]
            if rows:
                data = np.concatenate(rows, axis=0) if len(rows) > 1 else rows[0]
                self._probe_rows[probe] = [data]
            else:
                data = np.zeros((0, info.size, self.n_trials), dtype=np.float32)
            data = np.transpose(data, (2, 0, 1)).astype(np.float64)   # [trial, sample, size]
        else:
            snaps = self._snap[probe]
            if snaps:
                data = np.stack(snaps, axis=1)                        # [trial, sample, ...]
            else:
                shape = (self.n_trials, 0)
                data = np.zeros(shape)
        return data if self._batched else data[0]

    # ------------------------------------------------------------------ checker / bench conveniences
    def neuron_state(self, ens):
        """(voltage, refractory_time) arrays [n_trials, n_neurons] of an ensemble, unpacked from the
        one-word device state (s >= 0: voltage, not refractory; s < 0: refractory for -s more seconds)."""
        row0, n = self.plan.ens_state[ens]
        s = self._download("st", row0, n)[:, :self.n_trials].T
        return np.maximum(s, 0.0), np.maximum(-s, 0.0)

    def activities(self, ens):
        """Last-step output of a wide ensemble [n_trials, n_neurons] (0 or 1/dt when spiking)."""
        row0 = self.plan.ens_act[ens]
        return self._download("act", row0, ens.n_neurons)[:, :self.n_trials].T

    def learned_encoders(self, ens):
        row0, n, dims = self.plan.learned_enc[ens]
        return self._download("lenc", row0, n * dims)[:, :self.n_trials].T.reshape(self.n_trials, n, dims)

    def learned_decoders(self, conn):
        return self._download_decoders(conn)

    def cleanup_indices(self):
        """Last grid clean-up argmax per clean-up node: int array [n_nodes, n_trials]."""
        n = len(getattr(self.plan, "cleanup_nodes", []))
        if n == 0:
            return np.zeros((0, self.n_trials), dtype=np.int64)
        raw = self._download("cidx", 0, n)
        return raw.view(np.int32)[:, :self.n_trials].astype(np.int64)

    def filter_state(self, key):
        """Current Lowpass state of a filtered connection / probe: [n_trials, size]."""
        f0, size = self.plan.filters[key]
        nf = int(self.plan.scalars["nf"])
        half = nf if (self._n_steps & 1) else 0   # values the *next* step will read
        return self._download("vec", 1 + half + f0, size)[:, :self.n_trials].T

    def set_profiling(self, on=True, timeline=False):
        """Per-launch CUDA events.  Plain profiling serialises the launches on one stream (solo kernel times);
        ``timeline=True`` keeps the dependency streams, so ``timeline()`` shows what really overlaps."""
        cabi.check(self._lib.ssb_set_profiling(self._h, 2 if (on and timeline) else int(bool(on))), "ssb_set_profiling")

    def timeline(self):
        """[(kind, start_us, end_us)] of every launch since ``set_profiling(True, timeline=True)``, in launch order."""
        import ctypes as C
        n = C.c_int()
        cabi.check(self._lib.ssb_timeline(self._h, None, None, None, 0, C.byref(n)), "ssb_timeline")
        a, b, k = (C.c_float * n.value)(), (C.c_float * n.value)(), (C.c_int * n.value)()
        cabi.check(self._lib.ssb_timeline(self._h, a, b, k, n.value, C.byref(n)), "ssb_timeline")
        names = cabi.KERNEL_KINDS
        return [(names[k[i]] if k[i] < len(names) else str(k[i]), a[i] * 1e3, b[i] * 1e3) for i in range(n.value)]

    def kernel_times(self):
        import ctypes as C
        n = len(cabi.KERNEL_KINDS)
        ms = (C.c_float * n)()
        cnt = (C.c_longlong * n)()
        cabi.check(self._lib.ssb_kernel_times(self._h, ms, cnt, n), "ssb_kernel_times")
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(cabi.KERNEL_KINDS)}

    def last_run_ms(self):
        import ctypes as C
        ms = C.c_float()
        cabi.check(self._lib.ssb_last_run_ms(self._h, C.byref(ms)), "ssb_last_run_ms")
        return float(ms.value)

    def mark(self, slot):
        """Record a CUDA event on the library stream (bench timing; 4 slots)."""
        cabi.check(self._lib.ssb_mark(self._h, int(slot)), "ssb_mark")

    def mark_elapsed_ms(self, a, b):
        import ctypes as C
        ms = C.c_float()
        cabi.check(self._lib.ssb_mark_elapsed_ms(self._h, int(a), int(b), C.byref(ms)), "ssb_mark_elapsed_ms")
        return float(ms.value)

    def run_resident(self, n_steps):
        """Advance ``n_steps`` using input tables that are already resident on the device
        (``load_tables``); asynchronous, probes stay in the device probe buffer."""
        self._check_open()
        cabi.check(self._lib.ssb_run_steps(self._h, int(n_steps)), "ssb_run_steps")
        self._n_steps += int(n_steps)

    def load_tables(self, step0, n_steps):
        """Copy the staged input tables of steps [step0, step0+n_steps) to the device."""
        self._check_open()
        src = self._staged_ptr(step0, n_steps)
        if src is None:
            raise ValueError("load_tables: range not staged (call stage_inputs first)")
        cabi.check(self._lib.ssb_set_tables(self._h, src, int(step0), int(n_steps)), "ssb_set_tables")

    def sync(self):
        cabi.check(self._lib.ssb_sync(self._h), "ssb_sync")

    def total_launches(self):
        if self._h is None:
            return getattr(self, "_launches_at_close", 0)
        return int(self._lib.ssb_total_launches(self._h))
