"""ORACLE (test infrastructure, never shipped or measured as the product).

CPU restatement of the *nengo reference simulator* for the SSP-SLAM hot path.

The arithmetic of the reference's per-timestep path is not in ``/root/reference``: it
lives in the third-party package **nengo** (un-pinned in ``setup.py:21-26``; API use pins
it to >= 3.1, latest upstream 4.0.0), which is absent from this image and cannot be
installed (no network).  This file restates nengo's published operator semantics
(``nengo/builder/{operator,neurons,processes,learning_rules,
connection,ensemble,node,probe}.py``) as a plain NumPy operator-by-operator stepper and
anchors on the reference's own call sites:

* ``experiments/run_slam.py:198-233``  ``sim = nengo.Simulator(model); with sim: sim.run(T)``
* ``experiments/run_slam.py:243,250``  ``sim.trange()``, ``sim.data[probe]``
* ``experiments/run_pathint.py:147-163``, ``run_slamview.py:148-158``

**Parity unpinned**: the reference ships no tests, fixtures or golden vectors for this
path and real nengo cannot be run here, so this oracle is pinned only by
the analytic known-answer tests in ``tests/`` and by the reference's
own NumPy helpers executed unmodified (``tests/golden``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.

Semantics restated (file:line are into upstream nengo, quoted from memory):
  * step order = topological sort of per-signal  set -> inc -> read -> update
    (``nengo/builder/operator.py``; ``nengo/simulator.py:Simulator.step``)
  * ``TimeUpdate``: ``step += 1; t = step*dt``
  * LIF / LIFRate / RectifiedLinear steps (``nengo/neurons.py``)
  * ``Lowpass`` zero-order-hold, applied as an *update* (one-step delay)
  * ``SimPES`` / ``SimVoja`` deltas applied by ``Copy(inc)`` at the next step
  * probes sample after the operators of a step
"""
from __future__ import annotations

import numpy as np

from sspslam_b200 import nengo_shim as ns
from oracle.nengo_ref_build import build as ref_build


# ----------------------------------------------------------------------------- signals
class Sig:
    """A NumPy buffer or a view of one (``base`` is the owning buffer)."""

    def __init__(self, value=None, shape=None, name="", base=None, dtype=np.float64):
        if base is None:
            if value is None:
                value = np.zeros(shape, dtype=dtype)
            self.a = np.array(value, dtype=dtype)
            self.base = self
        else:
            self.a = value  # numpy view
            self.base = base.base
        self.name = name

    def __getitem__(self, key):
        if isinstance(key, (int, np.integer)):
            k = int(key) % self.a.shape[0]
            view = self.a[k:k + 1]
        else:
            view = self.a[key]
        if view.base is None and view is not self.a and view.size:
            raise ValueError(f"advanced indexing of signal {self.name} would copy")
        return Sig(view, base=self, name=f"{self.name}[{key}]")

    @property
    def shape(self):
        return self.a.shape


# ----------------------------------------------------------------------------- operators
class Op:
    sets = incs = reads = updates = ()

    def step(self):
        raise NotImplementedError


class TimeUpdate(Op):
    def __init__(self, step, time, dt):
        self.stepsig, self.time, self.dt = step, time, dt
        self.sets = (step, time)

    def step(self):
        self.stepsig.a[...] += 1
        self.time.a[...] = self.stepsig.a * self.dt


class Reset(Op):
    def __init__(self, dst, value=0.0):
        self.dst, self.value = dst, value
        self.sets = (dst,)

    def step(self):
        self.dst.a[...] = self.value


class Copy(Op):
    def __init__(self, src, dst, inc=False):
        self.src, self.dst, self.inc = src, dst, inc
        self.reads = (src,)
        if inc:
            self.incs = (dst,)
        else:
            self.sets = (dst,)

    def step(self):
        if self.inc:
            self.dst.a[...] += self.src.a
        else:
            self.dst.a[...] = self.src.a


class DotInc(Op):
    def __init__(self, A, X, Y):
        self.A, self.X, self.Y = A, X, Y
        self.incs = (Y,)
        self.reads = (A, X)

    def step(self):
        self.Y.a[...] += self.A.a.dot(self.X.a)


class ElementwiseInc(Op):
    def __init__(self, A, X, Y):
        self.A, self.X, self.Y = A, X, Y
        self.incs = (Y,)
        self.reads = (A, X)

    def step(self):
        self.Y.a[...] += self.A.a * self.X.a


class SimPyFunc(Op):
    def __init__(self, output, fn, t, x):
        self.output, self.fn, self.t, self.x = output, fn, t, x
        self.sets = (output,)
        self.reads = (t,) if x is None else (t, x)

    def step(self):
        t = float(self.t.a)
        y = self.fn(t) if self.x is None else self.fn(t, self.x.a.astype(np.float64).copy())
        y = np.asarray(y, dtype=np.float64)
        if not np.all(np.isfinite(y)):
            raise FloatingPointError(f"node function returned non-finite value at t={t}")
        self.output.a[...] = y.reshape(self.output.a.shape)


class SimNeurons(Op):
    def __init__(self, neuron_type, J, output, state, dt):
        self.nt, self.J, self.output, self.state, self.dt = neuron_type, J, output, state, dt
        self.sets = (output,) + tuple(state.values())
        self.reads = (J,)

    def step(self):
        nt, dt, J, out = self.nt, self.dt, self.J.a, self.output.a
        dtype = J.dtype.type
        if isinstance(nt, ns.LIF):
            v, ref = self.state["voltage"].a, self.state["refractory_time"].a
            tau_rc = dtype(nt.tau_rc)
            dtt = dtype(dt)
            ref -= dtt
            delta_t = np.clip(dtt - ref, dtype(0), dtt)
            v -= (J - v) * np.expm1(-delta_t / tau_rc)
            spiked = v > 1
            out[...] = spiked * dtype(nt.amplitude / dt)
            t_spike = dtt + tau_rc * np.log1p(-(v[spiked] - 1) / (J[spiked] - 1))
            v[v < nt.min_voltage] = nt.min_voltage
            v[spiked] = 0
            ref[spiked] = dtype(nt.tau_ref) + t_spike
        elif isinstance(nt, ns.LIFRate):
            j = J - 1
            out[...] = 0
            pos = j > 0
            out[pos] = dtype(nt.amplitude) / (dtype(nt.tau_ref) + dtype(nt.tau_rc) * np.log1p(1.0 / j[pos]))
        elif isinstance(nt, ns.RectifiedLinear):
            out[...] = dtype(nt.amplitude) * np.maximum(J, 0)
        else:
            raise NotImplementedError(type(nt).__name__)


class SimLowpass(Op):
    """``SimProcess(Lowpass(tau), mode='update')``: y <- a*y + (1-a)*u (App. A.9)."""

    def __init__(self, tau, dt, inp, out):
        self.inp, self.out = inp, out
        a = np.exp(-dt / tau) if tau > 0 else 0.0     # Lowpass(0): no state, the update-after-read delay remains
        self.a = out.a.dtype.type(a)
        self.b = out.a.dtype.type(1.0 - a)
        self.reads = (inp,)
        self.updates = (out,)

    def step(self):
        self.out.a[...] *= self.a
        self.out.a[...] += self.b * self.inp.a


class SimPES(Op):
    def __init__(self, pre_filtered, error, delta, learning_rate, dt):
        self.pre, self.error, self.delta = pre_filtered, error, delta
        self.alpha = delta.a.dtype.type(-learning_rate * dt / pre_filtered.a.shape[0])
        self.reads = (pre_filtered, error)
        self.updates = (delta,)

    def step(self):
        np.outer(self.alpha * self.error.a, self.pre.a, out=self.delta.a)


class SimVoja(Op):
    def __init__(self, pre_decoded, post_filtered, scaled_encoders, delta, scale, learning, learning_rate, dt):
        self.pre, self.post, self.enc, self.delta = pre_decoded, post_filtered, scaled_encoders, delta
        self.scale = scale.astype(delta.a.dtype)[:, None]
        self.learning = learning
        self.alpha = delta.a.dtype.type(learning_rate * dt)
        self.reads = (pre_decoded, post_filtered, scaled_encoders, learning)
        self.updates = (delta,)

    def step(self):
        post = self.post.a
        self.delta.a[...] = self.alpha * self.learning.a * (
            self.scale * np.outer(post, self.pre.a) - post[:, None] * self.enc.a)


# ----------------------------------------------------------------------------- toposort
def _toposort(ops):
    sets, incs, reads, ups = {}, {}, {}, {}
    for i, op in enumerate(ops):
        for table, sigs in ((sets, op.sets), (incs, op.incs), (reads, op.reads), (ups, op.updates)):
            for s in sigs:
                table.setdefault(id(s.base), []).append(i)
    succ = [set() for _ in ops]
    npred = [0] * len(ops)

    def edges(pre, post):
        for p in pre:
            for q in post:
                if p != q and q not in succ[p]:
                    succ[p].add(q)
                    npred[q] += 1

    for b in set(sets) | set(incs) | set(reads) | set(ups):
        s, i, r, u = sets.get(b, []), incs.get(b, []), reads.get(b, []), ups.get(b, [])
        edges(s, i)
        edges(s + i, r)
        edges(s + i + r, u)
    ready = [i for i, n in enumerate(npred) if n == 0]
    order = []
    while ready:
        i = ready.pop(0)
        order.append(i)
        for q in sorted(succ[i]):
            npred[q] -= 1
            if npred[q] == 0:
                ready.append(q)
    if len(order) != len(ops):
        raise RuntimeError("operator graph has a cycle (algebraic loop without a synapse)")
    return [ops[i] for i in order]


# ----------------------------------------------------------------------------- simulator
class RefSimulator:
    """``nengo.Simulator``-shaped reference stepper (single trial, CPU, NumPy)."""

    def __init__(self, network, dt=0.001, model=None, dtype=np.float64, trial_seed=None, node_tables=None):
        """``model=None``: the network is built by the oracle's OWN builder (``oracle/nengo_ref_build.py``).  The parity
        tests pass the product's built model instead so that both sides step the same numbers (the two builders are
        compared with each other in ``tests/test_oracle_build.py``)."""
        self.network = network
        self.dt = float(dt)
        self.model = model if model is not None else ref_build(network, dt)
        self.dtype = np.dtype(dtype)
        self.trial_seed = trial_seed
        self.node_tables = node_tables or {}  # node -> array [n_steps, size_out] replacing its callable
        self.closed = False
        self._build_ops()
        self.reset()

    # -- nengo.Simulator surface
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        self.closed = True

    @property
    def n_steps(self):
        return int(self.sig_step.a)

    @property
    def time(self):
        return float(self.sig_time.a)

    def trange(self, sample_every=None):
        period = 1 if sample_every is None else int(round(sample_every / self.dt))
        n = self.n_steps // period
        return self.dt * period * np.arange(1, n + 1)

    def run(self, t):
        self.run_steps(int(np.round(float(t) / self.dt)))

    def run_steps(self, n):
        for _ in range(int(n)):
            self.step()

    def step(self):
        for op in self.ops:
            op.step()
        n = self.n_steps
        for probe, (sig, period) in self._probe_sigs.items():
            if n % period < 1:
                self._probe_data[probe].append(sig.a.astype(np.float64).copy())

    @property
    def data(self):
        return _Data(self)

    # -- construction
    def _new(self, shape=None, value=None, name=""):
        return Sig(value=value, shape=shape, name=name, dtype=self.dtype)

    def _build_ops(self):
        net, model, dt = self.network, self.model, self.dt
        ops = []
        S = {}  # (obj, key) -> Sig
        self.sig_step = Sig(value=0, name="step", dtype=np.int64)
        self.sig_time = Sig(value=0.0, name="time", dtype=np.float64)
        ops.append(TimeUpdate(self.sig_step, self.sig_time, dt))
        self._tables = []

        # ---- nodes (App. A.5)
        for node in net.all_nodes:
            out = node.output
            if node in self.node_tables:
                tab = np.asarray(self.node_tables[node], dtype=np.float64)
                sig = self._new(shape=node.size_out, name=f"{node}.out")
                S[node, "out"] = sig
                ops.append(SimPyFunc(sig, _TableFn(tab, dt), self.sig_time, None))
            elif out is None:
                sig = self._new(shape=node.size_in, name=f"{node}.in")
                S[node, "in"] = S[node, "out"] = sig
                ops.append(Reset(sig))
            elif callable(out):
                osig = self._new(shape=node.size_out, name=f"{node}.out")
                S[node, "out"] = osig
                x = None
                if node.size_in > 0:
                    x = self._new(shape=node.size_in, name=f"{node}.in")
                    S[node, "in"] = x
                    ops.append(Reset(x))
                ops.append(SimPyFunc(osig, out, self.sig_time, x))
            else:
                S[node, "out"] = self._new(value=out, name=f"{node}.out")

        # ---- ensembles (App. A.3)
        self._neuron_state = {}
        for ens in net.all_ensembles:
            p = model.params[ens]
            n = ens.n_neurons
            S[ens, "in"] = self._new(shape=ens.dimensions, name=f"{ens}.in")
            ops.append(Reset(S[ens, "in"]))
            S[ens, "encoders"] = self._new(value=p.scaled_encoders, name=f"{ens}.encoders")
            bias = self._new(value=p.bias, name=f"{ens}.bias")
            J = self._new(shape=n, name=f"{ens}.J")
            out = self._new(shape=n, name=f"{ens}.neurons.out")
            S[ens.neurons, "in"], S[ens.neurons, "out"], S[ens, "out"] = J, out, out
            ops.append(Copy(bias, J))
            ops.append(DotInc(S[ens, "encoders"], S[ens, "in"], J))
            state = {}
            if isinstance(ens.neuron_type, ns.LIF):
                state["voltage"] = self._new(value=model.initial_voltage(ens, self.trial_seed), name=f"{ens}.v")
                state["refractory_time"] = self._new(shape=n, name=f"{ens}.ref")
            self._neuron_state[ens] = state
            ops.append(SimNeurons(ens.neuron_type, J, out, state, dt))

        # ---- learning-rule input signals must exist before connections target them
        for conn in net.all_connections:
            rule = conn.learning_rule
            if rule is None:
                continue
            lrt = rule.learning_rule_type
            if isinstance(lrt, ns.PES):
                S[rule, "in"] = self._new(shape=rule.size_in, name="PES:error")
                ops.append(Reset(S[rule, "in"]))
            elif isinstance(lrt, ns.Voja):
                S[rule, "in"] = self._new(shape=1, name="Voja:learning")
                ops.append(Reset(S[rule, "in"], value=1.0))
            else:
                raise NotImplementedError(type(lrt).__name__)

        # ---- probes get an input signal fed by an implicit connection (App. A.13)
        self._probe_sigs, self._probe_data = {}, {}
        implicit = []
        for probe in net.all_probes:
            period = 1 if probe.sample_every is None else probe.sample_every / dt
            self._probe_data[probe] = []
            if isinstance(probe.obj, (ns.Node, ns.Ensemble, ns.Neurons)):
                sig = self._new(shape=probe.size_in, name=f"{probe}.in")
                ops.append(Reset(sig))
                S[probe, "in"] = sig
                implicit.append(probe)
                self._probe_sigs[probe] = (sig, period)
            else:
                self._probe_sigs[probe] = (None, period)  # bound after connections exist

        # ---- connections (App. A.6)
        self._weights = {}
        for conn in net.all_connections:
            self._build_connection(conn, S, ops)
        for probe in implicit:
            self._build_probe_connection(probe, S, ops)
        for probe, (sig, period) in list(self._probe_sigs.items()):
            if sig is None:
                obj = probe.obj
                if isinstance(obj, ns.Connection) and probe.attr == "weights":
                    self._probe_sigs[probe] = (self._weights[obj], period)
                elif isinstance(obj, ns.LearningRule) and probe.attr == "scaled_encoders":
                    self._probe_sigs[probe] = (S[obj.connection.post_obj, "encoders"], period)
                else:
                    raise NotImplementedError(f"probe {probe!r}")

        self.signals = S
        self.ops = _toposort(ops)

    def _post_signal(self, conn, S):
        post = conn.post_obj
        sig = S[post, "in"]
        if conn.post_slice == slice(None):
            return sig
        return sig[conn.post_slice]

    def _weighted(self, in_sig, weights, size_out, ops, name):
        """Reset + (DotInc | ElementwiseInc) exactly as ``build_dense`` does."""
        wsig = self._new(value=weights, name=f"{name}.weights")
        weighted = self._new(shape=size_out, name=f"{name}.weighted")
        ops.append(Reset(weighted))
        ops.append((ElementwiseInc if wsig.a.ndim < 2 else DotInc)(wsig, in_sig, weighted))
        return weighted, wsig

    def _build_connection(self, conn, S, ops):
        model, dt = self.model, self.dt
        pre = conn.pre_obj
        name = repr(conn)
        if isinstance(pre, ns.Ensemble):
            in_sig = S[pre, "out"]
            weighted, wsig = self._weighted(in_sig, model.params[conn].weights, conn.size_out, ops, name)
        else:
            in_sig = S[pre, "out"]
            if conn.pre_slice != slice(None):
                in_sig = in_sig[conn.pre_slice]
            if conn.transform is None:
                weighted, wsig = in_sig, None
            else:
                weighted, wsig = self._weighted(in_sig, conn.transform, conn.size_out, ops, name)
        self._weights[conn] = wsig
        if conn.synapse is not None:
            filtered = self._new(shape=weighted.shape, name=f"{name}.filtered")
            ops.append(SimLowpass(conn.synapse.tau, dt, weighted, filtered))
            weighted = filtered

        post = conn.post_obj
        if isinstance(post, ns.Neurons):
            gains = self._new(value=model.params[post.ensemble].gain[conn.post_slice], name=f"{name}.gains")
            ops.append(ElementwiseInc(gains, weighted, self._post_signal(conn, S)))
        else:
            ops.append(Copy(weighted, self._post_signal(conn, S), inc=True))

        rule = conn.learning_rule
        if rule is not None:
            lrt = rule.learning_rule_type
            if isinstance(lrt, ns.PES):
                if wsig is None or wsig.a.ndim != 2:
                    raise NotImplementedError("PES needs a dense decoder matrix")
                delta = self._new(shape=wsig.shape, name="PES:delta")
                ops.append(Copy(delta, wsig, inc=True))
                acts = S[pre, "out"]
                if lrt.pre_synapse is not None:
                    filt = self._new(shape=acts.shape, name="PES:pre_filtered")
                    ops.append(SimLowpass(lrt.pre_synapse.tau, dt, acts, filt))
                    acts = filt
                ops.append(SimPES(acts, S[rule, "in"], delta, lrt.learning_rate, dt))
            elif isinstance(lrt, ns.Voja):
                post_ens = conn.post_obj
                enc = S[post_ens, "encoders"]
                delta = self._new(shape=enc.shape, name="Voja:delta")
                ops.append(Copy(delta, enc, inc=True))
                post_out = S[post_ens, "out"]
                if lrt.post_synapse is not None:
                    filt = self._new(shape=post_out.shape, name="Voja:post_filtered")
                    ops.append(SimLowpass(lrt.post_synapse.tau, dt, post_out, filt))
                    post_out = filt
                scale = model.params[post_ens].gain / post_ens.radius
                # nengo passes model.sig[conn]['out'] == the post ensemble's input signal as pre_decoded
                ops.append(SimVoja(S[post_ens, "in"], post_out, enc, delta, scale, S[rule, "in"], lrt.learning_rate, dt))

    def _build_probe_connection(self, probe, S, ops):
        obj = probe.obj
        if isinstance(obj, ns.Ensemble):
            weighted, _ = self._weighted(S[obj, "out"], self.model.probe_conns[probe], probe.size_in, ops, repr(probe))
        else:
            weighted = S[obj, "out"]
            if probe.slice != slice(None):
                weighted = weighted[probe.slice]
        if probe.synapse is not None:
            filtered = self._new(shape=weighted.shape, name=f"{probe}.filtered")
            ops.append(SimLowpass(probe.synapse.tau, self.dt, weighted, filtered))
            weighted = filtered
        ops.append(Copy(weighted, S[probe, "in"], inc=True))

    def reset(self, seed=None):
        if self.n_steps:
            self._build_ops()
        for k in self._probe_data:
            self._probe_data[k] = []

    # -- checker conveniences (not part of nengo's API)
    def spikes(self, ens):
        """Boolean spike mask of the last step."""
        return self.signals[ens, "out"].a > 0

    def voltage(self, ens):
        return self._neuron_state[ens]["voltage"].a

    def learned_weights(self, conn):
        return self._weights[conn].a

    def scaled_encoders(self, ens):
        return self.signals[ens, "encoders"].a


class _TableFn:
    def __init__(self, table, dt):
        self.table, self.dt = table, dt

    def __call__(self, t):
        n = int(round(t / self.dt))
        return self.table[n - 1]


class _Data:
    def __init__(self, sim):
        self.sim = sim

    def __getitem__(self, key):
        sim = self.sim
        if key in sim._probe_data:
            rows = sim._probe_data[key]
            if not rows:
                return np.zeros((0,) + tuple(np.shape(sim._probe_sigs[key][0].a)))
            return np.stack(rows)
        return sim.model.params[key]

    def __contains__(self, key):
        return key in self.sim._probe_data or key in self.sim.model.params
