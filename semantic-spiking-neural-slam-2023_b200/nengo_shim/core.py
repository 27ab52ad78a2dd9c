"""Minimal nengo-compatible *declaration* layer (front-end objects only).

``nengo`` is an external, un-vendored dependency of the reference
(``/root/reference/setup.py:21-26``) and is not installable here, so the network
declarations in ``sspslam/networks/*.py`` cannot even be imported without it.
This module provides exactly the front-end surface those files exercise
(SURVEY.md App. A.16): ``Network`` (context stack + ``config``), ``Ensemble``,
``Node``, ``Connection``, ``Probe``, slicing views, ``Neurons`` and
``LearningRule`` handles.  It declares graphs; it never simulates anything.

Attribute names follow upstream nengo (``pre_obj``, ``post_slice``, ``size_mid``,
``learning_rule`` ...) so that the builder also accepts real nengo objects.
"""
from __future__ import annotations

import numpy as np

from . import dists as _dists
from .params import Default, is_default
from .neurons import LIF, NeuronType
from .synapses import Lowpass, Synapse
from .solvers import LstsqL2
from .exceptions import ValidationError, NetworkContextError


# --------------------------------------------------------------------------- config
class _ClassParams:
    """Per-class default overrides stored in a ``Config`` (``net.config[Ensemble]``)."""

    def __init__(self):
        object.__setattr__(self, "_values", {})

    def __setattr__(self, key, value):
        self._values[key] = value

    def __getattr__(self, key):
        try:
            return object.__getattribute__(self, "_values")[key]
        except KeyError:
            raise AttributeError(key)

    def update(self, mapping):
        self._values.update(mapping)

    def __contains__(self, key):
        return key in self._values


class Config:
    """``network.config`` — only per-class default overrides are supported."""

    context: list["Config"] = []

    def __init__(self):
        self._params: dict[type, _ClassParams] = {}

    def __getitem__(self, cls):
        return self._params.setdefault(cls, _ClassParams())

    @classmethod
    def lookup(cls, klass, name, fallback):
        """Innermost-first search of the config stack (nengo ``Config.default``)."""
        for cfg in reversed(cls.context):
            for k in klass.__mro__:
                p = cfg._params.get(k)
                if p is not None and name in p:
                    return getattr(p, name)
        return fallback


# --------------------------------------------------------------------------- views
def _norm_slice(key, size):
    """Return an index array for ``obj[key]`` over a vector of length ``size``."""
    idx = np.arange(size)[key]
    return np.atleast_1d(idx)


class ObjView:
    """``obj[a:b]`` — a sliced handle usable as Connection pre/post and Probe target."""

    def __init__(self, obj, key):
        self.obj = obj
        self.key = key
        self.size_in = len(_norm_slice(key, obj.size_in)) if obj.size_in else 0
        self.size_out = len(_norm_slice(key, obj.size_out)) if obj.size_out else 0

    @property
    def slice(self):
        return self.key

    def __len__(self):
        return max(self.size_in, self.size_out)

    def __repr__(self):
        return f"{self.obj!r}[{self.key}]"


class _Sliceable:
    def __getitem__(self, key):
        return ObjView(self, key)

    def __len__(self):
        return self.size_out


# --------------------------------------------------------------------------- network
class Network:
    """Container with a ``with`` context stack (nengo ``Network``)."""

    context: list["Network"] = []

    def __init__(self, label=None, seed=None, add_to_container=None):
        self.label = label
        self.seed = seed
        self.ensembles: list = []
        self.nodes: list = []
        self.connections: list = []
        self.networks: list = []
        self.probes: list = []
        self.objects = {
            Ensemble: self.ensembles,
            Node: self.nodes,
            Connection: self.connections,
            Network: self.networks,
            Probe: self.probes,
        }
        self._config = Config()
        if add_to_container is None:
            add_to_container = len(Network.context) > 0
        if add_to_container:
            Network.add(self)

    @staticmethod
    def add(obj):
        if not Network.context:
            raise NetworkContextError(
                f"'{obj}' must either be created inside a ``with network:`` block, "
                "or set add_to_container=False in the object's constructor.")
        net = Network.context[-1]
        for cls in type(obj).__mro__:
            if cls in net.objects:
                net.objects[cls].append(obj)
                return
        raise NetworkContextError(f"Objects of type {type(obj).__name__} cannot be added to networks.")

    @property
    def config(self):
        return self._config

    def __enter__(self):
        Network.context.append(self)
        Config.context.append(self._config)
        return self

    def __exit__(self, *exc):
        Config.context.pop()
        Network.context.pop()
        return False

    # flattened views, depth first in declaration order (nengo ``all_*``)
    def _all(self, attr):
        out = list(getattr(self, attr))
        for sub in self.networks:
            out.extend(sub._all(attr))
        return out

    all_ensembles = property(lambda self: self._all("ensembles"))
    all_nodes = property(lambda self: self._all("nodes"))
    all_connections = property(lambda self: self._all("connections"))
    all_probes = property(lambda self: self._all("probes"))

    @property
    def all_networks(self):
        out = []
        for sub in self.networks:
            out.append(sub)
            out.extend(sub.all_networks)
        return out

    @property
    def all_objects(self):
        return self.all_ensembles + self.all_nodes + self.all_connections + self.all_networks + self.all_probes

    @property
    def n_neurons(self):
        return sum(e.n_neurons for e in self.all_ensembles)

    def __repr__(self):
        return f"<Network {self.label!r}>"


# --------------------------------------------------------------------------- ensemble
class Neurons(_Sliceable):
    """``ens.neurons`` — direct access to the neuron currents / spikes."""

    def __init__(self, ensemble):
        self.ensemble = ensemble

    @property
    def size_in(self):
        return self.ensemble.n_neurons

    @property
    def size_out(self):
        return self.ensemble.n_neurons

    def __repr__(self):
        return f"<Neurons of {self.ensemble!r}>"


class Ensemble(_Sliceable):
    def __init__(self, n_neurons, dimensions, radius=Default, encoders=Default, intercepts=Default,
                 max_rates=Default, eval_points=Default, n_eval_points=Default, neuron_type=Default,
                 gain=Default, bias=Default, noise=Default, normalize_encoders=Default,
                 label=None, seed=None):
        def d(name, val, fallback):
            return Config.lookup(Ensemble, name, fallback) if is_default(val) else val

        self.n_neurons = int(n_neurons)
        self.dimensions = int(dimensions)
        if self.n_neurons <= 0 or self.dimensions <= 0:
            raise ValidationError("n_neurons and dimensions must be positive", "n_neurons", self)
        self.radius = float(d("radius", radius, 1.0))
        self.encoders = d("encoders", encoders, _dists.ScatteredHypersphere(surface=True))
        self.intercepts = d("intercepts", intercepts, _dists.Uniform(-1.0, 0.9))
        self.max_rates = d("max_rates", max_rates, _dists.Uniform(200, 400))
        self.eval_points = d("eval_points", eval_points, _dists.ScatteredHypersphere(surface=False))
        self.n_eval_points = d("n_eval_points", n_eval_points, None)
        self.neuron_type = d("neuron_type", neuron_type, LIF())
        self.gain = d("gain", gain, None)
        self.bias = d("bias", bias, None)
        self.noise = d("noise", noise, None)
        self.normalize_encoders = d("normalize_encoders", normalize_encoders, True)
        if not isinstance(self.encoders, _dists.Distribution):
            self.encoders = np.array(self.encoders, dtype=np.float64)
            if self.encoders.shape != (self.n_neurons, self.dimensions):
                raise ValidationError(
                    f"encoders shape {self.encoders.shape} != ({self.n_neurons}, {self.dimensions})",
                    "encoders", self)
        for name in ("intercepts", "max_rates"):
            v = getattr(self, name)
            if not isinstance(v, _dists.Distribution):
                v = np.array(v, dtype=np.float64)
                if v.shape != (self.n_neurons,):
                    raise ValidationError(f"{name} must have shape ({self.n_neurons},)", name, self)
                setattr(self, name, v)
        if not isinstance(self.neuron_type, NeuronType):
            raise ValidationError("neuron_type must be a NeuronType", "neuron_type", self)
        self.label = label
        self.seed = seed
        self._neurons = Neurons(self)
        Network.add(self)

    @property
    def neurons(self):
        return self._neurons

    @property
    def size_in(self):
        return self.dimensions

    @property
    def size_out(self):
        return self.dimensions

    def __repr__(self):
        return f"<Ensemble {self.label!r} {self.n_neurons}x{self.dimensions}>"


# --------------------------------------------------------------------------- node
class Node(_Sliceable):
    def __init__(self, output=None, size_in=None, size_out=None, label=None, seed=None):
        self.label = label
        self.seed = seed
        self.size_in = 0 if size_in is None else int(size_in)
        self._size_out_arg = size_out
        self.size_out = 0
        self._output = None
        self.output = output  # validates / infers size_out
        Network.add(self)

    @property
    def output(self):
        return self._output

    @output.setter
    def output(self, value):
        """Assignment after construction is allowed (``pathintegration.py:167``)."""
        if value is None:
            self._output = None
            self.size_out = self.size_in if self._size_out_arg is None else int(self._size_out_arg)
            return
        if callable(value):
            if self._size_out_arg is not None:
                self.size_out = int(self._size_out_arg)
            else:
                # nengo infers size_out by evaluating the function at t=0 (App. A.5)
                args = (0.0,) if self.size_in == 0 else (0.0, np.zeros(self.size_in))
                res = value(*args)
                self.size_out = 0 if res is None else int(np.asarray(res).size)
            self._output = value
            return
        arr = np.array(value, dtype=np.float64)
        if self.size_in != 0:
            raise ValidationError("constant-output nodes cannot have size_in", "output", self)
        if arr.ndim > 1:
            raise ValidationError("node output must be 0-D or 1-D", "output", self)
        self._output = arr.reshape(-1)
        self.size_out = self._output.size

    def __repr__(self):
        return f"<Node {self.label!r}>"


# --------------------------------------------------------------------------- learning rules
class LearningRule:
    """``conn.learning_rule`` — connectable handle (error / learning-signal input)."""

    def __init__(self, connection, learning_rule_type):
        self.connection = connection
        self.learning_rule_type = learning_rule_type

    @property
    def modifies(self):
        return self.learning_rule_type.modifies

    @property
    def size_in(self):
        lrt = self.learning_rule_type
        if lrt.size_in == "post_state":
            post = self.connection.post_obj
            return post.ensemble.dimensions if isinstance(post, Neurons) else post.size_in
        if lrt.size_in == "scalar":
            return 1
        return int(lrt.size_in)

    size_out = 0

    def __repr__(self):
        return f"<LearningRule {type(self.learning_rule_type).__name__} of {self.connection!r}>"


# --------------------------------------------------------------------------- connection
def _split(target):
    if isinstance(target, ObjView):
        return target.obj, target.key
    return target, slice(None)


class Connection:
    def __init__(self, pre, post, synapse=Default, function=Default, transform=Default,
                 solver=Default, learning_rule_type=Default, eval_points=Default,
                 scale_eval_points=Default, label=None, seed=None):
        self.pre, self.post = pre, post
        self.pre_obj, self.pre_slice = _split(pre)
        self.post_obj, self.post_slice = _split(post)
        if not isinstance(self.pre_obj, (Ensemble, Neurons, Node)):
            raise ValidationError(f"invalid connection pre {pre!r}", "pre", self)
        if not isinstance(self.post_obj, (Ensemble, Neurons, Node, LearningRule, Probe)):
            raise ValidationError(f"invalid connection post {post!r}", "post", self)

        syn = Lowpass(0.005) if is_default(synapse) else synapse
        if syn is not None and not isinstance(syn, Synapse):
            syn = Lowpass(float(syn))
        self.synapse = syn
        self.function = None if is_default(function) else function
        self.solver = LstsqL2() if is_default(solver) else solver
        self.learning_rule_type = None if is_default(learning_rule_type) else learning_rule_type
        self.eval_points = None if is_default(eval_points) else eval_points
        self.scale_eval_points = True if is_default(scale_eval_points) else scale_eval_points
        self.label = label
        self.seed = seed

        # sizes: pre(size_in) -> function(size_mid) -> transform(size_out)
        self.size_in = len(_norm_slice(self.pre_slice, self.pre_obj.size_out))
        if self.function is None:
            self.size_mid = self.size_in
        elif callable(self.function):
            if not isinstance(self.pre_obj, (Ensemble, Node)):
                raise ValidationError("function only allowed on Ensemble/Node pre", "function", self)
            self.size_mid = int(np.asarray(self.function(np.zeros(self.size_in))).size)
        else:
            self.function = np.asarray(self.function, dtype=np.float64)
            self.size_mid = self.function.shape[1]

        if is_default(transform) or transform is None:
            self.transform = None
            self.size_out = self.size_mid
        else:
            t = np.array(transform, dtype=np.float64)
            if t.ndim == 0:
                self.size_out = self.size_mid
            elif t.ndim == 1:
                if t.shape[0] != self.size_mid:
                    raise ValidationError("vector transform length mismatch", "transform", self)
                self.size_out = self.size_mid
            elif t.ndim == 2:
                if t.shape[1] != self.size_mid:
                    raise ValidationError(
                        f"transform columns ({t.shape[1]}) != pre size ({self.size_mid})", "transform", self)
                self.size_out = t.shape[0]
            else:
                raise ValidationError("transform must be scalar, vector or matrix", "transform", self)
            self.transform = t

        post_size = len(_norm_slice(self.post_slice, self.post_obj.size_in))
        if post_size != self.size_out:
            raise ValidationError(
                f"connection output size ({self.size_out}) != post size ({post_size}) for {pre!r}->{post!r}",
                "transform", self)

        self._learning_rule = None
        if self.learning_rule_type is not None:
            self._learning_rule = LearningRule(self, self.learning_rule_type)
        Network.add(self)

    @property
    def learning_rule(self):
        return self._learning_rule

    def __repr__(self):
        return f"<Connection {self.label or ''} {self.pre!r}->{self.post!r}>"


# --------------------------------------------------------------------------- probe
class Probe:
    def __init__(self, target, attr=None, sample_every=None, synapse=Default, solver=Default,
                 label=None, seed=None):
        self.target = target
        self.obj, self.slice = _split(target)
        if attr is None:
            if isinstance(self.obj, Ensemble):
                attr = "decoded_output"
            elif isinstance(self.obj, (Node, Neurons)):
                attr = "output"
            elif isinstance(self.obj, Connection):
                attr = "output"
            else:
                raise ValidationError(f"no default probe attribute for {target!r}", "attr", self)
        self.attr = attr
        self.sample_every = sample_every
        syn = None if is_default(synapse) else synapse
        if syn is not None and not isinstance(syn, Synapse):
            syn = Lowpass(float(syn))
        self.synapse = syn
        self.solver = LstsqL2() if is_default(solver) else solver
        self.label = label
        self.seed = seed
        Network.add(self)

    @property
    def size_in(self):
        if isinstance(self.obj, (Connection, LearningRule)):
            return 0
        return len(_norm_slice(self.slice, self.obj.size_out))

    size_out = 0

    def __repr__(self):
        return f"<Probe {self.attr} of {self.target!r}>"
