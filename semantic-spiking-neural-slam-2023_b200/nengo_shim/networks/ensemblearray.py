"""``nengo.networks.EnsembleArray``: N equally sized ensembles
behind one pass-through input node, plus ``add_output`` fan-in nodes."""
import numpy as np

from ..core import Network, Node, Ensemble, Connection


class EnsembleArray(Network):
    def __init__(self, n_neurons, n_ensembles, ens_dimensions=1, label=None, seed=None,
                 add_to_container=None, **ens_kwargs):
        super().__init__(label=label, seed=seed, add_to_container=add_to_container)
        self.config[Ensemble].update(ens_kwargs)
        self.n_neurons_per_ensemble = int(n_neurons)
        self.n_ensembles = int(n_ensembles)
        self.dimensions_per_ensemble = int(ens_dimensions)
        self.ea_ensembles = []
        prefix = "" if label is None else label + "_"
        D = self.dimensions_per_ensemble
        with self:
            self.input = Node(size_in=self.dimensions, label="input")
            for i in range(self.n_ensembles):
                ens = Ensemble(self.n_neurons_per_ensemble, D, label=f"{prefix}{i}")
                Connection(self.input[i * D:(i + 1) * D], ens, synapse=None)
                self.ea_ensembles.append(ens)
        self.add_output("output", function=None)

    @property
    def dimensions(self):
        return self.n_ensembles * self.dimensions_per_ensemble

    def add_output(self, name, function, synapse=None, **conn_kwargs):
        D = self.dimensions_per_ensemble
        if function is None:
            width = D
        elif callable(function):
            width = int(np.asarray(function(np.zeros(D))).size)
        else:
            raise ValueError("add_output: function must be None or a callable")
        with self:
            out = Node(output=None, size_in=width * self.n_ensembles, label=name)
            setattr(self, name, out)
            for i, ens in enumerate(self.ea_ensembles):
                Connection(ens, out[i * width:(i + 1) * width], function=function,
                           synapse=synapse, **conn_kwargs)
        return out
