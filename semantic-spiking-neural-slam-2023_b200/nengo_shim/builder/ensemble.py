"""``nengo.builder.ensemble.get_activities`` as used at ``experiments/run_slam.py:266``."""
import numpy as np


def get_activities(built_ens, ens, eval_points):
    x = np.dot(np.asarray(eval_points, dtype=np.float64), built_ens.encoders.T / ens.radius)
    return ens.neuron_type.rates(x, built_ens.gain, built_ens.bias)
