"""BASELINE configs[2] / configs[3] on a RECORDED path (``-m gpu``): ``run_slam.py --path-data`` semantics
(``run_slam.py:100-112``) applied to the committed down-sampled slice of ``example_paths/twoRooms_path.npy``; every
trial walks the same path, trial ``i`` has its own landmark set and start state, the static weights are shared.
Network sizes: configs[2] = ``run_slam_map_gif.py:36-47`` defaults (d = 55, pi 800, mem 1000, length scale 0.1);
configs[3] = ``run_slamview.py:19-30`` defaults (d = 97, pi 800, mem 970, 100 landmarks, length scale 0.3)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle import ssp_ref
from oracle.nengo_ref_sim import RefSimulator
from sspslam_b200 import scenarios

pytestmark = pytest.mark.gpu
FIXTURE = os.path.join(GOLDEN_DIR, "twoRooms_path_ds20.npy")

SIZES = {
    "configs[2] slam d55 pi800 mem1000": dict(ssp_dim=55, pi_n_neurons=800, mem_n_neurons=1000, circonv_n_neurons=100,
                                               n_landmarks=50, length_scale=0.1, n_neurons=48800),
    "configs[3] slamview d97 pi800 mem970": dict(ssp_dim=97, pi_n_neurons=800, mem_n_neurons=970, circonv_n_neurons=100,
                                                  n_landmarks=100, length_scale=0.3, view=True, n_neurons=42110),
}


@pytest.fixture(scope="module", autouse=True)
def _built(lib):
    return lib


@pytest.mark.parametrize("name", list(SIZES))
def test_recorded_path_full_size_rate_mode_matches_oracle(name):
    from sspslam_b200.simulator import Simulator
    kw = dict(SIZES[name])
    n_neurons = kw.pop("n_neurons")
    n_steps, n_trials = 120, 32
    sc = scenarios.make_slam(n_trials=n_trials, n_steps=n_steps, neuron_type="lifrate", path_data=FIXTURE, data_dt=0.02,
                             view_rad=0.3, **kw)
    slam = sc.extra["slam"]
    with Simulator(sc.network, dt=sc.dt, n_trials=n_trials, trial_inputs=sc.trial_inputs) as sim:
        assert sim.plan.stats["n_neurons"] == n_neurons
        sim.run_steps(n_steps)
        idx = sim.cleanup_indices()[0].copy()
        dec = sim.learned_decoders(slam.assomemory.conn_out)
    got = sim.data[sc.probe]
    assert np.all(np.isfinite(got))
    assert np.array_equal(sc.paths[0], sc.paths[31])                                  # one recorded path for every trial
    lm = sc.extra["input_synthesis"]["landmarks"] if sc.extra["input_synthesis"] else None
    assert lm is None or not np.array_equal(lm[0], lm[31])
    for trial in (0, 31):
        tabs = {node: arr[trial] for node, arr in sc.trial_inputs.items()}
        ref = RefSimulator(sc.network, dt=sc.dt, model=sim.model, node_tables=tabs, trial_seed=sim.trial_seeds[trial])
        ref.run_steps(n_steps)
        want = ref.data[sc.probe]
        assert np.max(np.abs(want)) > 0.1
        assert np.max(np.abs(got[trial] - want)) < 1e-4 * np.max(np.abs(want))
        assert idx[trial] == ssp_ref.cleanup_index(slam.sample_ssps, ref.signals[slam.gridcells, "in"].a)
        want_dec = ref.learned_weights(slam.assomemory.conn_out)
        assert np.max(np.abs(dec[trial] - want_dec)) < 1e-4 * np.max(np.abs(want_dec)) + 1e-9
