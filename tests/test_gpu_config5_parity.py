"""BASELINE configs[4] at FULL network size (``-m gpu``): 3-D domain, HexagonalSSPSpace n_rotates = n_scales = 9
(d = 649), pi 500 / mem 970 / circonv 100 -> 426 380 neurons per trial (``run_slam.py --domain-dim 3``,
``slam.py:182-307``).  Every d = 649 kernel is on the path: K-blocked tcgen05 grid scan, static wide encode and
column-tiled decode, the row program's dense DFT blocks on tcgen05 (k_lin_tck), the streaming Voja kernel, deferred PES.
The oracle steps one trial of the same built model in rate mode (1e-4 per north_star); the clean-up index is bit-exact.

The clean-up grid has 30 points per axis (the resolution the reference drivers decode 3-D runs on, run_slam.py:248-250);
the reference's hard-coded 100 per axis (slam.py:209) is a 2.6 GB operand - see DESIGN.md section 8."""
import numpy as np
import pytest

from oracle import ssp_ref
from oracle.nengo_ref_sim import RefSimulator
from sspslam_b200 import scenarios

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("voja", ["cta", "stream"])
def test_config5_full_size_rate_mode_matches_oracle(lib, voja, monkeypatch):
    """``voja``: the memory ensemble (970 x 649 per-trial encoders) on the CTA-cooperative kernel (default) or on the per-warp
    streaming kernel."""
    from sspslam_b200.simulator import Simulator
    monkeypatch.setenv("SSB_VOJA", voja)
    n_steps, n_trials = 24, 32
    sc = scenarios.make_slam(n_trials=n_trials, n_steps=n_steps + 4, ssp_dim=649, pi_n_neurons=500, mem_n_neurons=970,
                             circonv_n_neurons=100, n_landmarks=50, T=20.0, domain_dim=3, grid_points_per_dim=30,
                             distinct_tables=4, neuron_type="lifrate", view_rad=0.6)
    assert sc.ssp_space.ssp_dim == 649
    slam = sc.extra["slam"]
    with Simulator(sc.network, dt=sc.dt, n_trials=n_trials, trial_inputs=sc.trial_inputs, chunk_steps=n_steps) as sim:
        assert sim.plan.stats["n_neurons"] == 426380
        sim.run_steps(n_steps)
        idx = sim.cleanup_indices()[0].copy()
        dec = sim.learned_decoders(slam.assomemory.conn_out)
        enc = sim.learned_encoders(slam.assomemory.memory)
        launches = sim.total_launches()
    got = sim.data[sc.probe]
    assert got.shape == (n_trials, n_steps, 649) and np.all(np.isfinite(got))
    assert launches > 0
    for trial in (1, 3):
        ref = RefSimulator(sc.network, dt=sc.dt, model=sim.model, trial_seed=sim.trial_seeds[trial],
                           node_tables={node: arr[trial] for node, arr in sc.trial_inputs.items()})
        ref.run_steps(n_steps)
        want = ref.data[sc.probe]
        assert np.max(np.abs(want)) > 0.02          # 24 steps from rest: the output filter has only started to charge
        assert np.max(np.abs(got[trial] - want)) < 1e-4 * np.max(np.abs(want))
        assert idx[trial] == ssp_ref.cleanup_index(slam.sample_ssps, ref.signals[slam.gridcells, "in"].a)
        want_dec = ref.learned_weights(slam.assomemory.conn_out)
        assert np.max(np.abs(dec[trial] - want_dec)) < 1e-4 * np.max(np.abs(want_dec)) + 1e-9
        want_enc = ref.scaled_encoders(slam.assomemory.memory)
        assert np.max(np.abs(enc[trial] - want_enc)) < 1e-4 * np.max(np.abs(want_enc))
