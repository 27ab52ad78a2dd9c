"""BASELINE configs[4] at FULL size on the CPU: the host logic (build, lowering to the device plan) for the 3-D d = 649
network - 426 380 neurons, 11.5 M entries of row program - stepped for 100 timesteps by the operator-merged executor of
the plan (``oracle/plan_cpu.py``) against the operator-level oracle on the same built model.  The GPU side of the same
configuration is ``tests/test_gpu_config5_parity.py`` (24 steps per kernel variant)."""
import numpy as np

from oracle import ssp_ref
from oracle.nengo_ref_sim import RefSimulator
from oracle.plan_cpu import MergedPlanSimulator
from sspslam_b200 import scenarios, lowering
from sspslam_b200.builder import build_model


def test_config5_full_size_plan_matches_the_oracle_for_100_steps():
    n_steps = 100
    sc = scenarios.make_slam(n_trials=1, n_steps=n_steps + 4, ssp_dim=649, pi_n_neurons=500, mem_n_neurons=970,
                             circonv_n_neurons=100, n_landmarks=50, T=20.0, domain_dim=3, grid_points_per_dim=30,
                             neuron_type="lifrate", view_rad=0.6)
    assert sc.ssp_space.ssp_dim == 649
    model = build_model(sc.network, dt=sc.dt)
    plan = lowering.lower(sc.network, model, chunk_cap=n_steps, n_trials=512)
    assert plan.stats["n_neurons"] == 426380 and plan.stats["n_levels"] == 2
    assert plan.scalars["n_lin_early"] > 0 and plan.scalars["n_lin_fused"] == 0
    tabs = {node: arr[0] for node, arr in sc.trial_inputs.items()}
    ms = MergedPlanSimulator(plan, model, sc.network, tabs)
    ms.run_steps(n_steps)
    ref = RefSimulator(sc.network, dt=sc.dt, model=model, node_tables=tabs)
    ref.run_steps(n_steps)
    info = [i for i in plan.probes if i.probe is sc.probe][0]
    got, want = ms.probe_data(info), ref.data[sc.probe]
    assert np.max(np.abs(want)) > 0.05
    assert np.max(np.abs(got - want)) < 1e-5 * np.max(np.abs(want))          # float32-rounded plan coefficients
    slam = sc.extra["slam"]
    assert ms.cidx[0] == ssp_ref.cleanup_index(slam.sample_ssps, ref.signals[slam.gridcells, "in"].a)
    conn = slam.assomemory.conn_out
    row0, so, n = plan.learned_dec[conn]
    want_dec = ref.learned_weights(conn)
    # the executor (like the kernels) applies the PES delta of step t at step t + 1: one delta is still pending
    assert np.max(np.abs(want_dec)) > 0
    got_dec = ms.ldec[row0:row0 + so * n].reshape(so, n)
    assert np.max(np.abs(got_dec - want_dec)) < 0.05 * np.max(np.abs(want_dec))
