"""The operator-merged CPU executor (``oracle/plan_cpu.py``, the companion CPU baseline of ``bench.py``) against the
operator-level port ``oracle/nengo_ref_sim.py`` and against the row-by-row plan interpreter: same built model, same
inputs, float64."""
import numpy as np
import pytest

from oracle.nengo_ref_sim import RefSimulator
from oracle.plan_cpu import MergedPlanSimulator
from sspslam_b200 import scenarios, lowering
from sspslam_b200.builder import build_model
from plan_interp import PlanInterpreter


def _run(sc, n_steps):
    model = build_model(sc.network, dt=sc.dt)
    plan = lowering.lower(sc.network, model, chunk_cap=n_steps)
    tabs = {node: arr[0] for node, arr in sc.trial_inputs.items()}
    ref = RefSimulator(sc.network, dt=sc.dt, model=model, node_tables=tabs)
    ref.run_steps(n_steps)
    it = PlanInterpreter(plan, model, sc.network, tabs)
    it.run_steps(n_steps)
    ms = MergedPlanSimulator(plan, model, sc.network, tabs)
    ms.run_steps(n_steps)
    info = [i for i in plan.probes if i.probe is sc.probe][0]
    return plan, ref.data[sc.probe], it, ms, info


@pytest.mark.parametrize("neuron_type", ["lifrate", "lif"])
def test_merged_executor_matches_the_port_on_slam(neuron_type):
    sc = scenarios.make_slam(n_trials=1, n_steps=90, ssp_dim=19, pi_n_neurons=30, mem_n_neurons=70, circonv_n_neurons=16,
                             n_landmarks=6, T=20.0, neuron_type=neuron_type, view_rad=0.6)
    plan, want, it, ms, info = _run(sc, 80)
    got = ms.probe_data(info)
    assert np.max(np.abs(want)) > 1e-3
    assert np.max(np.abs(got - want)) <= 2e-5 * np.max(np.abs(want))            # float32-rounded coefficients
    assert np.max(np.abs(got - it.probe_data(info))) <= 1e-12                   # same plan, same arithmetic, merged
    assert np.array_equal(ms.cidx, it.cidx)
    assert np.max(np.abs(ms.ldec - it.ldec)) <= 1e-12 and np.max(np.abs(ms.lenc - it.lenc)) <= 1e-12
    assert np.max(np.abs(ms.ldec)) > 0                                          # PES moved the decoders


def test_merged_executor_matches_the_port_on_pathint():
    sc = scenarios.make_pathint(n_trials=1, n_steps=90, ssp_dim=19, pi_n_neurons=40, neuron_type="lif")
    plan, want, it, ms, info = _run(sc, 80)
    got = ms.probe_data(info)
    assert np.max(np.abs(got - want)) <= 2e-5 * np.max(np.abs(want))
    assert np.max(np.abs(got - it.probe_data(info))) <= 1e-12


def test_merged_executor_refuses_per_trial_plans():
    sc = scenarios.make_pathint(n_trials=1, n_steps=10, ssp_dim=19, pi_n_neurons=40)
    model = build_model(sc.network, dt=sc.dt)
    plan = lowering.lower(sc.network, model, per_trial=True)
    with pytest.raises(NotImplementedError):
        MergedPlanSimulator(plan, model, sc.network, {node: arr[0] for node, arr in sc.trial_inputs.items()})
