"""Trials with their own NETWORK seed (north_star 'batched over seeds'; run_slam.py:151 ``nengo.Network(seed=args.seed)``,
aggregated over seeds by ``experiments/plot_trials_2d.py``): host logic on CPU, the per-trial-weight kernel on the GPU."""
import numpy as np
import pytest

from oracle.nengo_ref_sim import RefSimulator
from sspslam_b200 import builder, lowering, scenarios, simulator


def test_per_seed_models_share_the_plan_and_differ_only_in_weights():
    sc = scenarios.make_pathint(n_trials=6, n_steps=40, ssp_dim=19, pi_n_neurons=40, neuron_type="lifrate")
    seeds = [3, 4, 5, 3, 9, 11]
    models = simulator._build_models(sc.network, 0.001, seeds, workers=3)          # fork pool + re-keying
    assert models[0] is models[3]                                                   # equal seeds are built once
    direct = builder.build_model(sc.network, dt=0.001, seed_override=5)
    for ens in sc.network.all_ensembles:
        assert models[2].seeds[ens] == direct.seeds[ens]
        assert np.array_equal(models[2].params[ens].scaled_encoders, direct.params[ens].scaled_encoders)
    for conn in sc.network.all_connections:
        wa, wb = models[2].params[conn].weights, direct.params[conn].weights
        assert (wa is None and wb is None) or np.array_equal(np.asarray(wa), np.asarray(wb))
    own = builder.build_model(sc.network, dt=0.001)                                 # the network's own seed (0)
    assert own.seeds[sc.network] == 0 and direct.seeds[sc.network] == 5
    p0, p1 = lowering.lower(sc.network, models[0]), lowering.lower(sc.network, models[1])
    for name in p0.arrays:
        same = np.array_equal(p0.arrays[name], p1.arrays[name])
        assert same == (name != "weights"), name                                   # only the static weights depend on the seed
    assert np.array_equal(lowering.narrow_ensemble_weights(sc.network, models[1]), p1.arrays["weights"])


def test_per_trial_seeds_are_refused_for_wide_ensembles():
    sc = scenarios.make_slam(n_trials=1, n_steps=10, ssp_dim=19, pi_n_neurons=30, mem_n_neurons=64, circonv_n_neurons=16,
                             n_landmarks=6, T=20.0)
    with pytest.raises(NotImplementedError):
        lowering.narrow_ensemble_weights(sc.network, builder.build_model(sc.network))


@pytest.mark.gpu
@pytest.mark.parametrize("neuron_type,tol", [("lifrate", 1e-4), ("lif", 2e-3)])
def test_trials_with_their_own_network_seed_match_the_oracle_of_their_own_model(lib, neuron_type, tol):
    """Trial i = the reference driver started with --seed s_i: its own encoders / gains / decoders (per-trial weight arena,
    k_ens_small_pt) and nengo's own start voltages for that seed."""
    from sspslam_b200.simulator import Simulator
    n_steps, seeds = 150, [7, 8, 9, 10, 7]
    sc = scenarios.make_pathint(n_trials=5, n_steps=n_steps, ssp_dim=55, pi_n_neurons=120, neuron_type=neuron_type)
    with Simulator(sc.network, dt=sc.dt, n_trials=5, trial_inputs=sc.trial_inputs, trial_network_seeds=seeds) as sim:
        assert sim.plan.scalars["per_trial_weights"] == 1
        sim.run_steps(n_steps)
    got = sim.data[sc.probe]
    for trial in (0, 1, 3):
        tabs = {node: arr[trial] for node, arr in sc.trial_inputs.items()}
        ref = RefSimulator(sc.network, dt=sc.dt, model=builder.build_model(sc.network, dt=sc.dt, seed_override=seeds[trial]),
                           node_tables=tabs)
        ref.run_steps(n_steps)
        want = ref.data[sc.probe]
        assert np.max(np.abs(want)) > 0.1
        assert np.max(np.abs(got[trial] - want)) < tol * np.max(np.abs(want))
    assert not np.allclose(got[0], got[1], atol=1e-3)          # different seeds, different trajectories
    # trials 0 and 4 share the seed but not the input path; a shared-weight run with that seed reproduces trial 0
    with Simulator(sc.network, dt=sc.dt, n_trials=5, trial_inputs=sc.trial_inputs, trial_seeds=[None] * 5,
                   model=builder.build_model(sc.network, dt=sc.dt, seed_override=7)) as shared:
        shared.run_steps(n_steps)
    assert np.max(np.abs(shared.data[sc.probe][0] - got[0])) < tol * np.max(np.abs(got[0]))
    assert np.max(np.abs(shared.data[sc.probe][4] - got[4])) < tol * np.max(np.abs(got[4]))
