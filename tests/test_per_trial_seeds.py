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
    # lowered with per_trial=True the seed-dependent rows move to the per-trial weight arena, whose values for any other
    # model of the same graph come from trial_weights()
    q0, q1 = (lowering.lower(sc.network, m, per_trial=True) for m in models[:2])
    assert q0.scalars["per_trial_weights"] == 1
    for name in q0.arrays:
        assert np.array_equal(q0.arrays[name], q1.arrays[name]) == (name != "weights_pt"), name
    w1, enc, dec = lowering.trial_weights(sc.network, models[1])
    assert np.array_equal(w1, q1.arrays["weights_pt"]) and not enc and not dec
    assert q0.arrays["weights_pt"].size == p0.arrays["weights"].size                # narrow ensembles only: the same rows
    assert q0.stats["n_pt_weights"] == q0.arrays["weights_pt"].size - 8 and p0.stats["n_pt_weights"] == 0
    assert (lowering.algorithmic_bytes_per_trial_step(q0.stats) - lowering.algorithmic_bytes_per_trial_step(p0.stats)
            == 4 * q0.stats["n_pt_weights"])


def test_per_trial_plan_of_a_network_with_wide_ensembles():
    """SLAM with per-trial network seeds: narrow ensembles and wide bias / Voja scale in the per-trial weight arena, wide
    encoders as per-trial rows of the lenc arena (walked by the Voja kernel, alpha = 0 unless learned), wide static decoders
    in the ldec arena next to the PES-learned one; the clean-up grid and the row program stay shared."""
    sc = scenarios.make_slam(n_trials=4, n_steps=10, ssp_dim=19, pi_n_neurons=30, mem_n_neurons=64, circonv_n_neurons=16,
                             n_landmarks=6, T=20.0)
    m0, m1 = (builder.build_model(sc.network, dt=0.001, seed_override=s) for s in (3, 4))
    shared = lowering.lower(sc.network, m0, n_trials=4)
    p0, p1 = (lowering.lower(sc.network, m, n_trials=4, per_trial=True) for m in (m0, m1))
    for name in p0.arrays:
        assert np.array_equal(p0.arrays[name], p1.arrays[name]) == (name != "weights_pt"), name
    big = p0.arrays["ens_big"]
    assert len(big) == shared.stats["n_big"] and np.all(big[:, 9] & 1) and np.all(big[:, 9] & 4)
    n_voja = int(np.sum(shared.arrays["ens_big"][:, 9] & 1))
    assert len(p0.learned_enc) == n_voja and len(p0.pt_enc) == len(big) - n_voja
    assert np.sum(np.asarray(big[:, 15]) != 0) == n_voja                            # alpha = 0 for the static ones
    assert len(p0.pt_dec) == len(shared.static_dec) and not p0.static_dec
    assert np.all(p0.arrays["dec"][:, 6] == 1)                                      # no split-K on the per-trial decode
    # lenc / ldec rows: learned + per-trial static, without overlap
    spans = sorted((r0, r0 + n * d) for r0, n, d in list(p0.learned_enc.values()) + list(p0.pt_enc.values()))
    assert spans[0][0] == 0 and all(a[1] == b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] == p0.scalars["n_lenc"]
    spans = sorted((r0, r0 + n * (-(-so // 4) * 4)) for r0, so, n in list(p0.learned_dec.values()) + list(p0.pt_dec.values()))
    assert spans[0][0] == 0 and all(a[1] == b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] == p0.scalars["n_ldec"]
    # the shared array keeps only what does not depend on the network seed (grid, direct neuron-current weights)
    assert shared.arrays["weights"].size - p0.arrays["weights"].size > p0.arrays["weights_pt"].size
    w1, enc, dec = lowering.trial_weights(sc.network, m1, 4)
    assert np.array_equal(w1, p1.arrays["weights_pt"])
    assert set(enc) >= set(p0.pt_enc) and set(dec) == set(p0.pt_dec)
    for ens in p0.pt_enc:
        assert np.array_equal(enc[ens], np.asarray(m1.params[ens].scaled_encoders, dtype=np.float32))
    # trial 1's bias of a wide ensemble sits where the descriptor points
    ens = next(iter(p0.pt_enc))
    row = [r for r in big if r[5] == p0.pt_enc[ens][0]][0]
    assert np.array_equal(w1[row[6]:row[6] + row[0]], np.asarray(m1.params[ens].bias, dtype=np.float32))


@pytest.mark.parametrize("neuron_type", ["lifrate", "lif"])
def test_per_trial_slam_plan_interpreted_on_another_seeds_model_matches_its_oracle(neuron_type):
    """The plan is lowered from the model of seed 3 (layout + trial 0's values); executed with the arenas of the model of
    seed 4 by the NumPy plan interpreter it must reproduce the oracle stepped on the model of seed 4."""
    from plan_interp import PlanInterpreter
    n_steps = 70
    sc = scenarios.make_slam(n_trials=1, n_steps=n_steps, ssp_dim=19, pi_n_neurons=30, mem_n_neurons=70,
                             circonv_n_neurons=16, n_landmarks=6, T=20.0, neuron_type=neuron_type, view_rad=0.6)
    m3, m4 = (builder.build_model(sc.network, dt=sc.dt, seed_override=s) for s in (3, 4))
    plan = lowering.lower(sc.network, m3, chunk_cap=n_steps, per_trial=True)
    tabs = {node: arr[0] for node, arr in sc.trial_inputs.items()}
    ref = RefSimulator(sc.network, dt=sc.dt, model=m4, node_tables=tabs)
    ref.run_steps(n_steps)
    it = PlanInterpreter(plan, m4, sc.network, tabs)
    it.run_steps(n_steps)
    info = [i for i in plan.probes if i.probe is sc.probe][0]
    got, want = it.probe_data(info), ref.data[sc.probe]
    assert np.max(np.abs(want)) > 1e-3
    assert np.max(np.abs(got - want)) <= 2e-5 * np.max(np.abs(want))
    ref3 = RefSimulator(sc.network, dt=sc.dt, model=m3, node_tables=tabs)
    ref3.run_steps(n_steps)
    assert np.max(np.abs(ref3.data[sc.probe] - want)) > 1e-3 * np.max(np.abs(want))     # the seeds do differ


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


@pytest.mark.gpu
@pytest.mark.parametrize("neuron_type,tol", [("lifrate", 1e-4), ("lif", 5e-3)])
def test_slam_trials_with_their_own_network_seed_match_the_oracle_of_their_own_model(lib, neuron_type, tol):
    """SLAM (wide ensembles, Voja + PES, clean-up, gate) with one network seed per trial: per-trial encoders through the
    Voja-class kernel, per-trial static decoders through k_decode_pt, per-trial bias / scale rows; every checked trial
    against the oracle stepped on ITS model, incl. the learned matrices."""
    from sspslam_b200.simulator import Simulator
    n_steps, seeds = 120, [21, 22, 23, 21, 24, 25, 26]
    sc = scenarios.make_slam(n_trials=len(seeds), n_steps=n_steps, ssp_dim=31, pi_n_neurons=60, mem_n_neurons=120,
                             circonv_n_neurons=20, n_landmarks=8, T=20.0, neuron_type=neuron_type, view_rad=0.6)
    slam = sc.extra["slam"]
    with Simulator(sc.network, dt=sc.dt, n_trials=len(seeds), trial_inputs=sc.trial_inputs,
                   trial_network_seeds=seeds) as sim:
        assert sim.plan.scalars["per_trial_weights"] == 1 and sim.plan.pt_dec and sim.plan.pt_enc
        sim.run_steps(n_steps)
        dec = sim.learned_decoders(slam.assomemory.conn_out)
        enc = sim.learned_encoders(slam.assomemory.memory)
    got = sim.data[sc.probe]
    for trial in (0, 2, 3, 6):
        tabs = {node: arr[trial] for node, arr in sc.trial_inputs.items()}
        ref = RefSimulator(sc.network, dt=sc.dt, model=builder.build_model(sc.network, dt=sc.dt, seed_override=seeds[trial]),
                           node_tables=tabs)
        ref.run_steps(n_steps)
        want = ref.data[sc.probe]
        assert np.max(np.abs(want)) > 0.05
        assert np.max(np.abs(got[trial] - want)) < tol * np.max(np.abs(want))
        if neuron_type == "lifrate":
            want_dec = ref.learned_weights(slam.assomemory.conn_out)
            assert np.max(np.abs(dec[trial] - want_dec)) < 1e-4 * np.max(np.abs(want_dec)) + 1e-9
            want_enc = ref.scaled_encoders(slam.assomemory.memory)
            assert np.max(np.abs(enc[trial] - want_enc)) < 1e-4 * np.max(np.abs(want_enc))
    assert not np.allclose(got[0], got[1], atol=1e-3)
