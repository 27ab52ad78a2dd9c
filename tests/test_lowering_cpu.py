"""Host logic of the product path, checked WITHOUT a GPU: the lowered device plan, executed by the
NumPy plan interpreter in tests/plan_interp.py (float64), must reproduce the operator-level oracle
on the same built model.  Also structural checks of the plan (levels, chunk slots, traffic model)."""
import numpy as np
import pytest

from oracle.nengo_ref_sim import RefSimulator
from sspslam_b200 import scenarios, lowering
from sspslam_b200.builder import build_model
from plan_interp import PlanInterpreter


def _compare(sc, n_steps, n_trials_hint=1, tol=2e-5):
    model = build_model(sc.network, dt=sc.dt)
    plan = lowering.lower(sc.network, model, chunk_cap=n_steps, n_trials=n_trials_hint)
    tabs = {node: arr[0] for node, arr in sc.trial_inputs.items()}
    ref = RefSimulator(sc.network, dt=sc.dt, model=model, node_tables=tabs)
    ref.run_steps(n_steps)
    it = PlanInterpreter(plan, model, sc.network, tabs)
    it.run_steps(n_steps)
    info = [i for i in plan.probes if i.probe is sc.probe][0]
    got, want = it.probe_data(info), ref.data[sc.probe]
    scale = np.max(np.abs(want))
    assert scale > 1e-3
    # the plan stores weights / CSR coefficients in float32 (the interpreter's arithmetic is float64)
    assert np.max(np.abs(got - want)) <= tol * scale
    # step fusion: the rows the end-of-step launch evaluates for the NEXT step equal that step's own level-0 rows
    if plan.scalars["n_lin_fused"]:
        n0 = int(plan.arrays["stages"][0][11])
        assert plan.scalars["n_lin_fused"] == n0 - plan.scalars["n_lvl0_res"]
        assert it.fused_checked == plan.scalars["n_lin_fused"] * (n_steps - 1) and it.fused_err < 1e-4
    return plan, model, ref, it


@pytest.mark.parametrize("neuron_type", ["lifrate", "lif"])
def test_pathint_plan_matches_oracle(neuron_type):
    sc = scenarios.make_pathint(n_trials=1, n_steps=80, ssp_dim=19, pi_n_neurons=40, neuron_type=neuron_type)
    plan, *_ = _compare(sc, 80)
    assert plan.stats["n_levels"] == 1 and plan.stats["n_big"] == 0
    assert plan.stats["n_small"] == (sc.ssp_space.ssp_dim + 1) // 2


@pytest.mark.parametrize("neuron_type,hint", [("lifrate", 1), ("lif", 4096)])
def test_slam_plan_matches_oracle(neuron_type, hint):
    """Full SLAM graph: OVC, two circular convolutions, Voja + PES memory, clean-up, gate."""
    sc = scenarios.make_slam(n_trials=1, n_steps=70, ssp_dim=19, pi_n_neurons=30, mem_n_neurons=70,
                             circonv_n_neurons=16, n_landmarks=6, T=20.0, neuron_type=neuron_type, view_rad=0.6)
    plan, model, ref, it = _compare(sc, 70, n_trials_hint=hint)
    slam = sc.extra["slam"]
    # learned matrices moved and agree with the oracle (PES decoders, Voja encoders)
    conn = slam.assomemory.conn_out
    row0, so, n = plan.learned_dec[conn]
    D = it.ldec[row0:row0 + so * n].reshape(so, n)
    # the interpreter (like the kernel) applies the delta of step t at step t+1: one more delta is pending
    want = ref.learned_weights(conn)
    assert np.max(np.abs(want)) > 0
    mem = slam.assomemory.memory
    e0, n_e, dims = plan.learned_enc[mem]
    E = it.lenc[e0:e0 + n_e * dims].reshape(n_e, dims)
    assert not np.allclose(E, model.params[mem].scaled_encoders)          # Voja moved the encoders
    assert plan.stats["n_levels"] == 2
    assert plan.stats["n_learned"] == 2 * so * n
    # chunked decoders: the hint for a big batch needs fewer partial slots than a single-trial run
    chunks = int(plan.arrays["pes"][0][10])
    assert plan.scalars["n_part"] >= (chunks * so if chunks > 1 else 0)
    jtiles = -(-so // lowering.DEC_TILE)
    k_max = max(1, min(n // 32, lowering.MAX_DEC_CHUNKS))
    assert 1 <= chunks <= k_max
    # the split minimises (waves of resident CTAs) x (neurons per chunk)
    slots = lowering.N_SM * lowering.PES_CTAS_PER_SM
    cost = lambda k: -(-(jtiles * -(-hint // 32) * k) // slots) * (-(-n // k) + 8)
    assert cost(chunks) == min(cost(k) for k in range(1, k_max + 1))


def test_slamview_plan_matches_oracle():
    sc = scenarios.make_slam(n_trials=1, n_steps=60, ssp_dim=19, pi_n_neurons=30, mem_n_neurons=64,
                             circonv_n_neurons=16, n_landmarks=6, T=20.0, neuron_type="lifrate", view=True, view_rad=0.6)
    plan, *_ = _compare(sc, 60)
    assert plan.arrays["cleanup"].shape[0] == 1 and plan.arrays["gate"].shape[0] == 1


def test_unrecognised_python_node_is_rejected():
    """There is no host-callback path: a Python node with inputs that is not one of the three
    recognised device ops must make the lowering fail loudly."""
    from sspslam_b200 import nengo_shim as nengo
    with nengo.Network(seed=1) as net:
        a = nengo.Node(lambda t: [np.sin(t)])
        b = nengo.Node(lambda t, x: x ** 2 + 1.0, size_in=1)
        nengo.Connection(a, b, synapse=None)
        nengo.Probe(b)
    model = build_model(net)
    with pytest.raises(NotImplementedError):
        lowering.lower(net, model)


def test_traffic_model_config2_sizes():
    """config 2 has 40 280 neurons, 106 700 learned weights, ~1.5 MB per trial-step."""
    stats = dict(n_neurons=40280, n_filter_states=825, n_afilt=970, n_learned=106700, n_static_weights=0,
                 n_table_words=168, n_probe_words=55)
    b = lowering.algorithmic_bytes_per_trial_step(stats)
    assert 1.50e6 < b < 1.53e6


def test_slam_3d_plan_matches_oracle():
    """BASELINE configs[4] topology at toy size: 3-D domain (3-D velocity into the 3-D VCOs, 12^3 grid)."""
    sc = scenarios.make_slam(n_trials=1, n_steps=50, ssp_dim=55, pi_n_neurons=30, mem_n_neurons=64, circonv_n_neurons=16,
                             n_landmarks=6, T=20.0, neuron_type="lifrate", view_rad=0.6, domain_dim=3,
                             grid_points_per_dim=12)
    assert sc.ssp_space.domain_dim == 3
    plan, *_ = _compare(sc, 50)
    assert int(plan.arrays["cleanup"][0][0]) == 12 ** 3
    assert {int(r[1]) for r in plan.arrays["ens_small"]} == {1, 3}    # VCOs stay (Re, Im, frequency) in any domain
    assert sc.trial_inputs[[n for n in sc.trial_inputs if n.label == "vel_input"][0]].shape[2] == 3


def test_slam_grid_cell_ensemble_and_approx_velocity_plan_matches_oracle():
    """SURVEY.md §8f-4 topologies on the same kernels: clean-up -> grid-cell ensemble (wide, custom encoders,
    CosineSimilarity intercepts) -> circular convolution, and the velocity routed through a spiking ensemble."""
    sc = scenarios.make_slam(n_trials=1, n_steps=60, ssp_dim=19, pi_n_neurons=30, mem_n_neurons=64, circonv_n_neurons=16,
                             n_landmarks=6, T=20.0, neuron_type="lifrate", view_rad=0.6, gc_n_neurons=40, approx_vel=True,
                             vel_n_neurons=50)
    slam = sc.extra["slam"]
    assert slam.gridcells.n_neurons == 40
    plan, *_ = _compare(sc, 60)
    assert plan.stats["n_big"] == 5                      # OVC, memory, recall, error + the grid-cell ensemble
    assert plan.stats["n_levels"] >= 2                   # clean-up node -> grid cells in the same step


def test_neuron_slice_and_vco_probes_match_oracle():
    """run_pathint_gif.py:156-159: a sliced output probe and neuron-output probes with sample_every."""
    from sspslam_b200 import nengo_shim as nengo
    sc = scenarios.make_pathint(n_trials=1, n_steps=60, ssp_dim=19, pi_n_neurons=40, neuron_type="lif")
    pi = sc.extra["pathint"]
    with sc.network:
        vco_p = nengo.Probe(pi.oscillators.output[3:12], synapse=0.05)
        n1 = nengo.Probe(pi.oscillators.ea_ensembles[1].neurons[:25], synapse=None, sample_every=5 * sc.dt)
        n2 = nengo.Probe(pi.oscillators.ea_ensembles[2].neurons, synapse=None)
    plan, model, ref, it = _compare(sc, 60)
    assert plan.stats["n_big"] == 2                      # the two probed VCO populations keep activity rows
    for probe in (vco_p, n1, n2):
        info = [i for i in plan.probes if i.probe is probe][0]
        got, want = it.probe_data(info), ref.data[probe]
        assert got.shape == want.shape
        if probe is vco_p:
            assert np.max(np.abs(got - want)) <= 2e-5 * np.max(np.abs(want))
        else:
            assert np.max(np.abs(want)) > 0              # some neuron spiked
            assert np.array_equal(got != 0, want != 0)   # same spikes, step for step
            assert np.allclose(got, want, rtol=1e-6)
    assert ref.data[n1].shape == (12, 25)


def test_slam_loihi_plan_matches_oracle():
    """SLAMLoihiNetwork (slam_loihi.py) — neural gate, Lowpass(0) input, four wide populations."""
    sc = scenarios.make_slam(n_trials=1, n_steps=60, ssp_dim=19, pi_n_neurons=30, mem_n_neurons=48, circonv_n_neurons=12,
                             n_landmarks=6, T=20.0, neuron_type="lifrate", view_rad=0.5, grid_points_per_dim=12,
                             loihi=True, dotprod_n_neurons=16)
    plan, *_ = _compare(sc, 60)
    assert plan.stats["n_big"] == 4          # memory, recall, PES error, correction
    assert not plan.arrays["cleanup"].size and not plan.arrays["gate"].size     # no node functions: the gate is neural


def test_inverse_memory_topology_plan_matches_oracle():
    """experiments/slam_map_new.py:207-263 — two path integrators, two Voja + PES memories, and that
    script's probes (decoded ensemble output, node, weights, scaled_encoders)."""
    n = 60
    sc = scenarios.make_slam(n_trials=1, n_steps=n, ssp_dim=19, pi_n_neurons=30, mem_n_neurons=48, circonv_n_neurons=12,
                             n_landmarks=6, T=20.0, neuron_type="lifrate", view_rad=0.5, grid_points_per_dim=12,
                             inverse_memory=True)
    plan, model, ref, it = _compare(sc, n)
    d = sc.ssp_space.ssp_dim
    assert len(plan.arrays["pes"]) == 2 and plan.stats["n_learned"] == 2 * (48 * d + 48 * d)
    for name in ("ssp_pi_p", "newpos_p", "objssp_p", "recall_p", "isitem_p"):
        probe = sc.extra[name]
        info = [i for i in plan.probes if i.probe is probe][0]
        got, want = it.probe_data(info), ref.data[probe]
        assert got.shape == want.shape
        assert np.max(np.abs(got - want)) <= 2e-5 * max(1e-3, np.max(np.abs(want)))
    kinds = sorted(i.kind for i in plan.probes)
    assert kinds.count("weights") == 2 and kinds.count("scaled_encoders") == 2


def test_pathint_with_grid_cell_output_plan_matches_oracle():
    """pathintegration.py:150-154 (``with_gcs=True``): the probed output is a decoded grid-cell population."""
    sc = scenarios.make_pathint(n_trials=1, n_steps=80, ssp_dim=19, pi_n_neurons=40, neuron_type="lifrate",
                                with_gcs=True, n_gcs=64)
    plan, *_ = _compare(sc, 80)
    assert plan.stats["n_big"] == 1


def test_slam_without_voja_plan_matches_oracle():
    """run_slam.py --no-voja (slam.py:196-198): fixed landmark encoders, the memory becomes a static wide ensemble."""
    sc = scenarios.make_slam(n_trials=1, n_steps=60, ssp_dim=19, pi_n_neurons=30, mem_n_neurons=48, circonv_n_neurons=12,
                             n_landmarks=6, T=20.0, neuron_type="lifrate", view_rad=0.5, grid_points_per_dim=12, voja=False)
    plan, *_ = _compare(sc, 60)
    d = sc.ssp_space.ssp_dim
    assert plan.stats["n_learned"] == 48 * d          # only the PES decoders are per trial


def test_slamview_with_the_drivers_bound_view_input_matches_oracle():
    """run_slamview.py:103,123-130: the view vector is the normalised sum of SP (*) SSP(displacement) (slam_view.py:384-394)."""
    sc = scenarios.make_slam(n_trials=1, n_steps=60, ssp_dim=19, pi_n_neurons=30, mem_n_neurons=64,
                             circonv_n_neurons=16, n_landmarks=6, T=20.0, neuron_type="lifrate", view=True, view_rad=0.6,
                             view_bound=True)
    view = [arr for node, arr in sc.trial_inputs.items() if node.label == "lm_sp_input"][0][0]
    norms = np.linalg.norm(view, axis=1)
    assert np.all((np.abs(norms - 1.0) < 1e-9) | (norms == 0)) and norms.max() > 0     # unit view vectors when in view
    assert sc.extra["input_synthesis"] is None
    _compare(sc, 60)


def test_step_fusion_rows_equal_the_next_steps_level0_rows(monkeypatch):
    """SSB_LIN_FUSE=1 (opt-in; measured slower on B200): the end-of-step launch also evaluates the next
    step's level-0 sink rows from this step's columns; ``_compare`` checks every such row against the next step's own."""
    monkeypatch.setenv("SSB_LIN_FUSE", "1")
    sc = scenarios.make_slam(n_trials=1, n_steps=40, ssp_dim=19, pi_n_neurons=30, mem_n_neurons=70,
                             circonv_n_neurons=16, n_landmarks=6, T=20.0, neuron_type="lif", view_rad=0.6)
    plan, *_ = _compare(sc, 40)
    assert plan.scalars["n_lin_fused"] > 0 and plan.scalars["n_lvl0_res"] > 0
    sc = scenarios.make_pathint(n_trials=1, n_steps=40, ssp_dim=19, pi_n_neurons=40, neuron_type="lifrate")
    plan, *_ = _compare(sc, 40)
    assert plan.scalars["n_lin_fused"] > 0 and plan.scalars["n_lvl0_res"] == 0


def test_early_end_of_step_rows_read_only_what_exists_after_level0_narrow_ensembles():
    """The launch sequence runs the 'early' end-of-step rows right after level 0's narrow-ensemble kernel, next to every
    other chain of the step.  So none of their entries may point at a vec row that is produced later in the step: outputs
    of later levels' narrow ensembles, of static / PES decoders, of clean-up or gate nodes; and they must be filter rows
    (they write the half nobody reads in this step)."""
    sc = scenarios.make_slam(n_trials=1, n_steps=10, ssp_dim=19, pi_n_neurons=30, mem_n_neurons=70, circonv_n_neurons=16,
                             n_landmarks=6, T=20.0, view_rad=0.6)
    plan = lowering.lower(sc.network, build_model(sc.network, dt=sc.dt))
    a, s = plan.arrays, plan.scalars
    n_early, lin0 = int(s["n_lin_early"]), int(s["lin0"])
    assert 0 < n_early < s["n_lin"] and s["n_levels"] == 2
    late = set()
    for lvl, st in enumerate(a["stages"]):
        if lvl > 0:
            for d in a["ens_small"][st[0]:st[0] + st[1]]:
                late.update(range(int(d[6]), int(d[6]) + int(d[2])))
        for d in a["dec"][st[4]:st[4] + st[5]]:
            late.update(range(int(d[5]), int(d[5]) + int(d[1])))
        for d in a["cleanup"][st[6]:st[6] + st[7]]:
            late.update(range(int(d[5]), int(d[5]) + int(d[1])))
        for d in a["gate"][st[8]:st[8] + st[9]]:
            late.update(range(int(d[2]), int(d[2]) + int(d[0])))
    for d in a["pes"]:
        late.update(range(int(d[6]), int(d[6]) + int(d[1])))
    assert late
    ptr = a["csr_ptr"]
    early, rest = a["lin_rows"][lin0:lin0 + n_early], a["lin_rows"][lin0 + n_early:lin0 + int(s["n_lin"])]
    assert np.all(early[:, 1] == 0)
    for ent in (a["csr_ent0"], a["csr_ent1"]):
        for src, _, _ in early:
            rows = ent[ptr[src]:ptr[src + 1], 0]
            assert not late.intersection(int(r) for r in rows)
    # and the split is not trivial: some remaining filter row does read a late-produced row
    reads_late = False
    for src, kind, _ in rest:
        if kind == 0 and late.intersection(int(r) for r in a["csr_ent0"][ptr[src]:ptr[src + 1], 0]):
            reads_late = True
            break
    assert reads_late
    # a single-level plan keeps one end-of-step launch
    pi = scenarios.make_pathint(n_trials=1, n_steps=10, ssp_dim=19, pi_n_neurons=40)
    assert lowering.lower(pi.network, build_model(pi.network, dt=pi.dt)).scalars["n_lin_early"] == 0


def test_level_dependencies_of_slam():
    """Level 1 of SLAM (the landmark circular convolution) reads level 0's static decoders (bit 1: the OVC decode) and its
    clean-up node (bit 2), not its narrow ensembles (bit 0) nor the gate (bit 3): that is what lets its chain start before
    the 14 000 VCO neurons are done."""
    sc = scenarios.make_slam(n_trials=1, n_steps=10, ssp_dim=19, pi_n_neurons=30, mem_n_neurons=70, circonv_n_neurons=16,
                             n_landmarks=6, T=20.0, view_rad=0.6)
    plan = lowering.lower(sc.network, build_model(sc.network, dt=sc.dt))
    assert plan.arrays["level_deps"].tolist() == [[0, 0], [2 | 4, 0]]
