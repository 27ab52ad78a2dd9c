"""The oracle's own builder (``oracle/nengo_ref_build.py``, written against SURVEY.md App. A without importing the
product's builder) against (1) analytic known answers for App. A.2 / A.3 / A.7 and (2) the product's
``builder.build_model`` on the hot-path graphs: seeds exactly, sampled quantities to 1e-12, decoders to 1e-9 of scale."""
import numpy as np
import pytest

from oracle import nengo_ref_build as rb
from oracle.nengo_ref_sim import RefSimulator
from sspslam_b200 import builder, nengo_shim as nengo, scenarios

MAXINT = np.iinfo(np.int32).max


# ------------------------------------------------------------------------------- App. A.2: seeds
def test_seed_assignment_order_known_answer():
    with nengo.Network(seed=11) as net:
        n1 = nengo.Node(lambda t: [0.0])
        e1 = nengo.Ensemble(10, 1)
        with nengo.Network() as sub:
            e2 = nengo.Ensemble(10, 1, seed=77)              # own seed: the draw still happens, the own seed wins
            e3 = nengo.Ensemble(10, 1)
        c1 = nengo.Connection(n1, e1)
        c2 = nengo.Connection(e1, e3)
        p1 = nengo.Probe(e3)
    rng = np.random.RandomState(11)
    draws = [rng.randint(MAXINT) for _ in range(6)]          # type order: connections, ensembles, networks, nodes, probes
    want = {c1: draws[0], c2: draws[1], e1: draws[2], sub: draws[3], n1: draws[4], p1: draws[5]}
    sub_rng = np.random.RandomState(draws[3])
    sub_draws = [sub_rng.randint(MAXINT) for _ in range(2)]
    want.update({e2: 77, e3: sub_draws[1]})
    for seeds in (rb.assign_seeds(net, 11), builder.build_model(net).seeds):
        for obj, s in want.items():
            assert seeds[obj] == s, obj


# ------------------------------------------------------------------------------- App. A.3: evaluation points, gain / bias
def test_eval_point_counts_known_answer():
    for (n, d), want in {(50, 1): 750, (500, 3): 1500, (970, 55): 2500, (100, 2): 1000, (2000, 1): 4000,
                         (1500, 55): 3000}.items():
        assert rb.n_eval_points(n, d) == want == builder.n_eval_points_default(n, d)


def test_lif_gain_bias_known_answer_and_tuning_curve_endpoints():
    lif = nengo.LIF()                                         # tau_rc 0.02, tau_ref 0.002
    max_rates, intercepts = np.array([200.0, 400.0, 300.0]), np.array([0.0, -0.5, 0.9])
    gain, bias = rb.gain_bias(lif, max_rates, intercepts)
    # hand computation for (200 Hz, intercept 0): J_max = 1 + 1/(e^{(0.005-0.002)/0.02} - 1) = 1 + 1/(e^{0.15} - 1)
    assert gain[0] == pytest.approx(1.0 / np.expm1(0.15), rel=1e-14) and bias[0] == pytest.approx(1.0, abs=1e-15)
    assert gain[0] == pytest.approx(6.179, abs=1e-3)
    # the defining properties: rate(x = 1) = max_rate, threshold current at x = intercept
    np.testing.assert_allclose(rb.rates(lif, gain * 1.0 + bias), max_rates, rtol=1e-12)
    np.testing.assert_allclose(gain * intercepts + bias, 1.0, rtol=0, atol=1e-12)
    g2, b2 = lif.gain_bias(max_rates, intercepts)             # the declaration layer's formula (nengo's published one)
    np.testing.assert_allclose(gain, g2, rtol=1e-12)
    np.testing.assert_allclose(bias, b2, rtol=1e-12, atol=1e-12)
    relu = nengo.RectifiedLinear()
    g, b = rb.gain_bias(relu, np.array([100.0]), np.array([0.5]))
    assert g[0] == 200.0 and b[0] == -100.0 and rb.rates(relu, g * 1.0 + b)[0] == 100.0


# ------------------------------------------------------------------------------- App. A.7: LstsqL2
def test_lstsq_l2_known_answer():
    # one neuron, activities a_i, targets y_i: x = sum(a y) / (sum(a^2) + m sigma^2), sigma = reg * max(a)
    a = np.array([[0.0], [2.0], [4.0]])
    y = np.array([[0.0], [1.0], [2.0]])
    want = (2 * 1 + 4 * 2) / (4 + 16 + 3 * (0.1 * 4.0) ** 2)
    assert rb.lstsq_l2(a, y, 0.1)[0, 0] == pytest.approx(want, rel=1e-13)
    solver = nengo.solvers.LstsqL2(reg=0.1)
    assert solver(a, y)[0][0, 0] == pytest.approx(want, rel=1e-13)
    rng = np.random.default_rng(0)
    A, Y = np.abs(rng.standard_normal((300, 40))) * 50, rng.standard_normal((300, 3))
    m, sigma = 300, 0.1 * A.max()
    closed = np.linalg.solve(A.T @ A + m * sigma ** 2 * np.eye(40), A.T @ Y)
    np.testing.assert_allclose(rb.lstsq_l2(A, Y, 0.1), closed, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(solver(A, Y)[0], closed, rtol=1e-9, atol=1e-12)
    wide = np.abs(rng.standard_normal((20, 40))) * 50          # fewer evaluation points than neurons: A A^T branch
    np.testing.assert_allclose(solver(wide, Y[:20])[0], rb.lstsq_l2(wide, Y[:20], 0.1), rtol=1e-8, atol=1e-12)


# ------------------------------------------------------------------------------- oracle build == product build
def _assert_same_build(net, dt=0.001):
    ma, mb = rb.build(net, dt), builder.build_model(net, dt=dt)
    assert set(ma.seeds) == set(mb.seeds)
    for obj, s in ma.seeds.items():
        assert mb.seeds[obj] == s
    for ens in net.all_ensembles:
        pa, pb = ma.params[ens], mb.params[ens]
        for name in ("eval_points", "encoders", "scaled_encoders", "gain", "bias"):
            a, b = getattr(pa, name), getattr(pb, name)
            np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-12 * max(1.0, np.max(np.abs(b))), err_msg=name)
        np.testing.assert_allclose(ma.initial_voltage(ens), mb.initial_voltage(ens), rtol=0, atol=0)
        np.testing.assert_allclose(ma.initial_voltage(ens, 5), mb.initial_voltage(ens, 5), rtol=0, atol=1e-15)
    n_dec = 0
    for conn in net.all_connections:
        wa, wb = ma.params[conn].weights, mb.params[conn].weights
        if wa is None or wb is None:
            assert wa is None and wb is None
            continue
        wa, wb = np.asarray(wa, float), np.asarray(wb, float)
        assert wa.shape == wb.shape
        assert np.max(np.abs(wa - wb)) <= 1e-9 * max(1e-300, np.max(np.abs(wb)))
        n_dec += ma.params[conn].decoders is not None
    for probe, dec in ma.probe_conns.items():
        assert np.max(np.abs(dec - mb.probe_conns[probe])) <= 1e-9 * np.max(np.abs(dec))
    return ma, mb, n_dec


@pytest.mark.parametrize("kind", ["pathint", "slam", "slamview", "loihi", "gc", "slam3d"])
def test_oracle_build_equals_product_build(kind):
    kw = dict(n_trials=1, n_steps=20, ssp_dim=19, pi_n_neurons=30, mem_n_neurons=64, circonv_n_neurons=16, n_landmarks=6,
              T=20.0, view_rad=0.6)
    if kind == "pathint":
        sc = scenarios.make_pathint(n_trials=1, n_steps=20, ssp_dim=19, pi_n_neurons=60, with_gcs=True, n_gcs=50)
    elif kind == "slamview":
        sc = scenarios.make_slam(view=True, **kw)
    elif kind == "loihi":
        sc = scenarios.make_slam(loihi=True, dotprod_n_neurons=20, **kw)
    elif kind == "gc":
        sc = scenarios.make_slam(gc_n_neurons=48, approx_vel=True, vel_n_neurons=40, **kw)
    elif kind == "slam3d":
        sc = scenarios.make_slam(domain_dim=3, grid_points_per_dim=8, **{**kw, "ssp_dim": 33})
    else:
        sc = scenarios.make_slam(**kw)
    _, _, n_dec = _assert_same_build(sc.network)
    assert n_dec > 5


def test_oracle_stepping_on_its_own_build_equals_stepping_on_the_product_build():
    sc = scenarios.make_slam(n_trials=1, n_steps=60, ssp_dim=19, pi_n_neurons=30, mem_n_neurons=64, circonv_n_neurons=16,
                             n_landmarks=6, T=20.0, view_rad=0.6, neuron_type="lifrate")
    tabs = {n: a[0] for n, a in sc.trial_inputs.items()}
    a = RefSimulator(sc.network, dt=sc.dt, node_tables=tabs)                                         # oracle build
    b = RefSimulator(sc.network, dt=sc.dt, node_tables=tabs, model=builder.build_model(sc.network))  # product build
    a.run_steps(60)
    b.run_steps(60)
    assert np.max(np.abs(a.data[sc.probe])) > 1e-2
    np.testing.assert_allclose(a.data[sc.probe], b.data[sc.probe], rtol=0, atol=1e-7 * np.max(np.abs(b.data[sc.probe])))
