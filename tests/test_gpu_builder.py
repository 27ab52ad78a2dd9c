"""the decoder solves of the builder on the device (``libssb_builder.so``, cuBLAS / cuSOLVER batched,
float64) against the host solver and the oracle's own lstsq formulation (``-m gpu``)."""
import numpy as np
import pytest

from oracle import nengo_ref_build as rb
from sspslam_b200 import builder, cabi, scenarios

pytestmark = pytest.mark.gpu


def test_batched_solve_matches_the_closed_form(lib):
    rng = np.random.default_rng(0)
    for n_sys, m, n, k in ((5, 300, 40, 3), (3, 1500, 500, 6), (2, 64, 64, 1)):
        A = np.abs(rng.standard_normal((n_sys, m, n))) * 60.0 * (rng.random((n_sys, m, n)) < 0.6)
        Y = rng.standard_normal((n_sys, m, k))
        X = cabi.solve_decoders(A, Y, 0.1)
        for s in range(n_sys):
            sigma = 0.1 * A[s].max()
            want = np.linalg.solve(A[s].T @ A[s] + m * sigma ** 2 * np.eye(n), A[s].T @ Y[s])
            assert np.max(np.abs(X[s] - want)) <= 1e-9 * np.max(np.abs(want))
            assert np.max(np.abs(X[s] - rb.lstsq_l2(A[s], Y[s], 0.1))) <= 1e-8 * np.max(np.abs(want))
    with pytest.raises(cabi.SsbError):
        cabi.solve_decoders(np.ones((1, 4, 8)), np.ones((1, 4, 1)), 0.1)           # fewer evaluation points than neurons


@pytest.mark.parametrize("kind", ["pathint", "slam"])
def test_device_built_model_equals_host_built_model(lib, kind):
    if kind == "pathint":
        sc = scenarios.make_pathint(n_trials=1, n_steps=10, ssp_dim=55, pi_n_neurons=200)
    else:
        sc = scenarios.make_slam(n_trials=1, n_steps=10, ssp_dim=55, pi_n_neurons=100, mem_n_neurons=200, circonv_n_neurons=30,
                                 n_landmarks=20, T=20.0)
    host = builder.build_model(sc.network, dt=sc.dt)
    dev = builder.build_model(sc.network, dt=sc.dt, device_solver=0)
    n_dec = 0
    for conn in sc.network.all_connections:
        wa, wb = host.params[conn].weights, dev.params[conn].weights
        if host.params[conn].decoders is None:
            continue
        n_dec += 1
        assert np.max(np.abs(np.asarray(wa) - np.asarray(wb))) <= 1e-9 * np.max(np.abs(np.asarray(wa)))
    for probe, dec in host.probe_conns.items():
        assert np.max(np.abs(dec - dev.probe_conns[probe])) <= 1e-9 * np.max(np.abs(dec))
    assert n_dec > 20
