"""Fourier layouts, circular-convolution transforms, R_d sampling and the per-step input tables
against the unmodified reference (golden vectors) + KATs K3, K4, K5, K7 of SURVEY.md §4."""
import numpy as np

from sspslam_b200 import networks, inputs
from sspslam_b200.sspspace import HexagonalSSPSpace, SPSpace

BOUNDS2 = np.tile([-1.0, 1.0], (2, 1))


def test_fourier_layouts_match_reference(golden):
    for d in (7, 55):
        np.testing.assert_allclose(networks.get_to_Fourier(d), golden[f"toF{d}"], atol=1e-14)
        np.testing.assert_allclose(networks.get_from_Fourier(d), golden[f"fromF{d}"], atol=1e-14)


def test_k3_fourier_roundtrip():
    rng = np.random.default_rng(0)
    for d in (55, 97, 649):
        phi = rng.standard_normal(d)
        osc = networks.get_to_Fourier(d) @ phi
        osc[0] = np.fft.fft(phi)[0].real          # VCO 0 holds the DC term through its constant input
        np.testing.assert_allclose(networks.get_from_Fourier(d) @ osc, phi, atol=1e-12)


def test_circconv_transforms_match_reference(golden):
    for d in (7, 55):
        np.testing.assert_allclose(networks.transform_in(d, "A", False), golden[f"trA{d}"], atol=1e-13)
        np.testing.assert_allclose(networks.transform_in(d, "A", True), golden[f"trAinv{d}"], atol=1e-13)
        np.testing.assert_allclose(networks.transform_in(d, "B", False), golden[f"trB{d}"], atol=1e-13)
        np.testing.assert_allclose(networks.transform_out(d), golden[f"trOut{d}"], atol=1e-13)
    assert networks.transform_in(55, "A", False).shape == (112, 55)        # K4: all 4*(d//2+1) rows are kept
    assert networks.transform_out(55).shape == (55, 112)


def test_k4_neural_circconv_identity(golden):
    x, y = golden["cc_x"], golden["cc_y"]
    for inv in (False, True):
        p = networks.transform_in(55, "A", inv) @ x
        q = networks.transform_in(55, "B", False) @ y
        got = networks.transform_out(55) @ (p * q)
        want = golden["cc_xinv_y"] if inv else golden["cc_xy"]
        np.testing.assert_allclose(got, want, atol=1e-12)
        np.testing.assert_allclose(networks.circconv(x, y, invert_a=inv), want, atol=1e-13)


def test_k5_feedback_limit_cycle():
    fb = networks.oscillator_feedback(0.05, 2.0, 0.2, 1.0, True)
    out = fb(np.array([1.0, 0.0, 0.3]))
    np.testing.assert_allclose(out, [1.0, 0.05 * 0.3 / (2.0 * 0.2), 0.0], atol=1e-15)


def test_rd_sampling_matches_reference(golden):
    assert np.array_equal(inputs.rd_sampling(50, 2, seed=0), golden["rd_50_2_s0"])
    assert np.array_equal(inputs.rd_sampling(20, 2, seed=1000), golden["rd_20_2_s1000"])
    assert np.array_equal(inputs.rd_sampling(10, 3), golden["rd_10_3_s05"])


def test_k7_index_expressions(golden):
    _, i_prev, i_cur = inputs.step_indices(5000, 0.001, 20000)
    assert np.array_equal(i_prev, golden["k7_iprev"])                      # int((t-dt)/dt): n-1 or n-2
    assert np.array_equal(i_cur, golden["k7_icur"])
    n = np.arange(1, 5001)
    assert 0.7 < np.mean(i_prev == n - 1) < 0.95 and np.all((i_prev == n - 1) | (i_prev == n - 2))


def test_slam_input_tables_match_reference_closures(golden):
    """``inputs.slam_tables`` (vectorised) == get_slam_input_functions2 closures evaluated at t = n*dt."""
    dt, N = 0.001, 400
    sp = HexagonalSSPSpace(2, ssp_dim=55, domain_bounds=BOUNDS2, length_scale=0.2, backend="host")
    lm = SPSpace(50, 55, seed=0)
    path = inputs.random_path(20.0, dt, 0.1, 11, 2)
    assert np.array_equal(path[:N + 4], golden["in_path"])
    vels = inputs.velocities(path, dt)
    scale = inputs.velocity_scale(sp.phase_matrix, vels)
    np.testing.assert_allclose(scale, golden["in_vel_scale"], rtol=1e-15)
    obj = 1.8 * (inputs.rd_sampling(50, 2, seed=11) - 0.5)
    assert np.array_equal(obj, golden["in_obj"])
    vec_to = obj[None, :, :] - path[:, None, :]
    tb = inputs.slam_tables(sp.encode_host, lm.vectors, vels * scale, vec_to, 0.2, N, dt, real_ssp=sp.encode_host(path))
    np.testing.assert_allclose(tb["vel"], golden["in_vel"], rtol=0, atol=1e-15)
    assert np.array_equal(tb["nolm"][:, 0], golden["in_nolm"])
    np.testing.assert_allclose(tb["lm_sp"], golden["in_lm_sp"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(tb["lmvec_ssp"], golden["in_lmvec_ssp"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(tb["init"], golden["in_init"], rtol=0, atol=1e-14)
    assert golden["in_nolm"].min() == 0 and np.count_nonzero(golden["in_lm_sp"]) > 0   # the fixture sees landmarks


def test_white_signal_fixture(golden):
    from sspslam_b200.nengo_shim.processes import WhiteSignal
    got = WhiteSignal(20.0, high=0.1, seed=0).run(20.0, dt=0.001)[::100, 0]
    assert np.array_equal(got, golden["white_T20_h01_s0"])
