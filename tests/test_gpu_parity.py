"""Parity tests proper (``-m gpu``): the CUDA path, called through the C-ABI (ctypes ->
``libssb.so``), against the CPU oracle on the same built model, seeds and inputs.

Tolerances (BASELINE.json north_star): integer / index work (grid argmax, decode indices) is
bit-exact; decoded SSP trajectories agree within 1e-4 relative in rate mode over the stated horizon;
spiking (LIF) runs are fp32 vs fp64 and chaotic, so they are held to 2e-3 over 200 steps and to
identical spike masks over the first steps."""
import numpy as np
import pytest

from oracle import ssp_ref
from oracle.nengo_ref_sim import RefSimulator
from sspslam_b200 import scenarios, cabi
from sspslam_b200.sspspace import HexagonalSSPSpace

pytestmark = pytest.mark.gpu
BOUNDS2 = np.tile([-1.0, 1.0], (2, 1))


@pytest.fixture(scope="module", autouse=True)
def _built(lib):
    return lib


def _Simulator():
    from sspslam_b200.simulator import Simulator
    return Simulator


def _oracle(sc, sim, trial, n_steps):
    tabs = {node: arr[trial] for node, arr in sc.trial_inputs.items()}
    ref = RefSimulator(sc.network, dt=sc.dt, model=sim.model, node_tables=tabs, trial_seed=sim.trial_seeds[trial])
    ref.run_steps(n_steps)
    return ref


def _rel(got, want):
    return np.max(np.abs(got - want)) / max(1e-12, np.max(np.abs(want)))


# ------------------------------------------------------------------------------- SSP kernels
def test_ssp_encode_kernel_matches_oracle_and_reference(golden):
    sp = HexagonalSSPSpace(2, ssp_dim=55, domain_bounds=BOUNDS2, length_scale=0.2)        # backend='cuda'
    got = sp.encode(golden["enc_pts"])
    np.testing.assert_allclose(got, golden["enc55"], rtol=0, atol=1e-13)                  # unmodified reference output
    np.testing.assert_allclose(got, ssp_ref.encode(sp.phase_matrix, sp.length_scale, golden["enc_pts"]), atol=1e-13)
    sp3 = HexagonalSSPSpace(3, ssp_dim=55, domain_bounds=np.tile([-1.0, 1.0], (3, 1)), length_scale=0.3,
                            rng=np.random.default_rng(0))
    np.testing.assert_allclose(sp3.encode(golden["enc3d_pts"]), golden["enc3d"], rtol=0, atol=1e-13)
    one = sp.encode(np.array([[0.25, -0.5]]))
    assert one.shape == (1, 55) and abs(np.linalg.norm(one) - 1) < 1e-12
    assert cabi.ssp_encode(sp._scaled_phases(), np.zeros((0, 2))).shape == (0, 55)        # empty input


def test_ssp_decode_kernel_indices_are_bit_exact(golden):
    sp = HexagonalSSPSpace(2, ssp_dim=55, domain_bounds=BOUNDS2, length_scale=0.2)
    ssps, pts = sp.get_sample_pts_and_ssps(100, "grid")
    q = golden["dec55_in"]                                     # noisy, scaled, an all-zero row, a tiny-norm row
    idx = cabi.ssp_decode_argmax(ssps, q)
    assert np.array_equal(idx, golden["dec55_idx"])
    assert np.array_equal(pts[idx], golden["dec55_out"])
    assert np.array_equal(sp.decode(q, "from-set", "grid", 100), golden["dec55_out"])
    rng = np.random.default_rng(5)
    big = sp.encode_host(rng.uniform(-1, 1, (3001, 2))) + 0.2 * rng.standard_normal((3001, 55))   # ragged: 3001 rows
    assert np.array_equal(cabi.ssp_decode_argmax(ssps, big), ssp_ref.decode_indices(ssps, big))
    assert np.array_equal(cabi.ssp_decode_argmax(ssps, big[:1]), ssp_ref.decode_indices(ssps, big[:1]))
    assert cabi.ssp_decode_argmax(ssps, np.zeros((0, 55))).shape == (0,)


def test_ssp_decode_ties_first_maximum_wins():
    sp = HexagonalSSPSpace(2, ssp_dim=55, domain_bounds=BOUNDS2, length_scale=0.2, backend="host")
    ssps, _ = sp.get_sample_pts_and_ssps(30, "grid")
    dup = np.vstack([ssps, ssps[[17, 400, 899]]])              # exact duplicates later in the set
    q = ssps[[899, 17, 400, 5]] * 2.5
    assert list(cabi.ssp_decode_argmax(dup, q)) == [899, 17, 400, 5]
    assert list(ssp_ref.decode_indices(dup, q)) == [899, 17, 400, 5]
    # near ties: neighbouring grid rows, query exactly between them perturbed by 1e-9
    mid = 0.5 * (ssps[100] + ssps[101])
    qs = np.stack([mid + 1e-9 * (ssps[100] - ssps[101]), mid - 1e-9 * (ssps[100] - ssps[101])])
    assert np.array_equal(cabi.ssp_decode_argmax(ssps, qs), ssp_ref.decode_indices(ssps, qs))


@pytest.mark.parametrize("scan", ["tc", "ffma"])
def test_ssp_decode_generic_width_3d(scan, monkeypatch):
    if scan == "ffma":
        monkeypatch.setenv("SSB_SCAN", "ffma")             # the register-chunked generic-width FFMA scan (what d = 649 uses)
    sp = HexagonalSSPSpace(3, ssp_dim=55, domain_bounds=np.tile([-1.0, 1.0], (3, 1)), length_scale=0.3,
                           rng=np.random.default_rng(0), backend="host")       # d = 33 -> generic-width kernels
    ssps, pts = sp.get_sample_pts_and_ssps(12, "grid")
    rng = np.random.default_rng(1)
    q = sp.encode_host(rng.uniform(-1, 1, (257, 3))) + 0.05 * rng.standard_normal((257, sp.ssp_dim))
    assert np.array_equal(cabi.ssp_decode_argmax(ssps, q), ssp_ref.decode_indices(ssps, q))


def test_ssp_decode_d649_k_blocked_tensor_core_scan_is_bit_exact():
    """BASELINE configs[4] width (HexagonalSSPSpace n_rotates = n_scales = 9 -> d = 649): the operand tiles do not fit in
    shared memory, so the scan is the K-blocked tcgen05 kernel (k_cleanup_scan_tck); indices must still equal the float64
    argmax, ragged grid (15^3 = 3 375 rows: 26 tiles + a partial one) and ragged query count included."""
    sp = HexagonalSSPSpace(3, n_rotates=9, n_scales=9, domain_bounds=np.tile([-1.0, 1.0], (3, 1)), length_scale=0.3,
                           rng=np.random.default_rng(0), backend="host")
    assert sp.ssp_dim == 649
    ssps, pts = sp.get_sample_pts_and_ssps(15, "grid")
    rng = np.random.default_rng(2)
    q = sp.encode_host(rng.uniform(-1, 1, (300, 3))) + 0.02 * rng.standard_normal((300, 649))
    q[7] = 0.0                                                   # an all-zero query picks row 0
    q[8] = ssps[1234] * 3.0                                      # an exact grid point, scaled
    got = cabi.ssp_decode_argmax(ssps, q)
    assert np.array_equal(got, ssp_ref.decode_indices(ssps, q))
    assert got[8] == 1234


# ------------------------------------------------------------------------------- stepped path
@pytest.mark.parametrize("neuron_type,tol", [("lifrate", 1e-4), ("relu", 1e-4), ("lif", 2e-3)])
def test_pathintegration_matches_oracle(neuron_type, tol):
    n_steps = 300 if neuron_type != "lif" else 200
    sc = scenarios.make_pathint(n_trials=3, n_steps=n_steps, ssp_dim=55, pi_n_neurons=200, neuron_type=neuron_type)
    with _Simulator()(sc.network, dt=sc.dt, n_trials=3, trial_inputs=sc.trial_inputs) as sim:
        sim.run_steps(n_steps)
        assert sim.total_launches() > 0
    got = sim.data[sc.probe]                       # readable after close (run_slam.py:242-252)
    assert got.shape == (3, n_steps, 55)
    for trial in (0, 2):
        want = _oracle(sc, sim, trial, n_steps).data[sc.probe]
        assert np.max(np.abs(want)) > 0.1
        assert _rel(got[trial], want) < tol


def test_slam_rate_mode_matches_oracle_and_cleanup_indices_are_exact():
    n_steps = 200
    sc = scenarios.make_slam(n_trials=3, n_steps=n_steps, ssp_dim=55, pi_n_neurons=100, mem_n_neurons=200,
                             circonv_n_neurons=30, n_landmarks=20, T=20.0, neuron_type="lifrate")
    slam = sc.extra["slam"]
    sim = _Simulator()(sc.network, dt=sc.dt, n_trials=3, trial_inputs=sc.trial_inputs)
    idx_trace = []
    for _ in range(4):
        sim.run_steps(n_steps // 4)
        idx_trace.append(sim.cleanup_indices()[0].copy())
    got = sim.data[sc.probe]
    dec = sim.learned_decoders(slam.assomemory.conn_out)
    enc = sim.learned_encoders(slam.assomemory.memory)
    sim.close()
    for trial in (0, 1):
        tabs = {node: arr[trial] for node, arr in sc.trial_inputs.items()}
        ref = RefSimulator(sc.network, dt=sc.dt, model=sim.model, node_tables=tabs, trial_seed=sim.trial_seeds[trial])
        for k in range(4):
            ref.run_steps(n_steps // 4)
            x = ref.signals[slam.gridcells, "in"].a
            assert idx_trace[k][trial] == ssp_ref.cleanup_index(slam.sample_ssps, x)       # bit-exact index
        assert _rel(got[trial], ref.data[sc.probe]) < 1e-4
        # learned matrices: the device applies the delta of step t at step t+1, like nengo's Copy(inc)
        want_enc = ref.scaled_encoders(slam.assomemory.memory)
        assert _rel(enc[trial], want_enc) < 1e-4
        want_dec = ref.learned_weights(slam.assomemory.conn_out)
        assert np.max(np.abs(want_dec)) > 0
        assert np.max(np.abs(dec[trial] - want_dec)) < 1e-4 * np.max(np.abs(want_dec)) + 1e-9


def test_slam_spiking_matches_oracle_short_horizon():
    n_steps = 200
    sc = scenarios.make_slam(n_trials=3, n_steps=n_steps, ssp_dim=55, pi_n_neurons=100, mem_n_neurons=200,
                             circonv_n_neurons=30, n_landmarks=20, T=20.0, neuron_type="lif", weights_probe=True)
    slam = sc.extra["slam"]
    sim = _Simulator()(sc.network, dt=sc.dt, n_trials=3, trial_inputs=sc.trial_inputs)
    sim.run_steps(12)
    spikes12 = sim.activities(slam.ovc_ens) > 0
    sim.run_steps(n_steps - 12)
    got = sim.data[sc.probe]
    w = sim.data[sc.extra["weights_probe"]]
    sim.close()
    assert w.shape == (3, 1, 55, 200)                       # Probe(conn,'weights',sample_every=T): (1, 55, n) per trial
    for trial in (0, 1):
        tabs = {node: arr[trial] for node, arr in sc.trial_inputs.items()}
        ref = RefSimulator(sc.network, dt=sc.dt, model=sim.model, node_tables=tabs, trial_seed=sim.trial_seeds[trial])
        ref.run_steps(12)
        assert np.array_equal(spikes12[trial], ref.spikes(slam.ovc_ens))                   # identical spike mask
        ref.run_steps(n_steps - 12)
        assert _rel(got[trial], ref.data[sc.probe]) < 2e-3


def test_slamview_rate_mode_matches_oracle():
    n_steps = 150
    sc = scenarios.make_slam(n_trials=2, n_steps=n_steps, ssp_dim=55, pi_n_neurons=100, mem_n_neurons=200,
                             circonv_n_neurons=30, n_landmarks=20, T=20.0, neuron_type="lifrate", view=True)
    with _Simulator()(sc.network, dt=sc.dt, n_trials=2, trial_inputs=sc.trial_inputs) as sim:
        sim.run_steps(n_steps)
    got = sim.data[sc.probe]
    for trial in (0, 1):
        assert _rel(got[trial], _oracle(sc, sim, trial, n_steps).data[sc.probe]) < 1e-4


def test_generic_width_network_matches_oracle():
    """d = 31 is neither of the register-specialised widths (56 / 100): generic kernels."""
    sc = scenarios.make_slam(n_trials=2, n_steps=120, ssp_dim=31, pi_n_neurons=60, mem_n_neurons=120,
                             circonv_n_neurons=20, n_landmarks=8, T=20.0, neuron_type="lifrate")
    with _Simulator()(sc.network, dt=sc.dt, n_trials=2, trial_inputs=sc.trial_inputs) as sim:
        sim.run_steps(120)
    got = sim.data[sc.probe]
    assert _rel(got[1], _oracle(sc, sim, 1, 120).data[sc.probe]) < 1e-4


@pytest.mark.parametrize("env", [{}, {"SSB_LIN": "ffma", "SSB_ENCODE": "ffma", "SSB_SCAN": "ffma", "SSB_DECODE": "ffma"},
                                 {"SSB_VOJA": "cta"}])
def test_wide_d295_network_on_the_k_blocked_tensor_core_kernels(env, monkeypatch):
    """d = 295 (2-D, 7 x 7 scale / rotation pairs) at reduced neuron counts: wide enough for every K-blocked tcgen05 path of
    BASELINE configs[4] (d = 649) - grid scan (k_cleanup_scan_tck), static wide encode (k_wide_static_tck), column-tiled
    decode (k_decode_tc with 3 tiles), the 592 x 295 / 295 x 592 dense blocks of the row program (k_lin_tck) - against the
    oracle in rate mode, and the same network on the FFMA kernels; SSB_VOJA=cta puts the memory ensemble on the
    CTA-cooperative Voja kernel of d = 649 (k_wide_voja_cta, 48-row slices), incl. the learned encoders."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    n_steps = 60
    sc = scenarios.make_slam(n_trials=3, n_steps=n_steps, ssp_dim=295, pi_n_neurons=30, mem_n_neurons=96,
                             circonv_n_neurons=10, n_landmarks=8, T=20.0, neuron_type="lifrate", view_rad=0.6,
                             grid_points_per_dim=40)
    assert sc.ssp_space.ssp_dim == 295
    slam = sc.extra["slam"]
    with _Simulator()(sc.network, dt=sc.dt, n_trials=3, trial_inputs=sc.trial_inputs) as sim:
        sim.run_steps(n_steps)
        idx = sim.cleanup_indices()[0].copy()
        dec = sim.learned_decoders(slam.assomemory.conn_out)
        enc = sim.learned_encoders(slam.assomemory.memory)
    got = sim.data[sc.probe]
    for trial in (0, 2):
        ref = _oracle(sc, sim, trial, n_steps)
        want = ref.data[sc.probe]
        assert np.max(np.abs(want)) > 0.05
        assert _rel(got[trial], want) < 1e-4
        assert idx[trial] == ssp_ref.cleanup_index(slam.sample_ssps, ref.signals[slam.gridcells, "in"].a)
        want_dec = ref.learned_weights(slam.assomemory.conn_out)
        assert np.max(np.abs(dec[trial] - want_dec)) < 1e-4 * np.max(np.abs(want_dec)) + 1e-9
        want_enc = ref.scaled_encoders(slam.assomemory.memory)
        assert not np.allclose(want_enc, sim.model.params[slam.assomemory.memory].scaled_encoders)      # Voja moved them
        assert _rel(enc[trial], want_enc) < 1e-4


@pytest.mark.parametrize("env", [{"SSB_SCAN": "ffma"}, {"SSB_DECODE": "ffma"}, {"SSB_ENCODE": "tc"},
                                 {"SSB_SCAN": "ffma", "SSB_DECODE": "ffma", "SSB_SERIAL": "1"},
                                 {"SSB_PES_DEFER": "4"}, {"SSB_PES_FUSE": "1"}, {"SSB_LEVEL_DEPS": "0"},
                                 {"SSB_DECODE": "sparse"}, {"SSB_VOJA_NB": "3"}, {"SSB_PES_FOLD": "tasks"},
                                 {"SSB_LIN_FUSE": "1"}])
def test_alternate_kernel_paths_match_oracle(env, monkeypatch):
    """Every shared-weight GEMM has an FFMA and a tcgen05 (3xTF32) kernel; whichever is selected, the
    trajectory stays within the rate-mode tolerance and the clean-up index is the float64 argmax."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    n_steps = 120
    sc = scenarios.make_slam(n_trials=5, n_steps=n_steps, ssp_dim=55, pi_n_neurons=80, mem_n_neurons=160,
                             circonv_n_neurons=24, n_landmarks=12, T=20.0, neuron_type="lifrate")
    slam = sc.extra["slam"]
    with _Simulator()(sc.network, dt=sc.dt, n_trials=5, trial_inputs=sc.trial_inputs) as sim:
        sim.run_steps(n_steps)
        idx = sim.cleanup_indices()[0].copy()
    got = sim.data[sc.probe]
    for trial in (0, 4):
        ref = _oracle(sc, sim, trial, n_steps)
        assert _rel(got[trial], ref.data[sc.probe]) < 1e-4
        x = ref.signals[slam.gridcells, "in"].a
        assert idx[trial] == ssp_ref.cleanup_index(slam.sample_ssps, x)


def test_ssp_decode_ffma_scan_is_bit_exact_too(golden, monkeypatch):
    monkeypatch.setenv("SSB_SCAN", "ffma")
    sp = HexagonalSSPSpace(2, ssp_dim=55, domain_bounds=BOUNDS2, length_scale=0.2)
    ssps, _ = sp.get_sample_pts_and_ssps(100, "grid")
    assert np.array_equal(cabi.ssp_decode_argmax(ssps, golden["dec55_in"]), golden["dec55_idx"])


# ------------------------------------------------------------------------------- simulator surface
def test_simulator_surface_run_trange_reset_unbatched():
    sc = scenarios.make_pathint(n_trials=1, n_steps=120, ssp_dim=19, pi_n_neurons=50, neuron_type="lif")
    sim = _Simulator()(sc.network, dt=sc.dt)                      # unbatched: nengo's own API, closures evaluated
    with sim:
        sim.run(0.05)
        assert sim.n_steps == 50 and abs(sim.time - 0.05) < 1e-12
        first = sim.data[sc.probe].copy()
        assert first.shape == (50, sc.ssp_space.ssp_dim)
        np.testing.assert_allclose(sim.trange(), 0.001 * np.arange(1, 51))
        sim.step()
        assert sim.n_steps == 51
        sim.reset()
        assert sim.n_steps == 0
        sim.run(0.05)
        assert np.array_equal(sim.data[sc.probe], first)         # reset reproduces the run bit for bit
    assert np.array_equal(sim.data[sc.probe], first)
    with pytest.raises(Exception):
        sim.run_steps(1)                                         # closed
    ref = RefSimulator(sc.network, dt=sc.dt, model=sim.model)     # oracle calls the same closures
    ref.run_steps(50)
    assert _rel(first, ref.data[sc.probe]) < 2e-3
    built = sim.data[sc.extra["pathint"].oscillators.ea_ensembles[1]]
    assert built.encoders.shape == (50, 3) and built.gain.shape == (50,)


def test_batch_independence_and_determinism():
    """A trial's trajectory does not depend on the batch it runs in (ragged batch sizes, chunking)."""
    sc = scenarios.make_slam(n_trials=40, n_steps=64, ssp_dim=55, pi_n_neurons=60, mem_n_neurons=128,
                             circonv_n_neurons=20, n_landmarks=10, T=20.0, neuron_type="lif", distinct_tables=5)
    S = _Simulator()
    seeds = list(range(100, 140))
    with S(sc.network, dt=sc.dt, n_trials=40, trial_inputs=sc.trial_inputs, trial_seeds=seeds, chunk_steps=64) as a:
        a.run_steps(64)
    with S(sc.network, dt=sc.dt, n_trials=40, trial_inputs=sc.trial_inputs, trial_seeds=seeds, chunk_steps=10) as b:
        b.run_steps(64)                                           # different chunking (graph replay vs direct launches)
    assert np.array_equal(a.data[sc.probe], b.data[sc.probe])
    sub = [3, 17, 39]
    sub_inputs = {n: arr[sub] for n, arr in sc.trial_inputs.items()}
    with S(sc.network, dt=sc.dt, n_trials=3, trial_inputs=sub_inputs, trial_seeds=[seeds[i] for i in sub],
           model=a.model) as c:
        c.run_steps(64)
    got, want = c.data[sc.probe], a.data[sc.probe][sub]
    assert _rel(got, want) < 2e-3     # decoder chunking differs with batch size -> fp32 summation order, not bits
    assert np.all(np.isfinite(a.data[sc.probe]))


def test_full_size_config2_properties():
    """BASELINE configs[1] sizes (40 280 neurons / trial): finite, learning active, tracks the true SSP."""
    sc = scenarios.make_slam(n_trials=32, n_steps=300, T=200.0, distinct_tables=4)
    slam = sc.extra["slam"]
    with _Simulator()(sc.network, dt=sc.dt, n_trials=32, trial_inputs=sc.trial_inputs) as sim:
        assert sim.plan.stats["n_neurons"] == 40280 and sim.plan.stats["n_learned"] == 106700
        sim.run_steps(300)
        dec = sim.learned_decoders(slam.assomemory.conn_out)
        idx = sim.cleanup_indices()
    out = sim.data[sc.probe]
    assert np.all(np.isfinite(out)) and np.all(np.isfinite(dec))
    assert np.max(np.abs(dec)) > 0
    assert np.all((idx >= 0) & (idx < 10000))
    real = sc.real_ssp[:, :300]
    cos = np.sum(out[:, -1] * real[:, -1], axis=1) / (np.linalg.norm(out[:, -1], axis=1) * np.linalg.norm(real[:, -1], axis=1))
    assert np.mean(cos) > 0.6


def test_full_size_config1_pathint_d97_rate_mode_matches_oracle():
    """BASELINE configs[0]: run_pathint.py defaults (d = 97, 49 VCOs x 500): generic-width kernels at full size."""
    n_steps = 150
    sc = scenarios.make_pathint(n_trials=32, n_steps=n_steps, ssp_dim=97, pi_n_neurons=500, neuron_type="lifrate")
    with _Simulator()(sc.network, dt=sc.dt, n_trials=32, trial_inputs=sc.trial_inputs) as sim:
        assert sim.plan.stats["n_neurons"] == 24500
        sim.run_steps(n_steps)
    got = sim.data[sc.probe]
    assert got.shape == (32, n_steps, 97) and np.all(np.isfinite(got))
    for trial in (0, 31):
        assert _rel(got[trial], _oracle(sc, sim, trial, n_steps).data[sc.probe]) < 1e-4


def test_full_size_config4_slamview_d97_rate_mode_matches_oracle():
    """BASELINE configs[3] sizes: SLAMViewNetwork, d = 97, pi 800, mem 970, 100 landmarks, PES + Voja per trial."""
    n_steps = 100
    sc = scenarios.make_slam(n_trials=32, n_steps=n_steps, ssp_dim=97, pi_n_neurons=800, mem_n_neurons=970,
                             circonv_n_neurons=100, n_landmarks=100, T=20.0, length_scale=0.3, neuron_type="lifrate",
                             view=True, distinct_tables=4)
    slam = sc.extra["slam"]
    with _Simulator()(sc.network, dt=sc.dt, n_trials=32, trial_inputs=sc.trial_inputs) as sim:
        assert sim.plan.stats["n_neurons"] == 42110
        sim.run_steps(n_steps)
        dec = sim.learned_decoders(slam.assomemory.conn_out)
    got = sim.data[sc.probe]
    assert np.all(np.isfinite(got)) and np.all(np.isfinite(dec))
    assert _rel(got[1], _oracle(sc, sim, 1, n_steps).data[sc.probe]) < 1e-4


def test_spiking_error_distribution_matches_oracle():
    """Spiking runs are chaotic (fp32 vs fp64), so they are compared as a distribution over trials: the mean
    similarity to the true SSP and the mean decoded position error over the last 200 steps must agree with the
    oracle's within 0.02 / 0.03 (BASELINE north_star: 'reproduce the reference's RMSE distribution')."""
    from sspslam_b200 import results
    n_trials, n_steps = 8, 400
    sc = scenarios.make_pathint(n_trials=n_trials, n_steps=n_steps, ssp_dim=55, pi_n_neurons=100, neuron_type="lif")
    with _Simulator()(sc.network, dt=sc.dt, n_trials=n_trials, trial_inputs=sc.trial_inputs) as sim:
        sim.run_steps(n_steps)
    sp = HexagonalSSPSpace(2, ssp_dim=55, domain_bounds=BOUNDS2, length_scale=0.2, rng=np.random.default_rng(0))

    class _Ref:                                      # the oracle behind the same results.* interface
        def __init__(self, ref):
            self.data, self._n = ref.data, n_steps
        def trange(self):
            return sc.dt * np.arange(1, self._n + 1)

    gpu_sims, gpu_err, ref_sims, ref_err = [], [], [], []
    for trial in range(n_trials):
        r = results.pathint_results(sim, sc.probe, sp, sc.paths[trial], sc.real_ssp[trial], trial=trial)
        o = results.pathint_results(_Ref(_oracle(sc, sim, trial, n_steps)), sc.probe, sp, sc.paths[trial], sc.real_ssp[trial])
        gpu_sims.append(np.mean(r["pi_sims"][-200:])); gpu_err.append(np.mean(r["pi_error"][-200:]))
        ref_sims.append(np.mean(o["pi_sims"][-200:])); ref_err.append(np.mean(o["pi_error"][-200:]))
    assert np.mean(ref_sims) > 0.5                                   # the network does integrate the path
    assert abs(np.mean(gpu_sims) - np.mean(ref_sims)) < 0.02
    assert abs(np.mean(gpu_err) - np.mean(ref_err)) < 0.03
    assert np.max(np.abs(np.array(gpu_sims) - np.array(ref_sims))) < 0.06


def test_slam_results_with_landmark_estimates():
    """run_slam.py:236-293 on a finished batched run: decoded path, error, similarity, landmark recall."""
    from sspslam_b200 import results
    n_steps = 160
    sc = scenarios.make_slam(n_trials=2, n_steps=n_steps, ssp_dim=55, pi_n_neurons=60, mem_n_neurons=128,
                             circonv_n_neurons=20, n_landmarks=10, T=20.0, neuron_type="lif", weights_probe=True)
    with _Simulator()(sc.network, dt=sc.dt, n_trials=2, trial_inputs=sc.trial_inputs) as sim:
        sim.run_steps(n_steps)
    sp = HexagonalSSPSpace(2, ssp_dim=55, domain_bounds=BOUNDS2, length_scale=0.2, rng=np.random.default_rng(0))
    res = results.slam_results(sim, sc.probe, sp, sc.paths[1], sc.real_ssp[1], trial=1, slam=sc.extra["slam"],
                               weights_probe=sc.extra["weights_probe"], lm_vectors=sc.extra["lm_space"].vectors)
    assert res["slam_path"].shape == (n_steps, 2) and res["slam_error"].shape == (n_steps,)
    assert res["landmark_ssps_est"].shape == (10, 55) and res["landmark_loc_est"].shape == (10, 2)
    assert np.all(np.isfinite(res["slam_sims"])) and np.all(np.isfinite(res["landmark_ssps_est"]))


def test_slam_3d_domain_matches_oracle():
    """BASELINE configs[4] topology (3-D domain, HexagonalSSPSpace with random rotations) at reduced size."""
    n_steps = 120
    sc = scenarios.make_slam(n_trials=3, n_steps=n_steps, ssp_dim=55, pi_n_neurons=60, mem_n_neurons=128,
                             circonv_n_neurons=20, n_landmarks=8, T=20.0, neuron_type="lifrate", view_rad=0.6,
                             domain_dim=3, grid_points_per_dim=14)
    slam = sc.extra["slam"]
    with _Simulator()(sc.network, dt=sc.dt, n_trials=3, trial_inputs=sc.trial_inputs) as sim:
        sim.run_steps(n_steps)
        idx = sim.cleanup_indices()[0].copy()
    got = sim.data[sc.probe]
    for trial in (0, 2):
        ref = _oracle(sc, sim, trial, n_steps)
        assert _rel(got[trial], ref.data[sc.probe]) < 1e-4
        assert idx[trial] == ssp_ref.cleanup_index(slam.sample_ssps, ref.signals[slam.gridcells, "in"].a)


def test_deferred_pes_read_back_in_the_middle_of_a_window():
    """The PES history (window of 8 steps) is folded into the decoders before any read-back; reading at steps that
    are not multiples of the window, then continuing, must still follow the oracle's learned weights."""
    sc = scenarios.make_slam(n_trials=2, n_steps=64, ssp_dim=55, pi_n_neurons=60, mem_n_neurons=128,
                             circonv_n_neurons=20, n_landmarks=8, T=20.0, neuron_type="lifrate", view_rad=0.8)
    conn = sc.extra["slam"].assomemory.conn_out
    sim = _Simulator()(sc.network, dt=sc.dt, n_trials=2, trial_inputs=sc.trial_inputs)
    ref = RefSimulator(sc.network, dt=sc.dt, model=sim.model, trial_seed=sim.trial_seeds[1],
                       node_tables={node: arr[1] for node, arr in sc.trial_inputs.items()})
    seen_learning = False
    for n in (13, 16, 3, 21):                       # cumulative 13, 29, 32, 53
        sim.run_steps(n)
        ref.run_steps(n)
        got, want = sim.learned_decoders(conn)[1], ref.learned_weights(conn)
        seen_learning = seen_learning or np.max(np.abs(want)) > 0
        assert np.max(np.abs(got - want)) < 1e-4 * np.max(np.abs(want)) + 1e-9
    assert seen_learning
    assert _rel(sim.data[sc.probe][1], ref.data[sc.probe]) < 1e-4
    sim.close()


@pytest.mark.parametrize("kind", ["pathint", "slam", "slamview"])
def test_on_device_input_synthesis_matches_tables_and_oracle(kind):
    """the input closures (slam.py:442-497, run_pathint.py:134-136) evaluated by k_synth from the
    per-trial path / landmarks must drive the network like the host tables the oracle reads (fp32 sincos / IDFT
    instead of the float64 FFT: inside the rate-mode tolerance)."""
    n_steps = 130                                      # covers the t < 0.05 s init window and landmarks in view
    if kind == "pathint":
        sc = scenarios.make_pathint(n_trials=3, n_steps=n_steps, ssp_dim=55, pi_n_neurons=100, neuron_type="lifrate")
    else:
        sc = scenarios.make_slam(n_trials=3, n_steps=n_steps, ssp_dim=55, pi_n_neurons=80, mem_n_neurons=160,
                                 circonv_n_neurons=24, n_landmarks=12, T=20.0, neuron_type="lifrate", view_rad=0.5,
                                 view=(kind == "slamview"))
    S = _Simulator()
    with S(sc.network, dt=sc.dt, n_trials=3, trial_inputs=sc.trial_inputs) as a:
        a.run_steps(n_steps)
    with S(sc.network, dt=sc.dt, n_trials=3, input_synthesis=sc.extra["input_synthesis"], model=a.model,
           chunk_steps=48) as b:                       # 48-step chunks: indices re-uploaded per chunk, graph + direct steps
        b.run_steps(n_steps)
    got_tab, got_syn = a.data[sc.probe], b.data[sc.probe]
    assert np.max(np.abs(got_tab)) > 0.1
    assert _rel(got_syn, got_tab) < 1e-4
    for trial in (0, 2):
        assert _rel(got_syn[trial], _oracle(sc, a, trial, n_steps).data[sc.probe]) < 1e-4
    if kind != "pathint":                              # some landmark was in view, so the synthesised sums were exercised
        nolm = [n for n in sc.trial_inputs if n.label == "lm_in_view_input"][0]
        assert np.any(sc.trial_inputs[nolm][:, :n_steps] == 0.0)


def test_grid_cell_ensemble_and_approx_velocity_variants_match_oracle():
    """gc_n_neurons > 0 (slam.py:274-281) and --approx-vel (run_slam.py:154-160) on the same kernels."""
    n_steps = 120
    sc = scenarios.make_slam(n_trials=3, n_steps=n_steps, ssp_dim=55, pi_n_neurons=80, mem_n_neurons=160,
                             circonv_n_neurons=24, n_landmarks=12, T=20.0, neuron_type="lifrate", view_rad=0.5,
                             gc_n_neurons=200, approx_vel=True, vel_n_neurons=100)
    with _Simulator()(sc.network, dt=sc.dt, n_trials=3, trial_inputs=sc.trial_inputs) as sim:
        assert sim.plan.stats["n_big"] == 5
        sim.run_steps(n_steps)
    got = sim.data[sc.probe]
    for trial in (0, 2):
        assert _rel(got[trial], _oracle(sc, sim, trial, n_steps).data[sc.probe]) < 1e-4


@pytest.mark.parametrize("neuron_type", ["lifrate", "lif"])
def test_neuron_output_and_sliced_vco_probes_match_oracle(neuron_type):
    """the probes of run_pathint_gif.py:156-159 — a sliced, filtered VCO output probe and neuron-output
    probes (``ea_ensembles[k].neurons[:m]``, synapse=None, sample_every) served by the row program from the act arena."""
    from sspslam_b200 import nengo_shim as nengo
    n_steps = 120
    sc = scenarios.make_pathint(n_trials=3, n_steps=n_steps, ssp_dim=55, pi_n_neurons=200, neuron_type=neuron_type)
    pi = sc.extra["pathint"]
    with sc.network:
        vco_p = nengo.Probe(pi.oscillators.output[3:12], synapse=0.05)
        n1 = nengo.Probe(pi.oscillators.ea_ensembles[1].neurons[:150], synapse=None, sample_every=4 * sc.dt)
        n2 = nengo.Probe(pi.oscillators.ea_ensembles[2].neurons, synapse=None)
    with _Simulator()(sc.network, dt=sc.dt, n_trials=3, trial_inputs=sc.trial_inputs) as sim:
        assert sim.plan.stats["n_big"] == 2
        sim.run_steps(n_steps)
    assert sim.data[n1].shape == (3, n_steps // 4, 150) and sim.data[n2].shape == (3, n_steps, 200)
    assert sim.trange(sample_every=4 * sc.dt).shape == (n_steps // 4,)
    for trial in (0, 2):
        ref = _oracle(sc, sim, trial, n_steps)
        if neuron_type == "lifrate":
            for p in (sc.probe, vco_p, n1, n2):
                assert np.max(np.abs(ref.data[p])) > 0
                assert _rel(sim.data[p][trial], ref.data[p]) < 1e-4
        else:
            assert _rel(sim.data[sc.probe][trial], ref.data[sc.probe]) < 2e-3
            for p in (n1, n2):          # spikes: amplitude / dt or 0; a float32 voltage may cross one step early / late
                got, want = sim.data[p][trial], ref.data[p]
                assert np.allclose(got[got != 0], 1.0 / sc.dt, rtol=1e-6)     # float32 amplitude / dt
                assert want.sum() > 0
                assert abs(got.sum() - want.sum()) <= 0.02 * want.sum()
                assert np.mean((got != 0) != (want != 0)) < 2e-3


def test_slam_loihi_variant_matches_oracle():
    """SLAMLoihiNetwork (slam_loihi.py:190-293) on the same kernels — no node functions, the update gate is
    a threshold population that inhibits the correction neurons through a filtered ensemble -> neurons connection."""
    n_steps = 150
    sc = scenarios.make_slam(n_trials=3, n_steps=n_steps, ssp_dim=55, pi_n_neurons=80, mem_n_neurons=160,
                             circonv_n_neurons=24, n_landmarks=12, T=20.0, neuron_type="lifrate", view_rad=0.5,
                             loihi=True, dotprod_n_neurons=20)
    with _Simulator()(sc.network, dt=sc.dt, n_trials=3, trial_inputs=sc.trial_inputs) as sim:
        assert sim.plan.stats["n_big"] == 4
        sim.run_steps(n_steps)
    got = sim.data[sc.probe]
    for trial in (0, 2):
        want = _oracle(sc, sim, trial, n_steps).data[sc.probe]
        assert np.max(np.abs(want)) > 0.1
        assert _rel(got[trial], want) < 1e-4


def test_inverse_memory_topology_matches_oracle():
    """experiments/slam_map_new.py:207-263 — a second path integrator and a second Voja + PES memory on the
    same kernels (two PES rules, two Voja rules), with that script's decoded / node / weights / scaled_encoders probes."""
    n_steps = 128
    sc = scenarios.make_slam(n_trials=3, n_steps=n_steps, ssp_dim=55, pi_n_neurons=80, mem_n_neurons=160,
                             circonv_n_neurons=24, n_landmarks=12, T=20.0, neuron_type="lifrate", view_rad=0.5,
                             inverse_memory=True)
    with _Simulator()(sc.network, dt=sc.dt, n_trials=3, trial_inputs=sc.trial_inputs) as sim:
        assert len(sim.plan.arrays["pes"]) == 2
        sim.run_steps(n_steps)
    x = sc.extra
    assert sim.data[x["mem_weights"]].shape == (3, 2, 55, 160) and sim.data[x["meminv_encoders"]].shape == (3, 1, 160, 55)
    for trial in (0, 2):
        ref = _oracle(sc, sim, trial, n_steps)
        for name in ("ssp_pi_p", "newpos_p", "objssp_p", "recall_p", "isitem_p"):
            assert _rel(sim.data[x[name]][trial], ref.data[x[name]]) < 1e-4, name
        assert _rel(sim.data[sc.probe][trial], ref.data[sc.probe]) < 1e-4
        for name in ("mem_weights", "meminv_weights", "mem_encoders", "meminv_encoders"):
            want = ref.data[x[name]]
            assert np.max(np.abs(want)) > 0
            assert _rel(sim.data[x[name]][trial], want) < 1e-4, name


def test_pathintegration_with_grid_cell_output_matches_oracle():
    """pathintegration.py:150-154 (``with_gcs=True``): the VCO array drives a grid-cell output population whose decoded,
    filtered output is probed."""
    n_steps = 200
    sc = scenarios.make_pathint(n_trials=3, n_steps=n_steps, ssp_dim=55, pi_n_neurons=200, neuron_type="lifrate",
                                with_gcs=True, n_gcs=400)
    with _Simulator()(sc.network, dt=sc.dt, n_trials=3, trial_inputs=sc.trial_inputs) as sim:
        assert sim.plan.stats["n_big"] == 1
        sim.run_steps(n_steps)
    got = sim.data[sc.probe]
    for trial in (0, 2):
        want = _oracle(sc, sim, trial, n_steps).data[sc.probe]
        assert np.max(np.abs(want)) > 0.05
        assert _rel(got[trial], want) < 1e-4


def test_slam_without_voja_matches_oracle():
    """run_slam.py --no-voja: fixed landmark encoders (slam.py:196-198); only the PES decoders are learned."""
    n_steps = 120
    sc = scenarios.make_slam(n_trials=3, n_steps=n_steps, ssp_dim=55, pi_n_neurons=80, mem_n_neurons=160,
                             circonv_n_neurons=24, n_landmarks=12, T=20.0, neuron_type="lifrate", view_rad=0.5, voja=False)
    with _Simulator()(sc.network, dt=sc.dt, n_trials=3, trial_inputs=sc.trial_inputs) as sim:
        assert sim.plan.stats["n_learned"] == 160 * 55
        sim.run_steps(n_steps)
    got = sim.data[sc.probe]
    for trial in (0, 2):
        assert _rel(got[trial], _oracle(sc, sim, trial, n_steps).data[sc.probe]) < 1e-4


def test_slamview_with_bound_view_input_matches_oracle():
    """run_slamview.py:103,123-130 inputs (normalised sum of SP (*) SSP(displacement), slam_view.py:384-394) as tables."""
    n_steps = 120
    sc = scenarios.make_slam(n_trials=3, n_steps=n_steps, ssp_dim=55, pi_n_neurons=80, mem_n_neurons=160,
                             circonv_n_neurons=24, n_landmarks=12, T=20.0, neuron_type="lifrate", view=True, view_rad=0.6,
                             view_bound=True)
    with _Simulator()(sc.network, dt=sc.dt, n_trials=3, trial_inputs=sc.trial_inputs) as sim:
        sim.run_steps(n_steps)
    got = sim.data[sc.probe]
    for trial in (0, 2):
        assert _rel(got[trial], _oracle(sc, sim, trial, n_steps).data[sc.probe]) < 1e-4
