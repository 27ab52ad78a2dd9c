"""SSP algebra: oracle (oracle/ssp_ref.py) and the package's SSPSpace against vectors produced by
the UNMODIFIED reference ``sspslam/sspspace.py`` (tests/golden/make_golden.py), plus the analytic
known-answer identities K1, K2, K6 of SURVEY.md §4."""
import numpy as np
import pytest

from oracle import ssp_ref
from sspslam_b200.sspspace import HexagonalSSPSpace, SPSpace

BOUNDS2 = np.tile([-1.0, 1.0], (2, 1))


@pytest.fixture(scope="module")
def sp55():
    return HexagonalSSPSpace(2, ssp_dim=55, domain_bounds=BOUNDS2, length_scale=0.2, backend="host")


def test_phase_matrices_match_reference(golden, sp55):
    assert np.array_equal(sp55.phase_matrix, golden["hex55_phase"])
    sp97 = HexagonalSSPSpace(2, ssp_dim=97, domain_bounds=BOUNDS2, length_scale=0.2, backend="host")
    assert np.array_equal(sp97.phase_matrix, golden["hex97_phase"])
    sp3d = HexagonalSSPSpace(3, ssp_dim=55, domain_bounds=np.tile([-1.0, 1.0], (3, 1)), length_scale=0.3,
                             rng=np.random.default_rng(0), backend="host")
    np.testing.assert_allclose(sp3d.phase_matrix, golden["hex3d_phase"], rtol=0, atol=1e-14)
    dims = [sp55.ssp_dim, sp97.ssp_dim, sp3d.ssp_dim,
            HexagonalSSPSpace(3, n_rotates=9, n_scales=9, backend="host").ssp_dim]
    assert dims == list(golden["hex_dims"]) == [55, 97, 33, 649]          # K6


def test_phase_matrix_conjugate_symmetry(sp55):
    A = sp55.phase_matrix
    assert np.all(A[0] == 0)
    assert np.array_equal(A[1:], -A[1:][::-1])                             # K6: rows k and d-k are negatives


def test_oracle_encode_matches_reference(golden, sp55):
    got = ssp_ref.encode(golden["hex55_phase"], sp55.length_scale, golden["enc_pts"])
    np.testing.assert_allclose(got, golden["enc55"], rtol=0, atol=1e-15)
    got3 = ssp_ref.encode(golden["hex3d_phase"], 0.3, golden["enc3d_pts"])
    np.testing.assert_allclose(got3, golden["enc3d"], rtol=0, atol=1e-15)


def test_host_encode_matches_reference(golden, sp55):
    np.testing.assert_allclose(sp55.encode_host(golden["enc_pts"]), golden["enc55"], rtol=0, atol=2e-15)


def test_grid_matches_reference(golden, sp55):
    ssps, pts = sp55.get_sample_pts_and_ssps(100, "grid")
    assert pts.shape == (10000, 2) and ssps.shape == (10000, 55)
    assert np.array_equal(pts[:205], golden["grid55_pts_head"])
    np.testing.assert_allclose(ssps[[0, 1, 99, 100, 5050, 9999]], golden["grid55_ssps_rows"], rtol=0, atol=2e-15)
    assert np.array_equal(ssp_ref.grid_points(BOUNDS2, 100), pts)


def test_decode_matches_reference(golden, sp55):
    ssps, pts = sp55.get_sample_pts_and_ssps(100, "grid")
    idx = ssp_ref.decode_indices(ssps, golden["dec55_in"])
    assert np.array_equal(pts[idx], golden["dec55_out"])                  # bit-exact indices (integer work)
    assert np.array_equal(sp55.decode(golden["dec55_in"], "from-set", "grid", 100), golden["dec55_out"])
    # zero query -> argmax of all-equal similarities = first grid row (slam.py clean-up: x = 0 -> g = 0)
    assert idx[2] == 0


def test_k1_identities(sp55):
    rng = np.random.default_rng(1)
    x, y = rng.uniform(-1, 1, (5, 2)), rng.uniform(-1, 1, (5, 2))
    ex, ey = sp55.encode_host(x), sp55.encode_host(y)
    np.testing.assert_allclose(np.linalg.norm(ex, axis=1), 1.0, atol=1e-12)
    np.testing.assert_allclose(sp55.encode_host(np.zeros((1, 2)))[0], sp55.identity(), atol=1e-14)
    np.testing.assert_allclose(sp55.bind(ex, ey), sp55.encode_host(x + y), atol=1e-12)
    np.testing.assert_allclose(sp55.invert(ex), sp55.encode_host(-x), atol=1e-12)
    np.testing.assert_allclose(ssp_ref.bind(ex, ey), sp55.encode_host(x + y), atol=1e-12)
    np.testing.assert_allclose(ssp_ref.invert(ex), sp55.encode_host(-x), atol=1e-12)


def test_k2_decode_is_nearest_grid_point(sp55):
    rng = np.random.default_rng(2)
    x = rng.uniform(-1, 1, (200, 2))
    dec = sp55.decode(sp55.encode_host(x), "from-set", "grid", 100)
    assert np.max(np.abs(dec - x)) <= 0.5 * (2 / 99) + 1e-9


def test_algebra_matches_reference(golden, sp55):
    a, b = golden["enc55"][4:8], golden["enc55"][8:12]
    np.testing.assert_allclose(sp55.bind(a, b), golden["bind55"], atol=1e-15)
    np.testing.assert_allclose(ssp_ref.bind(a, b), golden["bind55"], atol=1e-15)
    assert np.array_equal(sp55.invert(a), golden["invert55"])
    assert np.array_equal(ssp_ref.invert(a), golden["invert55"])
    noisy = golden["dec55_in"][4:8]
    np.testing.assert_allclose(sp55.make_unitary(noisy), golden["unitary55"], atol=1e-14)
    np.testing.assert_allclose(np.stack([ssp_ref.make_unitary(v) for v in noisy]), golden["unitary55"], atol=1e-14)
    assert np.array_equal(sp55.identity(), golden["identity55"])


def test_spspace_matches_reference(golden):
    np.testing.assert_allclose(SPSpace(50, 55, seed=0).vectors, golden["sp50_vectors"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(SPSpace(8, 25, seed=3).vectors, golden["sp8_vectors"], rtol=0, atol=1e-14)
    v = SPSpace(50, 55, seed=0).vectors
    gram = v @ v.T
    assert np.max(np.abs(gram - np.diag(np.diag(gram)))) < 1e-14           # orthogonal ...
    assert np.linalg.norm(v[-1]) < 0.5                                     # ... but norms decay (App. C quirk)


def test_cuda_backend_fails_loudly_without_library(monkeypatch, sp55):
    """The product path must not fall back to NumPy when the CUDA library is missing."""
    from sspslam_b200 import cabi
    monkeypatch.setattr(cabi, "_lib", None)
    monkeypatch.setattr(cabi, "LIB_PATH", "/nonexistent/libssb.so")
    sp = HexagonalSSPSpace(2, ssp_dim=55, domain_bounds=BOUNDS2, length_scale=0.2)   # backend='cuda'
    with pytest.raises(cabi.SsbError):
        sp.encode(np.zeros((1, 2)))
    with pytest.raises(cabi.SsbError):
        from sspslam_b200.simulator import Simulator
        from sspslam_b200 import scenarios
        sc = scenarios.make_pathint(n_trials=1, n_steps=4, ssp_dim=7, pi_n_neurons=20, neuron_type="lifrate")
        Simulator(sc.network)
