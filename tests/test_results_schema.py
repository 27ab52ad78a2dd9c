"""Post-processing / result-file schema of the reference drivers (run_slam.py:236-293, run_pathint.py:168-203)."""
import numpy as np

from sspslam_b200 import results
from sspslam_b200.sspspace import HexagonalSSPSpace

BOUNDS2 = np.tile([-1.0, 1.0], (2, 1))
SLAM_KEYS = {"ts", "path", "real_ssp", "slam_sim_out", "slam_sims", "slam_path", "slam_error", "landmark_ssps_est",
             "landmark_loc_est"}                       # + run metadata (timesteps, obj_locs, view_rad, elapsed_time, args ...)
PI_KEYS = {"ts", "path", "real_ssp", "pi_sim_out", "pi_sims", "pi_path", "pi_error"}


class _Sim:
    def __init__(self, probe, data, dt=0.001):
        self.data = {probe: data}
        self._n = data.shape[-2]
        self.dt = dt

    def trange(self):
        return self.dt * np.arange(1, self._n + 1)


def _space():
    return HexagonalSSPSpace(2, ssp_dim=55, domain_bounds=BOUNDS2, length_scale=0.2, backend="host")


def test_slam_results_keys_and_arithmetic(tmp_path):
    sp = _space()
    rng = np.random.default_rng(0)
    path = rng.uniform(-0.9, 0.9, (300, 2))
    real = sp.encode_host(path)
    out = 0.7 * real + 0.05 * rng.standard_normal(real.shape)
    out[5] = 0.0                                                   # a silent step: the 1e-6 guard of run_slam.py:244
    res = results.slam_results(_Sim("p", out), "p", sp, path, real,
                               extra=dict(obj_locs=np.zeros((3, 2)), view_rad=0.2, elapsed_time=1.0))
    assert SLAM_KEYS <= set(res)
    est = sp.decode(out, "from-set", "grid", 100)
    assert np.array_equal(res["slam_path"], est)
    np.testing.assert_allclose(res["slam_error"], np.sqrt(np.sum((path - est) ** 2, axis=1)))
    np.testing.assert_allclose(res["slam_sims"], np.sum(out * real, axis=1) / np.maximum(1e-6, np.linalg.norm(out, axis=1)))
    assert res["slam_sims"][5] == 0.0 and np.all(np.isfinite(res["slam_sims"]))
    assert res["landmark_ssps_est"] is None and res["landmark_loc_est"] is None
    np.testing.assert_allclose(res["ts"], 0.001 * np.arange(1, 301))
    assert np.mean(res["slam_error"]) < 0.05                       # the noisy SSPs still decode next to the path
    fn = tmp_path / results.slam_filename(55, 500, 970, 100, 200.0, 0.1, 0)
    assert fn.name == "slam__backend_b200_sspdim_55_pinneurons_500_memnneurons_970_ccnneurons_100_T_200_limit_0.1_seed_0.npz"
    results.save(fn, res)
    back = np.load(fn, allow_pickle=True)
    assert SLAM_KEYS <= set(back.files) and np.array_equal(back["slam_path"], est)
    stats = results.trial_statistics(res)
    assert stats.shape == (5,) and stats[4] == 300


def test_pathint_results_batched_trial_and_skip_rule():
    sp = _space()
    rng = np.random.default_rng(1)
    T = 100300                                                     # > 1e5 rows: every 100th sample is kept
    ang = np.linspace(0, 4 * np.pi, T)
    path = 0.8 * np.stack([np.cos(ang), np.sin(ang)], axis=1)
    keep = path[::100]
    real = np.zeros((T, 55))
    real[::100] = sp.encode_host(keep)
    out = np.zeros((2, T, 55))
    out[1, ::100] = real[::100] + 0.02 * rng.standard_normal((keep.shape[0], 55))
    res = results.pathint_results(_Sim("p", out), "p", sp, path, real, trial=1)
    assert PI_KEYS <= set(res)
    n_keep = keep.shape[0]
    assert res["pi_path"].shape == (n_keep, 2) and res["ts"].shape == (n_keep,) and res["path"].shape == (n_keep, 2)
    np.testing.assert_allclose(res["ts"], 0.001 * np.arange(1, T + 1)[::100])
    assert np.max(res["pi_error"]) < 0.05
    assert results.pathint_filename(97, 500, 20.0, 0.1, 3) == "pi_backend_b200_sspdim_97_pinneurons_500_T_20_limit_0.1_seed_3.npz"
