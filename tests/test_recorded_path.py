"""``--path-data`` semantics of the drivers (``experiments/run_slam.py:80-112``): ``[:99999]`` slice, ``stretch_trajectory``
when ``data_dt != dt``, per-axis min-max rescale to +-0.9, ``T = len(path) * dt`` — and the recorded-path scenarios of
BASELINE configs[2] / [3] built on the committed fixture (a down-sampled slice of ``example_paths/twoRooms_path.npy``)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, has_reference
from sspslam_b200 import inputs, scenarios

FIXTURE = os.path.join(GOLDEN_DIR, "twoRooms_path_ds20.npy")


def _driver_path(raw, data_dt, dt=0.001, radius=1.0):
    """The driver's own statements (run_slam.py:80-89,100-112), written out step by step."""
    path = np.array(raw, dtype=np.float64)[:99999, :]
    if data_dt != dt:
        n_steps = path.shape[0]
        total_time = n_steps * data_dt
        n_timesteps = int(total_time / dt)
        original_times = np.linspace(0, total_time, n_steps)
        new_times = np.linspace(0, total_time, n_timesteps)
        new = np.zeros((n_timesteps, 2))
        new[:, 0] = np.interp(new_times, original_times, path[:, 0])
        new[:, 1] = np.interp(new_times, original_times, path[:, 1])
        path = new
    for i in range(path.shape[1]):
        x = path[:, i]
        path[:, i] = (0.9 * radius - -0.9 * radius) * (x - np.min(x)) / (np.max(x) - np.min(x)) + -0.9 * radius
    return path


def test_load_path_follows_the_driver_rules():
    raw = np.load(FIXTURE)
    assert raw.shape == (3000, 2) and raw.dtype == np.float32
    got = inputs.load_path(FIXTURE, data_dt=0.02)
    want = _driver_path(raw, 0.02)
    assert got.shape == (60000, 2)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)
    assert np.allclose(got.min(axis=0), -0.9) and np.allclose(got.max(axis=0), 0.9)
    same_dt = inputs.load_path(raw, data_dt=0.001)                     # an array works like a file; no re-sampling
    np.testing.assert_allclose(same_dt, _driver_path(raw, 0.001), rtol=0, atol=1e-12)
    long = np.cumsum(np.random.default_rng(0).standard_normal((120000, 2)), axis=0)
    assert inputs.load_path(long).shape == (99999, 2)                   # the [:99999] slice


@pytest.mark.skipif(not has_reference(), reason="reference checkout not mounted")
def test_fixture_tracks_the_full_recorded_path():
    full = np.load("/root/reference/example_paths/twoRooms_path.npy")
    a = inputs.load_path(full, data_dt=0.001)
    b = inputs.load_path(FIXTURE, data_dt=0.02)
    assert a.shape == (60000, 2) and b.shape == (60000, 2)
    # every 20th sample, re-sampled by the driver's stretch rule (whose linspace time axes skew by one sample over the
    # run): the fixture follows the recorded walk within 5 % of the 1.8-wide domain, 1.3 % rms
    assert np.max(np.abs(a - b)) < 0.09 and np.sqrt(np.mean((a - b) ** 2)) < 0.025


def test_recorded_path_scenario_shares_the_path_and_varies_landmarks():
    sc = scenarios.make_slam(n_trials=3, n_steps=80, ssp_dim=19, pi_n_neurons=30, mem_n_neurons=64, circonv_n_neurons=16,
                             n_landmarks=6, view_rad=0.6, path_data=FIXTURE, data_dt=0.02)
    assert sc.paths.shape == (3, 80, 2)
    assert np.array_equal(sc.paths[0], sc.paths[1]) and np.array_equal(sc.paths[0], sc.paths[2])
    want = inputs.load_path(FIXTURE, data_dt=0.02)[:80]
    np.testing.assert_allclose(sc.paths[0], want, rtol=0, atol=1e-12)
    syn = sc.extra["input_synthesis"]
    assert not np.array_equal(syn["landmarks"][0], syn["landmarks"][1])          # trial i: Rd_sampling(seed + 1000 i)
    # tables equal the closures of get_slam_input_functions2 evaluated on the same recorded path
    from sspslam_b200.sspspace import SPSpace
    path = inputs.load_path(FIXTURE, data_dt=0.02)[:82]
    vels = inputs.velocities(path)
    lm_space = SPSpace(6, sc.ssp_space.ssp_dim, seed=0)
    fns = inputs.get_slam_input_functions2(sc.ssp_space, lm_space, vels, syn["landmarks"][1][None] - path[:, None, :], 0.6)
    by_label = {n.label: a for n, a in sc.trial_inputs.items()}
    full_vels = inputs.velocities(inputs.load_path(FIXTURE, data_dt=0.02))
    scale = inputs.velocity_scale(sc.ssp_space.phase_matrix, full_vels)          # the driver scales by the WHOLE path
    assert sc.extra["vel_scale"] == pytest.approx(scale, rel=1e-14)
    for k in (1, 2, 17, 60, 80):
        t = k * 0.001
        np.testing.assert_allclose(by_label["lm_sp_input"][1, k - 1], fns[4](t), rtol=0, atol=1e-12)
        np.testing.assert_allclose(by_label["lm_vecssp_input"][1, k - 1], fns[6](t), rtol=0, atol=1e-12)
        np.testing.assert_allclose(by_label["vel_input"][1, k - 1], vels[int((t - 0.001) / 0.001)] * scale, rtol=0, atol=1e-12)
        assert by_label["lm_in_view_input"][1, k - 1, 0] == fns[2](t)
