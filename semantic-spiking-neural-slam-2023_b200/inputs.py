"""Vectorised synthesis of the input tables the reference produces with per-step Python
closures.

``get_slam_input_functions2`` (``sspslam/networks/slam.py:442-497``) and the driver
lambdas (``experiments/run_slam.py:164-169``, ``run_pathint.py:134-136``) are evaluated
by nengo once per step at ``t = n*dt``.  Their index expressions are float-fragile
(``int((t-dt)/dt)`` is ``n-1`` for only ~80 % of ``n``; K7) so the tables here reproduce
the *same expressions* element-wise in float64 rather than "fixing" them.
``tests/test_inputs_golden.py`` checks the tables against the unmodified closures.
"""
from __future__ import annotations

import numpy as np

from .nengo_shim.processes import WhiteSignal


def step_indices(n_steps, dt, pathlen, step0=0):
    """(i_prev, i_cur) used by the reference closures at steps step0+1 .. step0+n_steps."""
    t = np.arange(step0 + 1, step0 + n_steps + 1) * dt
    i_prev = np.trunc((t - dt) / dt).astype(np.int64)              # int((t-dt)/dt)
    i_cur = np.minimum(np.floor(t / dt), pathlen - 2).astype(np.int64)
    return t, i_prev, i_cur


def random_path(T, dt=0.001, limit=0.1, seed=0, domain_dim=2, radius=1.0):
    """Band-limited white-noise path rescaled to +-0.9*radius per axis (run_slam.py:95-112)."""
    cols = [WhiteSignal(T, high=limit, seed=seed + i).run(T, dt=dt) for i in range(domain_dim)]
    path = np.hstack(cols)
    lo, hi = path.min(axis=0), path.max(axis=0)
    return (1.8 * radius) * (path - lo) / (hi - lo) - 0.9 * radius


def velocities(path, dt=0.001):
    """``diff(path)/dt`` with a leading zero row (run_slam.py:114)."""
    return np.diff(path, axis=0, prepend=path[:1]) / dt


def rd_sampling(n, d, seed=0.5):
    """R_d low-discrepancy points with an additive offset (``sspslam/utils/utils.py:41-55``)."""
    g = 2.0
    for _ in range(10):
        g = pow(1 + g, 1 / (d + 1))
    alpha = np.array([pow(1 / g, j + 1) % 1 for j in range(d)])   # Python-float pow: bit-identical to the reference
    return (seed + alpha[None, :] * np.arange(1, n + 1)[:, None]) % 1


def velocity_scale(phase_matrix, vels):
    return 1.0 / np.max(np.abs(phase_matrix @ vels.T))


def pathint_tables(real_ssp, vels_scaled, n_steps, dt=0.001, step0=0, init_time=0.05):
    t, i_prev, _ = step_indices(n_steps, dt, len(vels_scaled), step0)
    vel = vels_scaled[i_prev]
    init = np.where((t < init_time)[:, None], real_ssp[i_prev], 0.0)
    return {"vel": vel, "init": init}


def slam_tables(encode, lm_vectors, vels_scaled, vec_to_landmarks, view_rad, n_steps, dt=0.001, step0=0,
                real_ssp=None, init_time=0.05, none_in_view_value=10.0):
    """Tables of ``get_slam_input_functions2`` (all in-view landmarks are superposed).

    ``encode`` maps (N, domain_dim) -> (N, d) float64; ``vec_to_landmarks`` is
    [time, landmark, dim] (the layout the code — not the docstring — uses)."""
    pathlen = vec_to_landmarks.shape[0]
    t, i_prev, i_cur = step_indices(n_steps, dt, pathlen, step0)
    d = lm_vectors.shape[1]
    dists = np.linalg.norm(vec_to_landmarks[i_prev], axis=2)          # [n, n_lm]
    in_view = dists <= view_rad
    any_view = in_view.any(axis=1)
    lm_sp = in_view.astype(np.float64) @ lm_vectors
    steps, ids = np.nonzero(in_view)
    lmvec = np.zeros((n_steps, d))
    if len(steps):
        np.add.at(lmvec, steps, encode(vec_to_landmarks[i_cur[steps], ids]))
    out = {
        "vel": vels_scaled[i_prev],
        "lmvec_ssp": lmvec,
        "lm_sp": lm_sp,
        "nolm": np.where(any_view, 0.0, none_in_view_value)[:, None],
    }
    if real_ssp is not None:
        out["init"] = np.where((t < init_time)[:, None], real_ssp[i_prev], 0.0)
    return out


def sparsity_to_x_intercept(d, p):
    """Intercept at which a neuron with a unit encoder in ``d`` dimensions is active for a fraction ``p`` of the points of
    the unit sphere (``sspslam/utils/utils.py:5-10``): the cap with relative area p has cos(angle) = sqrt(1 - I^-1), with
    I the regularised incomplete beta function of ((d-1)/2, 1/2)."""
    from scipy.special import betaincinv
    flip = p > 0.5
    if flip:
        p = 1.0 - p
    x = np.sqrt(1.0 - betaincinv((d - 1) / 2.0, 0.5, 2.0 * p))
    return -x if flip else x

