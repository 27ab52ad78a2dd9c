"""Vectorised synthesis of the input tables the reference produces with per-step Python
closures.

``get_slam_input_functions2`` (``sspslam/networks/slam.py:442-497``) and the driver
lambdas (``experiments/run_slam.py:164-169``, ``run_pathint.py:134-136``) are evaluated
by nengo once per step at ``t = n*dt``.  Their index expressions are float-fragile
(``int((t-dt)/dt)`` is ``n-1`` for only ~80 % of ``n``; K7) so the tables here reproduce
the *same expressions* element-wise in float64 rather than "fixing" them.
``tests/test_transforms_inputs_golden.py`` / ``tests/test_networks_parity.py`` check the tables against the unmodified
closures.
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


def stretch_trajectory(traj, original_dt=0.02, new_dt=0.001):
    """Linear re-sampling of a recorded path to the simulation step (``run_slam.py:80-89``): ``int(n*original_dt/new_dt)``
    points on the same time span, per axis."""
    traj = np.asarray(traj, dtype=np.float64)
    n = traj.shape[0]
    total = n * original_dt
    t_old, t_new = np.linspace(0, total, n), np.linspace(0, total, int(total / new_dt))
    return np.stack([np.interp(t_new, t_old, traj[:, k]) for k in range(traj.shape[1])], axis=1)


def load_path(path_data, data_dt=0.001, dt=0.001, radius=1.0):
    """The ``--path-data`` branch of the drivers (``run_slam.py:100-112``): the first 99 999 rows of the recorded path
    (a ``.npy`` file name or an array ``[n, domain_dim]``), re-sampled when ``data_dt != dt``, then every axis min-max
    rescaled to +-0.9*radius.  The simulated time is ``len(result) * dt``."""
    raw = np.load(path_data) if isinstance(path_data, (str, bytes)) or hasattr(path_data, "__fspath__") else path_data
    path = np.array(raw, dtype=np.float64)[:99999, :]
    if data_dt != dt:
        path = stretch_trajectory(path, original_dt=data_dt, new_dt=dt)
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


class _SlamInputs:
    """Per-path lookups behind the node callables of ``get_slam_input_functions`` / ``get_slam_input_functions2``.

    The in-view masks and nearest-landmark ids of the whole path are computed once (vectorised); the callables only
    index them with the reference's two time -> row rules: ``int((t-dt)/dt)`` for the velocity, the in-view set, the
    landmark SP and the displacement vector, and ``min(floor(t/dt), pathlen-2)`` for the displacement that is SSP-encoded
    (slam.py:393,408,434 / :452,457,485)."""

    def __init__(self, ssp_space, lm_space, velocity_data, vec_to_landmarks_data, view_rad, dt, superpose):
        self.space, self.dt, self.superpose = ssp_space, dt, superpose
        self.vec = np.asarray(vec_to_landmarks_data, dtype=np.float64)
        self.pathlen, _, self.domain_dim = self.vec.shape
        self.d = ssp_space.ssp_dim
        self.lm = np.asarray(lm_space.vectors, dtype=np.float64)
        self.scale = velocity_scale(ssp_space.phase_matrix, velocity_data)
        self.vels_scaled = np.asarray(velocity_data) * self.scale
        dists = np.linalg.norm(self.vec, axis=2)
        self.in_view = dists <= view_rad
        self.any_view = self.in_view.any(axis=1)
        self.nearest = np.argmin(dists, axis=1)
        self.encode = getattr(ssp_space, "encode_host", ssp_space.encode)

    def prev(self, t):
        return int((t - self.dt) / self.dt)

    def cur(self, t):
        return int(np.minimum(np.floor(t / self.dt), self.pathlen - 2))

    def ids(self, t):
        i = self.prev(t)
        if not self.any_view[i]:
            return None if self.superpose else -1
        return np.where(self.in_view[i])[0] if self.superpose else self.nearest[i]

    def _members(self, t):
        i = self.prev(t)
        if not self.any_view[i]:
            return np.zeros(0, dtype=int)
        return np.where(self.in_view[i])[0] if self.superpose else self.nearest[i:i + 1]

    def velocity(self, t):
        return self.vels_scaled[self.prev(t)]

    def vec_to(self, t):
        return self.vec[self.prev(t), self._members(t)].sum(axis=0) if self.any_view[self.prev(t)] else np.zeros(self.domain_dim)

    def sp(self, t):
        m = self._members(t)
        return self.lm[m].sum(axis=0) if len(m) else np.zeros(self.d)

    def vec_ssp(self, t):
        m = self._members(t)
        if not len(m):
            return np.zeros(self.d)
        return np.asarray(self.encode(self.vec[self.cur(t), m])).reshape(len(m), -1).sum(axis=0)

    def none_in_view(self, t):
        return 0 if self.any_view[self.prev(t)] else 10

    def functions(self):
        return (self.velocity, self.scale, self.none_in_view, self.ids, self.sp, self.vec_to, self.vec_ssp)


def get_slam_input_functions(ssp_space, lm_space, velocity_data, vec_to_landmarks_data, view_rad, dt=0.001):
    """Node callables for recorded data, nearest in-view landmark only (``sspslam/networks/slam.py:312-440``).  Returns
    ``(velocity_func, vel_scaling_factor, is_landmark_in_view, landmark_id_func, landmark_sp_func, landmark_vec_func,
    landmark_vecssp_func)`` — ``is_landmark_in_view`` is 10 when NO landmark is in view, as in the reference."""
    return _SlamInputs(ssp_space, lm_space, velocity_data, vec_to_landmarks_data, view_rad, dt, False).functions()


def get_slam_input_functions2(ssp_space, lm_space, velocity_data, vec_to_landmarks_data, view_rad, dt=0.001):
    """Same, with every in-view landmark superposed (``slam.py:442-497``; what ``run_slam.py:139-141`` uses).
    ``slam_tables`` is the batched table form of the same arithmetic."""
    return _SlamInputs(ssp_space, lm_space, velocity_data, vec_to_landmarks_data, view_rad, dt, True).functions()


def slamview_tables(ssp_space, lm_vectors, vels_scaled, vec_to_landmarks, view_rad, n_steps, dt=0.001, step0=0):
    """Batched table form of ``get_slamview_input_functions`` (``slam_view.py:352-403``): the local view is the normalised
    sum over in-view landmarks of ``SP_l (*) encode(displacement_l)``.  Note the index rules are the opposite of the SLAM
    variant: velocity and the in-view set (strict ``<``) use ``min(floor(t/dt), pathlen-2)``, the displacement
    ``int((t-dt)/dt)``; the none-in-view flag is 1, not 10."""
    vec = np.asarray(vec_to_landmarks, dtype=np.float64)
    _, i_prev, i_cur = step_indices(n_steps, dt, vec.shape[0], step0)
    d = lm_vectors.shape[1]
    in_view = np.linalg.norm(vec[i_cur], axis=2) < view_rad
    steps, ids = np.nonzero(in_view)
    view = np.zeros((n_steps, d))
    if len(steps):
        encode = getattr(ssp_space, "encode_host", ssp_space.encode)
        ssps = np.asarray(encode(vec[i_prev[steps], ids])).reshape(len(steps), d)
        bound = np.fft.ifft(np.fft.fft(np.asarray(lm_vectors)[ids], axis=1) * np.fft.fft(ssps, axis=1), axis=1).real
        np.add.at(view, steps, bound)
    norm = np.linalg.norm(view, axis=1, keepdims=True)
    view = np.where(norm > 1e-8, view / np.maximum(norm, 1e-300), view)
    return {"vel": np.asarray(vels_scaled)[i_cur], "view": view,
            "nolm": np.where(in_view.any(axis=1), 0.0, 1.0)[:, None]}


def get_slamview_input_functions(ssp_space, lm_space, velocity_data, vec_to_landmarks_data, view_rad, dt=0.001):
    """Node callables of the local-view variant (``sspslam/networks/slam_view.py:281-404``): returns
    ``(velocity_func, vel_scaling_factor, is_landmark_in_view, landmark_func)``."""
    vec = np.asarray(vec_to_landmarks_data, dtype=np.float64)
    pathlen, d = vec.shape[0], ssp_space.ssp_dim
    lm = np.asarray(lm_space.vectors, dtype=np.float64)
    scale = velocity_scale(ssp_space.phase_matrix, velocity_data)
    vels_scaled = np.asarray(velocity_data) * scale
    in_view = np.linalg.norm(vec, axis=2) < view_rad
    encode = getattr(ssp_space, "encode_host", ssp_space.encode)

    def cur(t):
        return int(np.minimum(np.floor(t / dt), pathlen - 2))

    def landmark_func(t):
        ids = np.where(in_view[cur(t)])[0]
        out = np.zeros(d)
        if len(ids):
            ssps = np.asarray(encode(vec[int((t - dt) / dt), ids])).reshape(len(ids), d)
            out = np.fft.ifft(np.fft.fft(lm[ids], axis=1) * np.fft.fft(ssps, axis=1), axis=1).real.sum(axis=0)
        norm = np.linalg.norm(out)
        return out / norm if norm > 1e-8 else out

    return (lambda t: vels_scaled[cur(t)]), scale, (lambda t: 0 if in_view[cur(t)].any() else 1), landmark_func

