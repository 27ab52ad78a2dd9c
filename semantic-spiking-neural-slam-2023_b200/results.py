"""Post-processing and result files of the reference drivers, on top of a finished ``Simulator``.

``experiments/run_slam.py:236-293`` and ``experiments/run_pathint.py:168-203`` decode the probed SSP
trajectory against the sample grid, compute the per-step position error and the (un-normalised-by-truth)
cosine similarity with the true SSP, sub-sample by 100 when the run is longer than 1e5 steps, and write an
``.npz`` whose key names are consumed by ``experiments/plot_trials_2d.py:68-77``.  The functions below
restate exactly that arithmetic (same key names, same ``skip`` rule, same ``1e-6`` guard in the SLAM
variant and none in the path-integration variant) for one trial or for a batch of trials.

The grid decode runs through ``SSPSpace.decode`` (``backend='cuda'`` -> ``ssb_ssp_decode_argmax``).
"""
from __future__ import annotations

import numpy as np

SKIP_THRESHOLD = 100000      # run_slam.py:237 / run_pathint.py:169
SKIP = 100


def _grid_points(domain_dim, kind):
    # run_slam.py:239,248 uses 30 points per axis in 3-D, run_pathint.py:171,180 uses 50
    if domain_dim < 3:
        return 100
    return 30 if kind == "slam" else 50


def _one_trial(sim_out, ts, path, real_ssp, ssp_space, kind):
    sim_out = np.asarray(sim_out, dtype=np.float64)
    path = np.asarray(path, dtype=np.float64)[:sim_out.shape[0]]
    real_ssp = np.asarray(real_ssp, dtype=np.float64)[:sim_out.shape[0]]
    ts = np.asarray(ts, dtype=np.float64)[:sim_out.shape[0]]
    if path.shape[0] > SKIP_THRESHOLD:
        sim_out, ts, path, real_ssp = sim_out[::SKIP], ts[::SKIP], path[::SKIP], real_ssp[::SKIP]
    est = ssp_space.decode(sim_out, "from-set", "grid", _grid_points(ssp_space.domain_dim, kind))
    norms = np.linalg.norm(sim_out, axis=1)
    if kind == "slam":
        sims = np.sum(sim_out * real_ssp, axis=1) / np.maximum(1e-6, norms)
    else:
        with np.errstate(divide="ignore", invalid="ignore"):
            sims = np.sum(sim_out * real_ssp, axis=1) / norms
    error = np.sqrt(np.sum((path - est) ** 2, axis=1))
    return ts, path, real_ssp, sim_out, sims, est, error


def _landmark_estimates(sim, slam, weights_probe, lm_vectors, ssp_space, trial):
    """run_slam.py:263-268: recall every landmark SP through the learned decoders."""
    from .nengo_shim.builder.ensemble import get_activities
    w = sim.data[weights_probe]
    w = w[trial] if w.ndim == 4 else w
    decoders = w[-1].T                                                   # (n_mem, d)
    memory = slam.assomemory.memory
    acts = get_activities(sim.data[memory], memory, lm_vectors)
    ssps = np.dot(acts, decoders)
    locs = ssp_space.decode(ssps, "from-set", "grid", _grid_points(ssp_space.domain_dim, "slam"))
    return ssps, locs


def slam_results(sim, probe, ssp_space, path, real_ssp, *, trial=None, slam=None, weights_probe=None,
                 lm_vectors=None, extra=None):
    """Dictionary with the keys of ``run_slam.py``'s ``np.savez`` call for one trial.

    ``path`` / ``real_ssp`` are that trial's ``[T, dim]`` / ``[T, d]`` arrays; for a batched simulator pass
    ``trial``.  Landmark estimates are filled when ``slam``, ``weights_probe`` and ``lm_vectors`` are given
    (``None`` otherwise, as the reference does for back-ends without weight probes)."""
    data = sim.data[probe]
    if data.ndim == 3:
        if trial is None:
            raise ValueError("batched simulator: pass trial=")
        data = data[trial]
    ts, path, real_ssp, out, sims, est, err = _one_trial(data, sim.trange(), path, real_ssp, ssp_space, "slam")
    res = dict(ts=ts, path=path, real_ssp=real_ssp, slam_sim_out=out, slam_sims=sims, slam_path=est, slam_error=err,
               landmark_ssps_est=None, landmark_loc_est=None)
    if slam is not None and weights_probe is not None and lm_vectors is not None:
        res["landmark_ssps_est"], res["landmark_loc_est"] = _landmark_estimates(
            sim, slam, weights_probe, lm_vectors, ssp_space, 0 if trial is None else trial)
    if extra:
        res.update(extra)
    return res


def pathint_results(sim, probe, ssp_space, path, real_ssp, *, trial=None, extra=None):
    """Dictionary with the keys of ``run_pathint.py``'s ``np.savez`` call for one trial."""
    data = sim.data[probe]
    if data.ndim == 3:
        if trial is None:
            raise ValueError("batched simulator: pass trial=")
        data = data[trial]
    ts, path, real_ssp, out, sims, est, err = _one_trial(data, sim.trange(), path, real_ssp, ssp_space, "pi")
    res = dict(ts=ts, path=path, real_ssp=real_ssp, pi_sim_out=out, pi_sims=sims, pi_path=est, pi_error=err)
    if extra:
        res.update(extra)
    return res


def slam_filename(d, pi_n_neurons, mem_n_neurons, circonv_n_neurons, T, limit, seed, domain_dim=2, backend="b200",
                  extra_name=""):
    """File name rule of run_slam.py:271-278."""
    if domain_dim != 2:
        extra_name = "_dim_" + str(domain_dim)
    if backend != "cpu":
        extra_name = "_backend_" + backend + extra_name
    return (f"slam_{extra_name}_sspdim_{d}_pinneurons_{pi_n_neurons}_memnneurons_{mem_n_neurons}"
            f"_ccnneurons_{circonv_n_neurons}_T_{int(T)}_limit_{limit}_seed_{seed}.npz")


def pathint_filename(d, pi_n_neurons, T, limit, seed, domain_dim=2, backend="b200", extra_name=""):
    """File name rule of run_pathint.py:189-196."""
    if domain_dim != 2:
        extra_name = "_dim_" + str(domain_dim)
    if backend != "cpu":
        extra_name = "_backend_" + backend + extra_name
    return f"pi{extra_name}_sspdim_{d}_pinneurons_{pi_n_neurons}_T_{int(T)}_limit_{limit}_seed_{seed}.npz"


def save(filename, results):
    """``np.savez`` with the reference's key names (``None`` entries are stored as object scalars, as NumPy does)."""
    np.savez(filename, **results)


def trial_statistics(results):
    """Per-trial summary gathered across GPUs at the end of a sharded run:
    [mean error, max error, final error, mean similarity, samples]."""
    err = results["slam_error"] if "slam_error" in results else results["pi_error"]
    sims = results["slam_sims"] if "slam_sims" in results else results["pi_sims"]
    return np.array([np.mean(err), np.max(err), err[-1], np.nanmean(sims), float(err.shape[0])], dtype=np.float32)
