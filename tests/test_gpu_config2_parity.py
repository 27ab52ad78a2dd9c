"""Parity of the BENCHMARKED configuration (BASELINE configs[1]: ``run_slam.py`` d = 55, pi 500, mem 970, circonv 100,
50 landmarks; 40 280 neurons per trial) and long-horizon / distribution checks (``-m gpu``).

* full-size rate-mode trajectories, learned matrices and clean-up indices against the float64 oracle;
* spiking: per-step spike COUNTS OF EVERY ENSEMBLE against the oracle run in float32 — equal over a stated horizon
  (the point where an fp32 kernel and an fp32 NumPy stepper with another summation order first part is printed);
* spiking SLAM compared as a distribution over trials (``run_slam.py:247-255``: ``slam_error``, ``slam_sims``);
* one long rate-mode run (5 000 steps: learning, gate open and closed, 625 deferred-PES folds).

The oracle is slow (~5 500 NumPy calls per step), so the long oracle runs are farmed out to a fork pool that is
started before the first CUDA call of this module and collected by the tests that need them.
"""
import multiprocessing as mp
import os

import numpy as np
import pytest

from oracle import ssp_ref
from oracle.nengo_ref_sim import RefSimulator
from sspslam_b200 import results, scenarios
from sspslam_b200.builder import build_model
from sspslam_b200.sspspace import HexagonalSSPSpace

pytestmark = pytest.mark.gpu
BOUNDS2 = np.tile([-1.0, 1.0], (2, 1))

DIST = dict(n_trials=16, n_steps=2000, ssp_dim=55, pi_n_neurons=100, mem_n_neurons=200, circonv_n_neurons=30,
            n_landmarks=20, T=20.0, neuron_type="lif")
LONG = dict(n_trials=2, n_steps=5000, ssp_dim=31, pi_n_neurons=60, mem_n_neurons=120, circonv_n_neurons=20,
            n_landmarks=12, T=20.0, neuron_type="lifrate", view_rad=0.3)


def _rel(got, want):
    return np.max(np.abs(got - want)) / max(1e-12, np.max(np.abs(want)))


def _Simulator():
    from sspslam_b200.simulator import Simulator
    return Simulator


# ------------------------------------------------------------------------------- oracle farm
_JOBS = {}


def _default_seeds(n):
    return [None] + list(range(n - 1))


def _oracle_job(key, trial, want_learned):
    sc, model, n_steps = _JOBS[key]
    tabs = {node: arr[trial] for node, arr in sc.trial_inputs.items()}
    ref = RefSimulator(sc.network, dt=sc.dt, model=model, node_tables=tabs,
                       trial_seed=_default_seeds(sc.paths.shape[0])[trial])
    ref.run_steps(n_steps)
    out = {"probe": ref.data[sc.probe]}
    if want_learned:
        slam = sc.extra["slam"]
        out["dec"] = ref.learned_weights(slam.assomemory.conn_out).copy()
        out["enc"] = ref.scaled_encoders(slam.assomemory.memory).copy()
        out["gate"] = np.asarray(ref.data[sc.extra["gate_probe"]])
    return key, trial, out


@pytest.fixture(scope="module")
def farm(lib):
    """Build the two long scenarios, fork the oracle workers (NumPy only), hand back async results."""
    from sspslam_b200 import nengo_shim as nengo
    for key, kw in (("dist", DIST), ("long", LONG)):
        sc = scenarios.make_slam(**kw)
        if key == "long":
            with sc.network:
                sc.extra["gate_probe"] = nengo.Probe(sc.extra["slam"].update_state, synapse=None)
        _JOBS[key] = (sc, build_model(sc.network, dt=sc.dt), kw["n_steps"])
    workers = max(1, min(18, len(os.sched_getaffinity(0))))
    pool = mp.get_context("fork").Pool(workers)
    pending = {("long", t): pool.apply_async(_oracle_job, ("long", t, True)) for t in range(LONG["n_trials"])}
    pending.update({("dist", t): pool.apply_async(_oracle_job, ("dist", t, False)) for t in range(DIST["n_trials"])})
    yield pending
    pool.terminate()


# ------------------------------------------------------------------------------- full-size rate mode
def test_full_size_config2_rate_mode_matches_oracle_with_exact_cleanup_indices():
    """BASELINE configs[1] at full size, LIFRate-equivalence mode: probed SSP trajectory within 1e-4 relative over 300
    steps, PES decoders / Voja encoders within 1e-4, grid clean-up argmax bit-exact at every checkpoint."""
    n_steps, n_trials, checks = 300, 3, 6
    sc = scenarios.make_slam(n_trials=n_trials, n_steps=n_steps, T=200.0, neuron_type="lifrate")
    slam = sc.extra["slam"]
    sim = _Simulator()(sc.network, dt=sc.dt, n_trials=n_trials, trial_inputs=sc.trial_inputs)
    assert sim.plan.stats["n_neurons"] == 40280 and sim.plan.stats["n_learned"] == 106700
    idx_trace = []
    for _ in range(checks):
        sim.run_steps(n_steps // checks)
        idx_trace.append(sim.cleanup_indices()[0].copy())
    got = sim.data[sc.probe]
    dec = sim.learned_decoders(slam.assomemory.conn_out)
    enc = sim.learned_encoders(slam.assomemory.memory)
    sim.close()
    for trial in (0, 2):
        tabs = {node: arr[trial] for node, arr in sc.trial_inputs.items()}
        ref = RefSimulator(sc.network, dt=sc.dt, model=sim.model, node_tables=tabs, trial_seed=sim.trial_seeds[trial])
        for k in range(checks):
            ref.run_steps(n_steps // checks)
            x = ref.signals[slam.gridcells, "in"].a
            assert idx_trace[k][trial] == ssp_ref.cleanup_index(slam.sample_ssps, x)
        want = ref.data[sc.probe]
        assert np.max(np.abs(want)) > 0.1
        assert _rel(got[trial], want) < 1e-4
        want_dec = ref.learned_weights(slam.assomemory.conn_out)
        assert np.max(np.abs(want_dec)) > 0                                 # a landmark was in view: PES did learn
        assert np.max(np.abs(dec[trial] - want_dec)) < 1e-4 * np.max(np.abs(want_dec)) + 1e-9
        assert _rel(enc[trial], ref.scaled_encoders(slam.assomemory.memory)) < 1e-4


# ------------------------------------------------------------------------------- spike counts
def _device_spike_counts(sim, n_steps, bounds):
    """Per-step spike count of every ensemble, [n_steps, n_ens, n_trials], from the packed one-word LIF state:
    a neuron spiked in step t exactly when its state is negative (refractory) and lower than before the step
    (refractory states only rise by dt per step; tau_ref >= dt)."""
    nn = int(sim.plan.scalars["nn"])
    prev = sim._download("st", 0, nn)[:, :sim.n_trials]
    out = np.zeros((n_steps, len(bounds), sim.n_trials), dtype=np.int64)
    starts = np.array([b[0] for b in bounds])
    order = np.argsort(starts)
    for t in range(n_steps):
        sim.run_steps(1)
        cur = sim._download("st", 0, nn)[:, :sim.n_trials]
        spiked = (cur < 0) & (cur < prev)
        sums = np.add.reduceat(spiked.astype(np.int64), starts[order], axis=0)
        out[t, order] = sums
        prev = cur
    return out


# Steps over which every per-ensemble spike count is asserted equal.  Measured on B200 (round 2): the first differing
# step was 106 (reduced size, trial 0) and 101 / none within 120 (full size, trials 1 / 0); one neuron of ~40 000 crossing
# the threshold a step early is all it takes, after which the totals still agree to 1e-4 (20 of 303 066 spikes).
SPIKE_HORIZON = {"small": 64, "full": 64}


@pytest.mark.parametrize("size", ["small", "full"])
def test_spike_counts_of_every_ensemble_match_the_fp32_oracle(size):
    """BASELINE north_star: 'integer work (spike counts per step) must match bit-exactly over a fixed horizon'.  The
    device computes in float32, so the checker is the oracle stepped in float32 (``RefSimulator(dtype=np.float32)``); the
    two still differ in summation order and in the polynomial expm1 / log1p, so a voltage within ~1e-7 of threshold
    eventually crosses one step apart.  Asserted: identical counts for every ensemble, every step, up to SPIKE_HORIZON;
    the first step at which any count differs is printed (``-s``)."""
    if size == "full":
        kw, n_steps, trials = dict(T=200.0), 120, (0, 1)
    else:
        kw = dict(ssp_dim=55, pi_n_neurons=100, mem_n_neurons=200, circonv_n_neurons=30, n_landmarks=20, T=20.0)
        n_steps, trials = 300, (0, 1, 2)
    sc = scenarios.make_slam(n_trials=3, n_steps=n_steps, neuron_type="lif", **kw)
    sim = _Simulator()(sc.network, dt=sc.dt, n_trials=3, trial_inputs=sc.trial_inputs)
    enss = list(sim.plan.ens_state)
    bounds = [sim.plan.ens_state[e] for e in enss]
    got = _device_spike_counts(sim, n_steps, bounds)
    sim.close()
    assert got.sum() > 0
    first_diff = []
    for trial in trials:
        tabs = {node: arr[trial] for node, arr in sc.trial_inputs.items()}
        ref = RefSimulator(sc.network, dt=sc.dt, model=sim.model, node_tables=tabs, trial_seed=sim.trial_seeds[trial],
                           dtype=np.float32)
        want = np.zeros((n_steps, len(enss)), dtype=np.int64)
        for t in range(n_steps):
            ref.step()
            want[t] = [np.count_nonzero(ref.signals[e, "out"].a) for e in enss]
        diff = np.nonzero(np.any(got[:, :, trial] != want, axis=1))[0]
        first = int(diff[0]) if diff.size else n_steps
        first_diff.append(first)
        total, mism = want.sum(), int(np.abs(got[:, :, trial] - want).sum())
        print(f"[spike counts {size}] trial {trial}: first differing step {first} of {n_steps}; "
              f"{total} oracle spikes, sum |count difference| {mism}")
        h = SPIKE_HORIZON[size]
        assert np.array_equal(got[:h, :, trial], want[:h]), f"spike counts differ before step {h} (first: {first})"
        # past the horizon the runs stay statistically together: total spike count within 0.5 %
        assert abs(int(got[:, :, trial].sum()) - int(total)) <= 0.005 * total
    print(f"[spike counts {size}] horizon asserted {SPIKE_HORIZON[size]}, measured first differences {first_diff}")


# ------------------------------------------------------------------------------- SLAM spiking distribution
class _RefView:
    def __init__(self, probe, data, n_steps, dt):
        self.data, self._n, self._dt = {probe: data}, n_steps, dt

    def trange(self):
        return self._dt * np.arange(1, self._n + 1)


def test_slam_spiking_error_distribution_matches_oracle(farm):
    """Spiking SLAM (PES + Voja learning, loop-closure gate) is chaotic in fp32 vs fp64, so it is compared the way the
    reference evaluates itself (``run_slam.py:247-255``, aggregated over seeds by ``plot_trials_2d.py``): per-trial mean
    of ``slam_error`` (decoded position error) and of ``slam_sims`` (similarity to the true SSP) over the second half of
    a 2 000-step run, 16 trials.  Tolerances: |difference of the trial means| < 0.02 (error, domain units; the domain is
    2 x 2) and < 0.02 (similarity); the 25 / 50 / 75 % quantiles over trials within 0.03; every single trial within 0.08."""
    sc, model, n_steps = _JOBS["dist"]
    n_trials = DIST["n_trials"]
    with _Simulator()(sc.network, dt=sc.dt, n_trials=n_trials, trial_inputs=sc.trial_inputs, model=model) as sim:
        sim.run_steps(n_steps)
    sp = HexagonalSSPSpace(2, ssp_dim=55, domain_bounds=BOUNDS2, length_scale=0.2, rng=np.random.default_rng(0))
    half = n_steps // 2
    g_err, g_sim, o_err, o_sim = [], [], [], []
    for trial in range(n_trials):
        r = results.slam_results(sim, sc.probe, sp, sc.paths[trial], sc.real_ssp[trial], trial=trial)
        _, _, out = farm[("dist", trial)].get(timeout=1800)
        o = results.slam_results(_RefView(sc.probe, out["probe"], n_steps, sc.dt), sc.probe, sp, sc.paths[trial],
                                 sc.real_ssp[trial])
        g_err.append(np.mean(r["slam_error"][half:])); g_sim.append(np.mean(r["slam_sims"][half:]))
        o_err.append(np.mean(o["slam_error"][half:])); o_sim.append(np.mean(o["slam_sims"][half:]))
    g_err, g_sim, o_err, o_sim = map(np.array, (g_err, g_sim, o_err, o_sim))
    q = [25, 50, 75]
    print(f"[slam distribution] error  mean gpu {g_err.mean():.4f} oracle {o_err.mean():.4f}; quantiles gpu "
          f"{np.percentile(g_err, q).round(4)} oracle {np.percentile(o_err, q).round(4)}; max trial diff {np.max(np.abs(g_err - o_err)):.4f}")
    print(f"[slam distribution] sims   mean gpu {g_sim.mean():.4f} oracle {o_sim.mean():.4f}; quantiles gpu "
          f"{np.percentile(g_sim, q).round(4)} oracle {np.percentile(o_sim, q).round(4)}; max trial diff {np.max(np.abs(g_sim - o_sim)):.4f}")
    assert o_sim.mean() > 0.5                                           # the network does track the path
    assert abs(g_err.mean() - o_err.mean()) < 0.02
    assert abs(g_sim.mean() - o_sim.mean()) < 0.02
    assert np.max(np.abs(np.percentile(g_err, q) - np.percentile(o_err, q))) < 0.03
    assert np.max(np.abs(np.percentile(g_sim, q) - np.percentile(o_sim, q))) < 0.03
    assert np.max(np.abs(g_err - o_err)) < 0.08 and np.max(np.abs(g_sim - o_sim)) < 0.08


# ------------------------------------------------------------------------------- long horizon
def test_long_rate_mode_run_5000_steps_matches_oracle(farm):
    """5 000 steps in rate mode at reduced size (d = 31): 625 deferred-PES folds, Voja drift, the loop-closure gate both
    open and closed.  Trajectory within 1e-4 relative over the whole run, learned matrices within 1e-4 at the end."""
    sc, model, n_steps = _JOBS["long"]
    slam = sc.extra["slam"]
    with _Simulator()(sc.network, dt=sc.dt, n_trials=LONG["n_trials"], trial_inputs=sc.trial_inputs, model=model,
                      chunk_steps=512) as sim:
        sim.run_steps(n_steps)
        dec = sim.learned_decoders(slam.assomemory.conn_out)
        enc = sim.learned_encoders(slam.assomemory.memory)
    got, gate = sim.data[sc.probe], sim.data[sc.extra["gate_probe"]]
    for trial in range(LONG["n_trials"]):
        _, _, out = farm[("long", trial)].get(timeout=1800)
        open_steps = np.any(out["gate"] != 0, axis=1)
        print(f"[long run] trial {trial}: gate open on {int(open_steps.sum())} of {n_steps} steps; "
              f"rel err {_rel(got[trial], out['probe']):.2e}")
        assert 0 < open_steps.sum() < n_steps                           # the gate was exercised both ways
        assert np.array_equal(np.any(gate[trial] != 0, axis=1), open_steps)
        assert _rel(got[trial], out["probe"]) < 1e-4
        assert _rel(gate[trial], out["gate"]) < 1e-4
        assert np.max(np.abs(out["dec"])) > 0
        assert np.max(np.abs(dec[trial] - out["dec"])) < 1e-4 * np.max(np.abs(out["dec"])) + 1e-9
        assert _rel(enc[trial], out["enc"]) < 1e-4
