"""BASELINE configs[4] at full network size (3-D, d = 649, 426 380 neurons / trial): run a few steps, check one trial
against the oracle in rate mode, then time a batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
g.build()
from sspslam_b200 import scenarios, lowering
from sspslam_b200.simulator import Simulator
from oracle.nengo_ref_sim import RefSimulator

B = int(os.environ.get("B", "64"))
steps = int(os.environ.get("STEPS", "16"))
nt = os.environ.get("NT", "lifrate")
t0 = time.time()
sc = scenarios.make_slam(n_trials=B, n_steps=4 * steps + 4, ssp_dim=649, pi_n_neurons=500, mem_n_neurons=970,
                         circonv_n_neurons=100, n_landmarks=50, T=20.0, domain_dim=3,
                         grid_points_per_dim=int(os.environ.get('GRID', '30')),
                         distinct_tables=int(os.environ.get('DISTINCT', '2')), neuron_type=nt, view_rad=0.6)
print(f"[cfg5] scenario {time.time()-t0:.1f}s d={sc.ssp_space.ssp_dim}", flush=True)
t0 = time.time()
sim = Simulator(sc.network, dt=sc.dt, n_trials=B, trial_inputs=sc.trial_inputs, chunk_steps=steps)
print(f"[cfg5] build+lower+upload {time.time()-t0:.1f}s stats={ {k: v for k, v in sim.plan.stats.items() if k != 'bytes_by_kind'} }", flush=True)
sim.run_steps(steps)
idx_dev = sim.cleanup_indices()[0].copy()
got = sim.data[sc.probe]
print("[cfg5] finite", bool(np.all(np.isfinite(got))), "shape", got.shape, flush=True)
if os.environ.get("ORACLE", "1") == "1":
    t0 = time.time()
    ref = RefSimulator(sc.network, dt=sc.dt, model=sim.model, trial_seed=sim.trial_seeds[1],
                       node_tables={node: arr[1] for node, arr in sc.trial_inputs.items()})
    ref.run_steps(steps)
    want = ref.data[sc.probe]
    print(f"[cfg5] oracle {time.time()-t0:.1f}s rel err {np.max(np.abs(got[1]-want))/np.max(np.abs(want)):.3e}", flush=True)
    from oracle import ssp_ref
    slam = sc.extra["slam"]
    want_idx = ssp_ref.cleanup_index(slam.sample_ssps, ref.signals[slam.gridcells, "in"].a)
    print(f"[cfg5] clean-up index device {idx_dev[1]} oracle {want_idx} equal {idx_dev[1] == want_idx}", flush=True)
bytes_ts = lowering.algorithmic_bytes_per_trial_step(sim.plan.stats)
for rep in range(2):
    sim.run_steps(steps)
    ms = sim.last_run_ms()
    tps = B * steps / (ms * 1e-3)
    print(f"[cfg5] B={B} {steps} steps: {ms/steps*1e3:.0f} us/step -> {tps/1e3:.1f} k trial-steps/s "
          f"({tps*bytes_ts/1e9/6535.4:.3f} of HBM model, {bytes_ts/1e6:.1f} MB / trial-step)", flush=True)
if os.environ.get("KERNELS"):
    sim.set_profiling(True)
    sim.run_steps(steps)
    for k, (ms, cnt) in sim.kernel_times().items():
        if cnt:
            print(f"[cfg5]   {k:13s} {ms/cnt*1e3:9.1f} us/launch x {cnt/steps:.1f}/step", flush=True)
sim.close()
