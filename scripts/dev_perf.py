"""Throughput probe only (no oracle): B trials of BASELINE configs[1]; prints us/step for a few repetitions."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.build()
from sspslam_b200 import scenarios, lowering
from sspslam_b200.simulator import Simulator

B = int(os.environ.get("B", "1024"))
steps = int(os.environ.get("STEPS", "64"))
reps = int(os.environ.get("REPS", "4"))
cfg = os.environ.get("CONFIG", "slam55")
import numpy as np
DISTINCT = int(os.environ.get("DISTINCT", str(B)))
if cfg == "pathint97":      # BASELINE configs[0]
    sc = scenarios.make_pathint(n_trials=B, n_steps=steps * (reps + 3), ssp_dim=int(os.environ.get("SSP_DIM", "97")), pi_n_neurons=500,
                                neuron_type="lif")
elif cfg == "slamview97":   # BASELINE configs[3] sizes
    sc = scenarios.make_slam(n_trials=B, n_steps=steps * (reps + 3), ssp_dim=97, pi_n_neurons=800, mem_n_neurons=970,
                             circonv_n_neurons=100, n_landmarks=100, T=200.0, length_scale=0.3, view=True, distinct_tables=DISTINCT, table_dtype=np.float32)
elif cfg == "slam55gif":    # BASELINE configs[2] sizes (run_slam_map_gif.py defaults)
    sc = scenarios.make_slam(n_trials=B, n_steps=steps * (reps + 3), ssp_dim=55, pi_n_neurons=800, mem_n_neurons=1000,
                             circonv_n_neurons=100, n_landmarks=50, T=200.0, length_scale=0.1, distinct_tables=DISTINCT, table_dtype=np.float32)
elif cfg == "slam55loihi":  # run_slam.py --backend loihi-sim sizes: SLAMLoihiNetwork, d=55, pi 500, mem 970, circonv 100, dot-product 50
    sc = scenarios.make_slam(n_trials=B, n_steps=steps * (reps + 3), T=200.0, distinct_tables=DISTINCT, table_dtype=np.float32, loihi=True, dotprod_n_neurons=50)
else:
    sc = scenarios.make_slam(n_trials=B, n_steps=steps * (reps + 3), T=200.0, distinct_tables=DISTINCT, table_dtype=np.float32)
if os.environ.get("PER_TRIAL_SEEDS"):   # every trial its own network seed (narrow-ensemble networks): per-trial static weights
    t0 = time.time()
    n_seeds = int(os.environ["PER_TRIAL_SEEDS"])
    sim = Simulator(sc.network, dt=sc.dt, n_trials=B, trial_inputs=sc.trial_inputs, chunk_steps=steps,
                    trial_network_seeds=[1000 + (i % n_seeds) for i in range(B)])
    print(f"[perf] built {n_seeds} models + upload in {time.time() - t0:.1f} s", flush=True)
elif os.environ.get("SYNTH"):      # on-device input synthesis instead of tables
    sim = Simulator(sc.network, dt=sc.dt, n_trials=B, input_synthesis=sc.extra["input_synthesis"], chunk_steps=steps)
else:
    sim = Simulator(sc.network, dt=sc.dt, n_trials=B, trial_inputs=sc.trial_inputs, chunk_steps=steps)
bytes_ts = lowering.algorithmic_bytes_per_trial_step(sim.plan.stats, per_trial_weights=bool(os.environ.get("PER_TRIAL_SEEDS")))
sim.run_steps(steps)
out = []
for rep in range(reps):
    sim.run_steps(steps)
    ms = sim.last_run_ms()
    out.append(ms / steps * 1e3)
tps = B / (min(out) * 1e-6)
print(f"[perf {cfg} {os.environ.get('TAG','')}] neurons {sim.plan.stats['n_neurons']} bytes/trial-step {bytes_ts} us/step " + " ".join(f"{x:.1f}" for x in out) +
      f" -> best {tps/1e6:.3f} M trial-steps/s ({tps*bytes_ts/1e9/6535.4:.3f} of HBM model)")
if os.environ.get("KERNELS"):
    sim.set_profiling(True)
    sim.run_steps(steps)
    for k, (ms, cnt) in sim.kernel_times().items():
        if cnt:
            print(f"[perf]   {k:13s} {ms/cnt*1e3:8.1f} us/launch x {cnt//steps}/step")
sim.close()
