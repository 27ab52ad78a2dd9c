"""Experiment: K independent sub-batches (B/K trials each, own streams and graphs) stepped concurrently, so that the
latency-bound phases of one sub-batch overlap the throughput-bound phases of another.  B=1024 SPLITS=1,2,4 python scripts/dev_split.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
g.build()
from sspslam_b200 import scenarios
from sspslam_b200.simulator import Simulator

B = int(os.environ.get("B", "1024"))
steps = int(os.environ.get("STEPS", "64"))
reps = int(os.environ.get("REPS", "6"))
DISTINCT = int(os.environ.get("DISTINCT", str(B)))
total = steps * (reps + 3)
sc = scenarios.make_slam(n_trials=B, n_steps=total + 2, T=200.0, distinct_tables=DISTINCT, table_dtype=np.float32)
model = None
for K in [int(x) for x in os.environ.get("SPLITS", "1,2,4").split(",")]:
    per = B // K
    sims = []
    for k in range(K):
        sl = slice(k * per, (k + 1) * per)
        ti = {n: a[sl] for n, a in sc.trial_inputs.items()}
        sim = Simulator(sc.network, dt=sc.dt, n_trials=per, trial_inputs=ti, trial_seeds=list(range(k * per, (k + 1) * per)),
                        chunk_steps=total, model=model)
        model = sim.model
        sim.stage_inputs(0, total)
        sim.load_tables(0, total)
        sims.append(sim)
    for _ in range(2):
        for s in sims:
            s.run_resident(steps)
    for s in sims:
        s.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        for s in sims:
            s.run_resident(steps)
    for s in sims:
        s.sync()
    dt = time.perf_counter() - t0
    print(f"[split] K={K} x {per} trials: {dt / (reps * steps) * 1e6:.1f} us per timestep of the whole batch -> "
          f"{B * reps * steps / dt / 1e6:.3f} M trial-steps/s", flush=True)
    for s in sims:
        s.close()
