"""Concurrency timeline of steady-state steps (per-launch CUDA events on the dependency streams; direct launches):
   B=1024 STEPS0=200 STEPS=24 python scripts/dev_timeline.py   -> prints one line per launch of the last few steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
g.build()
from sspslam_b200 import scenarios
from sspslam_b200.simulator import Simulator

B = int(os.environ.get("B", "1024"))
steps0 = int(os.environ.get("STEPS0", "200"))
steps = int(os.environ.get("STEPS", "24"))
DISTINCT = int(os.environ.get("DISTINCT", str(B)))
sc = scenarios.make_slam(n_trials=B, n_steps=steps0 + steps + 4, T=200.0, distinct_tables=DISTINCT, table_dtype=np.float32)
sim = Simulator(sc.network, dt=sc.dt, n_trials=B, trial_inputs=sc.trial_inputs, chunk_steps=max(steps0, steps))
sim.run_steps(steps0)
sim.sync()
sim.set_profiling(True, timeline=True)
sim.run_steps(steps)
tl = sim.timeline()
sim.set_profiling(False)
t_first = min(a for _, a, _ in tl)
t_last = max(b for _, _, b in tl)
print(f"[timeline] {len(tl)} launches over {steps} steps, span {t_last - t_first:.1f} us -> {(t_last - t_first) / steps:.1f} us/step")
# steps are delimited by the 'advance' launch (one per run_steps call) -> use the 'begin' launches instead
per_step = len(tl) // steps
show0 = (steps - 9) * per_step if steps > 9 else 0
base = tl[show0][1]
busy = np.zeros(int((t_last - base) * 10) + 2)
for kind, a, b in tl[show0:]:
    print(f"[timeline] {kind:13s} start {a - base:8.1f} end {b - base:8.1f} dur {b - a:6.1f}")
    busy[int((a - base) * 10):int((b - base) * 10) + 1] += 1
print("[timeline] concurrency histogram (fraction of time with k kernels in flight):",
      {int(k): round(float(np.mean(busy == k)), 3) for k in np.unique(busy)})
sim.close()
