"""Short full-size run for ncu captures: B trials of BASELINE configs[1], a few direct (non-graph) steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.build()
from sspslam_b200 import scenarios
from sspslam_b200.simulator import Simulator

B = int(os.environ.get("B", "1024"))
steps = int(os.environ.get("STEPS", "6"))
import numpy as np
DISTINCT = int(os.environ.get("DISTINCT", str(B)))
sc = scenarios.make_slam(n_trials=B, n_steps=steps + 4, T=200.0, distinct_tables=DISTINCT, table_dtype=np.float32)
sim = Simulator(sc.network, dt=sc.dt, n_trials=B, trial_inputs=sc.trial_inputs, chunk_steps=steps)
sim.run_steps(steps)      # fewer than 16 steps: direct launches, no graph replay
sim.sync()
print("ok", sim.data[sc.probe].shape, "launches", sim.total_launches())
sim.close()
