"""Smallest stepped run (debug aid for compute-sanitizer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.build()
from sspslam_b200 import scenarios
from sspslam_b200.simulator import Simulator
kind = sys.argv[1] if len(sys.argv) > 1 else "pi"
nt = sys.argv[2] if len(sys.argv) > 2 else "lifrate"
if kind == "pi":
    sc = scenarios.make_pathint(n_trials=3, n_steps=8, ssp_dim=19, pi_n_neurons=40, neuron_type=nt)
else:
    sc = scenarios.make_slam(n_trials=3, n_steps=8, ssp_dim=19, pi_n_neurons=30, mem_n_neurons=70,
                             circonv_n_neurons=16, n_landmarks=6, T=20.0, neuron_type=nt, view_rad=0.6)
sim = Simulator(sc.network, dt=sc.dt, n_trials=3, trial_inputs=sc.trial_inputs)
sim.run_steps(4)
print("ok", sim.data[sc.probe].shape)
sim.close()
