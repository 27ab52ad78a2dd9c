"""Development check on a GPU box: oracle comparisons + a quick throughput probe."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.build()
from sspslam_b200 import scenarios, lowering
from sspslam_b200.simulator import Simulator
from oracle.nengo_ref_sim import RefSimulator


def compare(sc, n_steps, n_trials, label, check_trials=(0,)):
    sim = Simulator(sc.network, dt=sc.dt, n_trials=n_trials, trial_inputs=sc.trial_inputs)
    t0 = time.time()
    sim.run_steps(n_steps)
    got = sim.data[sc.probe]
    print(f"[{label}] gpu run {time.time()-t0:.2f}s launches={sim.total_launches()}")
    for tr in check_trials:
        tabs = {node: arr[tr] for node, arr in sc.trial_inputs.items()}
        ref = RefSimulator(sc.network, dt=sc.dt, model=sim.model, node_tables=tabs, trial_seed=sim.trial_seeds[tr])
        t0 = time.time()
        ref.run_steps(n_steps)
        want = ref.data[sc.probe]
        scale = max(1e-12, np.max(np.abs(want)))
        err = np.max(np.abs(got[tr] - want), axis=1) / scale
        print(f"[{label}] trial {tr}: oracle {time.time()-t0:.2f}s  rel err max {err.max():.3e} "
              f"@10 {err[min(10,n_steps-1)]:.2e} @50 {err[min(50,n_steps-1)]:.2e} last {err[-1]:.2e} |want| {scale:.3f}")
    sim.close()
    return got


if __name__ == "__main__":
    which = sys.argv[1:] or ["pi_rate", "pi_lif", "slam_rate", "slam_lif", "perf"]
    if "pi_rate" in which:
        sc = scenarios.make_pathint(n_trials=3, n_steps=300, ssp_dim=55, pi_n_neurons=200, neuron_type="lifrate")
        compare(sc, 300, 3, "pi_rate", (0, 2))
    if "pi_lif" in which:
        sc = scenarios.make_pathint(n_trials=3, n_steps=300, ssp_dim=55, pi_n_neurons=200, neuron_type="lif")
        compare(sc, 300, 3, "pi_lif", (0, 2))
    if "slam_rate" in which:
        sc = scenarios.make_slam(n_trials=3, n_steps=200, ssp_dim=55, pi_n_neurons=100, mem_n_neurons=200,
                                 circonv_n_neurons=30, n_landmarks=20, T=20.0, neuron_type="lifrate")
        compare(sc, 200, 3, "slam_rate", (0, 1))
    if "slam_lif" in which:
        sc = scenarios.make_slam(n_trials=3, n_steps=200, ssp_dim=55, pi_n_neurons=100, mem_n_neurons=200,
                                 circonv_n_neurons=30, n_landmarks=20, T=20.0, neuron_type="lif")
        compare(sc, 200, 3, "slam_lif", (0, 1))
    if "perf" in which:
        B = int(os.environ.get("B", "1024"))
        steps = 50
        t0 = time.time()
        sc = scenarios.make_slam(n_trials=B, n_steps=steps * 6, T=200.0, distinct_tables=8)
        print(f"[perf] scenario {time.time()-t0:.1f}s")
        t0 = time.time()
        sim = Simulator(sc.network, dt=sc.dt, n_trials=B, trial_inputs=sc.trial_inputs, chunk_steps=steps)
        print(f"[perf] build+upload {time.time()-t0:.1f}s stats={sim.plan.stats}")
        bytes_ts = lowering.algorithmic_bytes_per_trial_step(sim.plan.stats)
        sim.run_steps(steps)
        for rep in range(3):
            t0 = time.time()
            sim.run_steps(steps)
            wall = time.time() - t0
            ms = sim.last_run_ms()
            tps = B * steps / (ms * 1e-3)
            print(f"[perf] B={B} {steps} steps: device {ms:.2f} ms ({ms/steps*1e3:.1f} us/step) wall {wall*1e3:.1f} ms "
                  f"-> {tps/1e6:.3f} M trial-steps/s, {tps*bytes_ts/1e9:.0f} GB/s algorithmic "
                  f"({tps*bytes_ts/1e9/6535.4:.3f} of measured HBM)")
        sim.set_profiling(True)
        sim.run_steps(steps)
        kt = sim.kernel_times()
        tot = sum(v[0] for v in kt.values())
        for k, (ms, cnt) in kt.items():
            if cnt:
                print(f"[perf]   {k:13s} {ms:8.2f} ms  {cnt:5d} launches  {ms/cnt*1e3:8.1f} us/launch  {ms/tot*100:5.1f}%")
        sim.close()
