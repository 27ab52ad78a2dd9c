"""N > 1 path on CPU: world_size-2 gloo processes shard the trials, step their own block with the
CPU checker standing in for the device, and all_gather the per-trial statistics.  The result must
equal the single-process run (trial outcome depends on the global trial id only)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT
from sspslam_b200 import sharding


def test_shard_range_covers_everything_once():
    for n, w in ((1024, 8), (10, 4), (3, 8), (4096, 8), (7, 2)):
        seen = []
        for r in range(w):
            lo, hi = sharding.shard_range(n, r, w)
            assert 0 <= lo <= hi <= n
            seen += list(range(lo, hi))
        assert seen == list(range(n))
        sizes = [sharding.shard_range(n, r, w)[1] - sharding.shard_range(n, r, w)[0] for r in range(w)]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)
    assert sharding.trial_seeds(10, 1, 4, base_seed=100) == [103, 104, 105]


def _trial_stat(scenario, model, trial, seed, n_steps):
    from oracle.nengo_ref_sim import RefSimulator
    tabs = {node: arr[trial] for node, arr in scenario.trial_inputs.items()}
    ref = RefSimulator(scenario.network, dt=scenario.dt, model=model, node_tables=tabs, trial_seed=seed)
    ref.run_steps(n_steps)
    out = ref.data[scenario.probe]
    real = scenario.real_ssp[trial, :n_steps]
    cos = np.sum(out[-1] * real[-1]) / (np.linalg.norm(out[-1]) * np.linalg.norm(real[-1]) + 1e-12)
    return [cos, float(np.linalg.norm(out[-1])), float(n_steps)]


def _worker(rank, world, port, n_trials, n_steps, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from sspslam_b200 import scenarios
    from sspslam_b200.builder import build_model
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = scenarios.make_pathint(n_trials=n_trials, n_steps=n_steps, ssp_dim=7, pi_n_neurons=20, neuron_type="lif")
        model = build_model(sc.network, dt=sc.dt)          # same seed on every rank -> same weights
        lo, hi = sharding.shard_range(n_trials, rank, world)
        seeds = sharding.trial_seeds(n_trials, rank, world)
        local = np.array([_trial_stat(sc, model, t, s, n_steps) for t, s in zip(range(lo, hi), seeds)], dtype=np.float32)
        local = local.reshape(hi - lo, 3)
        full = sharding.gather_trial_stats(local, n_trials)
        slow = sharding.max_over_ranks(10.0 + rank)
        if rank == 0:
            q.put((full, slow))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_world_size_2_matches_single_process():
    n_trials, n_steps = 5, 30          # ragged: blocks of 3 and 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_trials, n_steps, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, slow = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert full.shape == (n_trials, 3) and slow == 11.0
    from sspslam_b200 import scenarios
    from sspslam_b200.builder import build_model
    sc = scenarios.make_pathint(n_trials=n_trials, n_steps=n_steps, ssp_dim=7, pi_n_neurons=20, neuron_type="lif")
    model = build_model(sc.network, dt=sc.dt)
    want = np.array([_trial_stat(sc, model, t, t, n_steps) for t in range(n_trials)], dtype=np.float32)
    np.testing.assert_array_equal(full, want)
    assert len({tuple(r) for r in full.tolist()}) == n_trials      # trials really differ
