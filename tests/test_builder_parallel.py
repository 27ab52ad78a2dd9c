"""The host builder's fork pool (one job per ensemble: sampling, tuning curves, decoder solves of the connections and probes
leaving it) must be indistinguishable from the in-process build: same keys in the same order, bit-identical arrays."""
import dataclasses

import numpy as np

from sspslam_b200 import builder, scenarios


def _same(a, b):
    if a is None or b is None:
        return a is None and b is None
    if isinstance(a, np.ndarray):
        return isinstance(b, np.ndarray) and a.shape == b.shape and np.array_equal(a, b)
    if isinstance(a, dict):
        return list(a) == list(b) and all(_same(a[k], b[k]) for k in a)
    if dataclasses.is_dataclass(a):
        return all(_same(getattr(a, f.name), getattr(b, f.name)) for f in dataclasses.fields(a))
    return a == b


def test_parallel_build_is_bit_identical_to_the_in_process_build(monkeypatch):
    sc = scenarios.make_slam(n_trials=1, n_steps=10, ssp_dim=31, pi_n_neurons=40, mem_n_neurons=90, circonv_n_neurons=12,
                             n_landmarks=6, T=20.0)
    assert len(sc.network.all_ensembles) > 100
    monkeypatch.setenv("SSB_BUILDER_PROCS", "1")
    seq = builder.build_model(sc.network, dt=sc.dt, seed_override=5)
    monkeypatch.setenv("SSB_BUILDER_PROCS", "3")
    par = builder.build_model(sc.network, dt=sc.dt, seed_override=5)
    assert seq.seeds == par.seeds
    assert list(seq.params) == list(par.params)
    for k in seq.params:
        assert _same(seq.params[k], par.params[k]), k
    assert list(seq.probe_conns) == list(par.probe_conns)
    for k in seq.probe_conns:
        assert np.array_equal(seq.probe_conns[k], par.probe_conns[k])


def test_builder_process_count_rules(monkeypatch):
    monkeypatch.delenv("SSB_BUILDER_PROCS", raising=False)
    assert builder._builder_procs(20) == 1                       # small networks stay in-process
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "1024")               # one process per GPU shares the host cores
    assert builder._builder_procs(5000) == 1
    monkeypatch.setenv("SSB_BUILDER_PROCS", "4")
    assert builder._builder_procs(5000) == 4
