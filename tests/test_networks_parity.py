"""The from-scratch network declarations (sspslam_b200.networks) against the UNMODIFIED reference
``sspslam.networks`` loaded through the nengo shim: built from the same seed, both must give the same
object census, seeds, encoders, gains, decoders and transforms, and the oracle must produce the same
probe trajectory.  Runs only where /root/reference is mounted (the authoring container)."""
import numpy as np
import pytest

from conftest import has_reference
from sspslam_b200 import nengo_shim as nengo, networks, inputs
from sspslam_b200.builder import build_model
from sspslam_b200.sspspace import HexagonalSSPSpace, SPSpace

pytestmark = pytest.mark.skipif(not has_reference(), reason="reference checkout not mounted")
BOUNDS2 = np.tile([-1.0, 1.0], (2, 1))


def _ref():
    from sspslam_b200 import refload
    return refload.load_reference()


def _census(net):
    return (len(net.all_ensembles), len(net.all_nodes), len(net.all_connections),
            sum(e.n_neurons for e in net.all_ensembles))


def _assert_same_model(net_a, net_b):
    assert _census(net_a) == _census(net_b)
    ma, mb = build_model(net_a), build_model(net_b)
    for ea, eb in zip(net_a.all_ensembles, net_b.all_ensembles):
        assert (ea.n_neurons, ea.dimensions, ea.radius) == (eb.n_neurons, eb.dimensions, eb.radius)
        assert ma.seeds[ea] == mb.seeds[eb]
        pa, pb = ma.params[ea], mb.params[eb]
        np.testing.assert_allclose(pa.scaled_encoders, pb.scaled_encoders, rtol=0, atol=1e-12)
        np.testing.assert_allclose(pa.bias, pb.bias, rtol=0, atol=1e-12)
    for ca, cb in zip(net_a.all_connections, net_b.all_connections):
        assert ca.size_in == cb.size_in and ca.size_out == cb.size_out
        assert (ca.synapse is None) == (cb.synapse is None)
        if ca.synapse is not None:
            assert ca.synapse.tau == cb.synapse.tau
        wa, wb = ma.params[ca].weights, mb.params[cb].weights
        if wa is None or wb is None:
            assert wa is None and wb is None
        else:
            np.testing.assert_allclose(np.asarray(wa, dtype=float), np.asarray(wb, dtype=float), rtol=0, atol=1e-9)
    return ma, mb


def test_pathintegration_matches_reference():
    ref = _ref()
    space = HexagonalSSPSpace(2, ssp_dim=19, domain_bounds=BOUNDS2, length_scale=0.2, backend="host")
    rspace = ref.HexagonalSSPSpace(2, ssp_dim=19, domain_bounds=BOUNDS2, length_scale=0.2)
    assert np.array_equal(space.phase_matrix, rspace.phase_matrix)
    nets = []
    for mod, sp in ((networks, space), (ref.networks, rspace)):
        with nengo.Network(seed=4) as net:
            pi = mod.PathIntegration(sp, 40, 0.05, scaling_factor=0.7, stable=True, solver_weights=False)
            nengo.Probe(pi.output, synapse=0.05)
        nets.append(net)
    _assert_same_model(*nets)


def test_slam_network_matches_reference():
    ref = _ref()
    space = HexagonalSSPSpace(2, ssp_dim=19, domain_bounds=BOUNDS2, length_scale=0.2, backend="host")
    rspace = ref.HexagonalSSPSpace(2, ssp_dim=19, domain_bounds=BOUNDS2, length_scale=0.2)
    lm, rlm = SPSpace(6, space.ssp_dim, seed=2), ref.SPSpace(6, rspace.ssp_dim, seed=2)
    nets = []
    for mod, sp, l in ((networks, space, lm), (ref.networks, rspace, rlm)):
        np.random.seed(9)     # OVC encoders come from the global stream (slam.py:206)
        with nengo.Network(seed=4) as net:
            slam = mod.SLAMNetwork(sp, l, 0.2, 6, 30, 64, 16, tau_pi=0.05, update_thres=0.2, vel_scaling_factor=0.7,
                                   shift_rate=0.2, voja_learning_rate=1e-4, pes_learning_rate=5e-3, intercept=0.1)
            nengo.Probe(slam.pathintegrator.output, synapse=0.05)
        nets.append((net, slam))
    (na, sa), (nb, sb) = nets
    _assert_same_model(na, nb)
    np.testing.assert_allclose(sa.sample_ssps, sb.sample_ssps, atol=1e-14)
    # the reference's anonymous closures are recognised as the same device ops
    from sspslam_b200 import nodeops
    owners = [nb] + nb.all_networks
    kinds = sorted(nodeops.recognize(n, owners).kind for n in nb.all_nodes
                   if callable(n.output) and n.size_in > 0)
    assert kinds == ["cleanup", "gate", "identity"]
    gate = nodeops.recognize(sb.update_state, owners)
    assert (gate.d, gate.shift_rate, gate.update_thres) == (space.ssp_dim, 0.2, 0.2)


def test_slam_network_with_grid_cell_ensemble_matches_reference():
    """gc_n_neurons > 0 (slam.py:274-281, sample_grid_encoders sspspace.py:733-762, CosineSimilarity)."""
    ref = _ref()
    lm_seed = 2
    nets = []
    for mod, space_cls, sp_cls in ((networks, HexagonalSSPSpace, SPSpace), (ref.networks, ref.HexagonalSSPSpace, ref.SPSpace)):
        kw = dict(backend="host") if space_cls is HexagonalSSPSpace else {}
        sp = space_cls(2, ssp_dim=19, domain_bounds=BOUNDS2, length_scale=0.2, rng=np.random.default_rng(5), **kw)
        l = sp_cls(6, sp.ssp_dim, seed=lm_seed)
        np.random.seed(9)
        with nengo.Network(seed=4) as net:
            slam = mod.SLAMNetwork(sp, l, 0.2, 6, 30, 64, 16, tau_pi=0.05, update_thres=0.2, vel_scaling_factor=0.7,
                                   shift_rate=0.2, voja_learning_rate=1e-4, pes_learning_rate=5e-3, intercept=0.1,
                                   gc_n_neurons=48)
            nengo.Probe(slam.pathintegrator.output, synapse=0.05)
        nets.append((net, slam))
    (na, sa), (nb, sb) = nets
    _assert_same_model(na, nb)
    assert sa.gridcells.n_neurons == 48 and sa.gridcells.dimensions == sa.sample_ssps.shape[1]
    np.testing.assert_allclose(np.asarray(sa.gridcells.encoders), np.asarray(sb.gridcells.encoders), atol=1e-14)


def test_slamview_network_matches_reference():
    ref = _ref()
    space = HexagonalSSPSpace(2, ssp_dim=19, domain_bounds=BOUNDS2, length_scale=0.3, backend="host")
    rspace = ref.HexagonalSSPSpace(2, ssp_dim=19, domain_bounds=BOUNDS2, length_scale=0.3)
    lm, rlm = SPSpace(6, space.ssp_dim, seed=2), ref.SPSpace(6, rspace.ssp_dim, seed=2)
    nets = []
    for mod, sp, l in ((networks, space, lm), (ref.networks, rspace, rlm)):
        with nengo.Network(seed=4) as net:
            mod.SLAMViewNetwork(sp, l, 0.2, 6, 30, 64, 16, tau_pi=0.05, update_thres=0.2, vel_scaling_factor=0.7,
                                shift_rate=0.02, voja_learning_rate=5e-4, pes_learning_rate=1e-3)
        nets.append(net)
    _assert_same_model(*nets)


def test_reference_network_steps_identically_in_the_oracle():
    """Unmodified reference PathIntegration vs. this repo's declaration: same oracle trajectory."""
    from oracle.nengo_ref_sim import RefSimulator
    ref = _ref()
    space = HexagonalSSPSpace(2, ssp_dim=19, domain_bounds=BOUNDS2, length_scale=0.2, backend="host")
    rspace = ref.HexagonalSSPSpace(2, ssp_dim=19, domain_bounds=BOUNDS2, length_scale=0.2)
    path = inputs.random_path(20.0, 0.001, 0.1, 0, 2)
    vels = inputs.velocities(path)
    scale = inputs.velocity_scale(space.phase_matrix, vels)
    tb = inputs.pathint_tables(space.encode_host(path), vels * scale, 120)
    out = []
    for mod, sp in ((networks, space), (ref.networks, rspace)):
        with nengo.Network(seed=4) as net:
            vel = nengo.Node(lambda t: tb["vel"][int(round(t / 0.001)) - 1])
            init = nengo.Node(lambda t: tb["init"][int(round(t / 0.001)) - 1])
            pi = mod.PathIntegration(sp, 40, 0.05, scaling_factor=scale, stable=True, solver_weights=False)
            nengo.Connection(vel, pi.velocity_input, synapse=None)
            nengo.Connection(init, pi.input, synapse=None)
            p = nengo.Probe(pi.output, synapse=0.05)
        sim = RefSimulator(net)
        sim.run_steps(120)
        out.append(sim.data[p])
    np.testing.assert_allclose(out[0], out[1], rtol=0, atol=1e-9)
    assert np.max(np.abs(out[0])) > 0.05


def test_slam_loihi_network_matches_reference():
    """the all-neural variant (slam_loihi.py:190-293) — same census, seeds, encoders, decoders."""
    ref = _ref()
    space = HexagonalSSPSpace(2, ssp_dim=19, domain_bounds=BOUNDS2, length_scale=0.2, backend="host")
    rspace = ref.HexagonalSSPSpace(2, ssp_dim=19, domain_bounds=BOUNDS2, length_scale=0.2)
    lm, rlm = SPSpace(6, space.ssp_dim, seed=2), ref.SPSpace(6, rspace.ssp_dim, seed=2)
    nets = []
    for mod, sp, l in ((networks, space, lm), (ref.networks, rspace, rlm)):
        with nengo.Network(seed=4) as net:
            slam = mod.SLAMLoihiNetwork(sp, l, 0.2, 6, 30, 64, 16, 20, tau_pi=0.05, update_thres=0.2,
                                        vel_scaling_factor=0.7, shift_rate=0.1, pes_learning_rate=1e-3, seed=3)
            nengo.Probe(slam.pathintegrator.output, synapse=0.05)
        nets.append((net, slam))
    (na, sa), (nb, sb) = nets
    _assert_same_model(na, nb)
    assert [e.label for e in na.all_ensembles] == [e.label for e in nb.all_ensembles]
    from sspslam_b200 import nodeops
    owners = [nb] + nb.all_networks     # no clean-up / gate node functions in this variant, only the PI pass-through
    assert [nodeops.recognize(n, owners).kind for n in nb.all_nodes if callable(n.output) and n.size_in > 0] == ["identity"]
    np.testing.assert_allclose(np.asarray(sa.assomemory.memory.encoders), np.asarray(sb.assomemory.memory.encoders))


def test_pathintegration_with_grid_cell_output_matches_reference():
    """pathintegration.py:150-154: ``with_gcs=True`` — the output node becomes a grid-cell population whose intercept comes
    from ``sparsity_to_x_intercept`` (sspslam/utils/utils.py:5-10)."""
    ref = _ref()
    from sspslam_b200.inputs import sparsity_to_x_intercept
    for d, p in ((7, 0.1), (55, 0.1), (19, 0.7)):
        assert sparsity_to_x_intercept(d, p) == pytest.approx(ref.utils.sparsity_to_x_intercept(d, p), abs=1e-15)
    nets = []
    for mod, space_cls in ((networks, HexagonalSSPSpace), (ref.networks, ref.HexagonalSSPSpace)):
        kw = dict(backend="host") if space_cls is HexagonalSSPSpace else {}
        sp = space_cls(2, ssp_dim=19, domain_bounds=BOUNDS2, length_scale=0.2, rng=np.random.default_rng(5), **kw)
        with nengo.Network(seed=4) as net:
            pi = mod.PathIntegration(sp, 40, 0.05, scaling_factor=0.7, stable=True, solver_weights=False,
                                     with_gcs=True, n_gcs=64)
            nengo.Probe(pi.output, synapse=0.05)
        nets.append((net, pi))
    (na, pa), (nb, pb) = nets
    _assert_same_model(na, nb)
    assert pa.output.n_neurons == 64
    np.testing.assert_allclose(np.asarray(pa.output.encoders), np.asarray(pb.output.encoders), atol=1e-14)


def test_slam_network_without_voja_matches_reference():
    """run_slam.py --no-voja: the memory encoders are landmark SPs drawn with RandomState(seed) (slam.py:196-198)."""
    ref = _ref()
    space = HexagonalSSPSpace(2, ssp_dim=19, domain_bounds=BOUNDS2, length_scale=0.2, backend="host")
    rspace = ref.HexagonalSSPSpace(2, ssp_dim=19, domain_bounds=BOUNDS2, length_scale=0.2)
    lm, rlm = SPSpace(6, space.ssp_dim, seed=2), ref.SPSpace(6, rspace.ssp_dim, seed=2)
    nets = []
    for mod, sp, l in ((networks, space, lm), (ref.networks, rspace, rlm)):
        np.random.seed(9)
        with nengo.Network(seed=4) as net:
            slam = mod.SLAMNetwork(sp, l, 0.2, 6, 30, 64, 16, tau_pi=0.05, update_thres=0.2, vel_scaling_factor=0.7,
                                   shift_rate=0.2, pes_learning_rate=5e-3, intercept=0.1, voja=False, seed=5)
            nengo.Probe(slam.pathintegrator.output, synapse=0.05)
        nets.append((net, slam))
    (na, sa), (nb, sb) = nets
    _assert_same_model(na, nb)
    assert not [c for c in na.all_connections if c.learning_rule is not None
                and isinstance(c.learning_rule.learning_rule_type, nengo.Voja)]


@pytest.mark.parametrize("name", ["get_slam_input_functions", "get_slam_input_functions2"])
def test_slam_input_function_closures_match_reference(name):
    """The node callables handed to nengo.Node by the drivers (run_slam.py:139-141,164-169), at step times and between."""
    ref = _ref()
    dt = 0.001
    space = HexagonalSSPSpace(2, ssp_dim=19, domain_bounds=BOUNDS2, length_scale=0.2, backend="host")
    rspace = ref.HexagonalSSPSpace(2, ssp_dim=19, domain_bounds=BOUNDS2, length_scale=0.2)
    lm, rlm = SPSpace(8, space.ssp_dim, seed=2), ref.SPSpace(8, rspace.ssp_dim, seed=2)
    path = inputs.random_path(20.0, dt, 0.3, 3, 2)[:2000]
    vels = inputs.velocities(path, dt)
    obj = 1.8 * (inputs.rd_sampling(8, 2, seed=3) - 0.5)
    vec_to = obj[None, :, :] - path[:, None, :]
    mine = getattr(networks, name)(space, lm, vels, vec_to, 0.45, dt)
    theirs = getattr(ref.networks.slam, name)(rspace, rlm, vels, vec_to, 0.45, dt)
    assert mine[1] == pytest.approx(theirs[1], rel=1e-15)
    seen = set()
    times = np.concatenate([dt * np.arange(1, 1500, 7), np.random.default_rng(0).uniform(dt, 1.9, 100)])
    for t in times:
        for k in (0, 2, 4, 5, 6):
            np.testing.assert_allclose(np.asarray(mine[k](t), dtype=float), np.asarray(theirs[k](t), dtype=float),
                                       rtol=0, atol=1e-12)
        a, b = mine[3](t), theirs[3](t)
        assert (a is None and b is None) or np.array_equal(a, b)
        seen.add(0 if mine[2](t) else (1 if np.size(a) == 1 else 2))
    assert seen >= {0, 1} and (name.endswith("functions") or 2 in seen)    # none / one / several landmarks in view


def test_slamview_input_functions_and_tables_match_reference():
    """slam_view.py:281-404: the local-view closures, and their batched table form, against the reference closures."""
    ref = _ref()
    dt = 0.001
    space = HexagonalSSPSpace(2, ssp_dim=19, domain_bounds=BOUNDS2, length_scale=0.2, backend="host")
    rspace = ref.HexagonalSSPSpace(2, ssp_dim=19, domain_bounds=BOUNDS2, length_scale=0.2)
    lm, rlm = SPSpace(8, space.ssp_dim, seed=2), ref.SPSpace(8, rspace.ssp_dim, seed=2)
    path = inputs.random_path(20.0, dt, 0.3, 3, 2)[:2000]
    vels = inputs.velocities(path, dt)
    obj = 1.8 * (inputs.rd_sampling(8, 2, seed=3) - 0.5)
    vec_to = obj[None, :, :] - path[:, None, :]
    mine = networks.get_slamview_input_functions(space, lm, vels, vec_to, 0.45, dt)
    theirs = ref.networks.slam_view.get_slamview_input_functions(rspace, rlm, vels, vec_to, 0.45, dt)
    assert mine[1] == pytest.approx(theirs[1], rel=1e-15)
    n = 1500
    tb = inputs.slamview_tables(space, lm.vectors, vels * mine[1], vec_to, 0.45, n, dt)
    flags = set()
    for k in range(1, n + 1, 3):
        t = k * dt
        for i in (0, 2, 3):
            np.testing.assert_allclose(np.asarray(mine[i](t), dtype=float), np.asarray(theirs[i](t), dtype=float), rtol=0, atol=1e-12)
        np.testing.assert_allclose(tb["vel"][k - 1], theirs[0](t), rtol=0, atol=1e-15)
        np.testing.assert_allclose(tb["view"][k - 1], theirs[3](t), rtol=0, atol=1e-12)
        assert tb["nolm"][k - 1, 0] == theirs[2](t)
        flags.add(theirs[2](t))
    assert flags == {0, 1}
