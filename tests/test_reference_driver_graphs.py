"""North star: 'accepts the unchanged nengo.Network built by sspslam.networks'.  Driver-style graphs are built exactly as
``experiments/run_slam.py:151-195`` / ``run_slamview.py:107-141`` do — with the UNMODIFIED reference classes
(``sspslam.networks.SLAMNetwork`` / ``SLAMViewNetwork``), the unmodified ``get_slam_input_functions2`` /
``get_slamview_input_functions`` closures and the drivers' own lambdas — then built, lowered and stepped by this repo's
host path (the NumPy plan interpreter executes the device plan) and compared with the operator-level oracle, which calls
the reference's Python closures (clean-up, gate, identity, inputs) every step.  CPU only; needs /root/reference."""
import numpy as np
import pytest

from conftest import has_reference
from oracle.nengo_ref_sim import RefSimulator
from plan_interp import PlanInterpreter
from sspslam_b200 import inputs, lowering, nengo_shim as nengo
from sspslam_b200.builder import build_model

pytestmark = pytest.mark.skipif(not has_reference(), reason="reference checkout not mounted")
BOUNDS2 = np.tile([-1.0, 1.0], (2, 1))
DT = 0.001


def _ref():
    from sspslam_b200 import refload
    return refload.load_reference()


def _world(ref, n_lm, length_scale):
    space = ref.HexagonalSSPSpace(2, ssp_dim=19, domain_bounds=BOUNDS2, length_scale=length_scale)
    lm_space = ref.SPSpace(n_lm, space.ssp_dim, seed=0)
    path = inputs.random_path(20.0, DT, 0.3, 0, 2)[:400]
    vels = inputs.velocities(path, DT)
    obj_locs = 1.8 * (ref.utils.Rd_sampling(n_lm, 2, seed=0) - 0.5)
    return space, lm_space, path, vels, obj_locs[None, :, :] - path[:, None, :]


def _tables(net, n_steps):
    """What ``Simulator._node_table`` does for an unbatched run: the t-only node callables evaluated at t = n dt."""
    tabs = {}
    for node in net.all_nodes:
        if callable(node.output) and node.size_in == 0:
            tabs[node] = np.stack([np.asarray(node.output((k + 1) * DT), dtype=np.float64).reshape(-1)
                                   for k in range(n_steps)])
    return tabs


def _lower_step_compare(net, probe, n_steps, tol=2e-5):
    model = build_model(net, dt=DT)
    plan = lowering.lower(net, model, chunk_cap=n_steps)
    ref_sim = RefSimulator(net, dt=DT, model=model)           # the oracle calls every Python closure of the graph
    ref_sim.run_steps(n_steps)
    it = PlanInterpreter(plan, model, net, _tables(net, n_steps))
    it.run_steps(n_steps)
    info = [i for i in plan.probes if i.probe is probe][0]
    got, want = it.probe_data(info), ref_sim.data[probe]
    assert np.max(np.abs(want)) > 1e-2
    assert np.max(np.abs(got - want)) <= tol * np.max(np.abs(want))
    return plan


def test_unmodified_reference_slam_driver_graph_lowers_and_steps():
    ref = _ref()
    from sspslam.networks.slam import get_slam_input_functions2
    n_lm, view_rad = 6, 0.6
    space, lm_space, path, vels, vec_to = _world(ref, n_lm, 0.2)
    d = space.ssp_dim
    real_ssp = space.encode(path)
    (velocity_func, scale, is_landmark_in_view, _ids, landmark_sp_func, _vec,
     landmark_vecssp_func) = get_slam_input_functions2(space, lm_space, vels, vec_to, view_rad)
    np.random.seed(0)
    model = nengo.Network(seed=0)
    with model:                                                            # run_slam.py:153-195, default arguments
        vel_input = nengo.Node(velocity_func, label='vel_input')
        init_state = nengo.Node(lambda t: real_ssp[int((t - DT) / DT)] if t < 0.05 else np.zeros(d), label='init_state')
        landmark_vec = nengo.Node(landmark_vecssp_func, label='lm_vecssp_input')
        landmark_id = nengo.Node(landmark_sp_func, label='lm_sp_input')
        is_landmark = nengo.Node(is_landmark_in_view, label='lm_in_view_input')
        slam = ref.networks.SLAMNetwork(space, lm_space, view_rad, n_lm, 30, 64, 16, tau_pi=0.05, update_thres=0.2,
                                        vel_scaling_factor=scale, shift_rate=0.2, voja_learning_rate=1e-4,
                                        pes_learning_rate=5e-3, intercept=0.1, clean_up_method='grid', gc_n_neurons=0,
                                        encoders=None, voja=True, seed=0)
        nengo.Connection(landmark_vec, slam.landmark_vec_ssp, synapse=None)
        nengo.Connection(landmark_id, slam.landmark_id_input, synapse=None)
        nengo.Connection(is_landmark, slam.no_landmark_in_view, synapse=None)
        nengo.Connection(vel_input, slam.velocity_input, synapse=None)
        nengo.Connection(init_state, slam.pathintegrator.input, synapse=None)
        probe = nengo.Probe(slam.pathintegrator.output, synapse=0.05)
        nengo.Probe(slam.assomemory.conn_out, "weights", sample_every=0.09)
    plan = _lower_step_compare(model, probe, 90)
    assert plan.arrays["cleanup"].shape[0] == 1 and plan.arrays["gate"].shape[0] == 1      # the reference's closures
    assert len(plan.learned_dec) == 1 and len(plan.learned_enc) == 1                        # PES + Voja
    assert [p.kind for p in plan.probes] == ["rows", "weights"]


def test_unmodified_reference_slamview_driver_graph_lowers_and_steps():
    ref = _ref()
    from sspslam.networks.slam_view import get_slamview_input_functions
    n_lm, view_rad = 6, 0.6
    space, lm_space, path, vels, vec_to = _world(ref, n_lm, 0.3)
    d = space.ssp_dim
    real_ssp = space.encode(path)
    velocity_func, scale, is_landmark_in_view, landmark_func = get_slamview_input_functions(space, lm_space, vels, vec_to,
                                                                                             view_rad)
    model = nengo.Network(seed=0)
    with model:                                                            # run_slamview.py:107-141
        vel_input = nengo.Node(velocity_func, label='vel_input')
        init_state = nengo.Node(lambda t: real_ssp[int((t - DT) / DT)] if t < 0.05 else np.zeros(d), label='init_state')
        landmark_input = nengo.Node(landmark_func)
        landmark_inview = nengo.Node(is_landmark_in_view)
        slam = ref.networks.SLAMViewNetwork(space, lm_space, view_rad, n_lm, 30, 64, 16, tau_pi=0.05, update_thres=0.2,
                                            vel_scaling_factor=scale, shift_rate=0.02, voja_learning_rate=5e-4,
                                            pes_learning_rate=1e-3, clean_up_method='grid', gc_n_neurons=0, encoders=None,
                                            voja=True, seed=0)
        nengo.Connection(landmark_input, slam.view_input, synapse=None)
        nengo.Connection(landmark_inview, slam.no_landmark_in_view, synapse=None)
        nengo.Connection(vel_input, slam.velocity_input, synapse=None)
        nengo.Connection(init_state, slam.pathintegrator.input, synapse=None)
        probe = nengo.Probe(slam.pathintegrator.output, synapse=0.05)
        nengo.Probe(slam.assomemory.recall, synapse=0.05)
    plan = _lower_step_compare(model, probe, 90)
    assert plan.arrays["cleanup"].shape[0] == 1 and plan.arrays["gate"].shape[0] == 1
