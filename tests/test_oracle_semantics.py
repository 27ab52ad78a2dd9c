"""Known-answer tests that pin the operator-level oracle (oracle/nengo_ref_sim.py) to nengo's
published step semantics: time/step order, LIF / LIFRate equations, Lowpass
one-step delay, PES and Voja deltas applied at the next step, probe sampling.  The reference ships no
golden vectors for the stepped path, so these analytic cases are the pin."""
import numpy as np

from oracle.nengo_ref_sim import RefSimulator
from sspslam_b200 import nengo_shim as nengo
from sspslam_b200.builder import build_model


def test_time_and_node_function_sees_t_equal_dt_first():
    seen = []
    with nengo.Network(seed=0) as net:
        n = nengo.Node(lambda t: (seen.append(t), t)[1])
        p = nengo.Probe(n)
    seen.clear()
    sim = RefSimulator(net)
    sim.run_steps(3)
    np.testing.assert_allclose(seen, [0.001, 0.002, 0.003])
    np.testing.assert_allclose(sim.data[p][:, 0], [0.001, 0.002, 0.003])
    np.testing.assert_allclose(sim.trange(), [0.001, 0.002, 0.003])


def test_lowpass_is_zoh_with_one_step_delay():
    with nengo.Network(seed=0) as net:
        n = nengo.Node(lambda t: 1.0)
        out = nengo.Node(size_in=1)
        nengo.Connection(n, out, synapse=0.005)
        p = nengo.Probe(out)
    sim = RefSimulator(net)
    sim.run_steps(6)
    a = np.exp(-0.001 / 0.005)
    want = np.array([0.0] + [1 - a ** k for k in range(1, 6)])   # update is visible only to the next step
    np.testing.assert_allclose(sim.data[p][:, 0], want, atol=1e-15)


def test_lifrate_matches_tuning_curve_and_decoders_reconstruct():
    with nengo.Network(seed=3) as net:
        net.config[nengo.Ensemble].neuron_type = nengo.LIFRate()
        stim = nengo.Node(lambda t: [0.4, -0.3])
        ens = nengo.Ensemble(200, 2)
        nengo.Connection(stim, ens, synapse=None)
        p = nengo.Probe(ens)            # decoded output, no synapse
    model = build_model(net)
    sim = RefSimulator(net, model=model)
    sim.run_steps(2)
    prm = model.params[ens]
    J = prm.gain * (prm.encoders @ np.array([0.4, -0.3])) + prm.bias
    j = J - 1
    rates = np.where(j > 0, 1.0 / (0.002 + 0.02 * np.log1p(1.0 / np.where(j > 0, j, 1.0))), 0.0)
    np.testing.assert_allclose(sim.signals[ens, "out"].a, rates, rtol=1e-12)
    dec = sim.data[p][-1]
    assert np.linalg.norm(dec - [0.4, -0.3]) < 0.05
    # gain / bias solve max_rate at x.e = 1 and zero rate at the intercept
    top = prm.gain * 1.0 + prm.bias - 1
    np.testing.assert_allclose(1.0 / (0.002 + 0.02 * np.log1p(1.0 / top)), prm.max_rates, rtol=1e-9)
    np.testing.assert_allclose(prm.gain * prm.intercepts + prm.bias, 1.0, atol=1e-9)


def test_lif_step_hand_computed():
    """One neuron, constant current: voltage trajectory, spike time interpolation and refractory handling."""
    with nengo.Network(seed=1) as net:
        ens = nengo.Ensemble(1, 1, gain=[1.0], bias=[3.0], encoders=[[1.0]])
        p = nengo.Probe(ens)
    model = build_model(net)
    sim = RefSimulator(net, model=model)
    v0 = float(model.initial_voltage(ens)[0])
    dt, tau_rc, tau_ref, J = 0.001, 0.02, 0.002, 3.0
    v, ref = v0, 0.0
    outs = []
    for _ in range(40):
        ref -= dt
        delta = min(max(dt - ref, 0.0), dt)
        v = v - (J - v) * np.expm1(-delta / tau_rc)
        if v > 1:
            t_spike = dt + tau_rc * np.log1p(-(v - 1) / (J - 1))
            ref = tau_ref + t_spike
            v = 0.0
            outs.append(1.0 / dt)
        else:
            v = max(v, 0.0)
            outs.append(0.0)
    sim.run_steps(40)
    np.testing.assert_allclose(sim.signals[ens, "out"].a[0], outs[-1])
    np.testing.assert_allclose(sim.voltage(ens)[0], v, rtol=1e-12, atol=1e-15)
    assert sum(o > 0 for o in outs) >= 3     # several spikes in 40 ms at J = 3


def _learning_net(voja):
    with nengo.Network(seed=5) as net:
        key = nengo.Node(lambda t: [0.6, 0.2, -0.4])
        err = nengo.Node(lambda t: [0.3, -0.5])
        pre = nengo.Ensemble(30, 3, intercepts=[0.1] * 30)
        post = nengo.Ensemble(20, 2)
        kw = dict(learning_rule_type=nengo.Voja(learning_rate=1e-2, post_synapse=None)) if voja else {}
        cin = nengo.Connection(key, pre, synapse=None, **kw)
        cout = nengo.Connection(pre, post, function=lambda x: [0.0, 0.0], learning_rule_type=nengo.PES(1e-3))
        nengo.Connection(err, cout.learning_rule, synapse=None)
        wp = nengo.Probe(cout, "weights")
    return net, pre, cin, cout, wp


def test_pes_delta_is_applied_one_step_late():
    net, pre, cin, cout, wp = _learning_net(voja=False)
    model = build_model(net)
    sim = RefSimulator(net, model=model)
    n = pre.n_neurons
    a_f = np.zeros(n)
    W = np.zeros((2, n))
    decay = np.exp(-0.001 / 0.005)
    e = np.array([0.3, -0.5])
    for step in range(1, 30):
        sim.run_steps(1)
        acts = sim.signals[pre, "out"].a.copy()
        # weights probed at step k contain the deltas of steps 1..k-1 only
        np.testing.assert_allclose(sim.data[wp][-1], W, rtol=1e-12, atol=1e-18)
        delta = np.outer(-1e-3 * 0.001 / n * e, a_f)     # uses the trace as read this step (before its update)
        a_f = decay * a_f + (1 - decay) * acts
        W = W + delta
    assert np.max(np.abs(W)) > 0


def test_voja_moves_only_spiking_rows():
    net, pre, cin, cout, wp = _learning_net(voja=True)
    model = build_model(net)
    sim = RefSimulator(net, model=model)
    E = model.params[pre].scaled_encoders.copy()
    scale = model.params[pre].gain / pre.radius
    x = np.array([0.6, 0.2, -0.4])
    pending = np.zeros_like(E)
    for _ in range(25):
        sim.run_steps(1)
        E = E + pending                          # Copy(delta -> encoders, inc) runs at the start of the step
        np.testing.assert_allclose(sim.scaled_encoders(pre), E, rtol=1e-12, atol=1e-15)
        post = sim.signals[pre, "out"].a
        pending = 1e-2 * 0.001 * 1.0 * (scale[:, None] * np.outer(post, x) - post[:, None] * E)
        assert np.all(pending[post == 0] == 0)   # post_synapse=None: only rows that spiked move
    assert not np.allclose(E, model.params[pre].scaled_encoders)


def test_probe_sample_every_and_weights_shape():
    net, pre, cin, cout, wp0 = _learning_net(voja=False)
    with net:
        wp = nengo.Probe(cout, "weights", sample_every=0.01)
    sim = RefSimulator(net)
    sim.run_steps(35)
    assert sim.data[wp].shape == (3, 2, 30)
    assert sim.data[wp0].shape == (35, 2, 30)


def test_seeding_is_creation_order_dependent_and_reproducible():
    def make():
        with nengo.Network(seed=11) as net:
            a = nengo.Ensemble(10, 1)
            b = nengo.Ensemble(10, 1)
        return net, a, b
    n1, a1, b1 = make()
    n2, a2, b2 = make()
    m1, m2 = build_model(n1), build_model(n2)
    assert m1.seeds[a1] == m2.seeds[a2] and m1.seeds[a1] != m1.seeds[b1]
    assert np.array_equal(m1.params[a1].encoders, m2.params[a2].encoders)
    assert np.array_equal(m1.params[b1].gain, m2.params[b2].gain)


def test_spiking_lif_rate_converges_to_the_published_rate_formula():
    """A closed-form anchor that does not depend on our reading of nengo's step code: with the spike-time interpolation and
    the fractional refractory bookkeeping, a LIF neuron driven by a constant current J fires at exactly the LIFRate
    tuning-curve rate 1 / (tau_ref + tau_rc ln(1 + 1 / (J - 1))) (Eliasmith & Anderson 2003; nengo's documented
    ``LIFRate``), up to the discretisation of the count: |spikes / T - rate| <= 1 / T.  A step without the interpolation
    (spike only on step boundaries) misses this by several per cent at these rates."""
    gains = np.array([1.0, 1.0, 1.0, 1.0])
    biases = np.array([1.3, 2.0, 3.5, 8.0])                       # J = bias (input 0): 32 ... 214 Hz
    with nengo.Network(seed=2) as net:
        ens = nengo.Ensemble(4, 1, gain=gains, bias=biases, encoders=np.ones((4, 1)))
        nengo.Probe(ens)
    model = build_model(net)
    sim = RefSimulator(net, model=model)
    n_steps, dt, tau_rc, tau_ref = 4000, 0.001, 0.02, 0.002
    counts = np.zeros(4)
    for _ in range(n_steps):
        sim.run_steps(1)
        counts += sim.signals[ens, "out"].a > 0
    T = n_steps * dt
    rate = 1.0 / (tau_ref + tau_rc * np.log1p(1.0 / (biases - 1.0)))
    assert np.all(np.abs(counts / T - rate) <= 1.0 / T + 1e-9), (counts / T, rate)
    assert 30 < rate[0] < 35 and 200 < rate[3] < 230
