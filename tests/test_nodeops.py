"""``nodeops.recognize``: only the reference's own closures (same program, constants from the closure cells) or explicit
device-op instances are lowered; look-alikes that agree on a few probe inputs are rejected (there is no host callback)."""
import numpy as np
import pytest

from sspslam_b200 import lowering, nengo_shim as nengo, nodeops
from sspslam_b200.builder import build_model


def _node(fn, size_in):
    with nengo.Network(seed=1):
        return nengo.Node(fn, size_in=size_in)


def test_identity_lambda_and_explicit_ops_are_recognised():
    assert nodeops.recognize(_node(lambda t, x: x, 7), []).kind == "identity"
    assert nodeops.recognize(_node(nodeops.Identity(), 7), []).kind == "identity"
    S = np.random.default_rng(0).standard_normal((40, 5))
    assert nodeops.recognize(_node(nodeops.GridCleanup(S), 5), []).kind == "cleanup"


def test_reference_style_closures_are_recognised_with_their_constants():
    d, shift_rate, update_thres = 4, 0.3, 0.15
    sample_ssps = np.random.default_rng(1).standard_normal((50, 4))

    def clean_up_fun(x):
        sims = sample_ssps @ x
        return sample_ssps[np.argmax(sims), :]

    def update_state_func(t, x):
        if (np.allclose(x[-1], 0, atol=1e-3) & (np.sum(x[:d] * x[d:-1]) > update_thres)):
            return shift_rate * (x[:d] - x[d:-1])
        else:
            return np.zeros(d)

    op = nodeops.recognize(_node(lambda t, x: clean_up_fun(x), 4), [])
    assert op.kind == "cleanup" and np.array_equal(op.sample_ssps, sample_ssps)
    gate = nodeops.recognize(_node(update_state_func, 2 * d + 1), [])
    assert (gate.kind, gate.d, gate.shift_rate, gate.update_thres) == ("gate", 4, 0.3, 0.15)


@pytest.mark.parametrize("fn", [
    lambda t, x: np.clip(x, -5, 5),                       # identity on N(0,1) probes, saturates later
    lambda t, x: x if t < 0.5 else 0 * x,                 # a time switch
    lambda t, x: np.tanh(1e-3 * x) * 1e3,                 # numerically the identity for small x
    lambda t, x: x * 1.0,                                 # not the same program: rejected rather than guessed
])
def test_identity_lookalikes_are_rejected(fn):
    node = _node(fn, 6)
    assert nodeops.recognize(node, []) is None
    with nengo.Network(seed=1) as net:
        a = nengo.Node(lambda t: np.ones(6))
        b = nengo.Node(fn, size_in=6)
        nengo.Connection(a, b, synapse=None)
        nengo.Probe(b)
    with pytest.raises(NotImplementedError):
        lowering.lower(net, build_model(net))


def test_gate_and_cleanup_lookalikes_are_rejected():
    d, shift_rate, update_thres = 3, 0.2, 0.2
    sample_ssps = np.random.default_rng(2).standard_normal((30, 3))

    def update_state_func(t, x):                          # another threshold rule with the same closure cells
        if (np.allclose(x[-1], 0, atol=1e-3) & (np.sum(x[:d] * x[d:-1]) >= update_thres)):
            return shift_rate * (x[:d] - x[d:-1])
        else:
            return np.zeros(d)

    def clean_up_fun(x):                                  # argmin instead of argmax
        sims = sample_ssps @ x
        return sample_ssps[np.argmin(sims), :]

    assert nodeops.recognize(_node(update_state_func, 2 * d + 1), []) is None
    assert nodeops.recognize(_node(lambda t, x: clean_up_fun(x), 3), []) is None
