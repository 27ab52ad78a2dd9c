"""Neuron-model *parameters* and their static tuning maths (gain/bias, rate curves).

Restates nengo's published formulas (upstream
``nengo/neurons.py``) — only what the builder needs to solve decoders.  The
per-timestep update lives in the CUDA kernels (product) and, independently, in
``oracle/`` (checker).
"""
import numpy as np


class NeuronType:
    spiking = False
    state_names = ()

    def gain_bias(self, max_rates, intercepts):
        raise NotImplementedError

    def rates(self, x, gain, bias):
        """Steady-state rates for represented value(s) ``x`` (projected on the encoders)."""
        J = gain * x + bias
        return self.current_to_rate(J)

    def current_to_rate(self, J):
        raise NotImplementedError

    def __repr__(self):
        return f"{type(self).__name__}()"


class LIFRate(NeuronType):
    def __init__(self, tau_rc=0.02, tau_ref=0.002, amplitude=1):
        self.tau_rc = float(tau_rc)
        self.tau_ref = float(tau_ref)
        self.amplitude = float(amplitude)

    def gain_bias(self, max_rates, intercepts):
        max_rates = np.asarray(max_rates, dtype=np.float64)
        intercepts = np.asarray(intercepts, dtype=np.float64)
        if np.any(max_rates > 1.0 / self.tau_ref):
            raise ValueError("max_rates exceed 1/tau_ref")
        x = 1.0 / (1 - np.exp((self.tau_ref - (1.0 / max_rates)) / self.tau_rc))
        gain = (1 - x) / (intercepts - 1.0)
        bias = 1 - gain * intercepts
        return gain, bias

    def current_to_rate(self, J):
        J = np.asarray(J, dtype=np.float64)
        j = J - 1
        out = np.zeros_like(j)
        pos = j > 0
        out[pos] = self.amplitude / (self.tau_ref + self.tau_rc * np.log1p(1.0 / j[pos]))
        return out


class LIF(LIFRate):
    spiking = True
    state_names = ("voltage", "refractory_time")

    def __init__(self, tau_rc=0.02, tau_ref=0.002, min_voltage=0, amplitude=1):
        super().__init__(tau_rc=tau_rc, tau_ref=tau_ref, amplitude=amplitude)
        self.min_voltage = float(min_voltage)


class RectifiedLinear(NeuronType):
    def __init__(self, amplitude=1):
        self.amplitude = float(amplitude)

    def gain_bias(self, max_rates, intercepts):
        max_rates = np.asarray(max_rates, dtype=np.float64)
        intercepts = np.asarray(intercepts, dtype=np.float64)
        gain = max_rates / (1 - intercepts)
        bias = -intercepts * gain
        return gain, bias

    def current_to_rate(self, J):
        return self.amplitude * np.maximum(np.asarray(J, dtype=np.float64), 0.0)


class Direct(NeuronType):
    """Declared for API completeness; not supported by the B200 backend."""
