"""``nengo.processes.WhiteSignal`` (band-limited white noise).

The drivers synthesise random paths with it (``run_slam.py:98-99``,
``run_pathint.py:75``); it is host-side input synthesis, not part of the stepped path.
"""
import numpy as np


class Process:
    pass


class WhiteSignal(Process):
    def __init__(self, period, high, rms=0.5, y0=None, seed=None):
        self.period, self.high, self.rms, self.y0, self.seed = period, high, rms, y0, seed

    def _signal(self, dt, size_out, rng):
        n_coef = int(np.ceil(self.period / dt / 2.0))
        sigma = self.rms * np.sqrt(0.5)
        coef = 1j * rng.normal(0.0, sigma, size=(n_coef + 1, size_out))
        coef += rng.normal(0.0, sigma, size=(n_coef + 1, size_out))
        coef[0] = 0.0
        coef[-1].imag = 0.0
        freqs = np.fft.rfftfreq(2 * n_coef, d=dt)
        cut = freqs > self.high
        coef[cut] = 0.0
        coef /= np.sqrt(1 - np.sum(cut, dtype=float) / n_coef)
        coef *= np.sqrt(2 * n_coef)
        sig = np.fft.irfft(coef, axis=0)
        if self.y0 is not None:
            k = np.argmin(np.abs(self.y0 - sig), axis=0)
            sig = np.stack([np.roll(sig[:, i], -k[i]) for i in range(size_out)], axis=1)
        return sig

    def run_steps(self, n_steps, d=1, dt=0.001, rng=None):
        rng = np.random.RandomState(self.seed) if rng is None else rng
        sig = self._signal(dt, d, rng)
        idx = np.arange(1, n_steps + 1)
        t = idx * dt
        return sig[np.round(t / dt).astype(int) % sig.shape[0]]

    def run(self, t, d=1, dt=0.001, rng=None):
        return self.run_steps(int(np.round(float(t) / dt)), d=d, dt=dt, rng=rng)
