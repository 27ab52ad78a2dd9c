"""Distributions used while *building* (``nengo.dists`` surface of SURVEY.md App. A.8).

Upstream nengo is not available to diff against, so these follow
nengo's documented algorithms: Gaussian-normalised hypersphere samples, and the
R_d quasi-random sequence mapped through inverse-CDF spherical coordinates with a
random rotation for ``ScatteredHypersphere``.
"""
import numpy as np
import scipy.special


class Distribution:
    def sample(self, n, d=None, rng=np.random):
        raise NotImplementedError

    @staticmethod
    def _shape(n, d):
        return (n,) if d is None else (n, d)


def get_samples(dist_or_samples, n, d=None, rng=np.random):
    if isinstance(dist_or_samples, Distribution):
        return dist_or_samples.sample(n, d=d, rng=rng)
    return np.array(dist_or_samples, dtype=np.float64)


class Uniform(Distribution):
    def __init__(self, low, high, integer=False):
        self.low, self.high, self.integer = low, high, integer

    def sample(self, n, d=None, rng=np.random):
        shape = self._shape(n, d)
        if self.integer:
            return rng.randint(low=self.low, high=self.high, size=shape)
        return rng.uniform(low=self.low, high=self.high, size=shape)


class Gaussian(Distribution):
    def __init__(self, mean, std):
        self.mean, self.std = mean, std

    def sample(self, n, d=None, rng=np.random):
        return rng.normal(loc=self.mean, scale=self.std, size=self._shape(n, d))


class Choice(Distribution):
    def __init__(self, options, weights=None):
        self.options = np.array(options, dtype=np.float64)
        w = np.ones(len(self.options)) if weights is None else np.asarray(weights, dtype=np.float64)
        self.p = w / w.sum()

    def sample(self, n, d=None, rng=np.random):
        if d is not None and (self.options.ndim < 2 or self.options.shape[1] != d):
            raise ValueError("Choice options do not match requested dimensionality")
        i = np.searchsorted(np.cumsum(self.p), rng.rand(n))
        return self.options[np.minimum(i, len(self.options) - 1)]


class UniformHypersphere(Distribution):
    def __init__(self, surface=False, min_magnitude=0):
        self.surface = bool(surface)
        self.min_magnitude = float(min_magnitude)

    def sample(self, n, d=None, rng=np.random):
        if d is None or d < 1:
            raise ValueError("dimensions must be a positive integer")
        pts = rng.randn(n, d)
        pts /= np.linalg.norm(pts, axis=1, keepdims=True)
        if self.surface:
            return pts
        lo = self.min_magnitude ** d
        return pts * rng.uniform(low=lo, high=1, size=(n, 1)) ** (1.0 / d)


class QuasirandomSequence(Distribution):
    """Additive-recurrence R_d sequence (generalised golden ratio)."""

    @staticmethod
    def _phi(d):
        x = 1.0
        for _ in range(30):
            x -= (x ** (d + 1) - x - 1) / ((d + 1) * x ** d - 1)
        return x

    def sample(self, n, d=1, rng=np.random):
        if d == 1:
            return np.linspace(1.0 / n, 1, n)[:, None]
        inv = 1.0 / self._phi(d)
        alpha = inv ** np.arange(1, d + 1)
        z = (0.5 + alpha[None, :] * np.arange(1, n + 1)[:, None]) % 1.0
        return z


class ScatteredHypersphere(UniformHypersphere):
    def __init__(self, surface=False, min_magnitude=0, base=None, method="sct-approx"):
        super().__init__(surface=surface, min_magnitude=min_magnitude)
        self.base = QuasirandomSequence() if base is None else base
        self.method = method

    @staticmethod
    def _coord_ppf(dims, y):
        """Inverse CDF of one spherical angle (in half-turns) for a ``dims``-sphere slice."""
        refl = np.where(y < 0.5, y, 1 - y)
        z_sq = scipy.special.betaincinv(dims / 2.0, 0.5, 2 * refl)
        x = np.arcsin(np.sqrt(z_sq)) / np.pi
        return np.where(y < 0.5, x, 1 - x)

    @classmethod
    def _to_sphere(cls, cube):
        n, m = cube.shape  # m angles -> points on S^m in R^(m+1)
        ang = np.empty_like(cube)
        for j in range(m):
            ang[:, j] = cls._coord_ppf(m - j, cube[:, j])
        mult = np.ones(m)
        mult[-1] = 2.0
        s = np.sin(mult[None, :] * np.pi * ang)
        c = np.cos(mult[None, :] * np.pi * ang)
        out = np.ones((n, m + 1))
        out[:, 1:] = np.cumprod(s, axis=1)
        out[:, :-1] *= c
        return out

    @staticmethod
    def _rotation(d, rng):
        u, _, vt = np.linalg.svd(rng.standard_normal((d, d)))
        return u @ vt

    def sample(self, n, d=1, rng=np.random):
        if d == 1:
            return super().sample(n, d, rng)
        if self.surface:
            cube = self.base.sample(n, d - 1, rng)
            radius = 1.0
        else:
            cube = self.base.sample(n, d, rng)
            cube, radius = cube[:, :-1], cube[:, -1:].copy()
            if self.min_magnitude != 0:
                lo = self.min_magnitude ** d
                radius = radius * (1 - lo) + lo
            radius = radius ** (1.0 / d)
        pts = self._to_sphere(cube) * radius
        return pts @ self._rotation(d, rng)


class CosineSimilarity(Distribution):
    """Cosine similarity between a fixed and a uniformly random unit vector in R^d."""

    def __init__(self, dimensions):
        self.dimensions = int(dimensions)

    def sample(self, n, d=None, rng=np.random):
        shape = self._shape(n, d)
        mag = np.sqrt(rng.beta(0.5, (self.dimensions - 1) / 2.0, size=shape))
        return mag * np.where(rng.rand(*shape) < 0.5, -1.0, 1.0)
