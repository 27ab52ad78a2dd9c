"""SP / SSP representation spaces with the reference's ``sspspace.py`` API, B200-backed.

Mirrors the public surface of ``/root/reference/sspslam/sspspace.py`` that the hot path
uses: ``SPSpace`` (:11-182), ``SSPSpace`` (:183-636; ``encode``
:252-273, ``decode(...,'from-set','grid',num)`` :312-358, grid sampling :424-506,
``make_unitary`` :511-514, ``bind`` :525-528, ``invert`` :530-532), ``RandomSSPSpace``
(:638-668) and ``HexagonalSSPSpace`` (:678-731, ``conjsym`` :860-868).

``encode`` / ``decode`` / ``clean_up`` run the sm_100a kernels through the C-ABI
(``ssb_ssp_encode`` / ``ssb_ssp_decode_argmax``).  There is **no silent CPU path**: with
``backend='cuda'`` (default) a missing library or GPU raises.  ``backend='host'`` is an
explicit opt-in used for build-time constants (OVC encoders, the clean-up grid) and for
CPU-only unit tests; it is NumPy float64.
"""
from __future__ import annotations

import numpy as np
from scipy.special import gammainc
from scipy.stats import special_ortho_group

from .nengo_shim.dists import UniformHypersphere

__all__ = ["SPSpace", "SSPSpace", "RandomSSPSpace", "HexagonalSSPSpace", "conj_symmetric_phases"]


def _unitary_rows(v, floor=0.0):
    """Scale every Fourier coefficient of each row to unit magnitude."""
    f = np.fft.fft(v, axis=-1)
    mag = np.abs(f)
    if floor:
        mag = np.maximum(mag, floor)
    return np.fft.ifft(f / mag, axis=-1).real


def _cconv(a, b):
    return np.fft.ifft(np.fft.fft(a, axis=-1) * np.fft.fft(b, axis=-1), axis=-1).real


def _involution(a, dim):
    return a[..., (-np.arange(dim)) % dim]


# --------------------------------------------------------------------------------------
class SPSpace:
    """Discrete symbol space: ``domain_size`` mutually orthogonal semantic pointers.

    Follows ``sspspace.py:43-63``: unitary random vectors, then sequential Gram-Schmidt
    *without* re-normalisation (rows end orthogonal but with decaying norms; SURVEY.md
    App. C — reproduced on purpose)."""

    def __init__(self, domain_size, dim, seed=None, vectors=None, **_ignored):
        self.domain_size = int(domain_size)
        self.dim = int(dim)
        if seed is None:
            rng = np.random.RandomState()
        elif isinstance(seed, (int, np.integer)):
            rng = np.random.RandomState(int(seed))
        else:
            raise TypeError("seed must be an int or None")
        self.rng = rng
        if self.domain_size == 1:
            self.vectors = np.zeros((1, self.dim))
            self.vectors[0, 0] = 1.0
        elif vectors is not None:
            self.vectors = np.asarray(vectors)
        else:
            vecs = _unitary_rows(UniformHypersphere(surface=True).sample(self.domain_size, self.dim, rng=rng))
            for j in range(self.domain_size - 1):
                q = vecs[j] / np.linalg.norm(vecs[j])
                vecs[j + 1:] -= np.outer(vecs[j + 1:] @ q, q)
            self.vectors = vecs
        self.inverse_vectors = self.invert(self.vectors)

    def encode(self, i):
        return self.vectors[np.asarray(i).reshape(-1).astype(int)]

    def decode(self, v, **_kw):
        return np.argmax(self.vectors @ np.asarray(v).T, axis=0)

    def clean_up(self, v, **_kw):
        return self.vectors[self.decode(v)]

    def normalize(self, v):
        return v / np.sqrt(np.sum(v ** 2))

    def make_unitary(self, v):
        return _unitary_rows(np.atleast_2d(v))

    def identity(self):
        e = np.zeros(self.dim)
        e[0] = 1.0
        return e

    def bind(self, a, b):
        return _cconv(np.atleast_2d(a), np.atleast_2d(b))

    def invert(self, a):
        return _involution(np.atleast_2d(a), self.dim)


# --------------------------------------------------------------------------------------
def conj_symmetric_phases(K):
    """``[0; K; -flip(K)]`` — phase rows whose IFFT is real (``sspspace.py:860-868``)."""
    K = np.asarray(K, dtype=np.float64)
    return np.vstack([np.zeros((1, K.shape[1])), K, -K[::-1]])


class SSPSpace:
    """Continuous space: ``phi(x) = IFFT(exp(i A x / l)).real`` with phase matrix ``A``."""

    def __init__(self, domain_dim, ssp_dim, phase_matrix, domain_bounds=None, length_scale=1,
                 rng=None, backend="cuda"):
        self.domain_dim = int(domain_dim)
        self.ssp_dim = int(ssp_dim)
        self.length_scale = np.asarray(length_scale, dtype=np.float64) * np.ones((self.domain_dim, 1))
        self.rng = np.random.default_rng() if rng is None else rng
        if domain_bounds is not None:
            domain_bounds = np.asarray(domain_bounds, dtype=np.float64)
            if domain_bounds.shape != (self.domain_dim, 2):
                raise ValueError(f"domain_bounds must have shape ({self.domain_dim}, 2)")
        self.domain_bounds = domain_bounds
        phase_matrix = np.asarray(phase_matrix, dtype=np.float64)
        if phase_matrix.shape != (self.ssp_dim, self.domain_dim):
            raise ValueError("phase_matrix must be (ssp_dim, domain_dim)")
        self.phase_matrix = phase_matrix
        if backend not in ("cuda", "host"):
            raise ValueError("backend must be 'cuda' or 'host'")
        self.backend = backend
        self.decoder_model = None
        self._grid_cache = {}

    # ---- helpers
    def update_lengthscale(self, scale):
        scale = np.asarray(scale, dtype=np.float64)
        if scale.size == 1:
            self.length_scale = float(scale) * np.ones((self.domain_dim, 1))
        elif scale.size == self.domain_dim:
            self.length_scale = scale.reshape(self.domain_dim, 1)
        else:
            raise ValueError("length scale size mismatch")
        self._grid_cache.clear()

    def _scaled_phases(self):
        """``A / l`` (ssp_dim x domain_dim): the only constant the encode kernel needs."""
        return self.phase_matrix / self.length_scale.reshape(1, -1)

    # ---- encode
    def encode_host(self, x):
        """float64 NumPy encode (build-time constants; same formula as :252-273)."""
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        theta = x @ self._scaled_phases().T
        return np.fft.ifft(np.exp(1j * theta), axis=1).real

    def encode(self, x):
        if self.backend == "host":
            return self.encode_host(x)
        from . import cabi
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        return cabi.ssp_encode(self._scaled_phases(), x).astype(np.float64)

    def encode_fourier(self, x):
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        return np.exp(1j * (x @ self._scaled_phases().T))

    # ---- sampling grid
    def get_sample_points(self, samples_per_dim=100, method="length-scale"):
        bounds = self.domain_bounds
        if bounds is None:
            bounds = np.tile([-10.0, 10.0], (self.domain_dim, 1))
        if method == "grid":
            counts = [int(samples_per_dim)] * self.domain_dim
        elif method == "length-scale":
            counts = [2 * int(np.ceil((hi - lo) / self.length_scale[i, 0])) for i, (lo, hi) in enumerate(bounds)]
        elif method == "sobol":   # sspspace.py:469-474: scrambled Sobol points from the space's own generator
            from scipy.stats import qmc
            sampler = qmc.Sobol(d=self.domain_dim, seed=self.rng)
            u = sampler.random(int(np.prod(samples_per_dim)))
            return qmc.scale(u, np.asarray(bounds)[:, 0], np.asarray(bounds)[:, 1])
        else:
            raise NotImplementedError(f"sampling method {method!r} is outside the hot path")
        axes = [np.linspace(bounds[i, 0], bounds[i, 1], counts[i]) for i in range(self.domain_dim)]
        mesh = np.meshgrid(*axes)  # 'xy' indexing, like the reference (:460-464)
        return np.stack([m.reshape(-1) for m in mesh], axis=1)

    def sample_grid_encoders(self, n_neurons, method="sobol"):
        """Grid-cell encoders (behaviour of ``sspspace.py:733-762``): neuron ``i`` is the SSP of sample point ``p_i`` with
        every frequency switched off except ONE simplex sub-lattice ``g_i`` of the phase matrix (its ``n + 1`` rows) and
        the DC (and Nyquist) term; rows are normalised.  Sub-lattices are dealt out in equal blocks, the remainder at
        random from the space's generator (drawn after the sample points, as the reference does)."""
        d, n = self.ssp_dim, self.domain_dim
        half = (d - 1) // 2
        n_lattices = ((d - 2) // 2 if d % 2 == 0 else half) // (n + 1)
        n_pts = int(np.ceil(n_neurons ** (1.0 / n))) if method == "grid" else n_neurons
        points = self.get_sample_points(n_pts, method=method)[:n_neurons]
        block = n_neurons // n_lattices
        lattice = np.concatenate([np.repeat(np.arange(n_lattices), block),
                                  self.rng.integers(0, n_lattices, size=n_neurons - n_lattices * block)])
        # spectrum[i, k]: exp(i A_k . p_i) on the rows of lattice g_i, conjugate-mirrored, 1 at DC / Nyquist
        rows = 1 + lattice[:, None] * (n + 1) + np.arange(n + 1)[None, :]            # [n_neurons, n + 1]
        phase = np.einsum("ijk,ik->ij", self.phase_matrix[rows], points)
        spectrum = np.zeros((n_neurons, d), dtype=complex)
        np.put_along_axis(spectrum, rows, np.exp(1j * phase), axis=1)
        spectrum[:, half + 1:] = np.conj(spectrum[:, half:0:-1]) if d % 2 else np.conj(spectrum[:, 1:half + 1][:, ::-1])
        spectrum[:, 0] = 1.0
        if d % 2 == 0:
            spectrum[:, d // 2] = 1.0
        enc = np.fft.ifft(spectrum, axis=1).real
        return enc / np.linalg.norm(enc, axis=1, keepdims=True)

    def get_sample_ssps(self, num_points, **kwargs):
        return self.encode_host(self.get_sample_points(num_points, **kwargs))

    def get_sample_pts_and_ssps(self, num_points_per_dim=100, method="grid"):
        key = (int(num_points_per_dim), method)
        if key not in self._grid_cache:
            pts = self.get_sample_points(samples_per_dim=num_points_per_dim, method=method)
            self._grid_cache[key] = (self.encode_host(pts), pts)
        ssps, pts = self._grid_cache[key]
        return ssps, pts

    # ---- decode
    def decode(self, ssp, method="from-set", sampling_method="grid", num_samples=300, samples=None, **_kw):
        if method != "from-set":
            raise NotImplementedError("only method='from-set' is on the hot path (SURVEY.md §2)")
        ssp = np.atleast_2d(np.asarray(ssp, dtype=np.float64))
        if samples is None:
            sample_ssps, sample_points = self.get_sample_pts_and_ssps(num_samples, sampling_method)
        else:
            sample_ssps, sample_points = samples
        if sample_ssps.shape[1] != ssp.shape[1]:
            raise ValueError("sample / query dimensionality mismatch")
        idx = self.decode_indices(ssp, sample_ssps)
        return sample_points[idx]

    def decode_indices(self, ssp, sample_ssps):
        """argmax_g  S[g] . unit(ssp)  (rows with norm < 1e-6 are left unscaled, :350-354)."""
        if self.backend == "host":
            norms = np.linalg.norm(ssp, axis=1)
            unit = np.where((norms < 1e-6)[:, None], ssp, ssp / np.maximum(norms, 1e-300)[:, None])
            return np.argmax(sample_ssps @ unit.T, axis=0)
        from . import cabi
        return cabi.ssp_decode_argmax(sample_ssps, ssp)

    def clean_up(self, ssp, method="from-set", sampling_method="grid", num_samples=300):
        return self.encode(self.decode(ssp, method, sampling_method, num_samples))

    # ---- algebra
    def normalize(self, ssp):
        return ssp / np.maximum(np.sqrt(np.sum(ssp ** 2)), 1e-8)

    def make_unitary(self, ssp):
        return _unitary_rows(np.asarray(ssp, dtype=np.float64), floor=1e-8)

    def identity(self):
        e = np.zeros(self.ssp_dim)
        e[0] = 1.0
        return e

    def bind(self, a, b):
        return _cconv(np.atleast_2d(a), np.atleast_2d(b))

    def invert(self, a):
        return _involution(np.atleast_2d(a), self.ssp_dim)


class RandomSSPSpace(SSPSpace):
    """Random phase rows (``sspspace.py:638-668``)."""

    def __init__(self, domain_dim, ssp_dim, domain_bounds=None, scale_min=0.25, scale_max=2.0,
                 length_scale=1, rng=None, sampler="unif", norm_scale=None, backend="cuda", **_ignored):
        rng = np.random.default_rng() if rng is None else rng
        n_rows = (ssp_dim - 1) // 2
        if sampler == "unif":
            g = rng.normal(size=(n_rows, domain_dim))
            ssq = np.sum(g ** 2, axis=1)
            radial = scale_max * gammainc(domain_dim / 2, ssq / 2) ** (1 / domain_dim) / np.sqrt(ssq)
            phases = g * radial[:, None]
        elif sampler == "norm":
            if norm_scale is None:
                norm_scale = np.sqrt(np.pi / 2) * ((scale_max - scale_min) / 2 + scale_min)
            phases = rng.normal(loc=0.0, scale=norm_scale, size=(n_rows, domain_dim))
        else:
            raise ValueError(f"unknown sampler {sampler!r}")
        A = conj_symmetric_phases(phases)
        super().__init__(domain_dim, A.shape[0], A, domain_bounds=domain_bounds,
                         length_scale=length_scale, rng=rng, backend=backend)


class HexagonalSSPSpace(SSPSpace):
    """Phase rows on scaled + rotated copies of a regular simplex (``sspspace.py:678-731``)."""

    def __init__(self, domain_dim, ssp_dim=151, n_rotates=5, n_scales=5, scale_min=1, scale_max=np.pi,
                 scale_sampling="lin", domain_bounds=None, length_scale=1, rng=None, backend="cuda",
                 **_ignored):
        rng = np.random.default_rng() if rng is None else rng
        n = int(domain_dim)
        if n_rotates == 5 and n_scales == 5 and ssp_dim != 151:
            # caller specified the total dimension, not the simplex counts (:683-686)
            n_rotates = n_scales = int(np.sqrt((ssp_dim - 1) / (2 * (n + 1))))
        simplex = np.vstack([np.sqrt(1 + 1 / n) * np.eye(n) - n ** -1.5 * (np.sqrt(n + 1) + 1),
                             n ** -0.5 * np.ones((1, n))])  # (n+1) x n
        self.grid_basis_dim = n + 1
        self.num_grids = n_rotates * n_scales
        self.scale_min, self.scale_max = scale_min, scale_max
        self.n_scales, self.n_rotates = n_scales, n_rotates

        ns_eff = n_scales * n_rotates * n_rotates if n == 1 else n_scales  # (:699-703 doubles up in 1-D)
        golden = (1 + np.sqrt(5)) / 2
        if scale_sampling == "lin":
            lo = scale_max / (ns_eff * (golden - 1) + 1) if scale_min is None else scale_min
            scales = np.linspace(lo, scale_max, ns_eff)
        elif scale_sampling == "log":
            lo = scale_max / golden ** (ns_eff - 1) if scale_min is None else scale_min
            scales = np.geomspace(lo, scale_max, ns_eff)
        elif scale_sampling == "rand":
            scales = rng.uniform(0 if scale_min is None else scale_min, scale_max, ns_eff)
        else:
            raise ValueError(f"unknown scale_sampling {scale_sampling!r}")
        stacked = np.vstack([simplex * s for s in scales])

        if n_rotates == 1 or n == 1:
            rows = stacked
        else:
            if n == 2:
                ang = np.linspace(0, 2 * np.pi / 3, n_rotates, endpoint=False)
                rots = np.stack([np.stack([np.cos(ang), -np.sin(ang)], axis=1),
                                 np.stack([np.sin(ang), np.cos(ang)], axis=1)], axis=1)
            else:
                rots = special_ortho_group.rvs(n, size=n_rotates, random_state=rng)
                rots = rots.reshape(n_rotates, n, n)
            rows = (rots @ stacked.T).transpose(0, 2, 1).reshape(-1, n)
        A = conj_symmetric_phases(rows)
        super().__init__(n, A.shape[0], A, domain_bounds=domain_bounds, length_scale=length_scale,
                         rng=rng, backend=backend)
