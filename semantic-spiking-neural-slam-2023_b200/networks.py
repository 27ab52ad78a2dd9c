"""From-scratch declarations of the reference's hot-path network topologies.

The GPU box has no ``/root/reference`` checkout, so benchmarks, ``smoke()`` and the
``-m gpu`` tests declare the networks through this module.  Each class reproduces the
*graph* (objects, sizes, transforms, synapses, functions and — because nengo's seeding is
creation-order dependent, SURVEY.md App. A.2 — the creation order) of:

* ``PathIntegration``      <- ``sspslam/networks/pathintegration.py:108-191``
* ``Product`` / ``CircularConvolution`` <- ``sspslam/networks/binding.py:189-218,288-324``
* ``AssociativeMemory``    <- ``sspslam/networks/associativememory.py:12-54``
* ``SLAMNetwork``          <- ``sspslam/networks/slam.py:182-307``
* ``SLAMViewNetwork``      <- ``sspslam/networks/slam_view.py:181-276``
* ``SLAMLoihiNetwork``     <- ``sspslam/networks/slam_loihi.py:190-293`` (all-neural gating)

``tests/test_networks_parity.py`` checks (in the authoring container, where the
reference is mounted) that a model built from these declarations is *identical* —
same seeds, encoders, decoders, transforms — to one built from the unmodified reference
sources.  Node callables are explicit device-op objects (:mod:`nodeops`).
"""
from __future__ import annotations

import numpy as np

from . import nengo_shim as nengo
from .nengo_shim import Network, Node, Ensemble, Connection
from .nengo_shim.networks import EnsembleArray
from .nengo_shim.params import Default
from .nodeops import Identity, GridCleanup, GatedCorrection
from .inputs import (get_slam_input_functions, get_slam_input_functions2,      # re-exported like sspslam.networks does
                     get_slamview_input_functions)

__all__ = ["PathIntegration", "Product", "CircularConvolution", "AssociativeMemory", "SLAMNetwork",
           "SLAMViewNetwork", "SLAMLoihiNetwork", "get_slam_input_functions", "get_slam_input_functions2", "get_slamview_input_functions",
           "get_to_Fourier", "get_from_Fourier", "transform_in", "transform_out",
           "dft_half", "circconv", "oscillator_feedback"]


# ------------------------------------------------------------------------- Fourier layouts
def get_to_Fourier(d):
    """SSP (d) -> oscillator layout (3 per VCO: [Re F_k, Im F_k, 0]); k = 0 row stays zero
    (``pathintegration.py:816-822``)."""
    n_osc = (d + 1) // 2
    W = np.fft.fft(np.eye(d))
    M = np.zeros((3 * n_osc, d))
    for k in range(1, n_osc):
        M[3 * k] = W[k].real
        M[3 * k + 1] = W[k].imag
    return M


def get_from_Fourier(d):
    """Oscillator layout -> SSP, restoring the conjugate half (``pathintegration.py:824-844``).

    x[m] = (1/d) * [ Re F_0 + 2 * sum_k ( Re F_k cos(2 pi k m/d) - Im F_k sin(2 pi k m/d) ) ].
    Only odd ``d`` is meaningful (every hexagonal SSP dimension is odd)."""
    if d % 2 == 0:
        raise ValueError("oscillator layout is only defined for odd ssp_dim")
    n_osc = (d + 1) // 2
    m = np.arange(d)
    M = np.zeros((d, 3 * n_osc))
    M[:, 0] = 1.0 / d
    for k in range(1, n_osc):
        ang = 2 * np.pi * k * m / d
        M[:, 3 * k] = 2 * np.cos(ang) / d
        M[:, 3 * k + 1] = -2 * np.sin(ang) / d
    return M


def dft_half(n):
    w = np.arange(n // 2 + 1)[:, None]
    x = np.arange(n)[None, :]
    return np.exp(-2j * np.pi * w * x / n)


def circconv(a, b, invert_a=False, invert_b=False, axis=-1):
    A, B = np.fft.fft(a, axis=axis), np.fft.fft(b, axis=axis)
    if invert_a:
        A = A.conj()
    if invert_b:
        B = B.conj()
    return np.fft.ifft(A * B, axis=axis).real


def transform_in(dims, align, invert):
    """Input transform of the neural circular convolution: four product channels per
    half-spectrum coefficient (``binding.py:23-54``; all 4*(d//2+1) rows are kept)."""
    if align not in ("A", "B"):
        raise nengo.exceptions.ValidationError("'align' must be either 'A' or 'B'", "align")
    rows = dft_half(dims)
    if invert:
        rows = rows.conj()
    pick_real = {"A": (True, False, True, False), "B": (True, False, False, True)}[align]
    out = np.zeros((4 * rows.shape[0], dims))
    for c, real in enumerate(pick_real):
        out[c::4] = rows.real if real else rows.imag
    return out


def transform_out(dims):
    """Recombine the four product channels and apply the inverse half-DFT (``binding.py:57-74``)."""
    rows = dft_half(dims).conj()
    n_half = rows.shape[0]
    weight = np.full(n_half, 2.0)
    weight[0] = 1.0
    if dims % 2 == 0:
        weight[-1] = 1.0
    rows = rows * weight[:, None]
    tr = np.zeros((n_half, 4, dims))
    tr[:, 0] = rows.real
    tr[:, 1] = -rows.real
    tr[:, 2] = -rows.imag
    tr[:, 3] = -rows.imag
    return (tr.reshape(4 * n_half, dims) / dims).T


def oscillator_feedback(recurrent_tau, scaling_factor, length_scale0, max_radius=1.0, stable=True):
    """Target function of the recurrent VCO connection (``pathintegration.py:119-135``)."""
    denom = scaling_factor * length_scale0

    def stable_fb(x):
        w = x[2] / denom
        r = max(np.sqrt(x[0] ** 2 + x[1] ** 2), 1e-9)
        pull = (max_radius ** 2 - r ** 2) / r
        return np.array([x[0] + recurrent_tau * (x[0] * pull - x[1] * w),
                         x[1] + recurrent_tau * (x[1] * pull + x[0] * w), 0.0])

    def linear_fb(x):
        w = x[2] / denom
        return np.array([x[0] - recurrent_tau * x[1] * w, x[1] + recurrent_tau * x[0] * w, 0.0])

    return stable_fb if stable else linear_fb


# ------------------------------------------------------------------------- path integrator
class PathIntegration(Network):
    def __init__(self, ssp_space, n_neurons, recurrent_tau=0.05, scaling_factor=1, stable=True,
                 max_radius=1, with_gcs=False, n_gcs=1000, solver_weights=False, label="pathint",
                 **ens_kwargs):
        super().__init__(label=label)
        d, n_dom = ssp_space.ssp_dim, ssp_space.domain_dim
        n_osc = (d + 1) // 2
        if callable(stable):
            fb = stable
        else:
            fb = oscillator_feedback(recurrent_tau, scaling_factor, float(np.ravel(ssp_space.length_scale)[0]),
                                     max_radius, bool(stable))
        self.to_SSP = get_from_Fourier(d)
        self.to_Fourier = get_to_Fourier(d)
        with self:
            self.velocity_input = Node(size_in=n_dom, label=label + "_vel_input")
            self.input = Node(size_in=d, label=label + "_input")
            if with_gcs:    # pathintegration.py:150-154: the SSP is represented by a grid-cell population
                from .inputs import sparsity_to_x_intercept
                self.output = Ensemble(n_gcs, d, encoders=ssp_space.sample_grid_encoders(n_gcs),
                                       intercepts=nengo.dists.Choice([sparsity_to_x_intercept(d, 0.1)]),
                                       label=label + "_output")
            else:
                self.output = Node(size_in=d, label=label + "_output")
            self.oscillators = EnsembleArray(n_neurons, n_osc, ens_dimensions=3, radius=np.sqrt(2),
                                             label=label + "_vco", **ens_kwargs)
            self.oscillators.output.output = Identity()
            Connection(self.input, self.oscillators.input, transform=self.to_Fourier)
            self.recur_conns = []
            for k in range(1, n_osc):
                vco = self.oscillators.ea_ensembles[k]
                freq_row = np.vstack([np.zeros((2, n_dom)), ssp_space.phase_matrix[k].reshape(1, -1)])
                Connection(self.velocity_input, vco, transform=freq_row, synapse=None)
                self.recur_conns.append(Connection(vco, vco, function=fb, synapse=recurrent_tau,
                                                   solver=nengo.solvers.LstsqL2(weights=solver_weights)))
            dc = Node([1, 0, 0], label=label + "_zerofreq")
            Connection(dc, self.oscillators.ea_ensembles[0], synapse=None)
            Connection(self.oscillators.output, self.output, transform=self.to_SSP)


# ------------------------------------------------------------------------- binding
class Product(Network):
    """Element-wise product through two squaring populations: ab = ((a+b)^2 - (a-b)^2)/4."""

    def __init__(self, n_neurons, dimensions, input_magnitude=1.0, label="product", solver=Default, **kwargs):
        super().__init__(label=label, **kwargs)
        half = max(1, n_neurons // 2)
        radius = input_magnitude * np.sqrt(2)
        s = 1.0 / np.sqrt(2.0)
        with self:
            self.input_a = Node(size_in=dimensions, label=label + "_input_a")
            self.input_b = Node(size_in=dimensions, label=label + "_input_b")
            self.output = Node(size_in=dimensions, label=label + "_output")
            self.sq1 = EnsembleArray(half, n_ensembles=dimensions, ens_dimensions=1, radius=radius,
                                     label=label + "_sq1")
            self.sq2 = EnsembleArray(half, n_ensembles=dimensions, ens_dimensions=1, radius=radius,
                                     label=label + "_sq2")
            Connection(self.input_a, self.sq1.input, transform=s, synapse=None)
            Connection(self.input_b, self.sq1.input, transform=s, synapse=None)
            Connection(self.input_a, self.sq2.input, transform=s, synapse=None)
            Connection(self.input_b, self.sq2.input, transform=-s, synapse=None)
            plus = self.sq1.add_output("square", np.square, solver=solver)
            minus = self.sq2.add_output("square", np.square, solver=solver)
            Connection(plus, self.output, transform=0.5, synapse=None)
            Connection(minus, self.output, transform=-0.5, synapse=None)


class CircularConvolution(Network):
    def __init__(self, n_neurons, dimensions, invert_a=False, invert_b=False, input_magnitude=1.0,
                 label="circonv", solver=Default, **kwargs):
        super().__init__(label=label, **kwargs)
        tr_a = transform_in(dimensions, "A", invert_a)
        tr_b = transform_in(dimensions, "B", invert_b)
        tr_out = transform_out(dimensions)
        with self:
            self.input_a = Node(size_in=dimensions, label=label + "_input_a")
            self.input_b = Node(size_in=dimensions, label=label + "_input_b")
            self.product = Product(n_neurons, tr_out.shape[1], input_magnitude=input_magnitude * 2,
                                   label=label + "_product", solver=solver)
            self.output = Node(size_in=dimensions, label=label + "_output")
            Connection(self.input_a, self.product.input_a, transform=tr_a, synapse=None)
            Connection(self.input_b, self.product.input_b, transform=tr_b, synapse=None)
            Connection(self.product.output, self.output, transform=tr_out, synapse=None)


# ------------------------------------------------------------------------- associative memory
class AssociativeMemory(Network):
    """Voja-learned keys + PES-learned values with a gated error population."""

    def __init__(self, n_neurons, d_key, d_value, intercept, voja_learning_rate=5e-2,
                 pes_learning_rate=1e-3, encoders=None, radius=1, voja=True, tau=0.05, **kwargs):
        super().__init__(**kwargs)
        with self:
            self.key_input = Node(size_in=d_key, label="memory_input")
            self.value_input = Node(size_in=d_value)
            self.learning = Node(size_in=1)
            self.recall = Ensemble(n_neurons, d_value, label="memory_recall")
            mem_kwargs = dict(intercepts=[intercept] * n_neurons, radius=radius, label="memory")
            if encoders is not None:
                mem_kwargs["encoders"] = encoders
            self.memory = Ensemble(n_neurons, d_key, **mem_kwargs)
            if voja:
                rule = nengo.Voja(learning_rate=voja_learning_rate, post_synapse=None)
                self.conn_in = Connection(self.key_input, self.memory, synapse=None,
                                          learning_rule_type=rule, label="map_conn_in")
                Connection(self.learning, self.conn_in.learning_rule, synapse=None)
            else:
                self.conn_in = Connection(self.key_input, self.memory, synapse=None, label="map_conn_in")
            self.conn_out = Connection(self.memory, self.recall,
                                       learning_rule_type=nengo.PES(pes_learning_rate),
                                       function=lambda x: np.zeros(d_value), label="map_conn_pes")
            self.error = Ensemble(n_neurons, d_value, label="memory_pes_error")
            Connection(self.learning, self.error.neurons, transform=[[-2.5]] * n_neurons, synapse=None)
            Connection(self.value_input, self.error, transform=-1, synapse=tau)
            Connection(self.recall, self.error, synapse=tau)
            Connection(self.error, self.conn_out.learning_rule, synapse=tau)


# ------------------------------------------------------------------------- SLAM
def _default_intercept(landmark_sps, n_landmarks, cap=None):
    off_diag = (landmark_sps @ landmark_sps.T - np.eye(n_landmarks)).max()
    return off_diag if cap is None else min(off_diag, cap)


class SLAMNetwork(Network):
    def __init__(self, ssp_space, lm_space, view_rad, n_landmarks, pi_n_neurons, mem_n_neurons,
                 circonv_n_neurons, tau=0.01, tau_pi=0.05, update_thres=0.2, vel_scaling_factor=1.0,
                 rad_scaling_factor=1.0, shift_rate=0.1, voja_learning_rate=5e-4, pes_learning_rate=1e-2,
                 clean_up_method="grid", gc_n_neurons=0, encoders=None, voja=True, seed=0,
                 landmark_sps=None, intercept=None, grid_points_per_dim=100):
        super().__init__()
        if clean_up_method != "grid":
            raise NotImplementedError("only the grid clean-up node is on the hot path (SURVEY.md §8f-4)")
        d, n_dom = ssp_space.ssp_dim, ssp_space.domain_dim
        rng = np.random.RandomState(seed=seed)
        if landmark_sps is None:
            landmark_sps = lm_space.vectors
        if not voja and encoders is None:
            encoders = landmark_sps[rng.randint(n_landmarks, size=mem_n_neurons)]
        if intercept is None:
            intercept = _default_intercept(landmark_sps, n_landmarks, cap=0.5)
        # object-vector cells: encoders are SSPs of quasi-random displacement vectors
        # (drawn from the *global* NumPy stream, exactly like slam.py:206 — SURVEY.md F7)
        ovc_pts = nengo.dists.ScatteredHypersphere(surface=False, min_magnitude=1e-3).sample(mem_n_neurons, n_dom)
        enc_fn = getattr(ssp_space, "encode_host", ssp_space.encode)
        ovc_encoders = enc_fn(ovc_pts)
        self.sample_ssps, self.sample_points = ssp_space.get_sample_pts_and_ssps(grid_points_per_dim)
        self.clean_up_fun = GridCleanup(self.sample_ssps)
        make_unitary = ssp_space.make_unitary

        with self:
            self.velocity_input = Node(size_in=n_dom, label="vel_input")
            self.landmark_id_input = Node(size_in=d, label="lm_id_input")
            self.landmark_vec_ssp = Node(size_in=d, label="lm_vecssp_input")
            self.no_landmark_in_view = Node(size_in=1, label="lm_in_view_input")
            self.update_state = Node(GatedCorrection(d, shift_rate, update_thres), size_in=2 * d + 1)
            Connection(self.no_landmark_in_view, self.update_state[-1], synapse=None)

            self.pathintegrator = PathIntegration(ssp_space, pi_n_neurons, tau_pi, max_radius=rad_scaling_factor,
                                                  scaling_factor=vel_scaling_factor, stable=True,
                                                  solver_weights=False, label="pathint")
            self.output = self.pathintegrator.output
            Connection(self.velocity_input, self.pathintegrator.velocity_input, synapse=None)
            Connection(self.update_state, self.pathintegrator.input, synapse=None)

            self.ovc_ens = Ensemble(mem_n_neurons, d, encoders=ovc_encoders)
            Connection(self.landmark_vec_ssp, self.ovc_ens, synapse=None)
            self.landmark_ssp_ens = CircularConvolution(circonv_n_neurons, dimensions=d, label="landmark_circonv")
            Connection(self.ovc_ens, self.landmark_ssp_ens.input_b, synapse=None)

            if gc_n_neurons <= 0:
                self.gridcells = Node(self.clean_up_fun, size_in=d)
                Connection(self.pathintegrator.output, self.gridcells, synapse=tau)
                Connection(self.gridcells, self.landmark_ssp_ens.input_a, synapse=None)
            else:   # slam.py:274-281: the cleaned-up SSP is represented by a grid-cell ensemble
                gc_encoders = ssp_space.sample_grid_encoders(gc_n_neurons)
                self.cleanup = Node(self.clean_up_fun, size_in=d)
                self.gridcells = Ensemble(gc_n_neurons, d, encoders=gc_encoders,
                                          intercepts=nengo.dists.CosineSimilarity(d + 2))
                Connection(self.pathintegrator.output, self.cleanup, synapse=tau)
                Connection(self.cleanup, self.gridcells, synapse=None)
                Connection(self.gridcells, self.landmark_ssp_ens.input_a, synapse=tau)

            self.assomemory = AssociativeMemory(mem_n_neurons, d, d, intercept,
                                                voja_learning_rate=voja_learning_rate,
                                                pes_learning_rate=pes_learning_rate, voja=voja, encoders=encoders)
            Connection(self.landmark_id_input, self.assomemory.key_input, synapse=None)
            Connection(self.landmark_ssp_ens.output, self.assomemory.value_input, synapse=tau)
            Connection(self.no_landmark_in_view, self.assomemory.learning, synapse=None)

            self.position_estimate = CircularConvolution(circonv_n_neurons, d, invert_a=True, label="newpos_circonv")
            Connection(self.ovc_ens, self.position_estimate.input_a, synapse=tau,
                       function=lambda x: make_unitary(x))
            Connection(self.assomemory.recall, self.position_estimate.input_b, synapse=tau,
                       function=lambda x: make_unitary(x))
            Connection(self.position_estimate.output, self.update_state[:d], synapse=tau)
            Connection(self.pathintegrator.output, self.update_state[d:-1], synapse=tau)


class SLAMViewNetwork(Network):
    def __init__(self, ssp_space, lm_space, view_rad, n_landmarks, pi_n_neurons, mem_n_neurons,
                 circonv_n_neurons, tau=0.01, tau_pi=0.05, update_thres=0.2, vel_scaling_factor=1.0,
                 rad_scaling_factor=1.0, shift_rate=0.1, voja_learning_rate=5e-4, pes_learning_rate=1e-2,
                 clean_up_method="grid", gc_n_neurons=0, encoders=None, voja=True, seed=0,
                 grid_points_per_dim=100):
        super().__init__()
        if clean_up_method != "grid" or gc_n_neurons > 0:
            raise NotImplementedError("only the grid clean-up node is on the hot path (SURVEY.md §8f-4)")
        d, n_dom = ssp_space.ssp_dim, ssp_space.domain_dim
        rng = np.random.RandomState(seed=seed)
        landmark_sps = lm_space.vectors
        if not voja and encoders is None:
            encoders = landmark_sps[rng.randint(n_landmarks, size=mem_n_neurons)]
        intercept = _default_intercept(landmark_sps, n_landmarks)
        self.sample_ssps, self.sample_points = ssp_space.get_sample_pts_and_ssps(grid_points_per_dim)
        self.clean_up_fun = GridCleanup(self.sample_ssps)
        make_unitary = ssp_space.make_unitary

        with self:
            self.velocity_input = Node(size_in=n_dom, label="vel_input")
            self.view_input = Node(size_in=d, label="lm_input")
            self.no_landmark_in_view = Node(size_in=1, label="lm_in_view_input")
            self.update_state = Node(GatedCorrection(d, shift_rate, update_thres), size_in=2 * d + 1)
            Connection(self.no_landmark_in_view, self.update_state[-1], synapse=None)

            self.pathintegrator = PathIntegration(ssp_space, pi_n_neurons, tau_pi, max_radius=rad_scaling_factor,
                                                  scaling_factor=vel_scaling_factor, stable=True, label="pathint")
            self.output = self.pathintegrator.output
            Connection(self.velocity_input, self.pathintegrator.velocity_input, synapse=None)
            Connection(self.update_state, self.pathintegrator.input, synapse=None)

            self.assomemory = AssociativeMemory(mem_n_neurons, d, d, intercept,
                                                voja_learning_rate=voja_learning_rate,
                                                pes_learning_rate=pes_learning_rate, voja=voja, encoders=encoders)
            Connection(self.view_input, self.assomemory.key_input, synapse=None)
            Connection(self.no_landmark_in_view, self.assomemory.learning, synapse=None)

            self.gridcells = Node(self.clean_up_fun, size_in=d)
            Connection(self.pathintegrator.output, self.gridcells, synapse=tau)
            Connection(self.gridcells, self.assomemory.value_input, synapse=None)

            Connection(self.assomemory.recall, self.update_state[:d], function=lambda x: make_unitary(x), synapse=tau)
            Connection(self.pathintegrator.output, self.update_state[d:-1], synapse=tau)


class SLAMLoihiNetwork(Network):
    """SLAM without Python node functions (``slam_loihi.py:190-293``): fixed landmark encoders instead of Voja, a
    neural correction population, and the update gate computed by neurons — ``|a + b|^2 - |a - b|^2`` from two arrays of
    squaring populations feeds a threshold population that inhibits the correction neurons.  Input nodes may be passed
    in (the driver's table nodes, ``run_slam.py:171-176``) or are created as pass-through nodes."""

    def __init__(self, ssp_space, lm_space, view_rad, n_landmarks, pi_n_neurons, mem_n_neurons, circonv_n_neurons,
                 dotprod_n_neurons, velocity_input=None, landmark_vecssp_input=None, landmark_sp_input=None,
                 no_landmark_in_view=None, tau=0.01, tau_pi=0.05, update_thres=0.2, vel_scaling_factor=1.0,
                 rad_scaling_factor=1, shift_rate=0.1, pes_learning_rate=1e-2, encoders=None, solver=Default,
                 pi_solver_weights=False, seed=0):
        super().__init__()
        d, n_dom = ssp_space.ssp_dim, ssp_space.domain_dim
        landmark_sps = lm_space.vectors
        rng = np.random.RandomState(seed=seed)
        if encoders is None:
            encoders = landmark_sps[rng.randint(n_landmarks, size=mem_n_neurons), :]
        intercept = _default_intercept(landmark_sps, n_landmarks)

        def given(node, size, label):
            return Node(size_in=size, label=label) if node is None else node

        with self:
            self.velocity_input = given(velocity_input, n_dom, "vel_input")
            self.landmark_vecssp_input = given(landmark_vecssp_input, d, "lm_vecssp_input")
            self.landmark_sp_input = given(landmark_sp_input, d, "lm_sp_input")
            self.no_landmark_in_view = given(no_landmark_in_view, 1, "lm_in_view_input")

            self.pathintegrator = PathIntegration(ssp_space, pi_n_neurons, tau_pi, max_radius=rad_scaling_factor,
                                                  scaling_factor=vel_scaling_factor, stable=True, with_gcs=False,
                                                  solver_weights=pi_solver_weights, label="pathint")
            Connection(self.velocity_input, self.pathintegrator.velocity_input, synapse=None)
            self.output = self.pathintegrator.output

            # landmark location = own position (*) displacement
            self.landmark_ssp_ens = CircularConvolution(circonv_n_neurons, dimensions=d, solver=solver,
                                                        label="landmark_circonv")
            Connection(self.pathintegrator.output, self.landmark_ssp_ens.input_a, synapse=tau)
            Connection(self.landmark_vecssp_input, self.landmark_ssp_ens.input_b, synapse=0)

            # environment map: landmark SP -> landmark location SSP, PES-learned
            mem = self.assomemory = nengo.Network(seed=seed)
            mem.memory = Ensemble(mem_n_neurons, d, intercepts=[intercept] * mem_n_neurons, encoders=encoders,
                                  radius=1, label="memory")
            mem.recall = Ensemble(mem_n_neurons, d, label="memory_recall")
            Connection(self.landmark_sp_input, mem.memory, synapse=None, label="map_conn_in")
            mem.conn_out = Connection(mem.memory, mem.recall, learning_rule_type=nengo.PES(pes_learning_rate),
                                      label="map_conn_pes", function=lambda x: np.zeros(d))
            mem_error = Ensemble(mem_n_neurons, d, label="memory_pes_error")
            Connection(self.no_landmark_in_view, mem_error.neurons, transform=[[-2.5]] * mem_n_neurons, synapse=None)
            Connection(self.landmark_ssp_ens.output, mem_error, transform=-1, synapse=tau)
            Connection(mem.recall, mem_error, synapse=tau)
            Connection(mem_error, mem.conn_out.learning_rule, synapse=tau)

            # position from the map: recalled location (*) inverse displacement
            self.position_estimate = CircularConvolution(circonv_n_neurons, d, input_magnitude=1, invert_a=True,
                                                         solver=solver, label="newpos_circonv")
            Connection(self.landmark_vecssp_input, self.position_estimate.input_a, synapse=None)
            Connection(mem.recall, self.position_estimate.input_b, synapse=tau)

            # correction population, fed back through a long synapse
            self.correction = Ensemble(mem_n_neurons, d, label="correction_ens")
            Connection(self.position_estimate.output, self.correction, synapse=tau, transform=1)
            Connection(self.pathintegrator.output, self.correction, synapse=tau, transform=-1)
            Connection(self.correction, self.pathintegrator.input, synapse=0.1, transform=shift_rate)

            # gate: the correction is inhibited unless <estimate, state> exceeds the threshold
            bias = Node(1, label="threshold_bias")
            self.threshold = Ensemble(circonv_n_neurons, 1, intercepts=nengo.dists.Choice([update_thres]),
                                      encoders=np.ones((circonv_n_neurons, 1)), label="threshold")
            Connection(bias, self.threshold, synapse=None)
            Connection(self.no_landmark_in_view, self.threshold, synapse=None)
            Connection(self.threshold, self.correction.neurons, transform=[[-5]] * mem_n_neurons, synapse=0.05)
            half = max(1, dotprod_n_neurons // 2)
            squares = [EnsembleArray(half, n_ensembles=d, ens_dimensions=1, radius=np.sqrt(2), label=f"dotprod_sq{k}")
                       for k in (1, 2)]
            tr = 1.0 / np.sqrt(2.0)
            for arr, sign in zip(squares, (1.0, -1.0)):
                Connection(self.position_estimate.output, arr.input, transform=tr, synapse=tau)
                Connection(self.pathintegrator.output, arr.input, transform=sign * tr, synapse=tau)
            for i in range(d):
                Connection(squares[0].ea_ensembles[i], self.threshold, function=lambda x: -0.5 * x ** 2, synapse=tau)
                Connection(squares[1].ea_ensembles[i], self.threshold, function=lambda x: 0.5 * x ** 2, synapse=tau)

