"""Generate ``reference_golden.npz`` by executing the UNMODIFIED reference sources.

Run in the authoring container (``/root/reference`` mounted):

    python tests/golden/make_golden.py

The reference (``sspslam/sspspace.py``, ``networks/{pathintegration,binding,slam}.py``,
``utils/utils.py``) is imported through ``sspslam_b200.refload`` — its ``nengo`` imports
resolve to this repo's declaration shim (nengo itself is not installable here), and no
reference file is edited or copied.  The vectors pin the NumPy part of the hot path
 for ``oracle/ssp_ref.py``, ``sspslam_b200.sspspace``,
``sspslam_b200.networks`` and ``sspslam_b200.inputs``.  The arithmetic of the stepped
path itself lives in third-party nengo and stays unpinned.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from sspslam_b200 import refload  # noqa: E402
from sspslam_b200.nengo_shim.processes import WhiteSignal  # noqa: E402


def main():
    ref = refload.load_reference()
    from sspslam.networks.pathintegration import get_to_Fourier, get_from_Fourier
    from sspslam.networks import binding as rb
    from sspslam.networks.slam import get_slam_input_functions2, get_slam_input_functions
    from sspslam.utils import Rd_sampling, sparsity_to_x_intercept

    out = {}
    rng = np.random.RandomState(7)
    bounds2 = np.tile([-1.0, 1.0], (2, 1))

    # ---- SSP spaces (sspspace.py:678-731, 252-273, 312-358)
    sp55 = ref.HexagonalSSPSpace(2, ssp_dim=55, domain_bounds=bounds2, length_scale=0.2)
    sp97 = ref.HexagonalSSPSpace(2, ssp_dim=97, domain_bounds=bounds2, length_scale=0.2)
    sp3d = ref.HexagonalSSPSpace(3, ssp_dim=55, domain_bounds=np.tile([-1.0, 1.0], (3, 1)), length_scale=0.3,
                                 rng=np.random.default_rng(0))
    out["hex55_phase"] = sp55.phase_matrix
    out["hex97_phase"] = sp97.phase_matrix
    out["hex3d_phase"] = sp3d.phase_matrix
    out["hex_dims"] = np.array([sp55.ssp_dim, sp97.ssp_dim, sp3d.ssp_dim,
                                ref.HexagonalSSPSpace(3, n_rotates=9, n_scales=9).ssp_dim])
    pts = rng.uniform(-1, 1, size=(64, 2))
    pts[0] = 0.0
    out["enc_pts"] = pts
    out["enc55"] = sp55.encode(pts)
    out["enc97"] = sp97.encode(pts)
    pts3 = rng.uniform(-1, 1, size=(16, 3))
    out["enc3d_pts"] = pts3
    out["enc3d"] = sp3d.encode(pts3)
    # decode of noisy (and scaled, and zero) SSPs on the 100x100 grid
    noisy = out["enc55"] + 0.08 * rng.standard_normal(out["enc55"].shape)
    noisy[1] *= 3.7
    noisy[2] = 0.0
    noisy[3] = 1e-8 * noisy[3]
    out["dec55_in"] = noisy
    out["dec55_out"] = sp55.decode(noisy, "from-set", "grid", 100)
    s_ssps, s_pts = sp55.get_sample_pts_and_ssps(100, "grid")
    out["grid55_pts_head"] = s_pts[:205]
    out["grid55_ssps_rows"] = s_ssps[[0, 1, 99, 100, 5050, 9999]]
    out["dec55_idx"] = np.argmax(s_ssps @ (noisy / np.maximum(np.linalg.norm(noisy, axis=1, keepdims=True), 1e-300)).T,
                                 axis=0)
    a, b = out["enc55"][4:8], out["enc55"][8:12]
    out["bind55"] = sp55.bind(a, b)
    out["invert55"] = sp55.invert(a)
    out["unitary55"] = np.stack([sp55.make_unitary(v) for v in noisy[4:8]])
    out["identity55"] = sp55.identity()

    # ---- SP space (sspspace.py:43-81)
    lm = ref.SPSpace(50, 55, seed=0)
    out["sp50_vectors"] = lm.vectors
    lm8 = ref.SPSpace(8, 25, seed=3)
    out["sp8_vectors"] = lm8.vectors

    # ---- Fourier layouts (pathintegration.py:816-844) and circular-convolution transforms (binding.py:23-89)
    for d in (7, 55):
        out[f"toF{d}"] = get_to_Fourier(d)
        out[f"fromF{d}"] = get_from_Fourier(d)
        out[f"trA{d}"] = rb.transform_in(d, "A", False)
        out[f"trAinv{d}"] = rb.transform_in(d, "A", True)
        out[f"trB{d}"] = rb.transform_in(d, "B", False)
        out[f"trOut{d}"] = rb.transform_out(d)
    x, y = rng.standard_normal(55), rng.standard_normal(55)
    out["cc_x"], out["cc_y"] = x, y
    out["cc_xy"] = rb.circconv(x, y)
    out["cc_xinv_y"] = rb.circconv(x, y, invert_a=True)

    # ---- utils (utils.py:5-10, 41-55)
    out["rd_50_2_s0"] = Rd_sampling(50, 2, seed=0)
    out["rd_20_2_s1000"] = Rd_sampling(20, 2, seed=1000)
    out["rd_10_3_s05"] = Rd_sampling(10, 3)
    out["x_intercept_55_01"] = np.array(sparsity_to_x_intercept(55, 0.1))

    # ---- input closures (slam.py:442-497) evaluated like nengo does: t = n*dt, n = 1..N
    dt, T, N = 0.001, 20.0, 400
    path = np.hstack([WhiteSignal(T, high=0.1, seed=11 + i).run(T, dt=dt) for i in range(2)])
    path = 1.8 * (path - path.min(0)) / (path.max(0) - path.min(0)) - 0.9
    vels = np.diff(path, axis=0, prepend=path[:1]) / dt
    obj = 1.8 * (Rd_sampling(50, 2, seed=11) - 0.5)
    vec_to = obj[None, :, :] - path[:, None, :]
    out["in_path"], out["in_obj"] = path[:N + 4], obj
    fns = get_slam_input_functions2(sp55, lm, vels, vec_to, 0.2)
    velocity_func, vel_scale, is_in_view, lm_id_func, lm_sp_func, lm_vec_func, lm_vecssp_func = fns
    out["in_vel_scale"] = np.array(vel_scale)
    ts = np.arange(1, N + 1) * dt
    out["in_vel"] = np.stack([velocity_func(t) for t in ts])
    out["in_nolm"] = np.array([is_in_view(t) for t in ts], dtype=np.float64)
    out["in_lm_sp"] = np.stack([lm_sp_func(t) for t in ts])
    out["in_lmvec_ssp"] = np.stack([lm_vecssp_func(t) for t in ts])
    real_ssp = sp55.encode(path)
    out["in_init"] = np.stack([real_ssp[int((t - dt) / dt)] if t < 0.05 else np.zeros(55) for t in ts])
    # index quirk K7 over a long horizon (int((t-dt)/dt) vs n-1)
    n_all = np.arange(1, 200001)
    t_all = n_all * dt
    out["k7_iprev"] = np.array([int((t - dt) / dt) for t in t_all[:5000]])
    out["k7_icur"] = np.array([int(np.minimum(np.floor(t / dt), 20000 - 2)) for t in t_all[:5000]])

    # ---- WhiteSignal restatement is ours (nengo absent): store its output so a change is caught
    out["white_T20_h01_s0"] = WhiteSignal(20.0, high=0.1, seed=0).run(20.0, dt=dt)[::100, 0]

    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst) // 1024, "KiB;", len(out), "arrays")


if __name__ == "__main__":
    main()
