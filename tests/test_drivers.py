"""The driver scripts BASELINE.json's configs name (swraytracing_b200/drivers.py) against their oracle
restatements (oracle/swrt_oracle.py: *_driver), plus CPU known answers for the restatements themselves.

  config 1  SW_zero_background_raytracing.m   ode23 with output times (dense output), RHS with gH k/omega
  config 2  symplectic_full_fourier.m         ode_symplectic + absolute-frequency drift
  config 5  ray_trace_sw/raytrace_sw.m        geostrophic projection + step_packet_xka
            ray_trace_sw/raytrace.m           Childress-Soward + step_packet (the caller of step_packet)
(config 3 = drivers.qgsw_raytrace is covered in test_gpu_parity.py / test_reference_goldens.py.)
"""
from pathlib import Path

import numpy as np
import pytest

from oracle import swrt_oracle as O

L = 2 * np.pi


def _pv_frame(nx, U_g=0.5, seed=146):
    x = O.matlab_linspace(-L / 2, L / 2, nx)
    X, Y = np.meshgrid(x, x)
    return O.initial_q(X, Y, U_g, 3.0, O.matlab_rand_stream(seed), k_max=4)


def _sw_state(nx, seed=3):
    """balanced random-phase vortical part + a small unbalanced wave part, schema S(:,:,1:3) = [u,v,eta]"""
    rs = np.random.RandomState(seed)
    kx_, ky_ = O.wavenumbers(nx)
    K2 = kx_ ** 2 + ky_ ** 2
    amp = np.where((K2 > 0) & (K2 <= 36), 1.0 / (1.0 + K2) ** 1.5, 0.0)
    f, Cg = 3.0, 1.0
    psik = 0.05 * amp * np.exp(2j * np.pi * rs.rand(*K2.shape))
    etak = (f / Cg ** 2) * psik
    uk, vk = -1j * ky_ * psik, 1j * kx_ * psik
    wave = 0.005 * amp * np.exp(2j * np.pi * rs.rand(*K2.shape))
    S = np.stack([O.k2g(uk + wave), O.k2g(vk - 0.5 * wave), O.k2g(etak + 0.3 * wave)], axis=2)
    return S, f, Cg


# ------------------------------------------------------------------------------------------------
# CPU: known answers for the restated scripts
# ------------------------------------------------------------------------------------------------
def test_matlab_startup_rand_stream():
    """raytrace.m / raytrace_sw.m never call rng: MATLAB starts with mt19937ar seed 0 == 5489, whose first
    draws are the familiar 0.8147, 0.9058, 0.1270"""
    r = O.matlab_rand_stream(5489).rand(3)
    assert np.allclose(r, [0.814723686393179, 0.905791937075619, 0.126986816293506], atol=1e-15)


def test_childress_soward_as_written_differs_only_in_vx():
    _, U, G = O.childress_soward(32, L, 0.1, 4.0, 0.25)
    _, U2, G2 = O.childress_soward_as_written(32, L, 0.1, 4.0, 0.25)
    for n in ("u_x", "u_y", "v_y"):
        assert np.array_equal(G[n], G2[n])
    assert np.abs(G["v_x"] - G2["v_x"]).max() > 1e-2          # raytrace.m:36 '*' vs '.*'
    _, _, G0 = O.childress_soward(32, L, 0.1, 4.0, 0.0)
    _, _, G0w = O.childress_soward_as_written(32, L, 0.1, 4.0, 0.0)
    assert np.array_equal(G0["v_x"], G0w["v_x"])               # identical when a = 0


def test_geostrophic_projection_known_answers():
    """a balanced state is its own geostrophic part; the reference's bundled inertia-gravity-wave state
    (rsw/matlab.mat) has none (raytrace_sw.m:25-35 = rsw/wavevortdecomp.m:41-45)"""
    nx = 32
    kx_, ky_ = O.wavenumbers(nx)
    K2 = kx_ ** 2 + ky_ ** 2
    rs = np.random.RandomState(0)
    f, Cg = 3.0, 1.0
    psik = np.where((K2 > 0) & (K2 < 40), 1.0, 0.0) * np.exp(2j * np.pi * rs.rand(*K2.shape)) / (1 + K2)
    S = np.stack([O.k2g(-1j * ky_ * psik), O.k2g(1j * kx_ * psik), O.k2g(f / Cg ** 2 * psik)], axis=2)
    U, GradU, H = O.geostrophic_fields(S, f, Cg)
    assert np.abs(U["u"] - S[:, :, 0]).max() < 1e-14 and np.abs(U["v"] - S[:, :, 1]).max() < 1e-14
    assert np.abs(H - 1 - S[:, :, 2]).max() < 1e-14
    assert np.abs(GradU["u_x"] + GradU["v_y"]).max() < 1e-13    # non-divergent
    d = np.load(Path(__file__).parent / "golden" / "rsw_workspace_frame.npz")
    Sw = np.stack([O.k2g(d["damask"] * d["Sk"][:, :, j]) for j in range(3)], axis=2)
    Uw, _, Hw = O.geostrophic_fields(Sw, float(d["f"]), float(d["Cg"]))
    assert np.abs(Sw[:, :, 0]).max() > 0.03
    assert max(np.abs(Uw["u"]).max(), np.abs(Uw["v"]).max(), np.abs(Hw - 1).max()) < 1e-10


def test_sw_zero_background_is_free_propagation():
    nx = 32
    r = O.sw_zero_background_driver(np.zeros((nx, nx)), nx, 3.0, 1.0, Nparticles=10, Tend=0.5)
    k0, x0 = r["solver_k"][0], r["solver_x"][0]
    assert np.abs(r["solver_k"] - k0).max() == 0.0
    w = np.sqrt(9.0 + (k0 ** 2).sum(axis=0))
    for j, t in enumerate(r["t_hist"]):
        assert np.abs(r["solver_x"][j] - (x0 + k0 / w * t)).max() < 1e-14
    assert np.abs(r["solver_error"]).max() < 1e-15


def test_ode23_tableau_and_dense_output_match_scipy_rk23():
    """independent check of the restated Bogacki-Shampine pair: one step and its cubic interpolant against
    scipy.integrate.RK23 (same tableau, same dense-output polynomial P = ntrp23's BI)"""
    from scipy.integrate import RK23
    from scipy.integrate._ivp.rk import rk_step
    fun = lambda t, y: np.array([y[1] + 0.1 * t, -np.sin(y[0]) - 0.05 * y[1]])
    y0 = np.array([0.7, -0.2]); h = 0.05
    f0 = fun(0.0, y0)
    K = np.empty((4, 2))
    ynew, fnew = rk_step(fun, 0.0, y0, f0, h, RK23.A, RK23.B, RK23.C, K)
    assert np.allclose(RK23.P, [[1, -4 / 3, 5 / 9], [0, 1, -2 / 3], [0, 4 / 3, -8 / 9], [0, -1, 1]])
    assert np.allclose(-RK23.E, [-5 / 72, 1 / 12, 1 / 9, -1 / 8])
    # force exactly one step of size h with output at interior points
    ts = np.array([0.0, 0.3 * h, 0.8 * h, h])
    Y, st = O.ode23(fun, ts, y0, rtol=1e-1, atol=1e-1)
    if st["nsteps"] == 1:
        assert np.allclose(Y[-1], ynew, rtol=0, atol=1e-15)
        for j in (1, 2):
            s = ts[j] / h
            assert np.allclose(Y[j], y0 + h * (K.T @ (RK23.P @ np.array([s, s * s, s ** 3]))), rtol=0, atol=1e-15)
    else:                                                       # controller chose smaller steps: compare with the exact flow
        from scipy.integrate import solve_ivp
        ref = solve_ivp(fun, [0, h], y0, rtol=1e-12, atol=1e-14, t_eval=ts).y.T
        assert np.abs(Y - ref).max() < 1e-6


def test_parse_data_reads_the_reference_header(tmp_path):
    from swraytracing_b200.drivers import parse_data
    header = ["x"] * 10 + ["Resolution: 256x256", "Number of packets: 50", "Initial wavenumber radius: 6.000000",
                           "Time step: 0.004845", "Simulation time: 666.666667", "Spin-up time: 400.000000", "Steps per save: 50",
                           "Steps per packet save: 5", "Coriolis parameter: 3.000000", "Group velocity: 1.000000",
                           "Background velocity (parameter,computed): (0.500000,0.506570)", "Froude Number: 0.506570",
                           "Deformation wavenumber: 3.000000", "Simulation progress:  0.00%  0.04%"]
    p = tmp_path / "run.log"
    p.write_text("\n".join(header))
    assert parse_data(p) == (256, 50, 3.0, 1.0, 0.5)            # symplectic_full_fourier.m:3-4 on run-4/run.log (= run.log:10-22)


# ------------------------------------------------------------------------------------------------
# GPU: product drivers vs the restated scripts
# ------------------------------------------------------------------------------------------------
def _relerr(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(1.0, np.abs(b).max()))


@pytest.mark.gpu
@pytest.mark.parametrize("a", [0.25, 0.0])
def test_raytrace_script(a):
    from swraytracing_b200 import drivers
    ref, dt, ns = O.raytrace_driver(a=a, nx=64, nsteps=25)
    got = drivers.raytrace(a=a, nx=64, nsteps=25)
    assert got["dt"] == dt and got["nsteps"] == ns
    for n in ("x", "y", "k", "l"):
        assert _relerr(got["P"][n], ref[n]) < 1e-9, n
    # strided history = the same columns
    got4 = drivers.raytrace(a=a, nx=64, nsteps=25, save_stride=4)
    for n in ("x", "k"):
        assert np.array_equal(got4["P"][n], got["P"][n][:, ::4])
    assert np.allclose(got["omega"], np.sqrt(16.0 + got["K"] ** 2))


@pytest.mark.gpu
def test_raytrace_sw_script():
    from swraytracing_b200 import drivers
    S, f, Cg = _sw_state(64)
    ref, info = O.raytrace_sw_driver(S, f, Cg, np_=6, nsteps=20)
    got = drivers.raytrace_sw(S, f, Cg, np_=6, nsteps=20)
    assert abs(got["U0"] - info["U0"]) < 1e-13 and abs(got["dt"] - info["dt"]) < 1e-15
    for n in ("u", "v"):
        assert np.abs(got["U"][n] - info["U"][n]).max() < 1e-13
    assert np.abs(got["H"] - info["H"]).max() < 1e-13
    for n in ("x", "y", "k", "l", "a"):
        assert _relerr(got["P"][n], ref[n]) < 1e-9, n
    assert np.all(got["omega"] > f)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["lagrange", "spectral"])
def test_symplectic_full_fourier_script(mode):
    import swraytracing_b200 as S
    from swraytracing_b200 import drivers
    nx = 32
    q = _pv_frame(nx)
    ref = O.symplectic_full_fourier_driver(q, nx, 3.0, 1.0, Nparticles=10, Tend=0.6, mode=mode)
    got = drivers.symplectic_full_fourier(q, nx=nx, f=3.0, Cg=1.0, Nparticles=10, Tend=0.6,
                                          mode=S.MODE_LAGRANGE6 if mode == "lagrange" else S.MODE_SPECTRAL)
    assert abs(got["U0"] - ref["U0"]) < 1e-12 and abs(got["dt"] - ref["dt"]) < 1e-14
    assert got["solver_x"].shape == ref["solver_x"].shape and np.array_equal(got["solver_t"], ref["solver_t"])
    assert _relerr(got["solver_x"], ref["solver_x"]) < 1e-9
    assert _relerr(got["solver_k"], ref["solver_k"]) < 1e-9
    assert np.abs(got["solver_error"] - ref["solver_error"]).max() < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["lagrange", "spectral"])
def test_SW_zero_background_raytracing_script(mode):
    import swraytracing_b200 as S
    from swraytracing_b200 import drivers
    nx = 32
    q = _pv_frame(nx)
    ref = O.sw_zero_background_driver(q, nx, 3.0, 1.0, Nparticles=10, Tend=0.5, mode=mode)
    got = drivers.SW_zero_background_raytracing(q, nx=nx, f=3.0, Cg=1.0, Nparticles=10, Tend=0.5,
                                                mode=S.MODE_LAGRANGE6 if mode == "lagrange" else S.MODE_SPECTRAL)
    assert got["Nsteps"] == ref["Nsteps"] and np.array_equal(got["t_hist"], ref["t_hist"])
    for key in ("nsteps", "nfailed", "nfevals"):
        assert got["stats"][key] == ref["stats"][key], key       # identical accept/reject decisions
    assert _relerr(got["solver_x"], ref["solver_x"]) < 1e-9
    assert _relerr(got["solver_k"], ref["solver_k"]) < 1e-9
    assert np.abs(got["solver_error"] - ref["solver_error"]).max() < 1e-9


@pytest.mark.gpu
def test_SW_zero_background_true_zero_flow_is_analytic():
    from swraytracing_b200 import drivers
    nx = 32
    got = drivers.SW_zero_background_raytracing(np.zeros((nx, nx)), nx=nx, f=3.0, Cg=2.0, Nparticles=1000, Tend=0.5)
    k0, x0 = got["solver_k"][0], got["solver_x"][0]
    assert np.abs(got["solver_k"] - k0).max() == 0.0
    w = np.sqrt(9.0 + 4.0 * (k0 ** 2).sum(axis=0))              # gH = Cg^2 = 4: tells gH k/omega from Cg k/omega
    for j, t in enumerate(got["t_hist"]):
        assert np.abs(got["solver_x"][j] - (x0 + 4.0 * k0 / w * t)).max() < 1e-13
    with pytest.raises(ValueError):
        drivers.SW_zero_background_raytracing(np.zeros((nx, nx)), nx=nx, f=3.0, Cg=1.0)   # Tend = 1/(f Fr^2) is infinite


# ------------------------------------------------------------------------------------------------
# config 4: qg2layersw_raytrace.m (two-layer QG + packets through the top layer)
# ------------------------------------------------------------------------------------------------
def _qg2_setup(nx=32, L=20.0):
    kx_, ky_ = O.wavenumbers(nx)
    kap = 2 * np.pi / L
    return kx_ * kap, ky_ * kap


def test_qg2_inversion_matrix_and_exponential():
    """B is the inverse of the two-layer PV operator [-K2-F, F; F, -K2-F] (F = K_d2/2), zero at K = 0; the
    eig-based exp(L t) of the restatement equals scipy's expm"""
    from scipy.linalg import expm
    kx_, ky_ = _qg2_setup()
    K2 = kx_ ** 2 + ky_ ** 2
    B, FL, LV, LD, LV1 = O.qg2_operators(kx_, ky_, 3.0, 0.0, 0.5, 0.4, 0.1 * (20 / 32) ** 8, 4)
    F = 1.5
    for a, b in ((3, 2), (15, 0), (20, 7), (30, 15)):
        A = np.array([[-K2[a, b] - F, F], [F, -K2[a, b] - F]])
        want = np.zeros((2, 2)) if K2[a, b] == 0 else np.eye(2)
        assert np.allclose(B[:, :, a, b] @ A, want, atol=1e-13)
    E = O.qg2_expL(LV, LD, LV1, 0.05)
    for a in range(0, 31, 5):
        for b in range(0, 16, 3):
            assert np.abs(E[:, :, a, b] - expm(0.05 * FL[:, :, a, b])).max() < 1e-13
    assert np.abs(E[:, :, 15, 0] - np.eye(2)).max() == 0.0       # K = 0: factor_L = 0


def test_qg2_restated_driver_runs_and_changes_dt():
    r = O.qg2layersw_driver(32, 8, 2, 30.0, 0.0, 0.3, 3.0, 1.0, max_steps=4, k_max=6)
    assert r["steps"] == 4 and r["packet_steps"] == 4 and r["ode23_failed"] == 0
    assert np.all(np.isfinite(r["packets"][0])) and np.isfinite(np.abs(r["qk"]).max())


@pytest.mark.gpu
def test_qg2_device_steps_match_restatement():
    from swraytracing_b200.engine import QG2Flow
    nx, L = 32, 20.0
    kx_, ky_ = _qg2_setup(nx, L)
    K2 = kx_ ** 2 + ky_ ** 2
    x = O.matlab_linspace(-L / 2, L / 2, nx)
    X, Y = np.meshgrid(x, x, indexing="ij")
    q1 = O.initial_q(X, Y, 0.3, 3.0, O.matlab_rand_stream(5), k_min=10, k_max=6)
    qk = np.stack([O.g2k(q1), O.g2k(-q1)], axis=2)
    nu = 0.1 * (L / nx) ** 8
    B, _, LV, LD, LV1 = O.qg2_operators(kx_, ky_, 3.0, 0.0, 0.5, 0.4, nu, 4)
    qg = QG2Flow(nx, L, qk[:, :, 0], qk[:, :, 1], 3.0, 0.0, 0.5, 0.4, nu, 4)
    assert abs(qg.max_speed() - O.qg2_max_speed(qk, 3.0, K2, kx_, ky_, 0.5)) < 1e-13
    Qm = [np.zeros_like(qk), np.zeros_like(qk)]
    dts = [0.05, 0.05, 0.05, 0.05, 0.02, 0.02, 0.02]              # a dt change re-builds expLdt/expL2dt (:156-165)
    for step, dt in enumerate(dts, 1):
        E1, E2 = O.qg2_expL(LV, LD, LV1, dt), O.qg2_expL(LV, LD, LV1, 2 * dt)
        Qn = O.qg2_update(qk, B, kx_, ky_)
        if step == 1:
            dq = dt * Qn
        elif step == 2:
            dq = dt / 2 * (3 * Qn - O.mmult3(E1, Qm[0]))
        else:
            dq = dt / 12 * (23 * Qn - 16 * O.mmult3(E1, Qm[0]) + 5 * O.mmult3(E2, Qm[1]))
        Qm = [Qn, Qm[0]]
        qk = O.mmult3(E1, qk + dq)
        qg.step(dt)
        scale = np.abs(qk).max()
        for layer in range(2):
            assert np.abs(qg.get(layer) - O.symmetrise_ky0(qk[:, :, layer])).max() / scale < 1e-12 or \
                   np.abs(qg.get(layer) - qk[:, :, layer]).max() / scale < 1e-12, (step, layer)
    qg.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["lagrange", "spectral"])
def test_qg2layersw_raytrace_script(mode, tmp_path):
    import swraytracing_b200 as S
    from swraytracing_b200 import drivers, fieldio
    ref = O.qg2layersw_driver(32, 12, 2, 30.0, 0.0, 0.3, 3.0, 1.0, max_steps=5, k_max=6, eval_mode=mode)
    got = drivers.qg2layersw_raytrace(32, 12, 2, 30.0, 0.0, 0.3, 3.0, 1.0, outdir=str(tmp_path), max_steps=5, k_max=6,
                                      mode=S.MODE_LAGRANGE6 if mode == "lagrange" else S.MODE_SPECTRAL, log=lambda s: None)
    assert got["steps"] == ref["steps"] and abs(got["t"] - ref["t"]) < 1e-13 and abs(got["dt"] - ref["dt"]) < 1e-14
    for key in ("packet_steps", "ode23_steps", "ode23_failed", "dt_changes"):
        assert got[key] == ref[key], key
    for g, r in zip(got["packets"], ref["packets"]):
        assert _relerr(g, r) < 1e-9
    scale = np.abs(ref["qk"]).max()
    for layer in range(2):
        assert np.abs(got["qk"][layer] - ref["qk"][:, :, layer]).max() / scale < 1e-10
    t, x, k = fieldio.load_packet_frames(tmp_path, 12)
    assert x.shape == (12, 2, got["packet_frames"]) and np.all(np.abs(x) <= 10.0)


@pytest.mark.gpu
def test_load_data_reads_back_driver_output(tmp_path):
    """analysis/load_data.m on the files the device driver wrote: log header -> parse_data, packet frames -> omega,
    window histograms bit-exact against the restated histcounts, energy = center .* counts"""
    from swraytracing_b200 import drivers
    lines = []
    out = drivers.qgsw_raytrace(32, 40, 2, 1.0, 0.0, 0.3, 3.0, 1.0, outdir=str(tmp_path), max_steps=40, r_drag=0.01,
                                integrator="leapfrog", log=lines.append)
    (tmp_path / "run.log").write_text("\n".join(["banner"] * 10 + lines))
    assert out["packet_frames"] >= 5
    ld = drivers.load_data(tmp_path, times=[3, 5], offset=2)
    assert (ld["nx"], ld["Npackets"], ld["f"], ld["Cg"], ld["Ug"]) == (32, 40, 3.0, 1.0, 0.3)
    assert ld["omega"].shape == (40, out["packet_frames"]) and ld["edges"].size == 300 and ld["edges"][-1] == ld["omega"].max()
    assert np.allclose(ld["omega"][:, 0], 6.0)                       # initial ring: omega0 = near_inertial_factor * f
    for win in ld["windows"]:
        lo, hi = win["frames"]
        ref = O.histcounts(ld["omega"][:, lo - 1:hi].ravel(order="F"), ld["edges"])
        assert np.array_equal(win["distribution"], ref) and int(ref.sum()) == 40 * (hi - lo + 1)
        assert np.array_equal(win["energy"], ld["center"] * ref)
    assert ld["mean_omega"].shape == (out["packet_frames"],) and abs(ld["mean_omega"][0] - 2.0) < 1e-12


@pytest.mark.gpu
def test_production_driver_in_nufft_mode_matches_dense_spectral(tmp_path):
    """qgsw_raytrace end to end (device QG frames -> flow slots -> per-step ode23) with the NUFFT evaluation instead of
    the dense contraction: same accept/reject decisions, packets equal to 1e-9"""
    import swraytracing_b200 as S
    from swraytracing_b200 import drivers
    outs = []
    for mode, sub in ((S.MODE_SPECTRAL, "a"), (S.MODE_NUFFT, "b")):
        outs.append(drivers.qgsw_raytrace(32, 64, 2, 1.0, 0.0, 0.3, 3.0, 1.0, outdir=str(tmp_path / sub), max_steps=12, r_drag=0.01,
                                          mode=mode, log=lambda s: None))
    a, b = outs
    assert (a["ode23_steps"], a["ode23_failed"], a["packet_steps"]) == (b["ode23_steps"], b["ode23_failed"], b["packet_steps"])
    for pa, pb in zip(a["packets"], b["packets"]):
        assert np.abs(pa - pb).max() < 1e-9


@pytest.mark.gpu
def test_qg_get_grid_is_k2g_of_the_state():
    import swraytracing_b200 as S
    nx = 32
    q = _pv_frame(nx)
    qk = O.g2k(q)
    qg = S.QGFlow(nx, L, qk, 3.0, 0.01, 3.0, 1.0, r_drag=0.0, force_strength=0.0)
    qg.step(3)
    assert np.abs(qg.get_grid() - O.k2g(qg.get())).max() < 1e-13
    qg.close()


@pytest.mark.gpu
@pytest.mark.parametrize("nx", [12, 20, 36, 50])
def test_qg_solvers_at_grid_sizes_that_are_not_powers_of_two(nx):
    """both frame producers against their restatements where cuFFT runs mixed-radix transforms"""
    import swraytracing_b200 as S
    rs = np.random.RandomState(nx)
    Ld = 2 * np.pi
    kx_, ky_ = O.wavenumbers(nx)
    amp = 1.0 / (1 + kx_ ** 2 + ky_ ** 2)
    qk = O.symmetrise_ky0((rs.randn(nx - 1, nx // 2) + 1j * rs.randn(nx - 1, nx // 2)) * amp)
    dt = 0.02
    qg = S.QGFlow(nx, Ld, qk, 3.0, dt, 3.0, 1.0, beta=0.3, r_drag=0.01, force_strength=0.05)
    qg.step(6)
    ref = O.qg_run(qk, 6, dt, nx, Ld, 3.0, 3.0, 1.0, beta=0.3, r_drag=0.01, force_strength=0.05)
    got = qg.get()
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-12
    qg.close()
    # two-layer, L = 20
    Lq = 20.0
    kap = 2 * np.pi / Lq
    kxs, kys = kx_ * kap, ky_ * kap
    nu = 0.1 * (Lq / nx) ** 8
    B, _, LV, LD, LV1 = O.qg2_operators(kxs, kys, 3.0, 0.0, 0.5, 0.4, nu, 4)
    q2 = np.stack([qk, -0.7 * qk], axis=2)
    g2 = S.QG2Flow(nx, Lq, q2[:, :, 0], q2[:, :, 1], 3.0, 0.0, 0.5, 0.4, nu, 4)
    Qm = [np.zeros_like(q2), np.zeros_like(q2)]
    for step in range(1, 5):
        E1, E2 = O.qg2_expL(LV, LD, LV1, dt), O.qg2_expL(LV, LD, LV1, 2 * dt)
        Qn = O.qg2_update(q2, B, kxs, kys)
        dq = dt * Qn if step == 1 else (dt / 2 * (3 * Qn - O.mmult3(E1, Qm[0])) if step == 2 else
                                        dt / 12 * (23 * Qn - 16 * O.mmult3(E1, Qm[0]) + 5 * O.mmult3(E2, Qm[1])))
        Qm = [Qn, Qm[0]]
        q2 = O.mmult3(E1, q2 + dq)
        g2.step(dt)
    for layer in range(2):
        assert np.abs(g2.get(layer) - q2[:, :, layer]).max() / np.abs(q2).max() < 1e-12
    g2.close()
