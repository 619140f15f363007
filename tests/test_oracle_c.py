"""The C restatement (oracle/swrt_oracle.c) against the numpy restatement and the golden fixture."""
from pathlib import Path

import numpy as np
import pytest

from oracle import swrt_oracle as O
from oracle import c_oracle as CO

GOLD = np.load(Path(__file__).parent / "golden" / "hotpath_nx32.npz")


def _planes(psik, nx):
    kx_, ky_ = O.wavenumbers(nx)
    return O.velocity_planes_k(psik, kx_, ky_)


def test_interpolate_bit_exact(small_flow, packets):
    dx = small_flow["dx"]
    for g in small_flow["grids"][:2]:
        a = CO.interpolate(packets["x"], packets["y"], g, dx, dx)
        b = O.interpolate(packets["x"], packets["y"], g, dx, dx)
        assert np.array_equal(a, b)          # same operation order -> same doubles
    a = CO.interpolate(packets["x"], packets["y"], small_flow["grids"][0], dx, dx, bump=1e-10)
    assert np.array_equal(a, O.interpolate_par(packets["x"], packets["y"], small_flow["grids"][0], dx, dx))


def test_interpolate_empty_and_single():
    F = np.arange(64.0).reshape(8, 8)
    assert CO.interpolate(np.zeros(0), np.zeros(0), F, 1.0, 1.0).size == 0
    v = CO.interpolate(np.array([2.0]), np.array([3.0]), F, 1.0, 1.0)
    assert abs(v[0] - F[2, 3]) < 1e-10


def test_spectral_eval_c_vs_numpy(small_flow, packets):
    nx, dx = small_flow["nx"], small_flow["dx"]
    ref = O.spectral_eval_planes(packets["x"], packets["y"], small_flow["planes"], dx, nx)
    got = CO.spectral_eval(packets["x"], packets["y"], small_flow["planes"], dx, nx, precise=True)
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert (np.abs(got - ref) / scale).max() < 1e-15
    fast = CO.spectral_eval(packets["x"], packets["y"], small_flow["planes"], dx, nx, precise=False)
    assert (np.abs(fast - ref) / scale).max() < 1e-13


def test_golden_eval_and_interpU():
    nx = int(GOLD["nx"]); L = float(GOLD["L"]); dx = L / nx
    planes = _planes(GOLD["psik"], nx)
    grids = list(GOLD["grids"])
    assert np.array_equal(CO.interpolate6(GOLD["x"], GOLD["y"], grids, dx), GOLD["eval_lagrange"])
    got = CO.spectral_eval(GOLD["x"], GOLD["y"], planes, dx, nx)
    assert np.abs(got - GOLD["eval_spectral"]).max() < 1e-15
    # the numpy oracle still reproduces its own committed vectors (guards the oracle itself)
    assert np.array_equal(np.stack([O.interpolate(GOLD["x"], GOLD["y"], g, dx, dx) for g in grids]), GOLD["eval_lagrange"])


def test_golden_leapfrog():
    nx = int(GOLD["nx"]); L = float(GOLD["L"]); dx = L / nx
    f, gH, dt = float(GOLD["f"]), float(GOLD["gH"]), float(GOLD["dt"])
    st = CO.leapfrog_lagrange(GOLD["x"], GOLD["y"], GOLD["k"], GOLD["l"], list(GOLD["grids"]), dx, f, gH, dt, 20)
    assert np.abs(np.stack(st) - GOLD["leapfrog20_lagrange"]).max() < 1e-13
    st = CO.leapfrog_spectral(GOLD["x"], GOLD["y"], GOLD["k"], GOLD["l"], _planes(GOLD["psik"], nx), dx, nx, f, gH, dt, 20)
    assert np.abs(np.stack(st) - GOLD["leapfrog20_spectral"]).max() < 1e-12


def test_golden_rk4():
    nx = int(GOLD["nx"]); L = float(GOLD["L"]); dx = L / nx
    f, dt = float(GOLD["f"]), float(GOLD["dt"])
    grids = list(GOLD["grids"]) + [GOLD["H"]]
    a = np.ones(GOLD["x"].size)
    for xka, key in ((False, "rk4x3_packet_lagrange"), (True, "rk4x3_xka_lagrange")):
        st = CO.rk4_lagrange(GOLD["x"], GOLD["y"], GOLD["k"], GOLD["l"], a, grids, dx, f, 1.0, dt, 3, xka)
        assert np.abs(np.stack(st) - GOLD[key]).max() < 1e-13


def test_golden_rhs_and_hist():
    f = float(GOLD["f"])
    got = CO.rhs(GOLD["k"], GOLD["l"], GOLD["interpU_lagrange"], f, 1.0)
    assert np.abs(np.stack(got) - GOLD["rhs_lagrange"]).max() < 1e-15
    w = O.omega_of_k(GOLD["k"], GOLD["l"], f, float(GOLD["gH"]))
    assert np.array_equal(CO.histcounts(w, GOLD["hist_edges"]), GOLD["hist_counts"])
    assert np.array_equal(O.histcounts(w, GOLD["hist_edges"]), GOLD["hist_counts"])


def test_histcounts_c_vs_numpy_random():
    rs = np.random.RandomState(9)
    w = np.concatenate([rs.uniform(-1, 11, 20000), [0.0, 10.0, np.nan, np.inf]])
    edges = O.matlab_linspace(0, 10, 300)
    assert np.array_equal(CO.histcounts(w, edges), O.histcounts(w, edges))


def test_linearity_of_time_blend(small_flow, packets):
    # interpolate_U.m:19-23: blending evaluated fields == evaluating blended coefficients
    nx, dx = small_flow["nx"], small_flow["dx"]
    p1 = small_flow["planes"]; p2 = [p * np.exp(0.03j) for p in p1]
    al = 0.37
    x, y = packets["x"][:100], packets["y"][:100]
    a = (1 - al) * CO.spectral_eval(x, y, p1, dx, nx) + al * CO.spectral_eval(x, y, p2, dx, nx)
    b = CO.spectral_eval(x, y, [(1 - al) * u + al * v for u, v in zip(p1, p2)], dx, nx)
    assert np.abs(a - b).max() < 1e-15


def test_two_frame_leapfrog_c_port_equals_numpy_restatement(small_flow, packets):
    """orc_leapfrog_lagrange2 (the CPU arm of the C3 / C4 bench lines) == interpolate_U.m + ode_symplectic.m restated in
    numpy, bit for bit: both frames interpolated, blended (1-alpha)*F1 + alpha*F2, alpha_j = alpha0 + j*dalpha"""
    dx = small_flow["dx"]
    names = ("u", "v", "ux", "uy", "vx", "vy")
    g1 = small_flow["grids"]
    rs = np.random.RandomState(2)
    g2 = [g + 0.05 * rs.standard_normal(g.shape) for g in g1]
    bf1, bf2 = dict(zip(names, g1)), dict(zip(names, g2))
    x, y, k, l = (packets[c].copy() for c in "xykl")
    f, gH, dt, m = 3.0, 1.0, 0.01, 5
    a0, da = 0.5 / m, 1.0 / m
    for j in range(m):
        al = a0 + j * da

        def eval6(xx, yy):
            U, nab = O.interpolate_U(bf1, bf2, al, np.stack([xx, yy], axis=1), dx, bump=O.BUMP_LIVE)
            return U[:, 0], U[:, 1], nab["u_x"], nab["u_y"], nab["v_x"], nab["v_y"]
        x, y, k, l = O.leapfrog_step(x, y, k, l, dt, f, gH, eval6)
    got = CO.leapfrog_lagrange2(packets["x"], packets["y"], packets["k"], packets["l"], g1, g2, dx, f, gH, dt, m, a0, da)
    assert np.array_equal(np.stack(got), np.stack([x, y, k, l]))
    # dalpha = 0, alpha0 = 0 reduces to the steady port
    st = CO.leapfrog_lagrange(packets["x"], packets["y"], packets["k"], packets["l"], g1, dx, f, gH, dt, m)
    st2 = CO.leapfrog_lagrange2(packets["x"], packets["y"], packets["k"], packets["l"], g1, g2, dx, f, gH, dt, m, 0.0, 0.0)
    assert np.allclose(np.stack(st), np.stack(st2), rtol=0, atol=1e-14)
