"""Known-answer tests that pin the CPU oracle.

The reference ships no golden vectors and cannot run here, so the oracle is pinned against
closed-form answers derived from the reference's own scripts."""
import numpy as np
import pytest

from oracle import swrt_oracle as O


def test_g2k_k2g_roundtrip(small_flow):
    # KAT 5: g2k(k2g(fk)) == fk after the ky=0 symmetrisation of fulspec.m:16
    fk = small_flow["planes"][0]
    back = O.g2k(O.k2g(fk))
    assert np.abs(back - O.symmetrise_ky0(fk)).max() < 1e-14


def test_fulspec_is_hermitian(small_flow):
    full = O.fulspec(small_flow["psik"])
    nx = small_flow["nx"]
    # after ifftshift, F(-k) == conj F(k)
    f = np.fft.ifftshift(full)
    idx = (-np.arange(nx)) % nx
    d = f[np.ix_(idx, idx)] - np.conj(f)
    d[0, 0] = 0      # fulspec does not touch Im F(0,0) (it only shifts the mean; k2g takes the real part)
    assert np.abs(d).max() < 1e-15
    assert np.all(full[0, :] == 0) and np.all(full[:, 0] == 0)   # Nyquist row/column zero


def test_grid_node_identity(small_flow):
    # KAT 1: at grid nodes both evaluators return the gridded value
    nx, dx = small_flow["nx"], small_flow["dx"]
    ix, iy = np.meshgrid(np.arange(nx), np.arange(nx), indexing="ij")
    x = (ix * dx).ravel(); y = (iy * dx).ravel()
    for fk, fg in zip(small_flow["planes"][:3], small_flow["grids"][:3]):
        scale = np.abs(fg).max()
        assert np.abs(O.interpolate(x, y, fg, dx, dx) - fg.ravel()).max() / scale < 1e-12
        assert np.abs(O.spectral_eval(x, y, fk, dx, nx).astype(float) - fg.ravel()).max() / scale < 1e-14


def test_periodic_wrap_and_negative_positions(small_flow):
    nx, dx, L = small_flow["nx"], small_flow["dx"], small_flow["L"]
    rs = np.random.RandomState(0)
    x = rs.uniform(0, L, 50); y = rs.uniform(0, L, 50)
    fg, fk = small_flow["grids"][0], small_flow["planes"][0]
    a = O.interpolate(x, y, fg, dx, dx)
    b = O.interpolate(x - 3 * L, y + 7 * L, fg, dx, dx)
    assert np.abs(a - b).max() < 1e-12
    a = O.spectral_eval(x, y, fk, dx, nx); b = O.spectral_eval(x - 3 * L, y + 7 * L, fk, dx, nx)
    assert np.abs((a - b).astype(float)).max() < 1e-12


def test_matlab_mod_edge():
    # mod(-tiny, nx) rounds to nx (then i0 = nx+1, ax = 0, indices wrap): SURVEY Appendix A
    assert O.matlab_mod(-1e-20, 256.0) == 256.0
    F = np.arange(16.0).reshape(4, 4)
    v = O.interpolate(np.array([-1e-20]), np.array([0.0]), F, 1.0, 1.0)
    assert abs(v[0] - F[0, 0]) < 1e-11


def test_lagrange_weights_partition_of_unity():
    a = np.linspace(0, 1, 11)
    # interpolate.m:37 divides by (j - i), not (i - j): with five factors every 1-D weight is the
    # NEGATIVE of the Lagrange basis polynomial; the 2-D product wx*wy restores the sign.
    w = -O._lagrange_weights(a, 0.0)
    assert np.abs(w.sum(axis=0) - 1).max() < 1e-14
    # reproduces polynomials up to degree 5 exactly
    nodes = np.arange(-2, 4)[:, None]
    for p in range(6):
        assert np.abs((w * nodes ** p).sum(axis=0) - a ** p).max() < 1e-12


def test_childress_soward_closed_form():
    # KAT 2: analytic single-mode flow of raytrace.m:31-37
    nx = 64; L = 2 * np.pi; U0 = 0.1; km = 4; a = 0.25
    psi, U, G = O.childress_soward(nx, L, U0, km, a)
    rs = np.random.RandomState(1)
    x = rs.uniform(0, L, 200); y = rs.uniform(0, L, 200)
    exact = O.childress_soward_point(x, y, U0, km, a)
    fields = [U["u"], U["v"], G["u_x"], G["u_y"], G["v_x"], G["v_y"]]
    for F, ex in zip(fields, exact):
        sp = O.spectral_eval(x, y, O.g2k(F), L / nx, nx).astype(float)
        assert np.abs(sp - ex).max() < 1e-14           # exactly representable by modes (+-km, +-km)
        lg = O.interpolate(x, y, F, L / nx, L / nx)
        # degree-5 Lagrange bound (3.52/720) (k dx)^6 |F| per direction
        bound = 2 * (3.52 / 720) * (km * L / nx) ** 6 * np.abs(F).max() * 1.5
        assert np.abs(lg - ex).max() < bound
    # SpectralScheme built from psi reproduces u = -psi_y, v = psi_x
    sch = O.SpectralScheme(L, nx, psi, mode="spectral")
    Uo = sch.U(np.stack([x, y], axis=1))
    assert np.abs(Uo[:, 0] - exact[0]).max() < 1e-14 and np.abs(Uo[:, 1] - exact[1]).max() < 1e-14
    g = sch.grad_U(np.stack([x, y], axis=1))
    for n, ex in zip(("u_x", "u_y", "v_x", "v_y"), exact[2:]):
        assert np.abs(g[n] - ex).max() < 1e-13


def test_zero_flow_analytic_trajectory():
    # KAT 3 (config C1): x(t) = x0 + gH k/omega t, k constant
    f, gH, dt, n = 3.0, 1.0, 0.01, 100
    rs = np.random.RandomState(2)
    x0 = rs.uniform(-3, 3, 20); y0 = rs.uniform(-3, 3, 20)
    k = 3 * np.cos(np.arange(20.0)); l = 3 * np.sin(np.arange(20.0))
    zero = lambda xx, yy: np.zeros((6,) + xx.shape)
    x, y, kk, ll = x0.copy(), y0.copy(), k.copy(), l.copy()
    for _ in range(n):
        x, y, kk, ll = O.leapfrog_step(x, y, kk, ll, dt, f, gH, zero)
    w = np.sqrt(f * f + gH * (k * k + l * l))
    assert np.array_equal(kk, k) and np.array_equal(ll, l)
    assert np.abs(x - (x0 + gH * k / w * n * dt)).max() < 1e-13
    assert np.abs(y - (y0 + gH * l / w * n * dt)).max() < 1e-13


def test_direct_sum_matches_fourier_interpolate_test_pattern():
    # KAT 4: Velocity() of scratch/fourier_interpolate_test.m:125-136 (real amp/phase modes) equals
    # the half-plane complex evaluation of the same streamfunction
    nmode = 3; N = 2 * nmode + 1; nx = 32; L = 2 * np.pi
    rs = O.matlab_rand_stream(44)
    amp = 0.5 * rs.rand(N, N).T / N ** 2
    phase = 2 * np.pi * rs.rand(N, N).T
    xg = np.arange(nx) * L / nx
    X, Y = np.meshgrid(xg, xg, indexing="ij")
    psi = np.zeros_like(X)
    for k in range(-nmode, nmode + 1):
        for l in range(-nmode, nmode + 1):
            psi += amp[k + nmode, l + nmode] * np.cos(k * X + l * Y + phase[k + nmode, l + nmode])
    x = rs.uniform(-5, 5, 64); y = rs.uniform(-5, 5, 64)
    u = np.zeros_like(x); v = np.zeros_like(x)
    for k in range(-nmode, nmode + 1):
        for l in range(-nmode, nmode + 1):
            s = -np.sin(k * x + l * y + phase[k + nmode, l + nmode])
            u += -l * amp[k + nmode, l + nmode] * s
            v += k * amp[k + nmode, l + nmode] * s
    sch = O.SpectralScheme(L, nx, psi, mode="spectral")
    Uo = sch.U(np.stack([x, y], axis=1))
    assert np.abs(Uo[:, 0] - u).max() < 1e-14 and np.abs(Uo[:, 1] - v).max() < 1e-14


def test_omega_drift_bound_steady_flow():
    # KAT 6: leapfrog keeps Omega = omega + U.k bounded on a steady flow
    # (images/Symplectic_error: |dOmega/Omega0| <~ 5e-3 at dt = 0.01)
    nx = 32; L = 2 * np.pi; f = 3.0; gH = 1.0
    psi, U, G = O.childress_soward(nx, L, 0.2, 2, 0.25)
    planes = [O.g2k(F) for F in (U["u"], U["v"], G["u_x"], G["u_y"], G["v_x"], G["v_y"])]
    ev = lambda xx, yy: O.childress_soward_point(xx, yy, 0.2, 2, 0.25)
    rs = np.random.RandomState(3)
    n = 16
    x = rs.uniform(0, L, n); y = rs.uniform(0, L, n)
    k = 3 * np.cos(2 * np.pi * np.arange(n) / n); l = 3 * np.sin(2 * np.pi * np.arange(n) / n)
    def Om(x, y, k, l):
        e = ev(x, y)
        return np.sqrt(f * f + gH * (k * k + l * l)) + e[0] * k + e[1] * l
    O0 = Om(x, y, k, l)
    worst = 0.0
    for i in range(400):
        x, y, k, l = O.leapfrog_step(x, y, k, l, 0.01, f, gH, ev)
        worst = max(worst, np.abs((Om(x, y, k, l) - O0) / O0).max())
    assert worst < 5e-3


def test_histcounts_rule():
    edges = O.matlab_linspace(0, 10, 11)
    w = np.array([0.0, 0.999, 1.0, 9.999, 10.0, 10.0001, -0.1, np.nan, 5.5])
    c = O.histcounts(w, edges)
    assert c.tolist() == [2, 1, 0, 0, 0, 1, 0, 0, 0, 2]    # last bin closed; out-of-range/NaN dropped
    assert edges[-1] == 10.0 and len(edges) == 11


def test_ode_symplectic_shapes_and_row_convention(small_flow):
    nx, L = small_flow["nx"], small_flow["L"]
    psi = O.k2g(small_flow["psik"])
    sch = O.SpectralScheme(L, nx, psi, mode="lagrange")
    n = 5
    x0 = np.zeros((1, 2, n)); k0 = np.zeros((1, 2, n))
    x0[0, 0] = np.linspace(-1, 1, n); x0[0, 1] = np.linspace(1, 2, n)
    k0[0, 0] = 3.0
    dt = 0.01
    xs, ks, ts = O.ode_symplectic(x0, k0, dt, 0.055, 3.0, 1.0, sch)
    assert xs.shape == (5, 2, n) and ts.shape == (5,)        # Nsteps = floor(T/dt) rows, row 0 = initial
    assert np.array_equal(xs[0], x0[0]) and np.allclose(ts, np.arange(5) * dt)


def test_initial_q_always_true_bug_and_packet_init():
    # qgsw_raytrace.m:202 sums every |k|,|l| <= k_max mode; ring=True is the intended annulus
    nx = 32; L = 2 * np.pi
    xg = O.matlab_linspace(-L / 2, L / 2, nx)
    X, Y = np.meshgrid(xg, xg)
    q_all = O.initial_q(X, Y, 0.5, 3.0, O.matlab_rand_stream(146), k_min=2, k_max=3)
    q_ring = O.initial_q(X, Y, 0.5, 3.0, O.matlab_rand_stream(146), k_min=2, k_max=3, ring=True)
    assert np.abs(q_all - q_ring).max() > 1e-3
    x, y, k, l = O.init_packets(50, L, 3.0, O.matlab_rand_stream(123))
    assert np.allclose(k ** 2 + l ** 2, 9.0) and x.min() >= -L / 2 and x.max() < L / 2
    # first MATLAB rand after rng(123) is 0.696469185597861
    assert abs((x[0] + L / 2) / L - 0.696469185597861) < 1e-15


def test_step_packet_matches_batch_restatement(small_flow):
    nx, L, dx = small_flow["nx"], small_flow["L"], small_flow["dx"]
    g = small_flow["grids"]
    U = {"u": g[0], "v": g[1]}; G = {"u_x": g[2], "u_y": g[3], "v_x": g[4], "v_y": g[5]}
    H = 1 + 0.1 * O.k2g(small_flow["psik"])
    fields = {"u": g[0], "v": g[1], "u_x": g[2], "u_y": g[3], "v_x": g[4], "v_y": g[5], "H": H}
    P = {"x": 0.3, "y": -1.1, "k": 2.0, "l": -1.5, "a": 1.0}
    arr = lambda v: np.array([v])
    for xka in (False, True):
        lit = O.step_packet_xka(P, U, G, H, 1.0, 3.0, dx, dx, 0.02) if xka else O.step_packet(P, U, G, 1.0, 3.0, dx, dx, 0.02)
        bat = O.rk4_step_batch(arr(P["x"]), arr(P["y"]), arr(P["k"]), arr(P["l"]), arr(P["a"]), 0.02, 1.0, 3.0, fields, dx, xka)
        for i, name in enumerate(["x", "y", "k", "l"] + (["a"] if xka else [])):
            assert abs(lit[name] - bat[i][0]) < 1e-15


def test_ode23_restatement_known_answers():
    # Bogacki-Shampine 3(2) as MATLAB's ode23 drives it (parity unpinned: MATLAB is absent).  Pins what can be
    # pinned: third-order convergence of the propagated solution, the MaxStep = 0.1*(tf-t0) default (>= 10
    # steps), FSAL function-evaluation count 1 + 3*(accepted + rejected), exactness on cubics' derivatives.
    y, st = O.ode23(lambda t, y: -y, [0, 2.0], np.array([1.0]))
    assert st["nsteps"] >= 10 and st["nfevals"] == 1 + 3 * (st["nsteps"] + st["nfailed"])
    assert abs(y[0] - np.exp(-2.0)) < 1e-3
    errs = []
    for rtol in (1e-4, 1e-6, 1e-8):
        y, st = O.ode23(lambda t, y: np.array([y[1], -y[0]]), [0, 3.0], np.array([1.0, 0.0]), rtol=rtol, atol=rtol * 1e-3)
        errs.append(np.abs(y - np.array([np.cos(3.0), -np.sin(3.0)])).max())
    assert errs[0] > errs[1] > errs[2] and errs[2] < 1e-6
    # y' = 3 t^2 is integrated exactly by the third-order formula
    y, st = O.ode23(lambda t, y: np.array([3 * t * t]), [0, 1.0], np.array([0.0]))
    assert abs(y[0] - 1.0) < 1e-13
    # zero-flow packets: x advances with the constant group velocity, k is untouched
    n = 5; f, Cg = 3.0, 1.0
    zero = lambda xx, yy, al: np.zeros((6, xx.size))
    ode = O.generate_raytracing_ode(None, None, n, f, Cg, 1.0, 0.1, eval6=zero)
    y0 = np.concatenate([np.zeros(n), np.zeros(n), 3 * np.ones(n), 4 * np.ones(n)])
    y, st = O.ode23(ode, [0, 1.0], y0)
    w = np.sqrt(f * f + Cg * Cg * 25.0)
    assert np.allclose(y[:n], Cg * 3 / w, atol=1e-14) and np.allclose(y[2 * n:3 * n], 3.0)


def test_difference_scheme_cross_checks_spectral_scheme():
    """DifferenceScheme.m (finite differences of the analytic streamfunction, h = eps^(1/3)) and SpectralScheme.m
    (spectral derivatives of the gridded streamfunction) are the reference's two implementations of the same
    RaytracingScheme interface; on the Childress-Soward flow of raytrace.m:30 they must agree to the
    finite-difference error: ~eps/h = 4e-11 for U, ~eps/h^2 = 6e-6 (x |psi|) for grad U."""
    nx, L, U0, km, a = 64, 2 * np.pi, 0.1, 4.0, 0.25
    psi, _, _ = O.childress_soward(nx, L, U0, km, a)
    sch = O.SpectralScheme(L, nx, psi, mode="spectral")
    ds = O.DifferenceScheme(lambda x, y, t: U0 / km * (np.sin(km * x) * np.sin(km * y) + a * np.cos(km * x) * np.cos(km * y)))
    assert abs(ds.h - 6.055454452393343e-06) < 1e-20
    x = np.random.RandomState(0).uniform(-5, 5, (200, 2))
    assert np.abs(sch.U(x) - ds.U(x)).max() < 1e-10
    gs, gd = sch.grad_U(x), ds.grad_U(x)
    for n in ("u_x", "u_y", "v_x", "v_y"):
        assert np.abs(gs[n] - gd[n]).max() < 2e-6, n
    assert np.abs(gs["u_x"] + gs["v_y"]).max() < 1e-14 and np.array_equal(gd["u_x"], -gd["v_y"])
    # inherited diagnostics (RaytracingScheme.m:18-31) against the closed form: vorticity = -2 km^2 psi for this flow
    zeta = -2 * km ** 2 * ds.streamfunction(x[:, 0], x[:, 1])
    assert np.abs(sch.vorticity(x) - zeta).max() < 1e-13
    assert np.all(sch.strain(x) >= 0) and np.abs(sch.okuboWeiss(x) - (gs["v_y"] ** 2 + gs["v_x"] * gs["u_y"])).max() == 0
    # (T,2,Np) calling shape of ode_symplectic's histories
    x3 = np.transpose(x.reshape(4, 50, 2), (0, 2, 1))
    assert np.abs(sch.U(x3) - ds.U(x3)).max() < 1e-10


def test_interpolate_is_degree5_tensor_lagrange_interpolation_independent_check():
    """interpolate.m's weight products against an independent construction: scipy's Lagrange polynomial through the six
    nodes -2..3 of each axis, evaluated at the fractional position (the 1e-13 bump moves the abscissa by 1e-13)."""
    import warnings
    from scipy.interpolate import lagrange
    warnings.simplefilter("ignore")              # scipy.lagrange warns about its own conditioning for degree > ~20; degree 5 is fine
    rs = np.random.RandomState(4)
    nx = 16; dx = 2 * np.pi / nx
    F = rs.randn(nx, nx)
    x = rs.uniform(-20, 20, 40); y = rs.uniform(-20, 20, 40)
    got = O.interpolate(x, y, F, dx, dx)
    nodes = np.arange(-2, 4)
    for m in range(x.size):
        xl = np.mod(x[m] / dx, nx); yl = np.mod(y[m] / dx, nx)
        i0 = int(np.floor(xl)); j0 = int(np.floor(yl)); ax = xl - i0; ay = yl - j0
        block = F[np.ix_((i0 + nodes) % nx, (j0 + nodes) % nx)]
        col = np.array([lagrange(nodes, block[i, :])(ay) for i in range(6)])      # interpolate along y for each x node
        want = lagrange(nodes, col)(ax)
        assert abs(got[m] - want) < 1e-10 * max(1.0, np.abs(block).max()), (m, got[m], want)
    # the 1-D weights are minus the Lagrange basis (five negative denominators' worth of sign); the 2-D product restores it
    w = O._lagrange_weights(np.array([0.3]), 0.0)[:, 0]
    basis = np.array([lagrange(nodes, np.eye(6)[i])(0.3) for i in range(6)])
    assert np.allclose(w, -basis, atol=1e-14) and abs(w.sum() + 1.0) < 1e-14


def test_step_packet_xka_uniform_depth_zero_flow_closed_form():
    """step_packet_xka.m:38-91 + cg_sw.m:15-31 at rest over a uniform depth H0: the group velocity C = C0^2 H0 k/omega is
    constant, so RK4 in x is exact (x += dt*C), grad(omega) = 0 leaves k unchanged, and div C = -(Cx^2 + Cy^2)/omega is a
    constant z/dt, so the RK4 of da/dt = -a divC multiplies a by the degree-4 Taylor polynomial of exp(-divC dt)"""
    nx = 16; L = 2 * np.pi; dx = L / nx
    zero = np.zeros((nx, nx)); H0 = 1.3
    U = {"u": zero, "v": zero}; G = {"u_x": zero, "u_y": zero, "v_x": zero, "v_y": zero}
    H = np.full((nx, nx), H0)
    C0, f, dt = 0.9, 3.0, 0.07
    P = {"x": 0.4, "y": -2.0, "k": 2.5, "l": -1.25, "a": 0.8}
    out = O.step_packet_xka(P, U, G, H, C0, f, dx, dx, dt)
    gH = C0 ** 2 * H0
    om = np.sqrt(f ** 2 + gH * (P["k"] ** 2 + P["l"] ** 2))
    Cx, Cy = gH * P["k"] / om, gH * P["l"] / om
    z = (Cx ** 2 + Cy ** 2) / om * dt                       # -divC*dt
    assert abs(out["x"] - (P["x"] + dt * Cx)) < 1e-13 and abs(out["y"] - (P["y"] + dt * Cy)) < 1e-13
    assert out["k"] == P["k"] and out["l"] == P["l"]
    assert abs(out["a"] - P["a"] * (1 + z + z ** 2 / 2 + z ** 3 / 6 + z ** 4 / 24)) < 1e-14
    # and the batch restatement the GPU tests use agrees
    arr = lambda v: np.array([v])
    fields = {"u": zero, "v": zero, "u_x": zero, "u_y": zero, "v_x": zero, "v_y": zero, "H": H}
    bat = O.rk4_step_batch(arr(P["x"]), arr(P["y"]), arr(P["k"]), arr(P["l"]), arr(P["a"]), dt, C0, f, fields, dx, True)
    for i, name in enumerate(["x", "y", "k", "l", "a"]):
        assert abs(out[name] - bat[i][0]) < 1e-14


def test_ode23_initial_step_is_bounded_by_the_first_output_interval():
    """MATLAB's ode23 starts from absh = min(hmax, htspan, 1/rh) with htspan = |tspan(2) - tspan(1)| (odearguments): with the
    dense ``tspan = dt*(0:Nsteps)`` of SW_zero_background_raytracing.m:73-78 the first step is the first output interval
    whenever that is shorter than 1/rh.  y' = -y, y0 = 1: rh = 1/(0.8*rtol^(1/3)) -> 1/rh = 0.08 at RelTol 1e-3."""
    calls = []

    def f(t, y):
        calls.append(t)
        return -y
    y0 = np.array([1.0])
    O.ode23(f, [0.0, 1.0], y0)
    assert abs(calls[1] - 0.5 * 0.08) < 1e-15                 # two-point span: 1/rh decides (hmax = 0.1, span = 1)
    calls.clear()
    out = O.ode23(f, np.arange(0.0, 1.0 + 1e-12, 0.01), y0)
    assert abs(calls[1] - 0.5 * 0.01) < 1e-15                 # dense span: the first interval (0.01) decides
    Y = out[0] if isinstance(out, tuple) else out
    Y = np.asarray(Y["Y"] if isinstance(Y, dict) else Y)
    assert np.abs(Y[:, 0] - np.exp(-np.arange(0.0, 1.0 + 1e-12, 0.01))).max() < 1e-3
