"""The MATLAB drop-in layer EXECUTED end to end: ``matlab/shims/*.m`` -> ``swrt_mex`` -> ``matlab/swrt_mex.c`` -> libswrt -> GPU.

Round 1 could only syntax-check the shims: there is no MATLAB / Octave in the image.  Here ``oracle/minimat`` (the interpreter
that also executes the unmodified reference, tests/test_minimat.py) runs the shim ``.m`` files themselves -- ``interpolate``,
``interpolate_par``, ``interpolate_U``, the ``SpectralScheme`` classdef, ``ode_symplectic``, ``step_packet``,
``step_packet_xka`` (one packet at a time, as raytrace.m calls it, and as a struct array), ``grid_U``, ``swrt_ode23`` -- with
the shim folder as the only thing on the path; their ``swrt_mex(...)`` calls go to the REAL gateway: ``mexFunction`` of the
unmodified ``matlab/swrt_mex.c`` linked with the mx / mex harness of ``tests/mex_harness`` (tests/test_mex_gateway.py) and
libswrt.so.  The calling script below makes the same calls, with the same arguments, as ``tests/golden/make_octave_goldens.m``
makes to the reference's functions, so its results are compared with what the unmodified reference returned
(``tests/golden/octave_out``, ``reference_locals_nx32.npz``): a reference script that puts ``matlab/shims`` in front of its path
gets the reference's numbers from the GPU.  Nothing here reads /root/reference.

CPU part: the shims parse, and fail loudly through the gateway when there is no device.
"""
import io
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle.minimat import Interp, Frame, MatlabError, from_py, MStruct      # noqa: E402
from oracle.minimat.parser import parse_source                                # noqa: E402
from test_mex_gateway import Mex, MexError                                    # noqa: E402

SHIMS = ROOT / "matlab" / "shims"
GOLD = ROOT / "tests" / "golden"
NAMES = ("u", "v", "ux", "uy", "vx", "vy")
TOL_FIELD, TOL_TRAJ = 1e-12, 1e-9


class Handle:
    """an opaque uint64 handle travelling through the interpreter's variables"""

    def __init__(self, value):
        self.value = np.uint64(value)


def bridge(mex):
    """swrt_mex(cmd, ...) of the shims -> mexFunction of the real gateway"""
    def conv(a):
        if isinstance(a, Handle):
            return a.value
        if isinstance(a, bool):
            return float(a)
        return a

    def back(v):
        if isinstance(v, np.uint64):
            return Handle(v)
        if isinstance(v, np.ndarray) and v.dtype == np.uint64:
            return from_py(v.astype(np.float64))
        return from_py(v)

    def swrt_mex(I, args, nargout, frame):
        try:
            out = mex(args[0], *[conv(a) for a in args[1:]], nlhs=nargout)
        except MexError as ex:
            raise MatlabError(f"{ex.ident}: {ex.msg}")
        if nargout <= 1:
            return [] if out is None else [back(out)]
        return [back(v) for v in out]
    return swrt_mex


def interp(mex):
    I = Interp(cwd=str(ROOT / "tests"), out=io.StringIO())
    I.path.insert(0, str(SHIMS))
    I.overrides["swrt_mex"] = bridge(mex)
    return I


def read_in(name, *shape):
    return np.fromfile(GOLD / "octave_in" / f"{name}.bin").reshape(shape, order="F")


def read_out(name, *shape):
    return np.fromfile(GOLD / "octave_out" / f"{name}.bin").reshape(shape, order="F")


def workspace():
    p = read_in("params", 8)
    nx, n = int(p[0]), int(p[7])
    fr = Frame(None)
    v = dict(nx=float(nx), L=p[1], f=p[2], gH=p[3], alpha=p[4], dt=p[5], C0=p[6], np_=float(n), dx=p[1] / nx)
    for nm in "xykl":
        v[nm] = read_in(nm, n, 1)
    for b in ("bf1", "bf2"):
        v[b] = {nm: read_in(f"{b}_{nm}", nx, nx) for nm in NAMES}
    v["H"] = read_in("H", nx, nx)
    v["psi"] = read_in("psi", nx, nx)
    for k, a in v.items():
        fr.vars[k] = from_py(a)
    return fr, nx, n


SCRIPT = r"""
names = {'u', 'v', 'ux', 'uy', 'vx', 'vy'};
E = zeros(6, np_); Epar = zeros(6, np_);
for c = 1:6
    E(c, :) = interpolate(x, y, bf1.(names{c}), dx, dx);                 % make_octave_goldens.m: 'eval_lagrange'
    Epar(c, :) = interpolate_par(x, y, bf1.(names{c}), dx, dx);          % 'eval_lagrange_qg' (the same stencil, bump 1e-10)
end
[U, nab] = interpolate_U(bf1, bf2, alpha, [x y], dx);                    % 'interpU_lagrange'
IU = [U(:, 1)'; U(:, 2)'; nab.u_x(:)'; nab.u_y(:)'; nab.v_x(:)'; nab.v_y(:)'];
[U0, nab0] = interpolate_U(bf1, bf2, alpha, [x y], dx);                  % same pair again: the cached device flow is reused
IU0 = [U0(:, 1)'; U0(:, 2)'; nab0.u_x(:)'; nab0.u_y(:)'; nab0.v_x(:)'; nab0.v_y(:)'];
bf3 = bf2; bf3.vy = bf3.vy + 1;                                           % one gradient plane changed, u and v untouched
[~, nab3] = interpolate_U(bf1, bf3, 1, [x y], dx);
dvy = nab3.v_y - interpolate_par(x, y, bf2.vy, dx, dx);                   % must see the new plane: +1 everywhere

scheme = SpectralScheme(L, nx, psi, 1);                                  % mode 1 = the reference's 6x6 Lagrange semantics
x3 = zeros(1, 2, np_); x3(1, 1, :) = x; x3(1, 2, :) = y;
k3 = zeros(1, 2, np_); k3(1, 1, :) = k; k3(1, 2, :) = l;
Us = scheme.U(x3, 0);
g = scheme.grad_U(x3, 0);
gk = scheme.grad_U_times_k(x3, k3, 0);
SE = [squeeze(Us(1, 1, :))'; squeeze(Us(1, 2, :))'; g.u_x(:)'; g.u_y(:)'; g.v_x(:)'; g.v_y(:)'];        % 'scheme_eval'
GK = [squeeze(gk(1, 1, :))'; squeeze(gk(1, 2, :))'];                                                    % 'scheme_gradU_times_k'
SF = cat(3, scheme.U_field.u, scheme.U_field.v, scheme.GradU_field.u_x, scheme.GradU_field.u_y, ...
         scheme.GradU_field.v_x, scheme.GradU_field.v_y);                                               % 'scheme_fields'
nst = 100;
[xs, ks, ts] = ode_symplectic(x3, k3, dt, (nst + 1.5) * dt, f, gH, scheme);                             % 'leapfrog*_scheme'
nrows = size(xs, 1);
LF100 = [squeeze(xs(end, 1, :))'; squeeze(xs(end, 2, :))'; squeeze(ks(end, 1, :))'; squeeze(ks(end, 2, :))'];
LF20 = [squeeze(xs(21, 1, :))'; squeeze(xs(21, 2, :))'; squeeze(ks(21, 1, :))'; squeeze(ks(21, 2, :))'];
[xt, kt, tt] = ode_symplectic(x3, k3, dt, (nst + 1.5) * dt, f, gH, scheme, 20);                         % thinned history
scheme.delete();

Uf.u = bf1.u; Uf.v = bf1.v;
Gf.u_x = bf1.ux; Gf.u_y = bf1.uy; Gf.v_x = bf1.vx; Gf.v_y = bf1.vy;
S1 = zeros(4, npk); S2 = zeros(5, npk);
for m = 1:npk                                                            % one packet at a time, as raytrace.m:51-55 calls it
    P.x = x(m); P.y = y(m); P.k = k(m); P.l = l(m);
    Q = P; Q.a = 1;
    for s = 1:3
        P = step_packet(P, Uf, Gf, C0, f, dx, dx, dt);
        Q = step_packet_xka(Q, Uf, Gf, H, C0, f, dx, dx, dt);
    end
    S1(:, m) = [P.x; P.y; P.k; P.l];
    S2(:, m) = [Q.x; Q.y; Q.k; Q.l; Q.a];
end
for m = 1:np_                                                            % all packets as ONE struct array per call
    PA(m).x = x(m); PA(m).y = y(m); PA(m).k = k(m); PA(m).l = l(m); PA(m).a = 1;
end
QA = PA;
for s = 1:3
    PA = step_packet(PA, Uf, Gf, C0, f, dx, dx, dt);
    QA = step_packet_xka(QA, Uf, Gf, H, C0, f, dx, dx, dt);
end
A1 = [[PA.x]; [PA.y]; [PA.k]; [PA.l]];
A2 = [[QA.x]; [QA.y]; [QA.k]; [QA.l]; [QA.a]];
"""


@pytest.mark.gpu
@pytest.mark.parametrize("interleaved", [False, True], ids=["separate-complex", "interleaved-complex"])
def test_shims_through_the_gateway_return_the_references_numbers(interleaved):
    mex = Mex(interleaved)
    I = interp(mex)
    fr, nx, n = workspace()
    npk = 24                                             # packets stepped one call at a time (1 gateway call per packet-step)
    fr.vars["npk"] = float(npk)
    I.run(SCRIPT, fr)
    v = fr.vars

    def scaled(got, ref):
        sc = np.abs(ref).reshape(ref.shape[0], -1).max(axis=1)
        return float((np.abs(got - ref).reshape(ref.shape[0], -1).max(axis=1) / sc).max())
    assert np.array_equal(v["E"], read_out("eval_lagrange", 6, n))                   # LAGRANGE6 = the reference's arithmetic, bit for bit
    assert np.array_equal(v["Epar"], read_out("eval_lagrange_qg", 6, n))
    assert scaled(v["IU"], read_out("interpU_lagrange", 6, n)) <= TOL_FIELD
    assert np.array_equal(v["IU0"], v["IU"])
    assert np.abs(np.asarray(v["dvy"]) - 1.0).max() < 1e-12                          # a changed gradient plane is never served stale
    assert scaled(v["SE"], read_out("scheme_eval", 6, n)) <= TOL_FIELD
    assert scaled(v["GK"], read_out("scheme_gradU_times_k", 2, n)) <= TOL_FIELD
    assert scaled(np.moveaxis(v["SF"], 2, 0), np.moveaxis(read_out("scheme_fields", nx, nx, 6), 2, 0)) <= 1e-14
    assert v["nrows"] == 101.0 and np.array_equal(np.asarray(v["ts"]).ravel(), read_out("leapfrog_t", 101))
    assert np.abs(v["LF20"] - read_out("leapfrog20_scheme", 4, n)).max() <= TOL_TRAJ
    assert np.abs(v["LF100"] - read_out("leapfrog100_scheme", 4, n)).max() <= TOL_TRAJ
    assert v["xt"].shape == (6, 2, n) and np.array_equal(np.asarray(v["tt"]).ravel(), np.asarray(v["ts"]).ravel()[::20])
    assert np.array_equal(v["xt"][1], v["xs"][20]) and np.array_equal(v["kt"][5], v["ks"][100])   # fused 20-step launches == single steps
    r1, r2 = read_out("rk4x3_packet_lagrange", 4, n), read_out("rk4x3_xka_lagrange", 5, n)
    assert np.array_equal(v["S1"], r1[:, :npk]) and np.array_equal(v["S2"], r2[:, :npk])
    assert np.array_equal(v["A1"], r1) and np.array_equal(v["A2"], r2)


@pytest.mark.gpu
def test_grid_U_and_ode23_shims_through_the_gateway():
    """grid_U.m shim against the reference's six-argument grid_U (executed, reference_locals_nx32.npz); swrt_ode23.m -- the
    host controller written in MATLAB -- against the Python controller over the same device stages (identical decisions)"""
    import swraytracing_b200 as S
    from swraytracing_b200 import reference_api as A
    from oracle import swrt_oracle as O
    R = np.load(GOLD / "reference_locals_nx32.npz")
    mex = Mex(False)
    I = interp(mex)
    fr, nx, n = workspace()
    kx_, ky_ = O.wavenumbers(nx)
    for k_, a in dict(qk=R["update_qk_in"], K2=kx_ ** 2 + ky_ ** 2, kx_=kx_, ky_=ky_, shear=float(R["grid_U_shear_value"])).items():
        fr.vars[k_] = from_py(np.asfortranarray(a) if isinstance(a, np.ndarray) else a)
    I.run(r"""
        flow = grid_U(qk, 3, K2, kx_, ky_, shear);
        flow5 = grid_U(qk, 3, K2, kx_, ky_);                      % the five-argument call of qgsw_raytrace.m:63 works again
        du = flow.u - flow5.u;
        eng = swrt_mex('create', nx, L, f, 1, 1, 0, 1e-10);        % LAGRANGE6, bump of the copy beside interpolate_U.m
        swrt_mex('set_flow_grid', eng, 0, bf1.u, bf1.v, bf1.ux, bf1.uy, bf1.vx, bf1.vy);
        swrt_mex('set_flow_grid', eng, 1, bf2.u, bf2.v, bf2.ux, bf2.uy, bf2.vx, bf2.vy);
        swrt_mex('set_packets', eng, x, y, k, l);
        tmax = 0.4;
        stats = swrt_ode23(eng, [0 tmax], tmax);
        [px, py, pk, pl] = swrt_mex('get_packets', eng);
        swrt_mex('destroy', eng);
    """, fr)
    v = fr.vars
    for j, nm in enumerate(NAMES):
        ref = R["grid_U_shear"][j]
        assert np.abs(np.asarray(v["flow"].f[nm]) - ref).max() / np.abs(ref).max() <= 1e-14, nm
    assert np.abs(np.asarray(v["du"]) - float(R["grid_U_shear_value"])).max() < 1e-15
    bf1 = {nm: read_in(f"bf1_{nm}", nx, nx) for nm in NAMES}; bf2 = {nm: read_in(f"bf2_{nm}", nx, nx) for nm in NAMES}
    x, y, k, l = (read_in(c, n) for c in "xykl")
    with S.Engine(nx, float(v["L"]), float(v["f"]), 1.0, S.MODE_LAGRANGE6, bump=1e-10) as e:
        e.set_flow_grid(*[bf1[nm] for nm in NAMES], slot=0); e.set_flow_grid(*[bf2[nm] for nm in NAMES], slot=1)
        e.set_packets(x, y, k, l)
        st = A.ode23(e, [0.0, 0.4], 0.4)
        want = e.get_packets()
    assert (v["stats"].f["nsteps"], v["stats"].f["nfailed"]) == (float(st["nsteps"]), float(st["nfailed"])) and st["nsteps"] >= 5
    for got, ref in zip((v["px"], v["py"], v["pk"], v["pl"]), want):
        assert np.array_equal(np.asarray(got).ravel(), ref)


# ------------------------------------------------------------------------------------------------------------------------ CPU
def test_every_shim_parses_and_names_the_reference_function_it_replaces():
    files = sorted(SHIMS.glob("*.m"))
    assert {f.stem for f in files} >= {"interpolate", "interpolate_par", "interpolate2", "interpolate_U", "SpectralScheme", "ode_symplectic",
                                       "step_packet", "step_packet_xka", "grid_U", "swrt_ode23", "swrt_step_packets"}
    for f in files:
        u = parse_source(f.read_text(), str(f))
        assert u.kind in ("function", "class"), f
        name = u.classdef.name if u.kind == "class" else u.main.name
        assert name == f.stem


def test_shims_fail_loudly_without_a_device():
    import swraytracing_b200 as S
    if S.load_library().swrt_device_count() > 0:
        pytest.skip("a CUDA device is present")
    I = interp(Mex(False))
    fr, nx, n = workspace()
    with pytest.raises(MatlabError, match="no CPU path|swrt:call"):
        I.run("FI = interpolate(x, y, bf1.u, dx, dx);", fr)
    with pytest.raises(MatlabError, match="no CPU path|swrt:call"):
        I.run("scheme = SpectralScheme(L, nx, psi);", fr)
