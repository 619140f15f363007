"""The MEX gateway (matlab/swrt_mex.c) EXECUTED: its unmodified source is linked with tests/mex_harness (a small
implementation of the mx / mex functions it calls) and libswrt.so, and ``mexFunction`` is driven command by command from
here, in both complex-storage layouts (separate real/imag: GNU Octave and MATLAB -R2017b; interleaved: MATLAB -R2018a).

CPU part (``-m "not gpu"``): the harness builds, mexFunction runs, usage / handle / no-device errors surface as
``mexErrMsgIdAndTxt`` identifiers.  GPU part: every result is compared bit for bit with the ctypes path (engine.py), which
binds the same C-ABI symbols -- create -> set_flow_spectral -> set_packets -> step -> get_packets -> hist_omega -> destroy and
the rest of the command table.  Callers replaced: symplectic_full_fourier.m:20,44, raytrace_sw.m:128, qgsw_raytrace.m:141-150."""
import ctypes as C
import re
import sys
from pathlib import Path

import numpy as np
import pytest

import swraytracing_b200 as S
from swraytracing_b200 import workloads as W

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "mex_harness"))
import build as harness_build  # noqa: E402

_dp = C.POINTER(C.c_double)


class MexError(RuntimeError):
    def __init__(self, ident, msg):
        super().__init__(f"{ident}: {msg}")
        self.ident, self.msg = ident, msg


class Mex:
    """ctypes driver of one harness build: ``mex(cmd, *args, nlhs=k)`` = ``[o1..ok] = swrt_mex(cmd, args...)``"""

    def __init__(self, interleaved):
        S.load_library()                                   # libswrt.so first (the harness links against it)
        so = harness_build.build()[1 if interleaved else 0]
        lib = C.CDLL(str(so))
        lib.hx_double.restype = C.c_void_p; lib.hx_double.argtypes = [C.c_size_t, C.c_size_t, C.c_int]
        lib.hx_uint64.restype = C.c_void_p; lib.hx_uint64.argtypes = [C.c_uint64]
        lib.hx_string.restype = C.c_void_p; lib.hx_string.argtypes = [C.c_char_p]
        lib.hx_set.argtypes = [C.c_void_p, _dp, _dp]
        lib.hx_get.argtypes = [C.c_void_p, _dp, _dp]
        lib.hx_get_u64.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        lib.hx_dims.argtypes = [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.hx_free.argtypes = [C.c_void_p]
        lib.hx_call.restype = C.c_int; lib.hx_call.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)]
        lib.hx_error_id.restype = C.c_char_p; lib.hx_error_msg.restype = C.c_char_p
        self.lib = lib
        assert lib.hx_interleaved() == int(interleaved)

    def _to_mx(self, a):
        lib = self.lib
        if isinstance(a, str):
            return lib.hx_string(a.encode())
        if isinstance(a, np.uint64):
            return lib.hx_uint64(int(a))
        if a is None:
            return lib.hx_double(0, 0, 0)
        arr = np.asarray(a)
        cplx = np.iscomplexobj(arr)
        arr = np.atleast_1d(arr)
        m, n = (arr.shape[0], 1) if arr.ndim == 1 else arr.shape
        mx = lib.hx_double(m, n, int(cplx))
        flat = np.asfortranarray(arr.reshape(m, n)).ravel(order="F")
        re = np.ascontiguousarray(flat.real, dtype=np.float64)
        im = np.ascontiguousarray(flat.imag, dtype=np.float64) if cplx else None
        lib.hx_set(mx, re.ctypes.data_as(_dp), im.ctypes.data_as(_dp) if cplx else None)
        return mx

    def _from_mx(self, mx):
        lib = self.lib
        m, n, cls, cplx = C.c_size_t(), C.c_size_t(), C.c_int(), C.c_int()
        lib.hx_dims(mx, C.byref(m), C.byref(n), C.byref(cls), C.byref(cplx))
        m, n = m.value, n.value
        if cls.value == 13:                                # uint64
            out = np.zeros(m * n, dtype=np.uint64)
            lib.hx_get_u64(mx, out.ctypes.data_as(C.POINTER(C.c_uint64)))
            return out[0] if m * n == 1 else out
        re = np.zeros(m * n); im = np.zeros(m * n)
        lib.hx_get(mx, re.ctypes.data_as(_dp), im.ctypes.data_as(_dp) if cplx.value else None)
        out = (re + 1j * im) if cplx.value else re
        out = out.reshape((m, n), order="F")
        if n == 1:
            out = out[:, 0]
        return out[0] if out.size == 1 and not cplx.value and m == 1 else out

    def __call__(self, cmd, *args, nlhs=1):
        lib = self.lib
        rhs = [self._to_mx(cmd)] + [self._to_mx(a) for a in args]
        prhs = (C.c_void_p * len(rhs))(*rhs)
        nout = max(nlhs, 1)
        plhs = (C.c_void_p * (nout + 8))()                 # MATLAB gives mexFunction room for max(nlhs, 1) outputs; the 8 guard
        rc = lib.hx_call(nlhs, plhs, len(rhs), prhs)       # slots behind them must come back untouched
        try:
            overrun = [i for i in range(nout, nout + 8) if plhs[i]]
            assert not overrun, f"swrt_mex('{cmd}', ...) wrote plhs{overrun} with nlhs = {nlhs}"
            if rc != 0:
                raise MexError(lib.hx_error_id().decode(), lib.hx_error_msg().decode())
            outs = [self._from_mx(plhs[i]) for i in range(nlhs) if plhs[i]]
        finally:
            for r in rhs:
                lib.hx_free(r)
            for i in range(nout + 8):
                if plhs[i]:
                    lib.hx_free(plhs[i])
        if nlhs <= 1:
            return outs[0] if outs else None
        return outs


@pytest.fixture(scope="module", params=[False, True], ids=["separate-complex", "interleaved-complex"])
def mex(request):
    return Mex(request.param)


def _have_gpu():
    return S.load_library().swrt_device_count() > 0


# ------------------------------------------------------------------------------------------------
# CPU: the gateway runs; errors come back as MEX error identifiers
# ------------------------------------------------------------------------------------------------
def test_gateway_executes_and_reports_usage_errors(mex):
    assert mex("version") == 200.0
    assert mex.lib.hx_locked() >= 1                        # mexLock + mexAtExit were called on first use
    with pytest.raises(MexError) as ei:
        mex("no_such_command", np.uint64(1))
    assert ei.value.ident == "swrt:handle"                 # the handle is checked before the command table
    with pytest.raises(MexError) as ei:
        mex("step")
    assert ei.value.ident == "swrt:usage"
    with pytest.raises(MexError) as ei:
        mex("create", 32.0)                                # too few arguments
    assert ei.value.ident == "swrt:usage" and "create" in ei.value.msg
    with pytest.raises(MexError) as ei:
        mex("create", 31.0, 2 * np.pi, 3.0, 1.0)           # odd nx: the library's message travels through the gateway
    assert ei.value.ident == "swrt:call" and "nx" in ei.value.msg
    with pytest.raises(MexError) as ei:
        mex("qg_step", np.uint64(3))
    assert ei.value.ident == "swrt:handle"
    with pytest.raises(MexError) as ei:
        mex("interpolate", np.zeros(3), np.zeros(4), np.zeros((8, 8)), 1.0, 1.0)
    assert ei.value.ident == "swrt:usage"


def test_gateway_without_a_device_fails_loudly(mex):
    if _have_gpu():
        pytest.skip("a CUDA device is present")
    with pytest.raises(MexError) as ei:
        mex("create", 32.0, 2 * np.pi, 3.0, 1.0)
    assert ei.value.ident == "swrt:call" and "no CPU path" in ei.value.msg
    assert mex("device_count") == 0.0


def test_every_shim_command_exists_in_the_gateway():
    text = (ROOT / "matlab" / "swrt_mex.c").read_text()
    cmds = set()
    for m in (ROOT / "matlab" / "shims").glob("*.m"):
        cmds |= set(re.findall(r"swrt_mex\('([a-z0-9_]+)'", m.read_text()))
    assert cmds and all(f'"{c}"' in text for c in cmds), cmds


# ------------------------------------------------------------------------------------------------
# GPU: gateway == ctypes path, bit for bit
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gateway_leapfrog_session_equals_ctypes_path(mex):
    """create -> set_flow_spectral -> set_packets -> step -> get_packets -> hist_omega -> diag -> destroy"""
    w = W.make_workload("C3", n_packets=5003, nx=32)
    edges = np.linspace(0.0, 8.0, 300)
    h = mex("create", float(w.nx), w.L, w.f, w.gH, 0.0)
    assert isinstance(h, np.uint64)
    mex("set_flow_spectral", h, 0.0, w.psik, w.u_mean, nlhs=0)
    mex("set_flow_spectral", h, 1.0, w.psik2, nlhs=0)
    mex("set_packets", h, w.x, w.y, w.k, w.l, nlhs=0)
    assert mex("num_packets", h) == w.n_packets and mex("num_devices", h) == 1
    mex("step", h, 0.0, w.dt / 4, 4.0, 0.125, 0.25, nlhs=0)
    got = np.stack(mex("get_packets", h, nlhs=5))
    # fewer outputs than the command can give (shims/ode_symplectic.m asks for four): nothing may be written past plhs[nlhs-1]
    # -- the driver checks guard slots behind plhs on every call
    part = mex("get_packets", h, nlhs=2)
    assert len(part) == 2 and np.array_equal(part[0], got[0]) and np.array_equal(part[1], got[1])
    assert np.array_equal(mex("get_packets", h, nlhs=1), got[0]) and np.array_equal(mex("omega", h, 0.5), mex("omega", h, 0.5, nlhs=2)[0])
    counts = mex("hist_omega", h, 0.0, 0.0, edges)
    d = mex("diag", h, 0.5)
    ev = np.stack(mex("eval", h, 0.5, nlhs=6))
    rhs = np.stack(mex("rhs", h, 0.5, nlhs=4))
    om, Om = mex("omega", h, 0.5, nlhs=2)
    ea = np.stack(mex("eval_at", h, 0.25, w.x[:100], w.y[:100], nlhs=6))
    mex("destroy", h, nlhs=0)
    with pytest.raises(MexError) as ei:
        mex("get_packets", h, nlhs=5)
    assert ei.value.ident == "swrt:handle"

    eng = S.Engine(w.nx, w.L, w.f, w.gH, S.MODE_SPECTRAL)
    eng.set_flow_spectral(w.psik, 0, w.u_mean); eng.set_flow_spectral(w.psik2, 1)
    eng.set_packets(w.x, w.y, w.k, w.l)
    eng.step(S.SCHEME_LEAPFROG, w.dt / 4, 4, 0.125, 0.25)
    assert np.array_equal(got, np.stack(eng.get_packets(with_a=True)))
    assert np.array_equal(counts, eng.hist_omega(edges)) and counts.dtype == np.uint64 and int(counts.sum()) == w.n_packets
    assert np.array_equal(d, eng.diag(0.5))
    assert np.array_equal(ev, eng.eval(0.5)) and np.array_equal(rhs, np.stack(eng.rhs(0.5)))
    o2, O2 = eng.omega(0.5)
    assert np.array_equal(om, o2) and np.array_equal(Om, O2)
    assert np.array_equal(ea, eng.eval_at(w.x[:100], w.y[:100], 0.25))
    eng.close()


@pytest.mark.gpu
def test_gateway_spectral_kit_and_complex_storage(mex):
    """g2k returns a complex array, k2g takes one (and a REAL input gets a zero imaginary part, not a NULL pointer)"""
    rs = np.random.RandomState(4)
    nx = 32
    fg = rs.standard_normal((nx, nx))
    fk = mex("g2k", fg)
    assert np.iscomplexobj(fk) and fk.shape == (nx - 1, nx // 2)
    assert np.array_equal(fk, S.g2k_dev(fg))
    assert np.array_equal(mex("k2g", fk), S.k2g_dev(fk))
    real_fk = np.abs(fk)                                   # a real-valued spectrum handed over as a REAL array
    assert np.array_equal(mex("k2g", real_fk), S.k2g_dev(real_fk.astype(np.complex128)))
    x, y = rs.uniform(-9, 9, 200), rs.uniform(-9, 9, 200)
    dx = 2 * np.pi / nx
    assert np.array_equal(mex("interpolate", x, y, fg, dx, dx), S.interpolate_dev(x, y, fg, dx, dx))
    assert np.array_equal(mex("interpolate", x, y, fg, dx, dx, 1e-10), S.interpolate_dev(x, y, fg, dx, dx, bump=1e-10))
    # real psi-hat (imaginary part absent) through set_flow_spectral
    h = mex("create", float(nx), 2 * np.pi, 3.0, 1.0, 2.0)
    mex("set_flow_spectral", h, 0.0, real_fk * 1e-3, nlhs=0)
    u = mex("eval_at", h, 0.0, x, y, nlhs=6)[0]
    eng = S.Engine(nx, 2 * np.pi, 3.0, 1.0, S.MODE_NUFFT)
    eng.set_flow_spectral(real_fk * 1e-3 + 0j)
    assert np.array_equal(u, eng.eval_at(x, y)[0])
    eng.close(); mex("destroy", h, nlhs=0)


@pytest.mark.gpu
def test_gateway_step_host_planes_and_wave_action(mex):
    """step_packet_xka the way the shim calls it: LAGRANGE6 handle, gridded planes incl. H, one step_host call"""
    w = W.make_workload("C5", n_packets=3001, nx=32)
    planes = W.planes_from_psik(w.psik, w.L, w.u_mean, etak=w.extra["etak"])
    grids = [W._fulspec_ifft(p) for p in planes]
    a = np.linspace(0.5, 2.0, w.n_packets)
    h = mex("create", float(w.nx), w.L, w.f, w.gH, 1.0)
    mex("set_flow_grid", h, 0.0, *grids, nlhs=0)
    got = np.stack(mex("step_host", h, 2.0, w.dt, 1.0, w.x, w.y, w.k, w.l, a, nlhs=5))
    mex("destroy", h, nlhs=0)
    eng = S.Engine(w.nx, w.L, w.f, w.gH, S.MODE_LAGRANGE6)
    eng.set_flow_grid(*grids[:6], H=grids[6])
    eng.set_packets(w.x, w.y, w.k, w.l, a)
    eng.step(S.SCHEME_RK4_XKA, w.dt, 1)
    assert np.array_equal(got, np.stack(eng.get_packets(with_a=True)))
    eng.close()
    # spectral planes (seven complex arrays) + tuning flags + theoretical omega pdf
    h = mex("create", float(w.nx), w.L, w.f, w.gH, 0.0)
    mex("set_flow_planes_spectral", h, 0.0, *planes, nlhs=0)
    mex("set_tuning", h, 1.0, 0.0, nlhs=0)
    gx, gy = np.meshgrid(np.linspace(0, w.L, 16), np.linspace(0, w.L, 16))
    th = np.linspace(0, 2 * np.pi, 40)
    edges = np.linspace(3.0, 9.0, 60)
    ideal = mex("ideal_omega_hist", h, 0.0, gx.ravel(), gy.ravel(), 5 * np.cos(th), 5 * np.sin(th), np.sqrt(9 + 25.0), edges)
    mex("destroy", h, nlhs=0)
    eng = S.Engine(w.nx, w.L, w.f, w.gH, S.MODE_SPECTRAL)
    eng.set_flow_planes_spectral(planes)
    assert np.array_equal(ideal, eng.ideal_omega_hist(gx.ravel(), gy.ravel(), 5 * np.cos(th), 5 * np.sin(th), np.sqrt(9 + 25.0), edges))
    eng.close()


@pytest.mark.gpu
def test_gateway_ode23_building_blocks_and_qg_producer(mex):
    """the calls matlab/shims/swrt_ode23.m makes, and the on-device QG frame producer, through the gateway"""
    w = W.make_workload("C3", n_packets=1201, nx=32)
    h = mex("create", float(w.nx), w.L, w.f, w.gH, 1.0)
    eng = S.Engine(w.nx, w.L, w.f, w.gH, S.MODE_LAGRANGE6)
    for slot, psik in ((0.0, w.psik), (1.0, w.psik2)):
        mex("set_flow_spectral", h, slot, psik, nlhs=0)
        eng.set_flow_spectral(psik, int(slot))
    mex("set_packets", h, w.x, w.y, w.k, w.l, nlhs=0)
    eng.set_packets(w.x, w.y, w.k, w.l)
    thr = 1e-3
    assert mex("bs23_begin", h, 0.0, thr) == eng.bs23_begin(0.0, thr)
    hs = 0.3 * w.dt
    al = np.array([0.5 * hs, 0.75 * hs, hs]) / w.dt
    assert mex("bs23_attempt", h, hs, al, thr) == eng.bs23_attempt(hs, al, thr)
    assert np.array_equal(np.stack(mex("bs23_interp", h, hs, 0.5, nlhs=4)), np.stack(eng.bs23_interp(hs, 0.5)))
    mex("bs23_accept", h, nlhs=0); eng.bs23_accept()
    assert np.array_equal(np.stack(mex("get_packets", h, nlhs=5))[:4], np.stack(eng.get_packets()))
    with pytest.raises(MexError):
        mex("bs23_attempt", h, hs, al[:2], thr)
    # QG producer: create, step, read back, feed a flow slot
    kx, ky = W.wavenumbers(w.nx)
    qk = -(w.f / w.Cg + kx ** 2 + ky ** 2) * w.psik
    q = mex("qg_create", float(w.nx), w.L, w.f / w.Cg, 0.0, 0.01, 0.1, w.f, w.Cg, 1e-3, qk)
    mex("qg_step", q, 3.0, nlhs=0)
    qg = S.QGFlow(w.nx, w.L, qk, w.f / w.Cg, 1e-3, w.f, w.Cg, r_drag=0.01)
    qg.step(3)
    assert np.array_equal(mex("qg_get", q, float(w.nx)), qg.get())
    assert np.array_equal(mex("qg_get_grid", q, float(w.nx)), qg.get_grid())
    mex("set_flow_from_qg", h, 0.0, q, nlhs=0); qg.to_flow(eng, 0)
    assert np.array_equal(np.stack(mex("eval", h, 0.0, nlhs=6)), eng.eval(0.0))
    mex("qg_destroy", q, nlhs=0); qg.close()
    mex("destroy", h, nlhs=0); eng.close()


@pytest.mark.gpu
def test_gateway_at_exit_releases_every_handle(mex):
    h1 = mex("create", 16.0, 2 * np.pi, 3.0, 1.0)
    h2 = mex("create", 16.0, 2 * np.pi, 3.0, 1.0, 1.0)
    assert h1 != h2
    mex.lib.hx_run_atexit()                                # what MATLAB calls when the MEX file is cleared
    for h in (h1, h2):
        with pytest.raises(MexError) as ei:
            mex("num_packets", h)
        assert ei.value.ident == "swrt:handle"


@pytest.mark.gpu
@pytest.mark.skipif(not _have_gpu() or S.load_library().swrt_device_count() < 2, reason="needs >= 2 CUDA devices")
def test_gateway_multi_device_handle(mex):
    """swrt_mex('create', ..., ngpu): a MATLAB caller reaches every GPU through the same commands"""
    w = W.make_workload("C2", n_packets=4001, nx=32)
    out = []
    for ngpu in (1.0, 2.0):
        h = mex("create", float(w.nx), w.L, w.f, w.gH, 0.0, 0.0, 1e-13, 0.0, ngpu)
        assert mex("num_devices", h) == ngpu
        mex("set_flow_spectral", h, 0.0, w.psik, nlhs=0)
        out.append(np.stack(mex("step_host", h, 0.0, w.dt, 3.0, w.x, w.y, w.k, w.l, nlhs=5)))
        out.append(mex("hist_omega", h, 0.0, 0.0, np.linspace(0, 8, 100)))
        mex("destroy", h, nlhs=0)
    assert np.array_equal(out[0], out[2]) and np.array_equal(out[1], out[3])
