"""ctypes binding of libswrt.so (include/swrt.h) -- the tested skin of the C ABI.

``Engine`` is a thin 1:1 wrapper: every method is one C call on host numpy buffers.  There is no
CPU fallback: if the library is missing or no CUDA device is present the constructor raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libswrt.so"

MODE_SPECTRAL, MODE_LAGRANGE6, MODE_NUFFT = 0, 1, 2
SCHEME_LEAPFROG, SCHEME_RK4_PACKET, SCHEME_RK4_XKA = 0, 1, 2
HIST_INTRINSIC, HIST_ABSOLUTE = 0, 1
FLAG_RHS_GH = 1

_dp = C.POINTER(C.c_double)


class SwrtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libswrt error {code}: {msg}")
        self.code = code


class _Params(C.Structure):
    _fields_ = [("nx", C.c_int32), ("mode", C.c_int32), ("device", C.c_int32), ("flags", C.c_int32),
                ("L", C.c_double), ("f", C.c_double), ("gH", C.c_double), ("bump", C.c_double),
                ("ngpu", C.c_int32), ("reserved", C.c_int32)]


# name -> (restype, argtypes); every symbol include/swrt.h declares
SIGNATURES = {
    "swrt_version": (C.c_int, []),
    "swrt_device_count": (C.c_int, []),
    "swrt_create": (C.c_int, [C.POINTER(_Params), C.POINTER(C.c_void_p)]),
    "swrt_destroy": (C.c_int, [C.c_void_p]),
    "swrt_last_error": (C.c_char_p, [C.c_void_p]),
    "swrt_set_flow_spectral": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, C.c_int, C.c_int, C.c_double]),
    "swrt_set_flow_planes_spectral": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(_dp), C.POINTER(_dp), C.c_int, C.c_int, C.c_int]),
    "swrt_set_flow_grid": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_int]),
    "swrt_set_packets": (C.c_int, [C.c_void_p, C.c_int64, _dp, _dp, _dp, _dp, _dp]),
    "swrt_get_packets": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp, _dp]),
    "swrt_num_packets": (C.c_int64, [C.c_void_p]),
    "swrt_num_devices": (C.c_int, [C.c_void_p]),
    "swrt_shard_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "swrt_packets_alloc_dev": (C.c_int, [C.c_void_p, C.c_int64]),
    "swrt_packets_dev": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_void_p)] * 5),
    "swrt_eval": (C.c_int, [C.c_void_p, C.c_double] + [_dp] * 6),
    "swrt_eval_at": (C.c_int, [C.c_void_p, C.c_double, C.c_int64, _dp, _dp] + [_dp] * 7),
    "swrt_rhs": (C.c_int, [C.c_void_p, C.c_double] + [_dp] * 4),
    "swrt_interpolate": (C.c_int, [C.c_int, _dp, _dp, C.c_int64, _dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, _dp]),
    "swrt_step": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double]),
    "swrt_step_async": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double]),
    "swrt_step_host": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int64] + [_dp] * 10),
    "swrt_hist_omega": (C.c_int, [C.c_void_p, C.c_int, C.c_double, _dp, C.c_int, C.POINTER(C.c_uint64), C.c_int]),
    "swrt_bs23_begin": (C.c_int, [C.c_void_p, C.c_double, C.c_double, _dp]),
    "swrt_bs23_attempt": (C.c_int, [C.c_void_p, C.c_double, _dp, C.c_double, _dp]),
    "swrt_bs23_accept": (C.c_int, [C.c_void_p]),
    "swrt_bs23_interp": (C.c_int, [C.c_void_p, C.c_double, C.c_double] + [_dp] * 4),
    "swrt_hist_omega_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_double, _dp, C.c_int, C.POINTER(C.c_void_p)]),
    "swrt_hist_omega_launch": (C.c_int, [C.c_void_p, C.c_int, C.c_double, _dp, C.c_int, C.POINTER(C.c_void_p)]),
    "swrt_hist_omega_wait": (C.c_int, [C.c_void_p]),
    "swrt_ideal_omega_hist": (C.c_int, [C.c_void_p, C.c_double, C.c_int64, _dp, _dp, _dp, _dp, C.c_int, C.c_double, _dp, C.c_int,
                                        C.POINTER(C.c_uint64)]),
    "swrt_diag": (C.c_int, [C.c_void_p, C.c_double, _dp]),
    "swrt_omega": (C.c_int, [C.c_void_p, C.c_double, _dp, _dp]),
    "swrt_g2k": (C.c_int, [C.c_int, _dp, C.c_int, _dp, _dp]),
    "swrt_k2g": (C.c_int, [C.c_int, _dp, _dp, C.c_int, _dp]),
    "swrt_qg_create": (C.c_int, [C.c_int, C.c_int] + [C.c_double] * 8 + [_dp, _dp, C.POINTER(C.c_void_p)]),
    "swrt_qg_step": (C.c_int, [C.c_void_p, C.c_int]),
    "swrt_qg_get": (C.c_int, [C.c_void_p, _dp, _dp]),
    "swrt_qg_get_grid": (C.c_int, [C.c_void_p, _dp]),
    "swrt_qg_destroy": (C.c_int, [C.c_void_p]),
    "swrt_set_flow_from_qg": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_double]),
    "swrt_qg2_create": (C.c_int, [C.c_int, C.c_int] + [C.c_double] * 7 + [_dp] * 4 + [C.POINTER(C.c_void_p)]),
    "swrt_qg2_max_speed": (C.c_int, [C.c_void_p, _dp]),
    "swrt_qg2_step": (C.c_int, [C.c_void_p, C.c_double]),
    "swrt_qg2_get": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp]),
    "swrt_qg2_destroy": (C.c_int, [C.c_void_p]),
    "swrt_set_flow_from_qg2": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "swrt_launch_count": (C.c_int64, [C.c_void_p, C.c_int]),
    "swrt_last_kernel_ms": (C.c_double, [C.c_void_p, C.POINTER(C.c_int)]),
    "swrt_work_per_eval": (C.c_double, [C.c_void_p, C.c_int]),
    "swrt_synchronize": (C.c_int, [C.c_void_p]),
    "swrt_set_tuning": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "swrt_contracted_planes": (C.c_int, [C.c_void_p]),
    "swrt_spectral_geometry": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64)]),
    "swrt_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "swrt_gather_probe": (C.c_int, [C.c_int, C.c_int64, C.c_int, _dp]),
    "swrt_timer_start": (C.c_int, [C.c_void_p]),
    "swrt_timer_stop": (C.c_double, [C.c_void_p]),
}

_lib = None


def load_library(path: os.PathLike | None = None):
    """dlopen libswrt.so and declare every prototype.  Raises if the library is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else Path(os.environ.get("SWRT_LIB", LIB_PATH))   # SWRT_LIB: developer A/B builds
    if not p.exists():
        raise FileNotFoundError(f"{p} not found: build it with `make` or `python -c 'import __graft_entry__ as g; g.build()'`"
                                " (libswrt has no CPU fallback)")
    lib = C.CDLL(str(p))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


_nccl_preloaded = False


def _preload_nccl():
    """A multi-device handle makes libswrt dlopen("libnccl.so.2").  The dynamic loader keeps ONE library per soname in a
    process, so whichever NCCL is mapped first also serves everybody else -- e.g. a later ``import torch``, whose
    libtorch_cuda.so needs the symbols of the NCCL release it was built against.  When a pip-installed NCCL (the one torch
    bundles: site-packages/nvidia/nccl/lib) is present, map that one first; otherwise the system library is used."""
    global _nccl_preloaded
    if _nccl_preloaded or os.environ.get("SWRT_NCCL_LIB"):
        return
    _nccl_preloaded = True
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for root in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
            cand = Path(root) / "lib" / "libnccl.so.2"
            if cand.exists():
                C.CDLL(str(cand), mode=C.RTLD_GLOBAL)
                return
    except Exception:
        pass


def spectral_geometry(nx, nplanes=3, mtiles=1):
    """launch geometry of the dense kernel (host-only diagnostic, swrt_spectral_geometry)"""
    out = (C.c_int64 * 10)()
    rc = load_library().swrt_spectral_geometry(int(nx), int(nplanes), int(mtiles), out)
    if rc != 0:
        raise SwrtError(rc, load_library().swrt_last_error(None).decode())
    keys = ("ntiles", "npass", "ksteps", "kc", "nstages", "chunk_bytes", "twiddle_table", "table_bytes", "smem_bytes", "stack_bytes")
    return dict(zip(keys, (int(v) for v in out)))


def gather_probe(table_bytes=1 << 24, reps=5, device=0):
    """measured scattered-gather rate (GB/s) from an L2-resident table: the roofline denominator of the gather modes"""
    out = C.c_double(0.0)
    rc = load_library().swrt_gather_probe(int(device), int(table_bytes), int(reps), C.byref(out))
    if rc != 0:
        raise SwrtError(rc, load_library().swrt_last_error(None).decode())
    return out.value


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _colmajor(a):
    """host array -> flat column-major (MATLAB) float64 copy"""
    return np.asfortranarray(np.asarray(a, dtype=np.float64)).ravel(order="F").copy()


class Engine:
    """One handle = ``ngpu`` CUDA devices (default one) + device-resident packets + flow stacks.  With ``ngpu > 1`` the
    packets are sharded over the devices ``device .. device+ngpu-1`` inside libswrt (one host thread, in-library NCCL
    all-reduce of histograms / diagnostics / the ode23 norm); every method keeps its meaning for the whole ensemble."""

    def __init__(self, nx, L, f, gH, mode=MODE_SPECTRAL, device=0, bump=1e-13, flags=0, ngpu=1):
        self.lib = load_library()
        self.nx, self.L, self.f, self.gH, self.mode, self.device = int(nx), float(L), float(f), float(gH), int(mode), int(device)
        self.ngpu = int(ngpu)
        if self.ngpu > 1:
            _preload_nccl()
        self._h = C.c_void_p()
        prm = _Params(self.nx, self.mode, self.device, int(flags), self.L, self.f, self.gH, float(bump), self.ngpu, 0)
        rc = self.lib.swrt_create(C.byref(prm), C.byref(self._h))
        if rc != 0:
            raise SwrtError(rc, (self.lib.swrt_last_error(None) or b"").decode())
        self.n = 0

    # -- plumbing --
    def _check(self, rc):
        if rc != 0:
            raise SwrtError(rc, (self.lib.swrt_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.swrt_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- flow --
    def set_flow_spectral(self, psik, slot=0, u_mean=0.0):
        """psik: complex (nkx, nky) array in g2k layout (kx = first index)."""
        psik = np.asarray(psik, dtype=np.complex128)
        nkx, nky = psik.shape
        re, im = _colmajor(psik.real), _colmajor(psik.imag)
        self._check(self.lib.swrt_set_flow_spectral(self._h, slot, _ptr(re), _ptr(im), nkx, nky, float(u_mean)))

    def set_flow_planes_spectral(self, planes, slot=0):
        planes = [np.asarray(p, dtype=np.complex128) for p in planes]
        nkx, nky = planes[0].shape
        res = [_colmajor(p.real) for p in planes]
        ims = [_colmajor(p.imag) for p in planes]
        n = len(planes)
        tr = (_dp * n)(*[_ptr(a) for a in res])
        ti = (_dp * n)(*[_ptr(a) for a in ims])
        self._check(self.lib.swrt_set_flow_planes_spectral(self._h, slot, tr, ti, n, nkx, nky))

    def set_flow_grid(self, u, v, ux, uy, vx, vy, H=None, slot=0):
        """gridded planes F[ix, iy] (x = first index, as MATLAB)."""
        arrs = [_colmajor(a) for a in (u, v, ux, uy, vx, vy)]
        h = _colmajor(H) if H is not None else None
        nx = np.asarray(u).shape[0]
        self._check(self.lib.swrt_set_flow_grid(self._h, slot, *[_ptr(a) for a in arrs], _ptr(h), nx))

    # -- packets --
    def set_packets(self, x, y, k, l, a=None):
        x, y, k, l = map(_f64, (x, y, k, l))
        a = _f64(a) if a is not None else None
        self.n = x.size
        self._check(self.lib.swrt_set_packets(self._h, self.n, _ptr(x), _ptr(y), _ptr(k), _ptr(l), _ptr(a)))

    def get_packets(self, with_a=False):
        n = self.n
        out = [np.empty(n) for _ in range(5 if with_a else 4)]
        ptrs = [_ptr(o) for o in out] + ([] if with_a else [None])
        self._check(self.lib.swrt_get_packets(self._h, *ptrs))
        return tuple(out)

    def alloc_packets_dev(self, n):
        self._check(self.lib.swrt_packets_alloc_dev(self._h, int(n)))
        self.n = int(n)

    def packets_dev(self):
        """raw device pointers (ints) of the SoA buffers x,y,k,l,a"""
        ps = [C.c_void_p() for _ in range(5)]
        self._check(self.lib.swrt_packets_dev(self._h, *[C.byref(p) for p in ps]))
        return tuple(p.value for p in ps)

    # -- evaluation --
    def eval(self, alpha=0.0):
        outs = [np.empty(self.n) for _ in range(6)]
        self._check(self.lib.swrt_eval(self._h, float(alpha), *[_ptr(o) for o in outs]))
        return np.stack(outs)

    def eval_at(self, x, y, alpha=0.0, with_H=False):
        x, y = _f64(x).ravel(), _f64(y).ravel()
        n = x.size
        outs = [np.empty(n) for _ in range(7 if with_H else 6)]
        ptrs = [_ptr(o) for o in outs] + ([] if with_H else [None])
        self._check(self.lib.swrt_eval_at(self._h, float(alpha), n, _ptr(x), _ptr(y), *ptrs))
        return np.stack(outs)

    def rhs(self, alpha=0.0):
        outs = [np.empty(self.n) for _ in range(4)]
        self._check(self.lib.swrt_rhs(self._h, float(alpha), *[_ptr(o) for o in outs]))
        return tuple(outs)

    def step(self, scheme, dt, nsteps, alpha0=0.0, dalpha=0.0):
        self._check(self.lib.swrt_step(self._h, int(scheme), float(dt), int(nsteps), float(alpha0), float(dalpha)))

    def step_host(self, scheme, dt, nsteps, x, y, k, l, a=None, alpha0=0.0, dalpha=0.0, out=None):
        """host arrays in -> ``nsteps`` steps -> host arrays out in ONE C call (swrt_step_host: upload, compute and
        download pipelined over packet chunks).  ``out``: optional (x, y, k, l[, a]) arrays to receive the result."""
        x, y, k, l = map(_f64, (x, y, k, l))
        a = _f64(a) if a is not None else None
        self.n = x.size
        if out is None:
            out = [np.empty(self.n) for _ in range(5 if a is not None else 4)]
        ptrs = [_ptr(o) for o in out] + [None] * (5 - len(out))
        self._check(self.lib.swrt_step_host(self._h, int(scheme), float(dt), int(nsteps), float(alpha0), float(dalpha), self.n,
                                            _ptr(x), _ptr(y), _ptr(k), _ptr(l), _ptr(a), *ptrs))
        return tuple(out)

    def shard_info(self, i):
        """(device, lo, n) of shard ``i`` of a multi-device handle"""
        d, lo, n = C.c_int(0), C.c_int64(0), C.c_int64(0)
        self._check(self.lib.swrt_shard_info(self._h, int(i), C.byref(d), C.byref(lo), C.byref(n)))
        return d.value, lo.value, n.value

    def step_async(self, scheme, dt, nsteps, alpha0=0.0, dalpha=0.0):
        """queue the step's kernels and return at once (``synchronize()`` or any blocking call completes them)"""
        self._check(self.lib.swrt_step_async(self._h, int(scheme), float(dt), int(nsteps), float(alpha0), float(dalpha)))

    # -- ode23 building blocks (the controller lives in reference_api.ode23) --
    def bs23_begin(self, alpha, threshold):
        out = C.c_double(0.0)
        self._check(self.lib.swrt_bs23_begin(self._h, float(alpha), float(threshold), C.byref(out)))
        return out.value

    def bs23_attempt(self, hstep, alphas, threshold):
        al = (C.c_double * 3)(*[float(a) for a in alphas])
        out = C.c_double(0.0)
        self._check(self.lib.swrt_bs23_attempt(self._h, float(hstep), al, float(threshold), C.byref(out)))
        return out.value

    def bs23_accept(self):
        self._check(self.lib.swrt_bs23_accept(self._h))

    def bs23_interp(self, hstep, s):
        """ntrp23 dense output of the last attempted step at t + s*hstep -> (x, y, k, l)"""
        out = [np.empty(self.n) for _ in range(4)]
        self._check(self.lib.swrt_bs23_interp(self._h, float(hstep), float(s), *[_ptr(o) for o in out]))
        return tuple(out)

    # -- diagnostics --
    def hist_omega(self, edges, kind=HIST_INTRINSIC, alpha=0.0, counts=None):
        edges = _f64(edges)
        acc = counts is not None
        if counts is None:
            counts = np.zeros(edges.size - 1, dtype=np.uint64)
        self._check(self.lib.swrt_hist_omega(self._h, kind, float(alpha), _ptr(edges), edges.size,
                                             counts.ctypes.data_as(C.POINTER(C.c_uint64)), int(acc)))
        return counts

    def hist_omega_dev(self, edges, kind=HIST_INTRINSIC, alpha=0.0):
        """histogram left on the device: returns (device pointer, nbins) of the handle-owned u64 counts"""
        edges = _f64(edges)
        ptr = C.c_void_p()
        self._check(self.lib.swrt_hist_omega_dev(self._h, kind, float(alpha), _ptr(edges), edges.size, C.byref(ptr)))
        return ptr.value, edges.size - 1

    def hist_omega_launch(self, edges, kind=HIST_INTRINSIC, alpha=0.0):
        """non-blocking: queue the histogram kernel on the handle's stream -> (device pointer, nbins)"""
        edges = _f64(edges)
        ptr = C.c_void_p()
        self._check(self.lib.swrt_hist_omega_launch(self._h, kind, float(alpha), _ptr(edges), edges.size, C.byref(ptr)))
        return ptr.value, edges.size - 1

    def hist_omega_wait(self):
        self._check(self.lib.swrt_hist_omega_wait(self._h))

    def ideal_omega_hist(self, x, y, kvx, kvy, omega0, edges, alpha=0.0):
        x, y, kvx, kvy, edges = (_f64(a).ravel() for a in (x, y, kvx, kvy, edges))
        counts = np.zeros(edges.size - 1, dtype=np.uint64)
        self._check(self.lib.swrt_ideal_omega_hist(self._h, float(alpha), x.size, _ptr(x), _ptr(y), _ptr(kvx), _ptr(kvy), kvx.size,
                                                   float(omega0), _ptr(edges), edges.size, counts.ctypes.data_as(C.POINTER(C.c_uint64))))
        return counts

    def diag(self, alpha=0.0):
        out = np.empty(8)
        self._check(self.lib.swrt_diag(self._h, float(alpha), _ptr(out)))
        return out

    def omega(self, alpha=0.0, absolute=True):
        om = np.empty(self.n)
        Om = np.empty(self.n) if absolute else None
        self._check(self.lib.swrt_omega(self._h, float(alpha), _ptr(om), _ptr(Om)))
        return om, Om

    # -- instrumentation --
    def launch_count(self, reset=False):
        return int(self.lib.swrt_launch_count(self._h, int(reset)))

    def last_kernel_ms(self):
        n = C.c_int(0)
        ms = self.lib.swrt_last_kernel_ms(self._h, C.byref(n))
        return float(ms), int(n.value)

    def work_per_eval(self, nplanes=6):
        return float(self.lib.swrt_work_per_eval(self._h, nplanes))

    def synchronize(self):
        self._check(self.lib.swrt_synchronize(self._h))

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.swrt_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def timer_start(self):
        self._check(self.lib.swrt_timer_start(self._h))

    def timer_stop(self):
        ms = float(self.lib.swrt_timer_stop(self._h))
        if ms < 0:
            raise SwrtError(-3, "timer failed")
        return ms

    def set_tuning(self, mtiles=0, use_psi_moments=True, preblend_grid=False, unfused_rk4=False, twiddles=0):
        """``preblend_grid`` (LAGRANGE6): blend two frames on the grid before the gather (faster) instead of
        interpolating both frames and blending the results as interpolate_U.m does (default, bit-faithful);
        ``unfused_rk4`` (SPECTRAL, NUFFT): step_packet* as evaluation + stage launches instead of the fused kernel;
        ``twiddles`` (SPECTRAL): 0 automatic, 1 rotate in registers, 2 global (L2) table, 3 shared-memory table else global"""
        self._check(self.lib.swrt_set_tuning(self._h, int(mtiles), (0 if use_psi_moments else 1) | (2 if preblend_grid else 0)
                                             | (4 if unfused_rk4 else 0) | ((int(twiddles) & 3) << 3)))

    def contracted_planes(self):
        return int(self.lib.swrt_contracted_planes(self._h))


class QGFlow:
    """On-device one-layer QG solver producing the background-flow frames (qgsw_raytrace.m:111-137,
    :270-286): q-hat lives on the GPU; ``to_flow(engine, slot)`` fills a flow slot without a host copy."""

    def __init__(self, nx, L, qk, K_d2, dt, f, Cg, beta=0.0, r_drag=0.1, force_strength=0.1, device=0):
        self.lib = load_library()
        self.nx, self.L = int(nx), float(L)
        qk = np.asarray(qk, dtype=np.complex128)
        re, im = _colmajor(qk.real), _colmajor(qk.imag)
        self._q = C.c_void_p()
        rc = self.lib.swrt_qg_create(int(device), self.nx, self.L, float(K_d2), float(beta), float(r_drag), float(force_strength),
                                     float(f), float(Cg), float(dt), _ptr(re), _ptr(im), C.byref(self._q))
        if rc != 0:
            raise SwrtError(rc, (self.lib.swrt_last_error(None) or b"").decode())

    def step(self, nsteps=1):
        rc = self.lib.swrt_qg_step(self._q, int(nsteps))
        if rc != 0:
            raise SwrtError(rc, "swrt_qg_step failed")

    def get(self):
        nkx, nky = self.nx - 1, self.nx // 2
        re = np.empty(nkx * nky); im = np.empty(nkx * nky)
        rc = self.lib.swrt_qg_get(self._q, _ptr(re), _ptr(im))
        if rc != 0:
            raise SwrtError(rc, "swrt_qg_get failed")
        return (re + 1j * im).reshape((nkx, nky), order="F")

    def get_grid(self):
        """q = k2g(qk) (nx, nx), transformed on the device with the solver's own FFT plan"""
        out = np.empty(self.nx * self.nx)
        rc = self.lib.swrt_qg_get_grid(self._q, _ptr(out))
        if rc != 0:
            raise SwrtError(rc, "swrt_qg_get_grid failed")
        return out.reshape((self.nx, self.nx), order="F")

    def to_flow(self, engine, slot=0, u_mean=0.0):
        engine._check(self.lib.swrt_set_flow_from_qg(engine._h, int(slot), self._q, float(u_mean)))

    def close(self):
        if getattr(self, "_q", None) and self._q.value:
            self.lib.swrt_qg_destroy(self._q)
            self._q = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class QG2Flow:
    """On-device two-layer QG solver (qg2layersw_raytrace.m:120-181, :309-323): both layer PV spectra, the
    inversion matrix B and expm(factor_L*dt) live on the GPU; ``to_flow`` fills a flow slot with
    ``grid_U(qk(:,:,1), ..., shear_strength)`` of the top layer without a host copy."""

    def __init__(self, nx, L, q1k, q2k, K_d2, beta, shear_strength, r, nu, alpha, device=0):
        self.lib = load_library()
        self.nx, self.L = int(nx), float(L)
        parts = []
        for qk in (q1k, q2k):
            qk = np.asarray(qk, dtype=np.complex128)
            parts += [_colmajor(qk.real), _colmajor(qk.imag)]
        self._q = C.c_void_p()
        rc = self.lib.swrt_qg2_create(int(device), self.nx, self.L, float(K_d2), float(beta), float(shear_strength), float(r),
                                      float(nu), float(alpha), *[_ptr(a) for a in parts], C.byref(self._q))
        if rc != 0:
            raise SwrtError(rc, (self.lib.swrt_last_error(None) or b"").decode())

    def _check(self, rc, what):
        if rc != 0:
            raise SwrtError(rc, f"{what} failed")

    def max_speed(self):
        out = C.c_double(0.0)
        self._check(self.lib.swrt_qg2_max_speed(self._q, C.byref(out)), "swrt_qg2_max_speed")
        return out.value

    def step(self, dt):
        self._check(self.lib.swrt_qg2_step(self._q, float(dt)), "swrt_qg2_step")

    def get(self, layer):
        nkx, nky = self.nx - 1, self.nx // 2
        re = np.empty(nkx * nky); im = np.empty(nkx * nky)
        self._check(self.lib.swrt_qg2_get(self._q, int(layer), _ptr(re), _ptr(im)), "swrt_qg2_get")
        return (re + 1j * im).reshape((nkx, nky), order="F")

    def to_flow(self, engine, slot=0):
        engine._check(self.lib.swrt_set_flow_from_qg2(engine._h, int(slot), self._q))

    def close(self):
        if getattr(self, "_q", None) and self._q.value:
            self.lib.swrt_qg2_destroy(self._q)
            self._q = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# -- handle-free helpers ---------------------------------------------------------------------------

def interpolate_dev(x, y, F, dx, dy, bump=1e-13, device=0):
    lib = load_library()
    x = _f64(x); shp = x.shape
    x = x.ravel(); y = _f64(y).ravel()
    if x.size != y.size:
        raise ValueError(f"interpolate: x has {x.size} elements, y has {y.size}")
    F = np.asarray(F, dtype=np.float64)
    nx, ny = F.shape
    Ff = _colmajor(F)
    out = np.empty(x.size)
    rc = lib.swrt_interpolate(device, _ptr(x), _ptr(y), x.size, _ptr(Ff), nx, ny, float(dx), float(dy), float(bump), _ptr(out))
    if rc != 0:
        raise SwrtError(rc, (lib.swrt_last_error(None) or b"").decode())
    return out.reshape(shp)


def g2k_dev(fg, device=0):
    lib = load_library()
    fg = np.asarray(fg, dtype=np.float64)
    nx = fg.shape[0]
    nkx, nky = nx - 1, nx // 2
    re = np.empty(nkx * nky); im = np.empty(nkx * nky)
    f = _colmajor(fg)
    rc = lib.swrt_g2k(device, _ptr(f), nx, _ptr(re), _ptr(im))
    if rc != 0:
        raise SwrtError(rc, (lib.swrt_last_error(None) or b"").decode())
    return (re + 1j * im).reshape((nkx, nky), order="F")


def k2g_dev(fk, device=0):
    lib = load_library()
    fk = np.asarray(fk, dtype=np.complex128)
    nx = fk.shape[0] + 1
    re, im = _colmajor(fk.real), _colmajor(fk.imag)
    out = np.empty(nx * nx)
    rc = lib.swrt_k2g(device, _ptr(re), _ptr(im), nx, _ptr(out))
    if rc != 0:
        raise SwrtError(rc, (lib.swrt_last_error(None) or b"").decode())
    return out.reshape((nx, nx), order="F")
