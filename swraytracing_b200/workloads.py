"""Synthetic inputs for the BASELINE.json configs -- input generation only.

Seeded numpy; no oracle imports, no device work.  The same generator feeds bench.py, the GPU
parity tests and the CPU baseline so all three see identical packets and fields.

Field: random-phase QG streamfunction.  ``band`` reproduces the reference's ``initial_q``
(qgsw_raytrace.m:191-214: every |k|,|l| <= k_max mode because of the always-true chained
comparison on :202); ``full`` fills every (kx,ky) with |psi| ~ K^-3 so no sparsity can be
exploited (the timing default).  psi = -q/(K_d2 + K^2) (grid_U.m:2), normalised so max|U| = U_g.
Packets: x,y ~ U[-L/2, L/2), k on the ring k0 = sqrt((nif^2-1) f^2/Cg^2) (qgsw_raytrace.m:56-60).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


def wavenumbers(nx):
    kmax = nx // 2 - 1
    kx = np.arange(-kmax, kmax + 1, dtype=np.float64)[:, None] * np.ones((1, kmax + 1))
    ky = np.ones((2 * kmax + 1, 1)) * np.arange(0, kmax + 1, dtype=np.float64)[None, :]
    return kx, ky


def _fulspec_ifft(fk):
    """real grid of a half-plane spectrum (numpy; used only to normalise synthetic fields)"""
    nkx, nky = fk.shape
    nx = nkx + 1
    kmax = nky - 1
    full = np.zeros((nx, nx), dtype=np.complex128)
    up = fk.copy()
    up[kmax - 1::-1, 0] = np.conj(up[kmax + 1:, 0])
    up[kmax, 0] = up[kmax, 0].real
    # kx index i -> frequency i - kmax; place into FFT order
    fx = np.arange(-kmax, kmax + 1) % nx
    fy = np.arange(0, kmax + 1)
    full[np.ix_(fx, fy)] = up
    full[np.ix_((-np.arange(-kmax, kmax + 1)) % nx, (-fy[1:]) % nx)] = np.conj(up[:, 1:])
    return (nx * nx * np.fft.ifft2(full)).real


@dataclass
class Workload:
    name: str
    nx: int
    L: float
    f: float
    Cg: float
    n_packets: int
    psik: np.ndarray                 # frame 0, (nx-1, nx/2) complex
    psik2: np.ndarray | None         # frame 1 (time-dependent configs)
    u_mean: float
    x: np.ndarray
    y: np.ndarray
    k: np.ndarray
    l: np.ndarray
    dt: float
    U0: float
    scheme: str = "leapfrog"
    extra: dict = field(default_factory=dict)

    @property
    def gH(self):
        return self.Cg ** 2

    @property
    def dx(self):
        return self.L / self.nx


def synthetic_psik(nx, L, U_g, K_d2, seed, kind="full", k_max=8):
    rs = np.random.RandomState(seed)
    kx, ky = wavenumbers(nx)
    kappa = 2 * np.pi / L
    K2 = (kappa * kx) ** 2 + (kappa * ky) ** 2
    phase = 2 * np.pi * rs.rand(*kx.shape)
    if kind == "full":
        amp = 1.0 / np.maximum(K2, kappa ** 2) ** 1.5
        amp[(kx == 0) & (ky == 0)] = 0.0
    elif kind == "band":
        amp = ((np.abs(kx) <= k_max) & (np.abs(ky) <= k_max)).astype(np.float64)
        amp[(kx == 0) & (ky == 0)] = 0.0
        amp = amp / (K_d2 + K2)          # psi = -q/(K_d2+K2) with |q| ~ (K_d2 + K2)/(K_d2+K2)...
    else:
        raise ValueError(kind)
    psik = amp * np.exp(1j * phase)
    u = _fulspec_ifft(-1j * kappa * ky * psik)
    v = _fulspec_ifft(1j * kappa * kx * psik)
    U0 = np.sqrt((u * u + v * v).max())
    return psik * (U_g / U0), phase, amp * (U_g / U0)


def make_packets(n, L, f, Cg, nif, seed):
    rs = np.random.RandomState(seed)
    k0 = np.sqrt((nif ** 2 - 1) * f ** 2 / Cg ** 2)
    i = np.arange(1, n + 1)
    r = rs.rand(n, 2)
    return (L * r[:, 0] - L / 2, L * r[:, 1] - L / 2, k0 * np.cos(2 * np.pi * i / n), k0 * np.sin(2 * np.pi * i / n))


def make_workload(name, n_packets=None, nx=None, seed_field=146, seed_packets=123, kind="full"):
    """C1..C5 of BASELINE.json (sizes overridable for tests)."""
    f, Cg, nif = 3.0, 1.0, 2.0
    K_d2 = f / Cg
    if name == "C1":       # SW_zero_background_raytracing: zero flow, 1k packets
        nx = nx or 64; n = n_packets or 1000; L = 2 * np.pi
        psik = np.zeros((nx - 1, nx // 2), dtype=np.complex128)
        x, y, k, l = make_packets(n, L, f, Cg, nif, seed_packets)
        dt = 0.1 * (L / nx) / Cg
        return Workload(name, nx, L, f, Cg, n, psik, None, 0.0, x, y, k, l, dt, 0.0)
    if name == "C2":       # symplectic_full_fourier: steady 128^2, 64k packets
        nx = nx or 128; n = n_packets or 65536; L = 2 * np.pi; U_g = 0.5
        psik, _, _ = synthetic_psik(nx, L, U_g, K_d2, seed_field, kind)
        x, y, k, l = make_packets(n, L, f, Cg, nif, seed_packets)
        dt = 0.1 * (L / nx) / max(Cg, U_g)                 # symplectic_full_fourier.m:36
        return Workload(name, nx, L, f, Cg, n, psik, None, 0.0, x, y, k, l, dt, U_g)
    if name in ("C3", "C4"):   # time-dependent, two frames blended on the device
        if name == "C3":
            nx = nx or 256; n = n_packets or 1048576; L = 2 * np.pi; U_g = 0.5; cfl = 0.05; shear = 0.0
        else:
            nx = nx or 512; n = n_packets or 16777216; L = 20.0; U_g = 0.5; cfl = 0.25; shear = 0.5
        psik, phase, amp = synthetic_psik(nx, L, U_g, K_d2, seed_field if name == "C3" else 5, kind, k_max=8 if name == "C3" else 30)
        rs = np.random.RandomState(seed_field + 1)
        dphi = rs.uniform(-0.05, 0.05, size=phase.shape)
        psik2 = amp * np.exp(1j * (phase + dphi))
        x, y, k, l = make_packets(n, L, f, Cg, nif, seed_packets)
        dt = cfl * (L / nx) / (U_g + abs(shear))           # qgsw_raytrace.m:29,70 / qg2layersw_raytrace.m:31,78
        return Workload(name, nx, L, f, Cg, n, psik, psik2, shear, x, y, k, l, dt, U_g)
    if name == "C5":       # raytrace_sw + step_packet_xka on a synthetic geostrophic [u,v,eta] state
        nx = nx or 256; n = n_packets or 4194304; L = 2 * np.pi; U_g = 0.25
        psik, _, _ = synthetic_psik(nx, L, U_g, K_d2, seed_field, "band" if kind == "band" else kind, k_max=8)
        # geostrophic balance: eta_g = f/gH0 * psi  (raytrace_sw.m:30-35 inverted), H = 1 + eta_g
        etak = (f / Cg ** 2) * psik
        rs = np.random.RandomState(seed_packets)
        ki = 10 * f / Cg                                   # raytrace_sw.m:86-87
        i = np.arange(1, n + 1)
        x = rs.rand(n) * L; y = rs.rand(n) * L
        k = ki * np.cos(2 * np.pi * i / n); l = ki * np.sin(2 * np.pi * i / n)
        dt = 0.3 * (L / nx) / max(Cg, U_g)                 # raytrace_sw.m:102
        return Workload(name, nx, L, f, Cg, n, psik, None, 0.0, x, y, k, l, dt, U_g, scheme="rk4_xka", extra={"etak": etak})
    raise ValueError(f"unknown workload {name}")


def planes_from_psik(psik, L, u_mean=0.0, etak=None):
    """the six (seven) coefficient planes u,v,ux,uy,vx,vy[,H] of SpectralScheme.m:18-25 / grid_U.m:3-11"""
    nx = psik.shape[0] + 1
    kx, ky = wavenumbers(nx)
    kappa = 2 * np.pi / L
    kx = kappa * kx; ky = kappa * ky
    uk = -1j * ky * psik
    vk = 1j * kx * psik
    planes = [uk, vk, 1j * kx * uk, 1j * ky * uk, 1j * kx * vk, 1j * ky * vk]
    kmax = nx // 2 - 1
    planes[0] = planes[0].copy(); planes[0][kmax, 0] += u_mean
    if etak is not None:
        Hk = np.array(etak, dtype=np.complex128); Hk[kmax, 0] += 1.0
        planes.append(Hk)
    return planes
