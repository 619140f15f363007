"""The reference's production driver with everything on the device: ``qgsw_raytrace``.

``qgsw_raytrace(nx, Npackets, near_inertial_factor, T_Fr_days, packet_delay_days, U_g, f, Cg)`` keeps the
signature and control flow of qg_flow_ray_trace/qgsw_raytrace.m:1-180 -- initial PV, packet ring, dt from
the initial flow, the AB3 QG loop, per-step packet advection between the previous and the new flow frame,
packet frames every ``packet_steps_per_save`` steps, PV frames every ``steps_per_save`` steps -- but the QG
solver (engine.QGFlow), the flow-frame construction, and the packet integrator run on the GPU and the
packet state only comes back to the host when a frame is written.  Host work that remains is set-up
(random phases, file I/O); the reference's log header is reproduced so ``parse_data``
(symplectic_full_fourier.m:66-82) still reads it.
"""
from __future__ import annotations

import math
import time

import numpy as np

from . import fieldio
from .engine import Engine, QGFlow, MODE_SPECTRAL, SCHEME_LEAPFROG, k2g_dev, g2k_dev
from .reference_api import ode23


def wavenumber_grids(nx):
    """[kx_,ky_] = ndgrid(-kmax:kmax, 0:kmax) (qgsw_raytrace.m:18-20)"""
    kmax = nx // 2 - 1
    kx = np.arange(-kmax, kmax + 1, dtype=np.float64)[:, None] * np.ones((1, kmax + 1))
    ky = np.ones((2 * kmax + 1, 1)) * np.arange(0, kmax + 1, dtype=np.float64)[None, :]
    return kx, ky


def initial_q(X, Y, a_g, K_d2, rs, k_min=5, k_max=8, ring=False):
    """qgsw_raytrace.m:191-214.  The chained comparison on :202 is always true in MATLAB, so every
    |k|,|l| <= k_max mode is summed (``ring=True`` gives the evidently intended annulus).
    ``rs``: numpy RandomState seeded like ``rng(146)`` (same mt19937 stream, column-major fill)."""
    q = np.zeros_like(X); U = np.zeros_like(X); V = np.zeros_like(X)
    n = 2 * k_max + 1
    phase = 2 * np.pi * rs.rand(n, n).T
    for k in range(-k_max, k_max + 1):
        for l in range(-k_max, k_max + 1):
            K2 = k * k + l * l
            if (not ring) or (k_min ** 2 < K2 <= k_max ** 2):
                wp = k * X + l * Y + phase[k + k_max, l + k_max]
                U -= l * np.sin(wp); V += k * np.sin(wp)
                q -= (K_d2 + K2) * np.cos(wp)
    return a_g / np.sqrt((U * U + V * V).max()) * q


def qgsw_raytrace(nx, Npackets, near_inertial_factor, T_Fr_days, packet_delay_days, U_g, f, Cg, *, outdir="data",
                  integrator="ode23", mode=MODE_SPECTRAL, r_drag=0.1, beta=0.0, force_strength=0.1, max_steps=None,
                  seed=146, device=0, log=print, leapfrog_substeps=4):
    """qgsw_raytrace.m:1-180 on the device.  ``integrator``: 'ode23' (the reference) or 'leapfrog' (the fused
    symplectic stepper with time-centred frame blending).  ``max_steps`` truncates the run (tests)."""
    L = 2 * np.pi
    dx = L / nx
    xg = np.linspace(-L / 2, L / 2, nx)
    X, Y = np.meshgrid(xg, xg)
    kx_, ky_ = wavenumber_grids(nx)
    K2 = kx_ ** 2 + ky_ ** 2
    rs = np.random.RandomState(seed)                                   # rng(146)
    K_d2 = f / Cg
    T_days = T_Fr_days / f
    CFL_fraction = 0.05
    steps_per_save, packet_steps_per_save = 50, 5
    packet_delay = packet_delay_days / f

    q = initial_q(X, Y, U_g, K_d2, rs)
    qk = g2k_dev(q, device)
    wavenumber_factor = math.sqrt((near_inertial_factor ** 2 - 1) * f ** 2 / Cg ** 2)
    i = np.arange(1, Npackets + 1)
    pk = wavenumber_factor * np.cos(2 * np.pi * i / Npackets)
    pl = wavenumber_factor * np.sin(2 * np.pi * i / Npackets)
    r = rs.rand(Npackets, 2)                                           # rand(1,2) per packet, x then y
    px, py = L * r[:, 0] - L / 2, L * r[:, 1] - L / 2

    # time step and Froude number from the initial flow (qgsw_raytrace.m:62-73)
    psik = -qk / (K_d2 + K2)
    u0 = k2g_dev(-1j * ky_ * psik, device); v0 = k2g_dev(1j * kx_ * psik, device)
    U0 = math.sqrt((u0 * u0 + v0 * v0).max())
    Fr = U0 / Cg
    T = T_days / Fr ** 2
    dt = CFL_fraction * dx / U0
    Nsteps = int(math.ceil(T / dt))
    if max_steps is not None:
        Nsteps = min(Nsteps, int(max_steps))
    packet_step_start = int(math.ceil(packet_delay / dt))

    log("Resolution: %dx%d" % (nx, nx)); log("Number of packets: %d" % Npackets)
    log("Initial wavenumber radius: %f" % (near_inertial_factor * f)); log("Time step: %f" % dt)
    log("Simulation time: %f" % T); log("Spin-up time: %f" % packet_delay)
    log("Steps per save: %d" % steps_per_save); log("Steps per packet save: %d" % packet_steps_per_save)
    log("Coriolis parameter: %f" % f); log("Group velocity: %f" % Cg)
    log("Background velocity (parameter,computed): (%f,%f)" % (U_g, U0)); log("Froude Number: %f" % Fr)
    log("Deformation wavenumber: %f" % K_d2)

    writer = fieldio.PacketFrameWriter(outdir, L)
    writer.write(px, py, pk, pl, dt * (packet_step_start - 1))         # initial positions, :103-106
    fieldio.write_field(q, f"{outdir}/pv", 1); fieldio.write_field(0.0, f"{outdir}/pv_time", 1)

    qg = QGFlow(nx, L, qk, K_d2, dt, f, Cg, beta=beta, r_drag=r_drag, force_strength=force_strength, device=device)
    eng = Engine(nx, L, f, Cg ** 2, mode, device)
    eng.set_packets(px, py, pk, pl)
    t = 0.0
    tic = time.time()
    stats = {"packet_steps": 0, "ode23_steps": 0, "ode23_failed": 0}
    for step in range(1, Nsteps + 1):
        if Npackets > 0 and t + dt > packet_delay:
            qg.to_flow(eng, 0)                                         # background_flow1 = grid_U(prev_qk), :141
        qg.step(1)
        t += dt
        if Npackets > 0 and t > packet_delay:
            qg.to_flow(eng, 1)                                         # background_flow2 = grid_U(qk), :142
            if integrator == "ode23":
                st = ode23(eng, [0.0, dt], dt)                         # :143-150
                stats["ode23_steps"] += st["nsteps"]; stats["ode23_failed"] += st["nfailed"]
            else:
                m = leapfrog_substeps
                eng.step(SCHEME_LEAPFROG, dt / m, m, 0.5 / m, 1.0 / m)
            stats["packet_steps"] += 1
            if (step - packet_step_start + 1) % packet_steps_per_save == 0:
                writer.write(*eng.get_packets(), t)                    # wrapped on save only, :160
        if step % steps_per_save == 0:
            fieldio.write_field(k2g_dev(qg.get(), device), f"{outdir}/pv", 0)
            fieldio.write_field(t, f"{outdir}/pv_time", 0)
    log("Real time elapsed: %.3f seconds" % (time.time() - tic))
    out = {"dt": dt, "Nsteps": Nsteps, "U0": U0, "Fr": Fr, "t": t, "packets": eng.get_packets(), "qk": qg.get(), **stats,
           "packet_frames": writer.frames}
    qg.close(); eng.close()
    return out
