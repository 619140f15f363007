"""Generates tests/golden/hotpath_nx32.npz -- regression vectors for the hot path.

The reference is MATLAB and cannot be executed in this image (no MATLAB/Octave), so these are NOT
reference outputs: they are produced by the CPU oracle (oracle/swrt_oracle.py, the line-by-line
restatement pinned by tests/test_oracle_kat.py) on seeded inputs, and committed so that (a) the C
restatement and the CUDA path are checked against fixed numbers on any box and (b) an accidental
change of the oracle itself is caught.  Run from the repo root:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import swrt_oracle as O  # noqa: E402


def main():
    nx = 32; L = 2 * np.pi; dx = L / nx; f = 3.0; gH = 1.0
    rs = np.random.RandomState(2024)
    kx_, ky_ = O.wavenumbers(nx)
    K2 = kx_ ** 2 + ky_ ** 2
    psik = (rs.randn(nx - 1, nx // 2) + 1j * rs.randn(nx - 1, nx // 2)) / (1 + K2) ** 1.5 * 0.3
    psik2 = psik * np.exp(1j * rs.uniform(-0.05, 0.05, psik.shape))
    planes = O.velocity_planes_k(psik, kx_, ky_)
    planes2 = O.velocity_planes_k(psik2, kx_, ky_)
    grids = [O.k2g(p) for p in planes]
    grids2 = [O.k2g(p) for p in planes2]
    n = 203
    x = rs.uniform(-2 * L, 2 * L, n); y = rs.uniform(-2 * L, 2 * L, n)
    k = 3 * np.cos(2 * np.pi * np.arange(1, n + 1) / n); l = 3 * np.sin(2 * np.pi * np.arange(1, n + 1) / n)
    a = np.ones(n)
    out = dict(nx=nx, L=L, f=f, gH=gH, psik=psik, psik2=psik2, x=x, y=y, k=k, l=l)
    out["grids"] = np.stack(grids)
    out["eval_spectral"] = O.spectral_eval_planes(x, y, planes, dx, nx)
    out["eval_lagrange"] = np.stack([O.interpolate(x, y, g, dx, dx) for g in grids])
    alpha = 0.3
    bf1 = dict(zip(("u", "v", "ux", "uy", "vx", "vy"), grids)); bf2 = dict(zip(("u", "v", "ux", "uy", "vx", "vy"), grids2))
    U, nab = O.interpolate_U(bf1, bf2, alpha, np.stack([x, y], axis=1), dx, bump=O.BUMP_LIVE)     # engine-level vectors: the engine's default bump
    out["alpha"] = alpha
    out["interpU_lagrange"] = np.stack([U[:, 0], U[:, 1], nab["u_x"], nab["u_y"], nab["v_x"], nab["v_y"]])
    out["rhs_lagrange"] = np.stack(O.odefun_rhs(x, y, k, l, alpha, bf1, bf2, f, 1.0, dx, bump=O.BUMP_LIVE))
    sp1 = O.spectral_eval_planes(x, y, planes, dx, nx); sp2 = O.spectral_eval_planes(x, y, planes2, dx, nx)
    out["interpU_spectral"] = (1 - alpha) * sp1 + alpha * sp2
    dt = 0.1 * dx
    out["dt"] = dt
    for mode, ev in (("lagrange", lambda xx, yy: np.stack([O.interpolate(xx, yy, g, dx, dx) for g in grids])),
                     ("spectral", lambda xx, yy: O.spectral_eval_planes(xx, yy, planes, dx, nx))):
        xs, ys, ks, ls = x, y, k, l
        for _ in range(20):
            xs, ys, ks, ls = O.leapfrog_step(xs, ys, ks, ls, dt, f, gH, ev)
        out[f"leapfrog20_{mode}"] = np.stack([xs, ys, ks, ls])
    H = 1 + 0.1 * O.k2g(psik)
    out["H"] = H
    flds = dict(zip(("u", "v", "u_x", "u_y", "v_x", "v_y"), grids)); flds["H"] = H
    for xka in (False, True):
        st = (x, y, k, l, a)
        for _ in range(3):
            st = O.rk4_step_batch(*st, dt, 1.0, f, flds, dx, xka)
        out["rk4x3_xka_lagrange" if xka else "rk4x3_packet_lagrange"] = np.stack(st)
    w = O.omega_of_k(k, l, f, gH)
    edges = O.matlab_linspace(0, w.max(), 30)
    out["hist_edges"] = edges
    out["hist_counts"] = O.histcounts(w, edges)
    np.savez_compressed(Path(__file__).with_name("hotpath_nx32.npz"), **out)
    print("wrote", Path(__file__).with_name("hotpath_nx32.npz"))


if __name__ == "__main__":
    main()
