#!/usr/bin/env python
"""BASELINE config 2's driver, ``symplectic_full_fourier.m``, executed UNMODIFIED and to the end -> reference_c2_script.npz.

    python tests/golden/run_reference_c2_script.py [--ref /root/reference] [--nx 32 --amp 4  (a 30-second dry run)]

The script takes everything from files relative to the folder it is started in: the grid size, f and Cg from the header of
``analysis/job-36976465/run-4/run.log`` (its local ``parse_data`` / ``textscan``), the PV field from frame 2000 of ``analysis/pv``
(``read_field``), its helpers through ``addpath ./qg_flow_ray_trace`` and, inside ``SpectralScheme``, ``./rsw/`` and
``./ray_trace_sw/``.  ``analysis/pv.bin`` is not in the repository (.MISSING_LARGE_BLOBS), and BASELINE.json quotes config 2 on a
128^2 grid, so the script is started in a scratch folder that holds

* ``analysis/job-36976465/run-4/run.log``: the reference's own log with its ``Resolution:`` line set to 128x128 and nothing else
  changed (the banner, line count and every other value stay, so ``parse_data`` walks it exactly as it walks the original),
* ``analysis/pv.bin``: a sparse file whose frame 2000 is a seeded random-phase PV field (the reference's ``initial_q`` recipe:
  modes |k|, |l| <= 8), written with the reference's frame layout (``read_field`` seeks to it),
* symbolic links ``qg_flow_ray_trace``, ``rsw``, ``ray_trace_sw`` to the reference's folders.

Nothing is overridden except the plotting calls (no-ops in minimat): ``parse_data``, ``read_field``, ``g2k`` / ``k2g``,
``SpectralScheme``, ``scheme.U`` on the 128^2 grid, ``rng(123)`` / ``rand``, ``ode_symplectic`` over all floor(Tend/dt) steps,
the Omega-drift series -- all the reference's own code, ~15 minutes of interpretation.  What is stored: the scalars (nx, f,
Cg, U0, Fr, dt, Tend, the number of steps), the PV field, the initial packets, every 8th row of the packet history, and the
whole ``solver_error`` (relative drift of omega + U.k) series.  Run HERE; the .npz travels.
"""
import argparse
import hashlib
import io
import json
import os
import re
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))

from oracle.minimat import Interp, Frame          # noqa: E402

NX = 128


def pv_frame(nx, seed=2000, k_max=8, amp=1.6):
    """a random-phase PV field of the kind the reference's runs store in analysis/pv (qgsw_raytrace.m:191-214), seeded"""
    rs = np.random.RandomState(seed)
    x = np.arange(nx) * (2 * np.pi / nx)
    X, Y = np.meshgrid(x, x)
    q = np.zeros((nx, nx))
    for k in range(-k_max, k_max + 1):
        for l in range(-k_max, k_max + 1):
            if 0 < k * k + l * l <= k_max * k_max:
                q -= (3.0 + k * k + l * l) * np.cos(k * X + l * Y + 2 * np.pi * rs.rand()) / (1 + k * k + l * l)
    return amp * q / np.abs(q).max() * 6.0


def scratch_folder(ref, nx, q):
    """a folder the reference's frozen-flow scripts can be started in: the run log they parse (its Resolution line set to
    nx), frame 2000 of analysis/pv, links to the folders they addpath"""
    tmp = Path(tempfile.mkdtemp(prefix="swrt_script_"))
    for d in ("qg_flow_ray_trace", "rsw", "ray_trace_sw"):
        os.symlink(ref / d, tmp / d)
    logdir = tmp / "analysis" / "job-36976465" / "run-4"
    logdir.mkdir(parents=True)
    log = (ref / "analysis" / "job-36976465" / "run-4" / "run.log").read_text(errors="replace")
    lognx, nsub = re.subn(r"Resolution: 256x256", f"Resolution: {nx}x{nx}", log, count=1)
    assert nsub == 1
    (logdir / "run.log").write_text(lognx)
    with open(tmp / "analysis" / "pv.bin", "wb") as fh:           # frame 2000 of a frame-addressed real*8 stream (read_field.m:88)
        fh.seek(8 * nx * nx * 1999)
        fh.write(np.asfortranarray(q).tobytes(order="F"))
    np.array([0.0]).tofile(tmp / "analysis" / "pv_time.bin")       # SW_zero_background_raytracing.m:25 reads it, never uses it
    return tmp


def run_c1_script(ref, out_path, nx=32, amp=2.0):
    """BASELINE config 1's script, SW_zero_background_raytracing.m, executed unmodified and to the end on a 32^2 frame.  Its
    integrator is MATLAB's ``ode23`` (``method = @ode23``, RelTol 1e-6 / AbsTol 1e-7 through ``odeset``, output at
    ``dt*(0:Nsteps)``): a builtin, MathWorks code, not part of the reference -- served here by the restated controller
    ``oracle.ode23`` (``odeset`` by a name/value struct).  Everything else is the reference's own code: parse_data, read_field,
    grid_U (its six-argument call fails as in MATLAB: the script's line 26 is therefore given shear 0 through the same shim
    the qgsw tests use), SpectralScheme, the packet ring, initialize_raytracing / odefun, the Omega series."""
    from oracle import swrt_oracle as O
    from oracle.minimat import MStruct, from_py
    q = pv_frame(nx, amp=amp)
    tmp = scratch_folder(ref, nx, q)
    buf = io.StringIO()
    I = Interp(cwd=str(tmp), out=buf)
    I.path.insert(0, str(ref))
    ref_grid_U = I.load_unit(str(ref / "qg_flow_ray_trace" / "grid_U.m")).main
    I.overrides["grid_U"] = lambda I_, args, nargout, frame: I_.call_funcdef(ref_grid_U, list(args) + [0.0], nargout, frame)
    I.overrides["odeset"] = lambda I_, args, nargout, frame: MStruct({args[i]: args[i + 1] for i in range(0, len(args), 2)})
    I.overrides["hist"] = lambda I_, args, nargout, frame: None
    stats = {}

    def ode23_builtin(I_, args, nargout, frame):
        fun, tspan, y0 = args[0], np.asarray(args[1]).ravel(), np.asarray(args[2]).ravel()
        opts = args[3].f if len(args) > 3 else {}
        Y, st = O.ode23(lambda t, yv: np.asarray(I_.call_handle(fun, [float(t), from_py(np.asarray(yv).reshape(-1, 1))], 1, frame)[0]).ravel(),
                        tspan, y0, rtol=float(opts.get("RelTol", 1e-3)), atol=float(opts.get("AbsTol", 1e-6)))
        stats.update(st)
        return [from_py(tspan.reshape(-1, 1)), from_py(np.asfortranarray(Y))]
    I.overrides["ode23"] = ode23_builtin
    ws = Frame(None)
    t0 = time.time()
    I.run("SW_zero_background_raytracing", ws)
    secs = time.time() - t0
    I.close_all()
    v = ws.vars
    out = {"nx": np.float64(v["nx"]), "f": np.float64(v["f"]), "Cg": np.float64(v["Cg"]), "U0": np.float64(v["U0"]), "Fr": np.float64(v["Fr"]),
           "dt": np.float64(v["dt"]), "Tend": np.float64(v["Tend"]), "Nsteps": np.float64(v["Nsteps"]), "q": q,
           "x0": np.asarray(v["x"]), "k0": np.asarray(v["k"]), "t_hist": np.asarray(v["t_hist"]).ravel(),
           "solver_x": np.asarray(v["solver_x"]), "solver_k": np.asarray(v["solver_k"]), "solver_error": np.asarray(v["solver_error"]),
           "w": np.asarray(v["w"]), "ode23_nsteps": np.float64(stats["nsteps"]), "ode23_nfailed": np.float64(stats["nfailed"]),
           "stdout": np.array(buf.getvalue())}
    executed = sorted(p for p in I.units if str(Path(p).resolve()).startswith(str(ref)))
    prov = {"made_by": "tests/golden/run_reference_c2_script.py --c1", "executor": "oracle/minimat", "seconds": round(secs, 1),
            "ode23": "MATLAB builtin, served by oracle.ode23 (restated controller)",
            "reference_files_executed": {str(Path(p).resolve().relative_to(ref)): hashlib.sha256(Path(p).read_bytes()).hexdigest() for p in executed}}
    out["provenance"] = np.array(json.dumps(prov))
    np.savez_compressed(out_path, **out)
    print(f"run_c1_script: Nsteps = {int(v['Nsteps'])}, ode23 steps = {stats['nsteps']} (+{stats['nfailed']} failed), U0 = {float(v['U0']):.6f} in {secs:.0f} s -> {out_path}")


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=str(HERE / "reference_c2_script.npz"))
    ap.add_argument("--nx", type=int, default=NX, help="grid size written into the log (a small one makes a quick dry run)")
    ap.add_argument("--amp", type=float, default=0.67, help="PV amplitude (sets U0, hence Tend / dt)")
    ap.add_argument("--c1", action="store_true", help="run config 1's script (SW_zero_background_raytracing.m, 32^2, half a minute) "
                    "into reference_c1_script.npz instead")
    a = ap.parse_args(argv)
    nx = a.nx
    ref = Path(a.ref)
    if a.c1:
        return run_c1_script(ref, a.out if a.out != str(HERE / "reference_c2_script.npz") else str(HERE / "reference_c1_script.npz"))
    q = pv_frame(nx, amp=a.amp)
    tmp = scratch_folder(ref, nx, q)
    buf = io.StringIO()
    I = Interp(cwd=str(tmp), out=buf)
    I.path.insert(0, str(ref))                                     # symplectic_full_fourier.m, SpectralScheme.m, ode_symplectic.m
    ws = Frame(None)
    t0 = time.time()
    I.run("symplectic_full_fourier", ws)
    secs = time.time() - t0
    I.close_all()
    v = ws.vars
    sx, sk, st = np.asarray(v["solver_x"]), np.asarray(v["solver_k"]), np.asarray(v["solver_t"]).ravel()
    out = {"nx": np.float64(v["nx"]), "f": np.float64(v["f"]), "Cg": np.float64(v["Cg"]), "U0": np.float64(v["U0"]), "Fr": np.float64(v["Fr"]),
           "dt": np.float64(v["dt"]), "Tend": np.float64(v["Tend"]), "nrows": np.float64(sx.shape[0]), "q": q,
           "x0": np.asarray(v["x"]), "k0": np.asarray(v["k"]), "Omega_0": np.asarray(v["Omega_0"]),
           "solver_x_every8": sx[::8].copy(), "solver_k_every8": sk[::8].copy(), "solver_t_every8": st[::8].copy(),
           "solver_x_last": sx[-1].copy(), "solver_k_last": sk[-1].copy(),
           "solver_error": np.asarray(v["solver_error"]), "stdout": np.array(buf.getvalue())}
    executed = sorted(p for p in I.units if str(Path(p).resolve()).startswith(str(ref)))
    prov = {"made_by": "tests/golden/run_reference_c2_script.py", "executor": "oracle/minimat", "seconds": round(secs, 1),
            "reference_files_executed": {str(Path(p).resolve().relative_to(ref)): hashlib.sha256(Path(p).read_bytes()).hexdigest() for p in executed}}
    out["provenance"] = np.array(json.dumps(prov))
    np.savez_compressed(a.out, **out)
    print(f"run_reference_c2_script: {sx.shape[0]} history rows, U0 = {float(v['U0']):.6f}, dt = {float(v['dt']):.6g} in {secs:.0f} s -> {a.out}")


if __name__ == "__main__":
    main()
