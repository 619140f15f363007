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


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=str(HERE / "reference_c2_script.npz"))
    ap.add_argument("--nx", type=int, default=NX, help="grid size written into the log (a small one makes a quick dry run)")
    ap.add_argument("--amp", type=float, default=0.67, help="PV amplitude (sets U0, hence Tend / dt)")
    a = ap.parse_args(argv)
    nx = a.nx
    ref = Path(a.ref)
    tmp = Path(tempfile.mkdtemp(prefix="swrt_c2_"))
    for d in ("qg_flow_ray_trace", "rsw", "ray_trace_sw"):
        os.symlink(ref / d, tmp / d)
    logdir = tmp / "analysis" / "job-36976465" / "run-4"
    logdir.mkdir(parents=True)
    log = (ref / "analysis" / "job-36976465" / "run-4" / "run.log").read_text(errors="replace")
    log128, nsub = re.subn(r"Resolution: 256x256", f"Resolution: {nx}x{nx}", log, count=1)
    assert nsub == 1
    (logdir / "run.log").write_text(log128)
    q = pv_frame(nx, amp=a.amp)
    with open(tmp / "analysis" / "pv.bin", "wb") as fh:           # frame 2000 of a frame-addressed real*8 stream (read_field.m:88)
        fh.seek(8 * nx * nx * 1999)
        fh.write(np.asfortranarray(q).tobytes(order="F"))
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
