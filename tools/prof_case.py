"""Run a few calls of one hot-path kernel on a BASELINE workload -- the command ncu wraps for the per-kernel captures
committed under profiles/ (see profiles/README.md for the exact ncu command lines).

    python tools/prof_case.py C4 spectral --packets 262144 --substeps 2 --reps 2
"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import swraytracing_b200 as S
from swraytracing_b200 import workloads as W

ap = argparse.ArgumentParser()
ap.add_argument("workload")
ap.add_argument("mode", choices=["spectral", "lagrange6", "nufft"])
ap.add_argument("--packets", type=int, default=0)
ap.add_argument("--substeps", type=int, default=2)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--unfused", action="store_true")
ap.add_argument("--twiddles", type=int, default=0)
a = ap.parse_args()

w = W.make_workload(a.workload, n_packets=a.packets or None)
mode = {"spectral": S.MODE_SPECTRAL, "lagrange6": S.MODE_LAGRANGE6, "nufft": S.MODE_NUFFT}[a.mode]
eng = S.Engine(w.nx, w.L, w.f, w.gH, mode)
eng.set_tuning(unfused_rk4=a.unfused, twiddles=a.twiddles)
if w.scheme == "rk4_xka":
    eng.set_flow_planes_spectral(W.planes_from_psik(w.psik, w.L, w.u_mean, etak=w.extra["etak"]))
else:
    eng.set_flow_spectral(w.psik, 0, w.u_mean)
    if w.psik2 is not None:
        eng.set_flow_spectral(w.psik2, 1, w.u_mean)
td = w.psik2 is not None
scheme = {"leapfrog": S.SCHEME_LEAPFROG, "rk4_packet": S.SCHEME_RK4_PACKET, "rk4_xka": S.SCHEME_RK4_XKA}[w.scheme]
a0, da = (0.5 / a.substeps, 1.0 / a.substeps) if td else (0.0, 0.0)
dt = w.dt / a.substeps if td else w.dt
eng.set_packets(w.x, w.y, w.k, w.l)
for _ in range(a.reps):
    eng.step(scheme, dt, a.substeps, a0, da)
    ms, nl = eng.last_kernel_ms()
    tf = ""
    if a.mode == "spectral":
        pe = {"leapfrog": eng.contracted_planes(), "rk4_packet": eng.contracted_planes() + 6, "rk4_xka": 19}[w.scheme]
        tf = f" = {eng.work_per_eval(pe) * w.n_packets * a.substeps / ms * 1e-9:.2f} TFLOP/s"
    print(f"{a.workload} {a.mode} tw={a.twiddles}: {w.n_packets} packets x {a.substeps} steps: {ms:.3f} ms in {nl} launch(es) = "
          f"{w.n_packets * a.substeps / ms * 1e3:.4g} packet-steps/s{tf}", flush=True)
eng.close()
