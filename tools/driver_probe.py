"""Whole-driver timing on the reference's OWN problem size: qgsw_raytrace(256, 50, 2, ...) -- 50 packets, 256^2 one-layer QG,
ode23 over every QG step (qgsw_raytrace.m:121-176).  The reference's run logs give 3132.8 s for 137,599 steps (55,039 of
them advecting packets): ~5.8 ms per QG-only step and ~54 ms per step with packets.  Here: ms per
step of the device driver for both kinds of step, per evaluation mode."""
import sys, time, tempfile; sys.path.insert(0, '.')
import numpy as np
import swraytracing_b200 as S
from swraytracing_b200 import drivers
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
for label, delay in (("QG-only steps (packets not yet released)", 1e9), ("steps advecting 50 packets with ode23", 0.0)):
    for mname, mode in (("SPECTRAL", S.MODE_SPECTRAL), ("NUFFT", S.MODE_NUFFT), ("LAGRANGE6", S.MODE_LAGRANGE6)):
        if delay > 1 and mname != "SPECTRAL":
            continue
        with tempfile.TemporaryDirectory() as d:
            # set-up (initial_q over 289 modes, engine creation) is paid once: time two run lengths and take the difference
            drivers.qgsw_raytrace(256, 50, 2, 6000, delay, 0.5, 3.0, 1.0, outdir=d, max_steps=5, r_drag=0.0, force_strength=0.0, mode=mode, log=lambda s: None)
            t0 = time.time()
            drivers.qgsw_raytrace(256, 50, 2, 6000, delay, 0.5, 3.0, 1.0, outdir=d, max_steps=20, r_drag=0.0, force_strength=0.0, mode=mode, log=lambda s: None)
            t20 = time.time() - t0
            t0 = time.time()
            out = drivers.qgsw_raytrace(256, 50, 2, 6000, delay, 0.5, 3.0, 1.0, outdir=d, max_steps=nsteps + 20, r_drag=0.0, force_strength=0.0, mode=mode,
                                        log=lambda s: None)
            el = time.time() - t0 - t20
        fin = bool(np.isfinite(np.stack(out["packets"])).all())
        print(f"{label:45s} {mname:9s}: {1e3 * el / nsteps:7.3f} ms/step  (ode23 steps/QG step {out['ode23_steps'] / max(1, out['packet_steps']):.1f}, finite {fin})")
