"""The production integrator at scale, in the reference's own unit: one "packet-step" of qgsw_raytrace.m is ONE ode23 solve
over a QG step [0, dt] between two flow frames (qgsw_raytrace.m:141-150); the run logs give 610-960 of them per second
(BASELINE.md section 1).  Here: C3 (256^2, two frames, 1,048,576 packets), one solve per mode."""
import sys, time; sys.path.insert(0, '.')
import numpy as np
import swraytracing_b200 as S
from swraytracing_b200 import workloads as W, reference_api as R
w = W.make_workload(sys.argv[1] if len(sys.argv) > 1 else "C3")
for name, mode in (("NUFFT", S.MODE_NUFFT), ("LAGRANGE6", S.MODE_LAGRANGE6), ("SPECTRAL", S.MODE_SPECTRAL)):
    e = S.Engine(w.nx, w.L, w.f, w.gH, mode)
    e.set_flow_spectral(w.psik, 0, u_mean=w.u_mean); e.set_flow_spectral(w.psik2, 1, u_mean=w.u_mean)
    e.set_packets(w.x, w.y, w.k, w.l)
    R.ode23(e, [0.0, w.dt], w.dt)                       # warm-up (stacks, grids)
    e.set_packets(w.x, w.y, w.k, w.l)
    t0 = time.time(); st = R.ode23(e, [0.0, w.dt], w.dt); e.synchronize(); el = time.time() - t0
    print(f"{w.name} {name:9s}: one ode23 solve of {w.n_packets} packets: {el * 1e3:8.1f} ms ({st['nsteps']} steps, {st['nfailed']} rejected, "
          f"{st['nfevals']} RHS evaluations) -> {w.n_packets / el:.3e} packet-QG-steps/s")
    e.close()
