"""Small end-to-end exercise of every kernel in all three modes (written as a compute-sanitizer memcheck target; the sanitizer is
closed on this GPU pool, so it runs plain and the parity tests are the bounds check)."""
import sys; sys.path.insert(0, '.')
import numpy as np
import swraytracing_b200 as S
from swraytracing_b200 import workloads as W
for mode in (S.MODE_SPECTRAL, S.MODE_LAGRANGE6, S.MODE_NUFFT):
    w = W.make_workload("C5", n_packets=333, nx=32)
    planes = W.planes_from_psik(w.psik, w.L, etak=w.extra["etak"])
    e = S.Engine(w.nx, w.L, w.f, w.gH, mode)
    e.set_flow_planes_spectral(planes)
    e.set_packets(w.x, w.y, w.k, w.l)
    e.eval(); e.rhs(); e.step(S.SCHEME_RK4_XKA, w.dt, 2); e.step(S.SCHEME_RK4_PACKET, w.dt, 2); e.step(S.SCHEME_LEAPFROG, w.dt, 3)
    e.hist_omega(np.linspace(0, 40, 30), kind=S.HIST_ABSOLUTE); e.diag(); e.omega()
    w3 = W.make_workload("C3", n_packets=130, nx=32)
    e2 = S.Engine(w3.nx, w3.L, w3.f, w3.gH, mode)
    e2.set_flow_spectral(w3.psik, 0); e2.set_flow_spectral(w3.psik2, 1)
    e2.set_packets(w3.x, w3.y, w3.k, w3.l)
    for mt in (1, 2):
        e2.set_tuning(mt); e2.step(S.SCHEME_LEAPFROG, w3.dt, 3, 0.1, 0.2); e2.eval(0.5)
        e2.set_tuning(mt, use_psi_moments=False); e2.step(S.SCHEME_LEAPFROG, w3.dt, 2)
    # two frames with H (32-byte NUFFT nodes / odd-length Lagrange records, two-frame sweeps), awkward grid size
    w5 = W.make_workload("C5", n_packets=77, nx=36)
    p5 = W.planes_from_psik(w5.psik, w5.L, etak=w5.extra["etak"])
    e3 = S.Engine(w5.nx, w5.L, w5.f, w5.gH, mode)
    e3.set_flow_planes_spectral(p5, 0); e3.set_flow_planes_spectral([0.9 * p for p in p5], 1)
    e3.set_packets(w5.x, w5.y, w5.k, w5.l, np.ones(77))
    e3.eval_at(w5.x, w5.y, 0.3, with_H=True); e3.step(S.SCHEME_RK4_XKA, w5.dt, 2, 0.25, 0.5); e3.step(S.SCHEME_LEAPFROG, w5.dt, 2, 0.25, 0.5)
    print("mode", mode, "ok", np.isfinite(np.stack(e2.get_packets())).all(), np.isfinite(np.stack(e3.get_packets(with_a=True))).all())
S.g2k_dev(np.random.rand(32, 32)); S.interpolate_dev(np.zeros(5), np.zeros(5), np.random.rand(16, 16), 0.1, 0.1)
print("done")
