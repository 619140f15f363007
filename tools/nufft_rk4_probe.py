"""Developer probe: step_packet / step_packet_xka in NUFFT mode, fused kernel vs composed launches (run under gpurun)."""
import sys; sys.path.insert(0, '.')
import numpy as np
import swraytracing_b200 as S
from swraytracing_b200 import workloads as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4194304
w = W.make_workload("C5", n_packets=n)
planes = W.planes_from_psik(w.psik, w.L, w.u_mean, etak=w.extra["etak"])
for scheme, name in ((S.SCHEME_RK4_XKA, "step_packet_xka"), (S.SCHEME_RK4_PACKET, "step_packet")):
    for unfused in (True, False):
        with S.Engine(w.nx, w.L, w.f, w.gH, S.MODE_NUFFT) as e:
            e.set_tuning(unfused_rk4=unfused)
            e.set_flow_planes_spectral(planes)
            e.set_packets(w.x, w.y, w.k, w.l, np.ones(n))
            e.step(scheme, w.dt, 2)
            best = 1e9
            for r in range(3):
                e.timer_start(); e.step(scheme, w.dt, 4); best = min(best, e.timer_stop())
            print(f"{name} nx={w.nx} n={n} {'composed' if unfused else 'fused'}: {best:.3f} ms / 4 steps = {n * 4 / (best * 1e-3):.3e} packet-steps/s", flush=True)
