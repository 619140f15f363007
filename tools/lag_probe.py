"""Developer probe: LAGRANGE6 leapfrog, steady (C2, C3 field) and two-frame (C3), for A/B library builds (SWRT_LIB=...)."""
import sys; sys.path.insert(0, '.')
import numpy as np
import swraytracing_b200 as S
from swraytracing_b200 import workloads as W
for cfg, two in (("C2", False), ("C3", False), ("C3", True)):
    w = W.make_workload(cfg)
    n = w.n_packets
    with S.Engine(w.nx, w.L, w.f, w.gH, S.MODE_LAGRANGE6) as e:
        e.set_flow_spectral(w.psik, 0, u_mean=w.u_mean)
        if two:
            e.set_flow_spectral(w.psik2, 1, u_mean=w.u_mean)
        e.set_packets(w.x, w.y, w.k, w.l)
        args = (1 / 32, 1 / 16) if two else (0.0, 0.0)
        e.step(S.SCHEME_LEAPFROG, w.dt, 4, *args)
        best = 1e9
        for r in range(4):
            e.timer_start(); e.step(S.SCHEME_LEAPFROG, w.dt, 16, *args); best = min(best, e.timer_stop())
        st = np.stack(e.get_packets())
        print(f"{cfg} two_frames={two} nx={w.nx} n={n}: {best:.3f} ms / 16 steps = {n * 16 / (best * 1e-3):.3e} packet-steps/s  checksum {float(np.abs(st).sum()):.15e}", flush=True)
