"""Developer probe: NUFFT-mode leapfrog throughput (C2 / C3 fields) for A/B library builds (SWRT_LIB=...)."""
import sys; sys.path.insert(0, '.')
import numpy as np
import swraytracing_b200 as S
from swraytracing_b200 import workloads as W
for cfg in ("C2", "C3"):
    w = W.make_workload(cfg)
    n = w.n_packets
    with S.Engine(w.nx, w.L, w.f, w.gH, S.MODE_NUFFT) as e:
        e.set_flow_spectral(w.psik, 0, u_mean=w.u_mean)
        e.set_packets(w.x, w.y, w.k, w.l)
        e.step(S.SCHEME_LEAPFROG, w.dt, 4)
        best = 1e9
        for r in range(4):
            e.timer_start(); e.step(S.SCHEME_LEAPFROG, w.dt, 16); best = min(best, e.timer_stop())
        st = np.stack(e.get_packets())
        print(f"{cfg} nx={w.nx} n={n}: {best:.3f} ms / 16 steps = {n * 16 / (best * 1e-3):.3e} packet-steps/s  checksum {float(np.abs(st).sum()):.12e}", flush=True)
