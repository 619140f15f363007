"""Developer perf probe / ncu driver for the NUFFT-mode leapfrog kernel."""
import sys; sys.path.insert(0, '.')
import swraytracing_b200 as S
from swraytracing_b200 import workloads as W
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
npk = int(sys.argv[2]) if len(sys.argv) > 2 else None
w = W.make_workload(name, n_packets=npk)
e = S.Engine(w.nx, w.L, w.f, w.gH, S.MODE_NUFFT)
e.set_flow_spectral(w.psik, u_mean=w.u_mean)
e.set_packets(w.x, w.y, w.k, w.l)
for _ in range(3):
    e.step(S.SCHEME_LEAPFROG, w.dt, 16)
    ms, _n = e.last_kernel_ms()
    print(f"{name} n={w.n_packets}: kernel {ms:.3f} ms  {w.n_packets * 16 / (ms * 1e-3):.3e} packet-steps/s")
