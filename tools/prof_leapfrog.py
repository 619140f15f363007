"""Short driver for ncu captures: the fused spectral leapfrog kernel on the bench workload."""
import sys; sys.path.insert(0, '.')
import swraytracing_b200 as S
from swraytracing_b200 import workloads as W
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
sub = int(sys.argv[2]) if len(sys.argv) > 2 else 16
mode = S.MODE_LAGRANGE6 if (len(sys.argv) > 3 and sys.argv[3] == "lag") else S.MODE_SPECTRAL
npk = int(sys.argv[4]) if len(sys.argv) > 4 else None
w = W.make_workload(name, n_packets=npk)
e = S.Engine(w.nx, w.L, w.f, w.gH, mode)
e.set_flow_spectral(w.psik)
e.set_packets(w.x, w.y, w.k, w.l)
for _ in range(3):
    e.step(S.SCHEME_LEAPFROG, w.dt, sub)
    print("kernel ms", e.last_kernel_ms())
