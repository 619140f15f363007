"""Short driver for ncu captures of the LAGRANGE6 leapfrog kernel (reference semantics)."""
import sys; sys.path.insert(0, '.')
import swraytracing_b200 as S
from swraytracing_b200 import workloads as W
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
w = W.make_workload(name, n_packets=int(sys.argv[2]) if len(sys.argv) > 2 else None)
e = S.Engine(w.nx, w.L, w.f, w.gH, S.MODE_LAGRANGE6)
e.set_flow_spectral(w.psik)
e.set_packets(w.x, w.y, w.k, w.l)
for _ in range(3):
    e.step(S.SCHEME_LEAPFROG, w.dt, 16)
    print("kernel ms", e.last_kernel_ms())
