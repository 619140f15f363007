"""Does keeping the ensemble sorted by cell pay?  Same packets, random order vs sorted by (row, column) of the grid."""
import sys; sys.path.insert(0, '.')
import numpy as np
import swraytracing_b200 as S
from swraytracing_b200 import workloads as W
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
w = W.make_workload(name, n_packets=int(sys.argv[2]) if len(sys.argv) > 2 else None)
ix = np.floor(np.mod(w.x / w.dx, w.nx)).astype(np.int64); iy = np.floor(np.mod(w.y / w.dx, w.nx)).astype(np.int64)
order = np.lexsort((ix, iy))
for label, o in (("random", np.arange(w.n_packets)), ("sorted", order)):
    for mname, mode in (("nufft", S.MODE_NUFFT), ("lagrange6", S.MODE_LAGRANGE6)):
        e = S.Engine(w.nx, w.L, w.f, w.gH, mode)
        e.set_flow_spectral(w.psik, u_mean=w.u_mean)
        e.set_packets(w.x[o], w.y[o], w.k[o], w.l[o])
        best = 1e9
        for _ in range(3):
            e.step(S.SCHEME_LEAPFROG, w.dt, 16); best = min(best, e.last_kernel_ms()[0])
        print(f"{name} {label:6s} {mname:9s}: {best:.3f} ms  {w.n_packets * 16 / (best * 1e-3):.3e} packet-steps/s")
        e.close()
