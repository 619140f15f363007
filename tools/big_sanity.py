"""16,777,216 packets on ONE GPU (config 4's full ensemble) in every mode: indexing / memory sanity, NUFFT vs LAGRANGE6 vs
dense spot check on a subsample, histogram totals."""
import sys, time; sys.path.insert(0, '.')
import numpy as np
import swraytracing_b200 as S
from swraytracing_b200 import workloads as W
w = W.make_workload("C4", n_packets=16777216)
edges = np.linspace(0.0, 20.0, 300)
ref = None
for name, mode in (("nufft", S.MODE_NUFFT), ("lagrange6", S.MODE_LAGRANGE6)):
    e = S.Engine(w.nx, w.L, w.f, w.gH, mode)
    e.set_flow_spectral(w.psik, 0, u_mean=w.u_mean); e.set_flow_spectral(w.psik2, 1, u_mean=w.u_mean)
    e.set_packets(w.x, w.y, w.k, w.l)
    t0 = time.time(); e.step(S.SCHEME_LEAPFROG, w.dt / 8, 8, 1 / 16, 1 / 8); dt = time.time() - t0
    ms, nl = e.last_kernel_ms()
    c = e.hist_omega(edges)
    st = np.stack(e.get_packets())
    d = e.diag(1.0)
    print(f"{name}: 8 steps of 16,777,216 packets in {dt:.3f} s (kernels {ms:.1f} ms / {nl} launches) -> {16777216 * 8 / dt:.3e} packet-steps/s;"
          f" hist total {int(c.sum())}, nonfinite {int(d[4])}, n {int(d[6])}")
    assert np.isfinite(st).all() and int(d[6]) == 16777216 and int(c.sum()) <= 16777216
    if ref is None:
        ref = st
    else:
        print("   max |nufft - lagrange6| over all packets:", np.abs(st - ref).max(), "(interpolation-error level, not parity)")
    e.close()
# dense spot check on the last 8192 packets
sl = slice(16777216 - 8192, 16777216)
e = S.Engine(w.nx, w.L, w.f, w.gH, S.MODE_SPECTRAL)
e.set_flow_spectral(w.psik, 0, u_mean=w.u_mean); e.set_flow_spectral(w.psik2, 1, u_mean=w.u_mean)
e.set_packets(w.x[sl], w.y[sl], w.k[sl], w.l[sl]); e.step(S.SCHEME_LEAPFROG, w.dt / 8, 8, 1 / 16, 1 / 8)
print("dense vs nufft on the last 8192 packets:", np.abs(np.stack(e.get_packets()) - ref[:, sl]).max())
assert np.abs(np.stack(e.get_packets()) - ref[:, sl]).max() < 1e-9
print("ok")
