"""Developer smoke: CUDA path vs oracle on small seeded inputs (run under gpurun)."""
import sys, time; sys.path.insert(0, '.')
import numpy as np
from oracle import swrt_oracle as O
import swraytracing_b200 as S

def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = 2 * np.pi; dx = L / nx; f = 3.0; gH = 1.0
rs = np.random.RandomState(7)
kx_, ky_ = O.wavenumbers(nx)
K2 = kx_**2 + ky_**2
psik = (rs.randn(nx - 1, nx // 2) + 1j * rs.randn(nx - 1, nx // 2)) / (1 + K2) ** 1.5 * 0.3
planes = O.velocity_planes_k(psik, kx_, ky_)
n = 1000
x = rs.uniform(-3 * L, 3 * L, n); y = rs.uniform(-3 * L, 3 * L, n)
k = 3 * np.cos(np.arange(n)); l = 3 * np.sin(np.arange(n))
ref = O.spectral_eval_planes(x, y, planes, dx, nx)

for mt in (1, 2):
    e = S.Engine(nx, L, f, gH, S.MODE_SPECTRAL)
    e.set_tuning(mt)
    e.set_flow_spectral(psik)
    e.set_packets(x, y, k, l)
    got = e.eval()
    print(f"spectral eval mt={mt}:", [f"{rel(got[c], ref[c]):.2e}" for c in range(6)], "scaled max", f"{max(np.abs(got[c]-ref[c]).max()/np.abs(ref[c]).max() for c in range(6)):.2e}")
    # leapfrog 5 steps vs oracle
    dt = 0.1 * dx
    xo, yo, ko, lo = x.copy(), y.copy(), k.copy(), l.copy()
    ev = lambda xx, yy: O.spectral_eval_planes(xx, yy, planes, dx, nx)
    for _ in range(5):
        xo, yo, ko, lo = O.leapfrog_step(xo, yo, ko, lo, dt, f, gH, ev)
    e.step(S.SCHEME_LEAPFROG, dt, 5)
    xg, yg, kg, lg = e.get_packets()
    print(f"  leapfrog 5 steps mt={mt}: x {np.abs(xg-xo).max():.2e} y {np.abs(yg-yo).max():.2e} k {np.abs(kg-ko).max():.2e} l {np.abs(lg-lo).max():.2e}", "ms", e.last_kernel_ms())
    e.close()

# lagrange
fields = [O.k2g(p) for p in planes]
e = S.Engine(nx, L, f, gH, S.MODE_LAGRANGE6)
e.set_flow_grid(*fields)
e.set_packets(x, y, k, l)
got = e.eval()
refl = np.stack([O.interpolate(x, y, F, dx, dx) for F in fields])
print("lagrange eval:", [f"{rel(got[c], refl[c]):.2e}" for c in range(6)])
e2 = S.Engine(nx, L, f, gH, S.MODE_LAGRANGE6)
e2.set_flow_spectral(psik)
e2.set_packets(x, y, k, l)
got2 = e2.eval()
print("lagrange eval via device k2g:", [f"{rel(got2[c], refl[c]):.2e}" for c in range(6)])
print("g2k/k2g dev:", rel(S.k2g_dev(planes[0]), fields[0]), np.abs(S.g2k_dev(fields[0]) - O.symmetrise_ky0(planes[0])).max())
print("interpolate_dev:", rel(S.interpolate_dev(x, y, fields[0], dx, dx), refl[0]))
print("OK")
