"""Developer perf probe: fused leapfrog throughput for a config (run under gpurun)."""
import os, sys; sys.path.insert(0, '.')
import numpy as np
from oracle import swrt_oracle as O
import swraytracing_b200 as S
nx = int(sys.argv[1]); n = int(sys.argv[2]); steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
mode = S.MODE_LAGRANGE6 if (len(sys.argv) > 4 and sys.argv[4] == 'lag') else S.MODE_SPECTRAL
L = 2 * np.pi; dx = L / nx; f = 3.0; gH = 1.0
rs = np.random.RandomState(7)
kx_, ky_ = O.wavenumbers(nx); K2 = kx_**2 + ky_**2
psik = (rs.randn(nx - 1, nx // 2) + 1j * rs.randn(nx - 1, nx // 2)) / (1 + K2) ** 1.5 * 0.3
x = rs.uniform(-L / 2, L / 2, n); y = rs.uniform(-L / 2, L / 2, n)
k = 3 * np.cos(2 * np.pi * np.arange(n) / n); l = 3 * np.sin(2 * np.pi * np.arange(n) / n)
mts = [int(v) for v in sys.argv[5].split(",")] if len(sys.argv) > 5 else [1, 2]
for mt in (mts if mode == S.MODE_SPECTRAL else (1,)):
    e = S.Engine(nx, L, f, gH, mode)
    e.set_tuning(mt, use_psi_moments=os.environ.get('SWRT_NOPSI') is None)
    e.set_flow_spectral(psik)
    e.set_packets(x, y, k, l)
    dt = 0.1 * dx
    e.step(S.SCHEME_LEAPFROG, dt, 2)
    best = 1e9
    for r in range(3):
        e.step(S.SCHEME_LEAPFROG, dt, steps)
        ms, nl = e.last_kernel_ms(); best = min(best, ms)
    pps = n * steps / (best * 1e-3)
    w = e.work_per_eval(e.contracted_planes() or 6)
    print(f"nx={nx} n={n} steps={steps} mt={mt} mode={mode}: {best:.3f} ms  {pps:.3e} packet-steps/s  {pps*w*1e-12:.2f} T(work)/s")
    e.close()
