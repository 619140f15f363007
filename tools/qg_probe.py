import sys, time; sys.path.insert(0, '.')
import numpy as np
import swraytracing_b200 as S
from swraytracing_b200 import drivers
nx = 256; L = 2 * np.pi
xg = np.linspace(-L / 2, L / 2, nx); X, Y = np.meshgrid(xg, xg)
q = drivers.initial_q(X, Y, 0.5, 3.0, np.random.RandomState(146))
qk = S.g2k_dev(q)
qg = S.QGFlow(nx, L, qk, 3.0, 0.0012, 3.0, 1.0, r_drag=0.0, force_strength=0.0)
qg.step(10)
t0 = time.time(); qg.step(500); t1 = time.time() - t0
t0 = time.time()
for _ in range(500): qg.step(1)
t2 = time.time() - t0
print(f"one call of 500 steps: {2 * t1:.3f} ms/step; 500 calls of 1 step: {2 * t2:.3f} ms/step")
eng = S.Engine(nx, L, 3.0, 1.0, S.MODE_SPECTRAL)
t0 = time.time()
for _ in range(200): qg.to_flow(eng, 0)
print(f"to_flow (SPECTRAL): {5 * (time.time() - t0):.3f} ms")
for mode, name in ((S.MODE_NUFFT, "NUFFT"), (S.MODE_LAGRANGE6, "LAGRANGE6")):
    e2 = S.Engine(nx, L, 3.0, 1.0, mode); qg.to_flow(e2, 0)
    t0 = time.time()
    for _ in range(100): qg.to_flow(e2, 0)
    print(f"to_flow ({name}): {10 * (time.time() - t0):.3f} ms")
