"""Writes tests/golden/octave_in/*.bin -- the seeded INPUTS of tests/golden/hotpath_nx32.npz as raw files in the
reference's own on-disk format (qg_flow_ray_trace/write_field.m: native-endian real*8, MATLAB column-major, one frame), so
that tests/golden/make_octave_goldens.m can feed them to the UNMODIFIED reference functions under MATLAB / GNU Octave.
Inputs only (fields, packets, scalars): nothing here is a computed result of the path under test.
Run from the repo root:  python tests/golden/export_hotpath_inputs.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import swrt_oracle as O  # noqa: E402

HERE = Path(__file__).resolve().parent
OUT = HERE / "octave_in"
NAMES = ("u", "v", "ux", "uy", "vx", "vy")


def put(name, arr):
    a = np.asarray(arr, dtype=np.float64)
    (OUT / f"{name}.bin").write_bytes(np.asfortranarray(a).ravel(order="F").tobytes())


def main():
    g = np.load(HERE / "hotpath_nx32.npz")
    nx = int(g["nx"]); L = float(g["L"])
    kx_, ky_ = O.wavenumbers(nx)
    OUT.mkdir(exist_ok=True)
    for f in OUT.glob("*.bin"):
        f.unlink()
    for c, name in enumerate(NAMES):
        put(f"bf1_{name}", g["grids"][c])                       # frame 1 = the gridded planes of the fixture
    grids2 = [O.k2g(p) for p in O.velocity_planes_k(g["psik2"], kx_, ky_)]
    for c, name in enumerate(NAMES):
        put(f"bf2_{name}", grids2[c])                           # frame 2 (interpolate_U.m)
    put("H", g["H"])                                            # H = 1 + eta_g/H0 (step_packet_xka.m, cg_sw.m)
    put("psi", O.k2g(g["psik"]))                                # streamfunction grid (SpectralScheme.m ctor argument)
    for name in ("x", "y", "k", "l"):
        put(name, g[name])
    # scalars: nx, L, f, gH, alpha, dt, C0, n
    put("params", np.array([nx, L, float(g["f"]), float(g["gH"]), float(g["alpha"]), float(g["dt"]), 1.0, g["x"].size]))
    print("wrote", sorted(p.name for p in OUT.glob("*.bin")))


if __name__ == "__main__":
    main()
