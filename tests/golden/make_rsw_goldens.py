#!/usr/bin/env python
"""Extract tests/golden/rsw_workspace_frame.npz from the reference's bundled MATLAB workspace dump.

``/root/reference/rsw/matlab.mat`` (157 MB, too large to carry) is what the bare ``save`` at
rsw/swk.m:178 wrote at step n = 300 of a rotating-shallow-water run: the spectral state ``Sk``
(255 x 128 x 3, g2k layout, [u, v, eta]) TOGETHER WITH MATLAB's own outputs of the spectral kit on
that state, computed in swk.m:getrhs (:205-213) one line before the save:

    u    = k2gp(Sk(:,:,1))   real(u)    = nx^2*ifft2(ifftshift(fulspec(damask.*Sk(:,:,1))))  = k2g(damask.*Sk1)
    v,h  likewise for Sk2, Sk3           (k2gp = k2g with a half-cell-shifted copy packed in the imaginary part)
    zeta = k2gp(ikx_.*Sk2 - iky_.*Sk1)   real(zeta) = k2g of the spectral curl

so they are known answers for k2g/fulspec, the array layout and the ik-multiplication convention
(rsw/fulspec.m, rsw/g2k.m, rsw/k2g.m are the same files as qg_flow_ray_trace/'s).  The state itself is
the only bundled [u,v,eta] frame with the schema ray_trace_sw/raytrace_sw.m:11-28 loads
(``S(:,:,1:3)``, ``nx``, ``f``, ``Cg``) -- the stand-in for the missing
``wavevort_231058_restart_frame100.mat`` (.MISSING_LARGE_BLOBS:22) used by workload C5.

Kept: Sk (3 planes, exact), damask, f, Cg, t, dt, and every 8th row of real(u), real(v), real(h),
real(zeta) (MATLAB's outputs; 32 x 256 each).  Run HERE (needs /root/reference and scipy).
"""
from pathlib import Path

import numpy as np
import scipy.io as sio

SRC = Path("/root/reference/rsw/matlab.mat")
OUT = Path(__file__).resolve().parent / "rsw_workspace_frame.npz"
ROWS = slice(0, 256, 8)


def main():
    m = sio.loadmat(SRC, variable_names=["Sk", "damask", "u", "v", "h", "zeta", "f", "Cg", "t", "dt", "nx", "frame", "n", "Sout"])
    frame = int(m["frame"][0, 0])
    # Sout(:,:,:,frame) = real([u v h]) was stored in the same block (swk.m:163-165): cross-check the dump
    for j, name in enumerate(("u", "v", "h")):
        assert np.array_equal(m["Sout"][:, :, j, frame - 1], m[name].real), name
    np.savez_compressed(
        OUT,
        Sk=m["Sk"].astype(np.complex128), damask=m["damask"].astype(np.uint8),
        f=float(m["f"][0, 0]), Cg=float(m["Cg"][0, 0]), t=float(m["t"][0, 0]), dt=float(m["dt"][0, 0]),
        nx=int(m["nx"][0, 0]), n=int(m["n"][0, 0]), frame=frame, row_stride=8,
        u_rows=m["u"].real[ROWS].copy(), v_rows=m["v"].real[ROWS].copy(), h_rows=m["h"].real[ROWS].copy(),
        zeta_rows=m["zeta"].real[ROWS].copy())
    print(OUT, OUT.stat().st_size / 1e6, "MB")


if __name__ == "__main__":
    main()
