"""Frame-addressed binary field files of the reference: write_field.m / read_field.m.

Format (qg_flow_ray_trace/write_field.m:22-48): file ``<fname>.bin``, native-endian IEEE ``real*8``,
MATLAB column-major.  A real frame is nx*ny*nz doubles; a complex frame is the real block followed by the
imaginary block.  The reference opens the file with mode ``'a'`` (:31), so the ``fseek`` to ``frame`` is
ineffective and frames are strictly appended -- reproduced here.  ``read_field`` (read_field.m:58-99)
random-accesses frames of ``unit*nx*ny*nz`` (x2 if complex) bytes; a field is taken to be complex when
``nx == 2*ny-1`` unless ``is_real`` says otherwise (:37-41).

Host-side I/O only (the callers on either side of the hot path: packet_x / packet_k / packet_time / pv
streams of qgsw_raytrace.m:104-109,153-172 that analysis/load_data.m:29-32 reads back).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np


def write_field(field, fname, frame=1):
    """write_field(field, fname, frame): append one frame (the reference's 'a' mode ignores ``frame``)."""
    a = np.asarray(field)
    if a.ndim == 0:
        a = a.reshape(1, 1)
    with open(str(fname) + ".bin", "ab") as fh:
        if np.iscomplexobj(a):
            fh.write(np.asfortranarray(a.real, dtype=np.float64).tobytes(order="F"))
            fh.write(np.asfortranarray(a.imag, dtype=np.float64).tobytes(order="F"))
        else:
            fh.write(np.asfortranarray(a, dtype=np.float64).tobytes(order="F"))


def read_field(file, nx=1, ny=1, nz=1, frmvec=(1,), is_real=None):
    """field = read_field(file, nx, ny, nz, frmvec, is_real) -> array (nx, ny, nz, nframes) squeezed;
    with nx == 1 the whole file is returned as a (1, n) series (read_field.m:68-70)."""
    path = Path(str(file) + ".bin")
    if not path.exists():
        return 0                                  # read_field.m:61-66 displays the message and returns 0
    if is_real is None:
        is_real = not (nx == 2 * ny - 1)
    if nx == 1:
        return np.fromfile(path, dtype=np.float64).reshape(1, -1)
    frmvec = [int(f) for f in np.atleast_1d(frmvec)]
    per = nx * ny * nz
    out = np.zeros((nx, ny, nz, len(frmvec)), dtype=np.float64 if is_real else np.complex128)
    with open(path, "rb") as fh:
        for j, frm in enumerate(frmvec):
            fh.seek(8 * per * (1 if is_real else 2) * (frm - 1))
            re = np.fromfile(fh, dtype=np.float64, count=per)
            if re.size != per:
                raise EOFError(f"{path}: frame {frm} is beyond the end of the file")
            if is_real:
                out[..., j] = re.reshape((nx, ny, nz), order="F")
            else:
                im = np.fromfile(fh, dtype=np.float64, count=per)
                out[..., j] = (re + 1j * im).reshape((nx, ny, nz), order="F")
    return np.squeeze(out)


class PacketFrameWriter:
    """The packet output streams of qgsw_raytrace.m:35-39,104-106,153-163: ``packet_x`` holds
    [x(1..Np); y(1..Np)] wrapped to [-L/2, L/2) (:160), ``packet_k`` [k; l], ``packet_time`` one double."""

    def __init__(self, directory, L, prefix="packet"):
        self.dir = Path(directory); self.dir.mkdir(parents=True, exist_ok=True)
        self.L = float(L); self.prefix = prefix; self.frames = 0

    def write(self, x, y, k, l, t, wrap=True):
        """``wrap=False`` for the initial frame, which the drivers write as drawn (qgsw_raytrace.m:104)"""
        L = self.L
        if wrap:
            px = np.stack([np.mod(np.asarray(x) + L / 2, L) - L / 2, np.mod(np.asarray(y) + L / 2, L) - L / 2], axis=1)
        else:
            px = np.stack([np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)], axis=1)
        self.frames += 1
        write_field(px, self.dir / f"{self.prefix}_x", self.frames)
        write_field(np.stack([k, l], axis=1), self.dir / f"{self.prefix}_k", self.frames)
        write_field(float(t), self.dir / f"{self.prefix}_time", self.frames)


def load_packet_frames(directory, Npackets, prefix="packet"):
    """analysis/load_data.m:29-32: t, x(Np,2,frames), k(Np,2,frames)"""
    d = Path(directory)
    t = read_field(d / f"{prefix}_time")
    nfr = t.shape[1]
    x = read_field(d / f"{prefix}_x", Npackets, 2, 1, range(1, nfr + 1))
    k = read_field(d / f"{prefix}_k", Npackets, 2, 1, range(1, nfr + 1))
    return t.ravel(), x.reshape(Npackets, 2, nfr), k.reshape(Npackets, 2, nfr)
