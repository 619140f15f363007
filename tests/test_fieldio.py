"""write_field / read_field formats (qg_flow_ray_trace/write_field.m, read_field.m) -- CPU only."""
import numpy as np

from swraytracing_b200 import fieldio as F


def test_real_frames_append_and_random_access(tmp_path):
    rs = np.random.RandomState(0)
    frames = [rs.randn(5, 7) for _ in range(4)]
    for fr in frames:
        F.write_field(fr, tmp_path / "pv", 1)          # the reference's 'a' mode: frame index is ignored
    raw = np.fromfile(tmp_path / "pv.bin")
    assert raw.size == 4 * 35 and np.array_equal(raw[:35], frames[0].ravel(order="F"))   # column-major real*8
    got = F.read_field(tmp_path / "pv", 5, 7, 1, [3, 1])
    assert got.shape == (5, 7, 2) and np.array_equal(got[..., 0], frames[2]) and np.array_equal(got[..., 1], frames[0])
    assert F.read_field(tmp_path / "missing", 5, 7) == 0


def test_complex_spectral_frames(tmp_path):
    rs = np.random.RandomState(1)
    nky = 4; nkx = 2 * nky - 1                         # nx == 2*ny-1  => read as complex (read_field.m:37-41)
    fk = rs.randn(nkx, nky) + 1j * rs.randn(nkx, nky)
    F.write_field(fk, tmp_path / "qk")
    F.write_field(2 * fk, tmp_path / "qk")
    raw = np.fromfile(tmp_path / "qk.bin")
    n = nkx * nky
    assert np.array_equal(raw[:n], fk.real.ravel(order="F")) and np.array_equal(raw[n:2 * n], fk.imag.ravel(order="F"))
    assert np.array_equal(F.read_field(tmp_path / "qk", nkx, nky, 1, [2]), 2 * fk)
    assert np.array_equal(F.read_field(tmp_path / "qk", nkx, nky, 1, [1], is_real=True), fk.real)


def test_scalar_series_and_layers(tmp_path):
    for t in (0.5, 1.5, 2.5):
        F.write_field(t, tmp_path / "packet_time")
    assert np.array_equal(F.read_field(tmp_path / "packet_time"), np.array([[0.5, 1.5, 2.5]]))
    a = np.arange(2 * 3 * 2, dtype=float).reshape(2, 3, 2)
    F.write_field(a, tmp_path / "layers")
    assert np.array_equal(F.read_field(tmp_path / "layers", 2, 3, 2, [1]), a)


def test_packet_frame_writer_roundtrip(tmp_path):
    L = 2 * np.pi; n = 11
    rs = np.random.RandomState(2)
    w = F.PacketFrameWriter(tmp_path, L)
    xs = []
    for i in range(3):
        x = rs.uniform(-20, 20, n); y = rs.uniform(-20, 20, n); k = rs.randn(n); l = rs.randn(n)
        w.write(x, y, k, l, 0.1 * i); xs.append((x, y, k, l))
    t, x, k = F.load_packet_frames(tmp_path, n)
    assert np.allclose(t, [0.0, 0.1, 0.2]) and x.shape == (n, 2, 3)
    assert x.min() >= -L / 2 and x.max() < L / 2                      # wrapped only on save (qgsw_raytrace.m:160)
    d = x[:, 0, 1] - xs[1][0]
    assert np.allclose(d / L, np.round(d / L), atol=1e-12)            # equal modulo the period
    assert np.array_equal(k[:, 1, 2], xs[2][3])
