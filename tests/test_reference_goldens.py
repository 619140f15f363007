"""Known answers the reference's own MATLAB runs left behind (tests/golden/reference_runlogs.json,
extracted by tests/golden/make_reference_goldens.py from /root/reference/run.log,
analysis/job-*/run-*/run.log and the stored pv_time stream).

They pin, end to end, rng(146)/rand (mt19937ar, column-major fill) -> initial_q with the always-true
chained comparison (qgsw_raytrace.m:202) -> g2k -> grid_U -> k2g/fulspec -> max speed -> dt:
  * 19 log headers: "Background velocity (parameter,computed)", "Froude Number", "Time step" to the 6
    decimals MATLAB printed, for ten U_g and both CFL fractions the author used;
  * the 2,682-frame pv_time stream of a U_g = 0.2 run: `t = t + dt` accumulated over 134,050 steps,
    which fixes dt (hence U0) to a few ulps.
CPU tests check the oracle; the GPU tests check the product (cuFFT g2k/k2g + the device driver).
"""
import json
from pathlib import Path

import numpy as np
import pytest

GOLD = json.loads((Path(__file__).parent / "golden" / "reference_runlogs.json").read_text())
LOGS = GOLD["logs"]
PV_TIME = np.array([float.fromhex(h) for h in GOLD["pv_time"]["hex"]])
L = 2 * np.pi


def _fmt(v):
    return float("%f" % v)          # MATLAB fprintf('%f') == C printf("%f"): 6 decimals, round-half-even on the binary value


def _oracle_U0(nx, U_g, f, Cg):
    from oracle import swrt_oracle as O
    x = O.matlab_linspace(-L / 2, L / 2, nx)
    X, Y = np.meshgrid(x, x)                                   # qgsw_raytrace.m:15-16
    kx_, ky_ = O.wavenumbers(nx)
    K2 = kx_ ** 2 + ky_ ** 2
    q = O.initial_q(X, Y, U_g, f / Cg, O.matlab_rand_stream(146))
    flow = O.grid_U(O.g2k(q), f / Cg, K2, kx_, ky_)
    return float(np.sqrt((flow["u"] ** 2 + flow["v"] ** 2).max()))


@pytest.fixture(scope="module")
def oracle_U0_unit():
    """U0 for U_g = 1; initial_q scales linearly with a_g (qgsw_raytrace.m:213)"""
    return _oracle_U0(256, 1.0, 3.0, 1.0)


def test_golden_file_has_the_reference_runs():
    assert len(LOGS) == 19 and len(PV_TIME) == 2682
    assert {r["U_g"] for r in LOGS} == {0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0}
    assert all(r["nx"] == 256 and r["npackets"] == 50 and r["f"] == 3.0 and r["Cg"] == 1.0 for r in LOGS)


@pytest.mark.parametrize("row", LOGS, ids=[r["source"].replace("/run.log", "") for r in LOGS])
def test_oracle_reproduces_log_header(row, oracle_U0_unit):
    U0 = _oracle_U0(256, row["U_g"], row["f"], row["Cg"]) if row["U_g"] in (0.2, 0.5) else oracle_U0_unit * row["U_g"]
    assert _fmt(U0) == row["U0"], (U0, row)
    assert _fmt(U0 / row["Cg"]) == row["Fr"]
    dx = L / row["nx"]
    assert row["dt"] in (_fmt(0.05 * dx / U0), _fmt(0.1 * dx / U0)), (row["dt"], 0.05 * dx / U0)
    assert _fmt(row["f"] / row["Cg"]) == row["K_d2"]
    # near_inertial_factor*f (qgsw_raytrace.m:78) for the sweep's omega0/f in {2,4,8,16} (parameters.txt)
    assert row["wavenumber_radius"] / row["f"] in (2.0, 4.0, 8.0, 16.0)


def test_ring_bug_is_what_matlab_ran():
    """The annulus the comment on qgsw_raytrace.m:193 intends gives a different U0 (0.504825 at
    U_g = 0.5); the logs show 0.506570, i.e. the always-true chained comparison is what ran."""
    from oracle import swrt_oracle as O
    x = O.matlab_linspace(-L / 2, L / 2, 256)
    X, Y = np.meshgrid(x, x)
    kx_, ky_ = O.wavenumbers(256)
    q = O.initial_q(X, Y, 0.5, 3.0, O.matlab_rand_stream(146), ring=True)
    flow = O.grid_U(O.g2k(q), 3.0, kx_ ** 2 + ky_ ** 2, kx_, ky_)
    assert _fmt(np.sqrt((flow["u"] ** 2 + flow["v"] ** 2).max())) == 0.504825 != 0.506570


def _accumulate(dt, nframes, every=50):
    t, out = 0.0, [0.0]
    for s in range(1, (nframes - 1) * every + 1):
        t = t + dt                                           # qgsw_raytrace.m:134
        if s % every == 0:
            out.append(t)
    return np.array(out)


def test_oracle_dt_reproduces_stored_pv_time_stream():
    U0 = _oracle_U0(256, 0.2, 3.0, 1.0)
    dt = 0.05 * (L / 256) / U0                               # qgsw_raytrace.m:29,70
    got = _accumulate(dt, len(PV_TIME))
    ulp = np.spacing(PV_TIME)
    assert np.all(np.abs(got - PV_TIME) <= 2 * ulp)           # three early frames differ by 1-2 ulps, the rest are bit-identical
    assert (got == PV_TIME).mean() > 0.998
    # the stored stream is bit-exact for a dt within 4 ulps of the oracle's (FFTW vs pocketfft rounding)
    cands = [dt]
    for _ in range(4):
        cands.append(np.nextafter(cands[-1], np.inf))
    assert any(np.array_equal(_accumulate(c, len(PV_TIME)), PV_TIME) for c in cands)


# ------------------------------------------------------------------------------------------------
# the product against the same numbers (device cuFFT kit + device driver)
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("U_g", [0.2, 0.5])
def test_device_driver_prints_the_reference_log_header(U_g, tmp_path):
    from swraytracing_b200 import drivers, fieldio
    lines = []
    out = drivers.qgsw_raytrace(256, 50, 2, 6000, 1000, U_g, 3.0, 1.0, outdir=str(tmp_path), max_steps=150, log=lines.append)
    want = "run.log" if U_g == 0.5 else "analysis/job-37011720/run-1/run.log"
    row = next(r for r in LOGS if r["source"] == want)
    assert row["U_g"] == U_g
    text = "\n".join(lines)
    assert "Resolution: 256x256" in text and "Number of packets: 50" in text
    assert "Initial wavenumber radius: %f" % row["wavenumber_radius"] in text
    assert "Background velocity (parameter,computed): (%f,%f)" % (row["U_g"], row["U0"]) in text
    assert "Froude Number: %f" % row["Fr"] in text
    assert "Deformation wavenumber: %f" % row["K_d2"] in text
    if U_g == 0.2:
        assert "Time step: %f" % row["dt"] in text               # CFL 0.05, the tree as committed
        # pv_time frames written by the device driver == the reference's stored stream (<= 1 ulp)
        t = fieldio.read_field(str(tmp_path / "pv_time"), 1, 1, 1, [1, 2, 3, 4]).ravel()
        assert np.all(np.abs(t - PV_TIME[:4]) <= 4 * np.spacing(PV_TIME[:4])), (t, PV_TIME[:4])
        assert abs(out["dt"] - PV_TIME[1] / 50) < 2e-17


@pytest.mark.gpu
def test_device_spectral_kit_reproduces_logged_U0():
    """g2k / grid_U / k2g through the C ABI (cuFFT) on the reference's own initial PV."""
    from swraytracing_b200 import drivers, reference_api as R
    xg = R.matlab_linspace(-L / 2, L / 2, 256)
    X, Y = np.meshgrid(xg, xg)
    kx_, ky_ = drivers.wavenumber_grids(256)
    q = drivers.initial_q(X, Y, 1.0, 3.0, np.random.RandomState(146))
    flow = R.grid_U(R.g2k(q), 3.0, kx_ ** 2 + ky_ ** 2, kx_, ky_)
    U0 = float(np.sqrt((flow["u"] ** 2 + flow["v"] ** 2).max()))
    assert _fmt(U0) == 1.013140
    for row in LOGS:
        assert _fmt(U0 * row["U_g"]) == row["U0"]


# ------------------------------------------------------------------------------------------------
# MATLAB's own k2g / fulspec / ik-multiplication outputs, from the workspace dump rsw/matlab.mat
# (tests/golden/rsw_workspace_frame.npz, extracted by tests/golden/make_rsw_goldens.py)
# ------------------------------------------------------------------------------------------------
RSW = np.load(Path(__file__).parent / "golden" / "rsw_workspace_frame.npz")


def _rsw_cases():
    """(name, half-plane spectrum, MATLAB's gridded rows) -- rsw/swk.m:205-209 (getrhs)"""
    Sk, dm = RSW["Sk"], RSW["damask"].astype(np.float64)
    nx = int(RSW["nx"])
    kmax = nx // 2 - 1
    kx = np.arange(-kmax, kmax + 1, dtype=np.float64)[:, None]
    ky = np.arange(0, kmax + 1, dtype=np.float64)[None, :]
    return [("u", dm * Sk[:, :, 0], RSW["u_rows"]), ("v", dm * Sk[:, :, 1], RSW["v_rows"]), ("h", dm * Sk[:, :, 2], RSW["h_rows"]),
            ("zeta", dm * (1j * kx * Sk[:, :, 1] - 1j * ky * Sk[:, :, 0]), RSW["zeta_rows"])]


@pytest.mark.parametrize("case", _rsw_cases(), ids=lambda c: c[0])
def test_oracle_k2g_matches_matlab_workspace(case):
    from oracle import swrt_oracle as O
    name, fk, rows = case
    got = O.k2g(fk)[::int(RSW["row_stride"])]
    assert np.abs(rows).max() > 1e-2
    assert np.abs(got - rows).max() <= 1e-15, name          # measured 6e-17 (FFTW vs pocketfft round-off)


def test_oracle_g2k_inverts_matlab_grid():
    """g2k of MATLAB's gridded field returns MATLAB's (dealiased) spectrum: pins g2k.m's normalisation,
    fftshift and the rows 2:end / columns kmax+2:end window against a MATLAB-produced pair."""
    from oracle import swrt_oracle as O
    fk = _rsw_cases()[2][1]
    full = O.k2g(fk)                                         # == MATLAB's real(h) on the stored rows (test above)
    back = O.g2k(full)
    sym = O.symmetrise_ky0(fk)
    assert np.abs(back - sym).max() <= 1e-16


@pytest.mark.gpu
@pytest.mark.parametrize("case", _rsw_cases(), ids=lambda c: c[0])
def test_device_k2g_matches_matlab_workspace(case):
    from swraytracing_b200 import reference_api as R
    name, fk, rows = case
    got = R.k2g(fk)[::int(RSW["row_stride"])]
    assert np.abs(got - rows).max() <= 1e-15, name


# ------------------------------------------------------------------------------------------------
# the PRODUCT's initial condition against the reference's logs, with no oracle in the loop
# ------------------------------------------------------------------------------------------------
def test_product_initial_q_reproduces_logged_U0_without_the_oracle():
    """swraytracing_b200.drivers.initial_q is checked against the reference's printed ``Background velocity`` directly:
    the spectral kit between q and U0 (g2k.m:5-9, grid_U.m:2-4, k2g.m:5-6 / fulspec.m:10-19) is spelled out here in plain
    numpy FFT calls, so neither the oracle (whose initial_q has the same text as the product's) nor the device is involved.
    A wrong phase fill order, seed, mode set (the always-true comparison of qgsw_raytrace.m:202) or normalisation in the
    product changes the sixth decimal of U0."""
    from swraytracing_b200 import drivers
    nx, K_d2 = 256, 3.0
    kmax = nx // 2 - 1
    xg = drivers.matlab_linspace(-L / 2, L / 2, nx)
    X, Y = np.meshgrid(xg, xg)
    q = drivers.initial_q(X, Y, 1.0, K_d2, np.random.RandomState(146))
    # g2k.m: fk = fftshift(fft2(fg))/nx^2, rows 2:end (kx = -kmax..kmax), columns kmax+2:end (ky = 0..kmax)
    qk = (np.fft.fftshift(np.fft.fft2(q)) / nx ** 2)[1:, kmax + 1:]
    kx = np.arange(-kmax, kmax + 1, dtype=np.float64)[:, None]
    ky = np.arange(0, kmax + 1, dtype=np.float64)[None, :]
    psik = -qk / (K_d2 + kx ** 2 + ky ** 2)                       # grid_U.m:2

    def k2g(fk):                                                   # fulspec.m:10-19 + k2g.m:5-6
        full = np.zeros((nx, nx), dtype=np.complex128)
        up = fk.copy()
        up[:kmax, 0] = np.conj(up[:kmax:-1, 0])                    # conjugate-symmetrise the ky = 0 column from its kx > 0 half
        up[kmax, 0] = up[kmax, 0].real
        full[1:, kmax + 1:] = up
        full[1:, 1:kmax + 1] = np.conj(up[::-1, :0:-1])            # the ky < 0 half-plane
        return (nx ** 2 * np.fft.ifft2(np.fft.ifftshift(full))).real

    u, v = k2g(-1j * ky * psik), k2g(1j * kx * psik)               # grid_U.m:3-4
    U0 = float(np.sqrt((u * u + v * v).max()))
    assert _fmt(U0) == 1.013140
    for row in LOGS:
        assert _fmt(U0 * row["U_g"]) == row["U0"], row["source"]
    # and the product's mode set really is the full square: the intended annulus prints a different number
    q_ring = drivers.initial_q(X, Y, 1.0, K_d2, np.random.RandomState(146), ring=True)
    assert np.abs(q_ring - q).max() > 1e-3
