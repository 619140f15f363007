"""Host-side logic on CPU: synthetic workloads, sharding, the gloo world_size=2 reduction path."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import swrt_oracle as O
from oracle import c_oracle as CO
from swraytracing_b200 import workloads as W
from swraytracing_b200.distributed import ShardedEnsemble, shard_range

ROOT = Path(__file__).resolve().parent.parent


def test_workload_generator_matches_oracle_kit():
    w = W.make_workload("C2", n_packets=100, nx=32)
    planes = W.planes_from_psik(w.psik, w.L)
    kx_, ky_ = O.wavenumbers(32)
    ref = O.velocity_planes_k(w.psik, kx_, ky_)
    for a, b in zip(planes, ref):
        assert np.abs(a - b).max() < 1e-15
    u = O.k2g(ref[0]); v = O.k2g(ref[1])
    assert abs(np.sqrt((u * u + v * v).max()) - w.U0) < 1e-12       # normalised so max|U| = U_g
    assert abs(W._fulspec_ifft(ref[0]) - u).max() < 1e-13
    w2 = W.make_workload("C2", n_packets=100, nx=32)
    assert np.array_equal(w.psik, w2.psik) and np.array_equal(w.x, w2.x)   # seeded
    assert np.allclose(w.k ** 2 + w.l ** 2, 27.0)                  # k0^2 = (nif^2-1) f^2/Cg^2


@pytest.mark.parametrize("name", ["C1", "C3", "C4", "C5"])
def test_other_workloads_build_small(name):
    w = W.make_workload(name, n_packets=64, nx=32)
    assert w.x.shape == (64,) and np.isfinite(w.dt) and w.dt > 0
    if name in ("C3", "C4"):
        assert w.psik2 is not None and np.abs(w.psik2 - w.psik).max() > 0
    if name == "C4":
        assert w.L == 20.0 and w.u_mean == 0.5
    if name == "C5":
        assert len(W.planes_from_psik(w.psik, w.L, etak=w.extra["etak"])) == 7


def test_shard_range_partitions():
    for n in (0, 1, 7, 65536, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


class OracleEngine:
    """oracle-backed stand-in with the Engine interface (CPU tests of the multi-rank host logic)"""

    def __init__(self, grids, dx, f, gH):
        self.grids, self.dx, self.f, self.gH = grids, dx, f, gH

    def set_packets(self, x, y, k, l, a=None):
        self.s = [np.array(v, dtype=np.float64) for v in (x, y, k, l)]

    def get_packets(self, with_a=False):
        return tuple(self.s)

    def step(self, scheme, dt, nsteps, alpha0=0.0, dalpha=0.0):
        self.s = list(CO.leapfrog_lagrange(*self.s, self.grids, self.dx, self.f, self.gH, dt, nsteps))

    def hist_omega(self, edges, kind=0, alpha=0.0):
        return CO.histcounts(O.omega_of_k(self.s[2], self.s[3], self.f, self.gH), edges)

    # ode23 building blocks (same contract as swrt_bs23_*), numpy arithmetic
    def _F(self, alpha, st):
        e6 = CO.interpolate6(st[0], st[1], self.grids, self.dx)
        return list(O.rhs_from_eval(e6, st[2], st[3], self.f, np.sqrt(self.gH)))

    def bs23_begin(self, alpha, thr):
        self.f1 = self._F(alpha, self.s)
        return max(np.max(np.abs(f) / np.maximum(np.abs(y), thr)) for f, y in zip(self.f1, self.s))

    def bs23_attempt(self, h, alphas, thr):
        y = self.s
        f2 = self._F(alphas[0], [a + b * (h * 0.5) for a, b in zip(y, self.f1)])
        f3 = self._F(alphas[1], [a + b * (h * 0.75) for a, b in zip(y, f2)])
        self.yn = [a + (b * (h * 2 / 9) + c * (h / 3) + d * (h * 4 / 9)) for a, b, c, d in zip(y, self.f1, f2, f3)]
        self.f4 = self._F(alphas[2], self.yn)
        fE = [b * (-5 / 72) + c * (1 / 12) + d * (1 / 9) + e * (-1 / 8) for b, c, d, e in zip(self.f1, f2, f3, self.f4)]
        return max(np.max(np.abs(g) / np.maximum(np.maximum(np.abs(a), np.abs(b)), thr)) for g, a, b in zip(fE, y, self.yn))

    def bs23_accept(self):
        self.s, self.f1 = self.yn, self.f4

    def ideal_omega_hist(self, x, y, kvx, kvy, omega0, edges, alpha=0.0):
        e6 = CO.interpolate6(np.asarray(x), np.asarray(y), self.grids, self.dx)
        Uk = np.outer(e6[0], kvx) + np.outer(e6[1], kvy)           # U*kv, ideal_omega_distribution.m:9-10
        return CO.histcounts((omega0 + Uk).ravel(), edges)

    def diag(self, alpha=0.0):
        w = O.omega_of_k(self.s[2], self.s[3], self.f, self.gH)
        return np.array([w.sum(), w.sum(), w.max(), w.min(), 0.0, float(w.size), float(w.size), w.sum()])


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, tmp):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(ROOT))
    w = W.make_workload("C2", n_packets=1001, nx=32)
    kx_, ky_ = O.wavenumbers(32)
    grids = [O.k2g(p) for p in O.velocity_planes_k(w.psik, kx_, ky_)]
    ens = ShardedEnsemble(OracleEngine(grids, w.dx, w.f, w.gH), w.n_packets, rank, world, dist)
    ens.set_packets_global(w.x, w.y, w.k, w.l)
    ens.step(0, w.dt, 7)
    edges = O.matlab_linspace(0, 6.5, 300)
    counts = ens.hist_omega(edges)
    pipe = ens.hist_pipeline(edges)                 # pipelined form: same counts, handed out by rotate()/drain()
    pipe.launch(); pipe.launch()
    d = ens.diag()
    outs = pipe.drain()
    assert len(outs) == 2 and all(np.array_equal(o, counts) for o in outs)
    allp = ens.gather_packets()
    gx, gy = np.meshgrid(np.linspace(0, w.L, 32), np.linspace(0, w.L, 32))
    th = np.linspace(0, 2 * np.pi, 100)
    ideal = ens.ideal_omega_hist(gx.ravel(order="F"), gy.ravel(order="F"), 3 * np.cos(th), 3 * np.sin(th), np.sqrt(18.0),
                                 O.matlab_linspace(2.0, 6.5, 120))
    st = ens.ode23([0, 20 * w.dt], np.inf)            # global error norm: MAX all-reduce per attempted step
    allq = ens.gather_packets()
    if rank == 0:
        np.savez(tmp, counts=counts, ideal=ideal, diag=d, x=allp[0], k=allp[2], nsteps=st["nsteps"], nfailed=st["nfailed"], x23=allq[0], k23=allq[2])
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_two_ranks_match_single_rank(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "r.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    w = W.make_workload("C2", n_packets=1001, nx=32)
    kx_, ky_ = O.wavenumbers(32)
    grids = [O.k2g(p) for p in O.velocity_planes_k(w.psik, kx_, ky_)]
    ens = ShardedEnsemble(OracleEngine(grids, w.dx, w.f, w.gH), w.n_packets)
    ens.set_packets_global(w.x, w.y, w.k, w.l)
    ens.step(0, w.dt, 7)
    edges = O.matlab_linspace(0, 6.5, 300)
    assert np.array_equal(got["counts"], ens.hist_omega(edges))        # integer histogram: bit-exact
    assert int(got["counts"].sum()) == 1001
    x1, _, k1, _ = ens.gather_packets()
    assert np.array_equal(got["x"], x1) and np.array_equal(got["k"], k1)   # per-packet states identical
    d1 = ens.diag()
    assert np.allclose(got["diag"], d1, rtol=1e-13) and got["diag"][6] == 1001
    gx, gy = np.meshgrid(np.linspace(0, w.L, 32), np.linspace(0, w.L, 32))
    th = np.linspace(0, 2 * np.pi, 100)
    ideal1 = ens.ideal_omega_hist(gx.ravel(order="F"), gy.ravel(order="F"), 3 * np.cos(th), 3 * np.sin(th), np.sqrt(18.0),
                                  O.matlab_linspace(2.0, 6.5, 120))
    assert np.array_equal(got["ideal"], ideal1) and int(ideal1.sum()) == 32 * 32 * 100     # sharded grid points, same counts
    # ode23 over two ranks == ode23 over one: same accepted/rejected steps, bit-identical packets
    st = ens.ode23([0, 20 * w.dt], np.inf)
    assert (int(got["nsteps"]), int(got["nfailed"])) == (st["nsteps"], st["nfailed"]) and st["nsteps"] >= 10
    x2, _, k2, _ = ens.gather_packets()
    assert np.array_equal(got["x23"], x2) and np.array_equal(got["k23"], k2)


def test_cg_sw_mirror_matches_restatement():
    from swraytracing_b200 import reference_api as R
    rs = np.random.RandomState(1)
    U = {"u": rs.randn(8, 8) * 0.1, "v": rs.randn(8, 8) * 0.1}
    H = 1 + 0.1 * rs.randn(8, 8)
    C, om, oma, divC, grad = R.cg_sw(1.5, -2.0, 1.2, 3.0, U, H)
    Cx, Cy, om0, d0, gx, gy = O.cg_sw(1.5, -2.0, 1.2, 3.0, U, H)
    assert np.array_equal(C["x"], Cx) and np.array_equal(C["y"], Cy) and np.array_equal(om, om0) and np.array_equal(oma, np.abs(om0))
    assert np.array_equal(divC, d0) and np.array_equal(grad["x"], gx) and np.array_equal(grad["y"], gy)
    C, om, _, divC, grad = R.cg_sw(1.5, -2.0, 1.2, 3.0)
    assert divC is None and grad is None and om == np.sqrt(9 + 1.44 * 6.25)


def test_bench_reference_arm_prints_one_contract_line():
    """bench.py --impl reference (the CPU arm the driver times beside the GPU arm): exactly one JSON line on stdout with
    the contract's keys, zero GPU launches, the oracle port as `kind`, every host core as `cores` -- also when
    OMP_NUM_THREADS=1 is exported the way torchrun does for its ranks (TORCHELASTIC_RUN_ID present)"""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OMP_NUM_THREADS="1", TORCHELASTIC_RUN_ID="test", SWRT_BENCH_TARGET_S="0.3")
    env.pop("RANK", None)
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--packets", "2048"], capture_output=True, text=True, env=env, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "higher_is_better", "scaling", "dtype", "data",
                "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["gpu_launches"] == 0 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    ncores = len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["cores"] == ncores
    # other ranks of a torchrun launch print nothing and exit 0
    out2 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, env=dict(env, RANK="1"), timeout=120, cwd=root)
    assert out2.returncode == 0 and out2.stdout.strip() == ""
