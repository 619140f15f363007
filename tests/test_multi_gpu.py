"""Two real GPUs over NCCL: the sharded path against the single-GPU path (needs >= 2 CUDA devices; skipped on a
one-GPU box -- the same host logic runs on CPU with gloo in tests/test_host_logic.py).

Packets are independent and the flow is replicated, so a 2-rank run must give bit-identical per-packet states, identical
integer histograms (blocking and pipelined), identical ode23 accept/reject decisions (MAX all-reduce of the error
norm) and the same sharded theoretical omega pdf."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _ngpu():
    try:
        import swraytracing_b200 as S
        return int(S.load_library().swrt_device_count())
    except Exception:
        return 0


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _scenario(eng_factory, ens_factory):
    """the sequence both the 1-GPU and the 2-GPU run execute; returns everything that must agree"""
    import swraytracing_b200 as S
    from swraytracing_b200 import workloads as W
    w = W.make_workload("C3", n_packets=4099, nx=32)
    eng = eng_factory(w)
    eng.set_flow_spectral(w.psik, 0); eng.set_flow_spectral(w.psik2, 1)
    ens = ens_factory(eng, w.n_packets)
    ens.set_packets_global(w.x, w.y, w.k, w.l)
    ens.step(S.SCHEME_LEAPFROG, w.dt / 4, 4, 0.125, 0.25)
    edges = np.linspace(0.0, 8.0, 300)
    blocking = ens.hist_omega(edges)
    pipe = ens.hist_pipeline(edges)
    piped = []
    for _ in range(3):
        eng.step_async(S.SCHEME_LEAPFROG, w.dt / 4, 4, 0.125, 0.25)
        r = pipe.rotate()
        if r is not None:
            piped.append(r)
        pipe.launch()
    piped += pipe.drain()
    st = ens.ode23([0.0, w.dt], w.dt)
    diag = ens.diag(1.0)
    gx, gy = np.meshgrid(np.linspace(0, w.L, 24), np.linspace(0, w.L, 24))
    th = np.linspace(0, 2 * np.pi, 50)
    ideal = ens.ideal_omega_hist(gx.ravel(), gy.ravel(), 5 * np.cos(th), 5 * np.sin(th), np.sqrt(9 + 25.0), np.linspace(3.0, 9.0, 100))
    x, y, k, l = ens.gather_packets()
    return {"blocking": blocking, "piped": np.stack(piped), "nsteps": st["nsteps"], "nfailed": st["nfailed"], "diag": diag,
            "ideal": ideal, "x": x, "y": y, "k": k, "l": l}


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    import swraytracing_b200 as S
    from swraytracing_b200.distributed import ShardedEnsemble
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    res = _scenario(lambda w: S.Engine(w.nx, w.L, w.f, w.gH, S.MODE_SPECTRAL, device=rank),
                    lambda eng, n: ShardedEnsemble(eng, n, rank, world, dist, device=torch.device("cuda", rank)))
    if rank == 0:
        np.savez(out, **res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpus_nccl_match_one_gpu(tmp_path):
    if _ngpu() < 2:
        pytest.skip("needs two CUDA devices (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    import swraytracing_b200 as S
    from swraytracing_b200.distributed import ShardedEnsemble
    out = str(tmp_path / "r.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    ref = _scenario(lambda w: S.Engine(w.nx, w.L, w.f, w.gH, S.MODE_SPECTRAL, device=0), lambda eng, n: ShardedEnsemble(eng, n))
    for name in ("x", "y", "k", "l"):
        assert np.array_equal(got[name], ref[name]), name                  # per-packet states: bit-identical under sharding
    assert np.array_equal(got["blocking"], ref["blocking"]) and int(ref["blocking"].sum()) <= 4099
    assert got["piped"].shape == ref["piped"].shape == (3, 299) and np.array_equal(got["piped"], ref["piped"])
    assert (int(got["nsteps"]), int(got["nfailed"])) == (ref["nsteps"], ref["nfailed"])
    assert np.array_equal(got["ideal"], ref["ideal"]) and int(ref["ideal"].sum()) > 0
    assert np.allclose(got["diag"], ref["diag"], rtol=1e-12) and got["diag"][6] == 4099
