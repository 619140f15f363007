"""Multi-GPU layer: packets shard, the flow replicates, only histograms / diagnostics reduce.

One process per GPU (torchrun); ``torch.distributed`` is plumbing only.  The step has NO data-path
collective (packets are independent); the collectives are an integer SUM all-reduce
of the omega / energy histogram (analysis/load_data.m:39-49) and of a handful of diagnostic
scalars, once per diagnostic interval.  Integer histograms are bit-exact under any sharding.
"""
from __future__ import annotations

import numpy as np


def shard_range(n, rank, world):
    """contiguous packet slice [lo, hi) owned by ``rank``."""
    return (rank * n) // world, ((rank + 1) * n) // world


class _DeviceI64:
    """__cuda_array_interface__ view of n int64 at a raw device pointer (the engine's histogram counts;
    u64 counts reinterpreted as i64 -- identical bits for counts < 2^63)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i8", "data": (int(ptr), False), "version": 3}


class HistPipeline:
    """Histogram diagnostics pipelined against the packet kernel (the only collective of the path).

    Per diagnostic interval the caller does
        engine.step(...)          # asynchronous launch of the interval's packet kernel
        done = pipe.rotate()      # host work hidden under that kernel (see below)
        pipe.launch(alpha)        # local histogram kernel queued behind the packet kernel, no host wait
    ``rotate`` takes the histogram launched in the PREVIOUS interval (its kernel has long finished), snapshots the
    counts with a device-to-device copy, issues the SUM all-reduce asynchronously, and returns the globally reduced
    counts of the interval before that (or None while the pipeline fills).  No rank ever blocks on another inside an
    interval; ``drain()`` returns the results still in flight.  Without a process group (1 GPU, or the CPU stand-in
    engines of the gloo tests) the same calls degrade to local / synchronous reductions."""

    def __init__(self, ens, edges, kind=0):
        self.ens, self.edges, self.kind = ens, np.ascontiguousarray(edges, dtype=np.float64), kind
        self.nccl = (ens.dist is not None and ens.world > 1 and ens.device is not None
                     and getattr(ens.device, "type", "cpu") == "cuda")
        self.device_side = hasattr(ens.engine, "hist_omega_launch")
        self._launched = None      # (ptr, nbins) of a histogram kernel in flight
        self._reducing = None      # (work, tensor) of an all-reduce in flight
        self._ready = []           # finished results not yet handed out (non-NCCL paths)

    def launch(self, alpha=0.0):
        if self.device_side:
            self._launched = self.ens.engine.hist_omega_launch(self.edges, self.kind, alpha)
        else:
            self._ready.append(self.ens.hist_omega(self.edges, self.kind, alpha))

    def _collect(self):
        if self._reducing is None:
            return None
        work, t = self._reducing
        self._reducing = None
        if work is not None:
            work.wait()
        return t.cpu().numpy().astype(np.uint64)

    def rotate(self):
        out = self._collect()
        if self._launched is not None:
            import torch
            ptr, nb = self._launched
            self._launched = None
            self.ens.engine.hist_omega_wait()                         # that kernel only, not the work queued after it
            dev = self.ens.device if self.ens.device is not None else torch.device("cuda", self.ens.engine.device)
            t = torch.empty(nb, dtype=torch.int64, device=dev)
            t.copy_(torch.as_tensor(_DeviceI64(ptr, nb), device=dev))   # D2D snapshot: the engine reuses its buffer
            torch.cuda.current_stream(dev).synchronize()
            work = self.ens.dist.all_reduce(t, op=self.ens.dist.ReduceOp.SUM, async_op=True) if self.nccl else None
            self._reducing = (work, t)
        if out is None and self._ready:
            out = self._ready.pop(0)
        return out

    def drain(self):
        """results still in flight, oldest first"""
        outs = []
        for _ in range(3):
            r = self.rotate()
            if r is not None:
                outs.append(r)
        return outs


class ShardedEnsemble:
    """A rank-local engine holding this rank's packet shard.

    ``engine`` is any object with the Engine methods used here (set_packets/get_packets/step/
    hist_omega/diag); tests on CPU (gloo) pass an oracle-backed stand-in, production passes
    ``swraytracing_b200.Engine`` bound to the local CUDA device."""

    def __init__(self, engine, n_total, rank=0, world=1, dist=None, device=None):
        self.engine, self.n_total, self.rank, self.world = engine, int(n_total), int(rank), int(world)
        self.lo, self.hi = shard_range(self.n_total, self.rank, self.world)
        self.dist = dist
        self.device = device

    # -- packets --
    def set_packets_global(self, x, y, k, l, a=None):
        s = slice(self.lo, self.hi)
        self.engine.set_packets(x[s], y[s], k[s], l[s], None if a is None else a[s])

    def step(self, scheme, dt, nsteps, alpha0=0.0, dalpha=0.0):
        self.engine.step(scheme, dt, nsteps, alpha0, dalpha)

    def ode23(self, tspan, tmax, rtol=1e-3, atol=1e-6):
        """the reference's per-flow-step ode23 solve over ALL shards: the error norm is MAX-all-reduced so
        that every rank accepts/rejects the same steps (the reference's norm couples all packets)."""
        from .reference_api import ode23
        return ode23(self.engine, tspan, tmax, rtol, atol, reduce_max=self._allreduce_max)

    def _allreduce_max(self, v):
        if self.dist is None or self.world == 1:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64)
        if self.device is not None:
            t = t.to(self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.cpu()[0])

    # -- reductions --
    def _allreduce_sum(self, arr_i64):
        if self.dist is None or self.world == 1:
            return arr_i64
        import torch
        t = torch.from_numpy(arr_i64.copy())
        if self.device is not None:
            t = t.to(self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def hist_omega(self, edges, kind=0, alpha=0.0):
        """global histcounts over all shards: local u64 counts -> SUM all-reduce (bit-exact).
        With NCCL the counts are reduced in place in the engine's device buffer (no host round trip)."""
        on_gpu = (self.dist is not None and self.world > 1 and self.device is not None
                  and getattr(self.device, "type", "cpu") == "cuda" and hasattr(self.engine, "hist_omega_dev"))
        if on_gpu:
            import torch
            ptr, nb = self.engine.hist_omega_dev(edges, kind, alpha)
            t = torch.as_tensor(_DeviceI64(ptr, nb), device=self.device)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
            return t.cpu().numpy().astype(np.uint64)
        local = self.engine.hist_omega(edges, kind, alpha)
        return self._allreduce_sum(local.astype(np.int64)).astype(np.uint64)

    def hist_pipeline(self, edges, kind=0):
        return HistPipeline(self, edges, kind)

    def ideal_omega_hist(self, x, y, kvx, kvy, omega0, edges, alpha=0.0):
        """theoretical omega pdf (ideal_omega_distribution.m:3-11) over ALL grid points, sharded: each rank bins
        omega0 + U(x_i).k_j for its contiguous slice of the points, then the same integer SUM all-reduce as the packet
        histogram.  Every rank passes the full point list; counts are bit-exact under any sharding."""
        x = np.asarray(x, dtype=np.float64).ravel(); y = np.asarray(y, dtype=np.float64).ravel()
        lo, hi = shard_range(x.size, self.rank, self.world)
        local = self.engine.ideal_omega_hist(x[lo:hi], y[lo:hi], kvx, kvy, omega0, edges, alpha)
        return self._allreduce_sum(np.asarray(local).astype(np.int64)).astype(np.uint64)

    def energy_vs_omega(self, edges, kind=0, alpha=0.0):
        """energy = centre .* counts (analysis/load_data.m:40,49)"""
        counts = self.hist_omega(edges, kind, alpha)
        edges = np.asarray(edges, dtype=np.float64)
        centre = (edges[1:] + edges[:-1]) / 2
        return centre, centre * counts.astype(np.float64), counts

    def diag(self, alpha=0.0):
        """sum/max/min diagnostics all-reduced: sums add, extrema reduce with max/min."""
        d = np.asarray(self.engine.diag(alpha), dtype=np.float64)
        if self.dist is None or self.world == 1:
            return d
        import torch
        sums = torch.tensor([d[0], d[1], d[4], d[5], d[6], d[7]], dtype=torch.float64)
        mx = torch.tensor([d[2]], dtype=torch.float64)
        mn = torch.tensor([d[3]], dtype=torch.float64)
        if self.device is not None:
            sums, mx, mn = sums.to(self.device), mx.to(self.device), mn.to(self.device)
        self.dist.all_reduce(sums, op=self.dist.ReduceOp.SUM)
        self.dist.all_reduce(mx, op=self.dist.ReduceOp.MAX)
        self.dist.all_reduce(mn, op=self.dist.ReduceOp.MIN)
        s = sums.cpu().numpy()
        return np.array([s[0], s[1], float(mx.cpu()[0]), float(mn.cpu()[0]), s[2], s[3], s[4], s[5]])

    def gather_packets(self):
        """all shards' packets on every rank (for tests / small runs)."""
        loc = self.engine.get_packets()
        if self.dist is None or self.world == 1:
            return loc
        import torch
        out = []
        for arr in loc:
            pieces = [None] * self.world
            self.dist.all_gather_object(pieces, arr)
            out.append(np.concatenate(pieces))
        return tuple(out)
