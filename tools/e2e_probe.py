"""Developer probe: end-to-end cost of host buffers through the C ABI.
set_packets + step + get_packets against swrt_step_host, pageable and pinned numpy buffers, per mode / workload."""
import sys, time, json
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import swraytracing_b200 as S
from swraytracing_b200 import workloads as W

def pinned(arrs):
    import torch
    ts = [torch.from_numpy(a.copy()).pin_memory() for a in arrs]
    return ts, [t.numpy() for t in ts]

def run(name, mode, n=None, sub=16, reps=6):
    w = W.make_workload(name, n_packets=n)
    eng = S.Engine(w.nx, w.L, w.f, w.gH, mode)
    eng.set_flow_spectral(w.psik, 0, w.u_mean)
    td = w.psik2 is not None
    if td: eng.set_flow_spectral(w.psik2, 1, w.u_mean)
    a0, da = (0.5 / sub, 1.0 / sub) if td else (0.0, 0.0)
    out = {}
    eng.set_packets(w.x, w.y, w.k, w.l)
    for _ in range(2): eng.step(S.SCHEME_LEAPFROG, w.dt, sub, a0, da)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); eng.step(S.SCHEME_LEAPFROG, w.dt, sub, a0, da); ts.append(time.perf_counter() - t0)
    out["resident_ms"] = 1e3 * min(ts)
    keep = None
    for kind in ("pageable", "pinned"):
        src = [a.copy() for a in (w.x, w.y, w.k, w.l)]
        dst = [np.empty_like(a) for a in src]
        if kind == "pinned":
            keep = (pinned(src), pinned(dst)); src, dst = keep[0][1], keep[1][1]
        for label in ("set_step_get", "step_host", "set_only", "get_only"):
            ts = []
            for _ in range(reps):
                t0 = time.perf_counter()
                if label == "set_step_get":
                    eng.set_packets(*src); eng.step(S.SCHEME_LEAPFROG, w.dt, sub, a0, da)
                    import ctypes as C
                    eng._check(eng.lib.swrt_get_packets(eng._h, *[o.ctypes.data_as(C.POINTER(C.c_double)) for o in dst], None))
                elif label == "step_host":
                    eng.step_host(S.SCHEME_LEAPFROG, w.dt, sub, *src, alpha0=a0, dalpha=da, out=dst)
                elif label == "set_only":
                    eng.set_packets(*src)
                else:
                    import ctypes as C
                    eng._check(eng.lib.swrt_get_packets(eng._h, *[o.ctypes.data_as(C.POINTER(C.c_double)) for o in dst], None))
                ts.append(time.perf_counter() - t0)
            out[f"{kind}_{label}_ms"] = 1e3 * min(ts[1:])
    mb = 32 * w.n_packets / 1e6
    out["MB_each_way"] = mb
    eng.close()
    return out

if __name__ == "__main__":
    res = {}
    for name, n in (("C2", None), ("C3", None), ("C4", 2097152)):
        for mname, mode in (("lagrange6", S.MODE_LAGRANGE6), ("nufft", S.MODE_NUFFT), ("spectral", S.MODE_SPECTRAL)):
            if mname == "spectral" and name != "C2": continue
            r = run(name, mode, n)
            res[f"{name}_{mname}"] = r
            print(name, mname, json.dumps({k: round(v, 3) for k, v in r.items()}), flush=True)
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/e2e_probe.json").write_text(json.dumps(res, indent=1))
