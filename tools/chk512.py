import sys; sys.path.insert(0,'.')
import numpy as np
import swraytracing_b200 as S
from swraytracing_b200 import workloads as W
from oracle import c_oracle as CO
w = W.make_workload("C4", n_packets=512)
planes1 = W.planes_from_psik(w.psik, w.L, w.u_mean); planes2 = W.planes_from_psik(w.psik2, w.L, w.u_mean)
for alpha in (0.0, 0.4):
    blend=[(1-alpha)*a+alpha*b for a,b in zip(planes1,planes2)]
    ref = CO.spectral_eval(w.x, w.y, blend, w.dx, w.nx, precise=True)
    for name,mode in (("dense",S.MODE_SPECTRAL),("nufft",S.MODE_NUFFT)):
        with S.Engine(w.nx, w.L, w.f, w.gH, mode) as e:
            e.set_flow_spectral(w.psik, 0, u_mean=w.u_mean); e.set_flow_spectral(w.psik2, 1, u_mean=w.u_mean)
            got = e.eval_at(w.x, w.y, alpha)
            print(alpha, name, " ".join("%.1e" % (np.abs(got[c]-ref[c]).max()/np.abs(ref[c]).max()) for c in range(6)))
