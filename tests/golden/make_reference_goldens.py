#!/usr/bin/env python
"""Extract the known answers the reference itself left behind into tests/golden/reference_runlogs.json.

The reference (MATLAB R2020b, run.log:3) ships no test vectors, but its own production runs printed
and stored numbers that depend on the arithmetic this repo restates:

  * ``run.log`` and ``analysis/job-*/run-*/run.log`` -- the 13 header lines of qgsw_raytrace.m:76-88.
    "Background velocity (parameter,computed): (U_g,U0)" is U0 = sqrt(max(u.^2+v.^2)) after
    rng(146) -> rand(17,17) -> initial_q (qgsw_raytrace.m:191-214, including the always-true chained
    comparison on :202) -> g2k -> grid_U -> 6x k2g (qgsw_raytrace.m:23,52-53,63-65), printed with %f.
    "Time step" is CFL*dx/U0 (qgsw_raytrace.m:70; CFL 0.1 in job-36976465, 0.05 in job-37011720 and
    in the tree as committed).
  * ``qg_flow_ray_trace/data/.nfs00000000032a756700000024`` -- a ``pv_time`` frame stream written by
    write_field (qgsw_raytrace.m:109,169-170): t after every 50 steps of ``t = t + dt`` (:134), raw
    fp64.  Its longest run (2,682 frames, U_g = 0.2) pins dt -- hence U0 -- to the last few ulps.

Run HERE (the container that has /root/reference); the JSON is what travels.
"""
import json
import re
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "reference_runlogs.json"

PAT = {
    "nx": r"Resolution: (\d+)x\d+",
    "npackets": r"Number of packets: (\d+)",
    "wavenumber_radius": r"Initial wavenumber radius: ([\d.]+)",
    "dt": r"Time step: ([\d.]+)",
    "T": r"Simulation time: ([\d.]+)",
    "f": r"Coriolis parameter: ([\d.]+)",
    "Cg": r"Group velocity: ([\d.]+)",
    "U_g": r"Background velocity \(parameter,computed\): \(([\d.]+),",
    "U0": r"Background velocity \(parameter,computed\): \([\d.]+,([\d.]+)\)",
    "Fr": r"Froude Number: ([\d.]+)",
    "K_d2": r"Deformation wavenumber: ([\d.]+)",
}


def parse_log(path):
    text = path.read_text(errors="replace")[:4000]
    row = {"source": str(path.relative_to(REF))}
    for key, pat in PAT.items():
        m = re.search(pat, text)
        if not m:
            return None
        row[key] = float(m.group(1)) if key not in ("nx", "npackets") else int(m.group(1))
    return row


def main():
    logs = [REF / "run.log"] + sorted(REF.glob("analysis/job-*/run-*/run.log"))
    rows = [r for r in (parse_log(p) for p in logs) if r]
    stream = np.fromfile(REF / "qg_flow_ray_trace/data/.nfs00000000032a756700000024", dtype="<f8")
    # the file was opened 'a' by several runs (write_field.m:31); runs start where t returns to 0
    zeros = np.flatnonzero(stream == 0.0)
    best = (0, 0)
    for i, z in enumerate(zeros):
        end = zeros[i + 1] if i + 1 < len(zeros) else len(stream)
        if end - z > best[1] - best[0]:
            best = (int(z), int(end))
    seg = stream[best[0]:best[1]]
    out = {
        "note": "numbers printed / stored by the reference's own MATLAB R2020b runs; see make_reference_goldens.py",
        "logs": rows,
        "pv_time": {"source": "qg_flow_ray_trace/data/.nfs00000000032a756700000024", "offset": best[0],
                    "steps_per_save": 50, "hex": [float(v).hex() for v in seg]},
    }
    OUT.write_text(json.dumps(out, indent=1))
    print(f"{len(rows)} log headers, pv_time run of {len(seg)} frames (t_end = {seg[-1]:.3f}) -> {OUT}")


if __name__ == "__main__":
    sys.exit(main())
