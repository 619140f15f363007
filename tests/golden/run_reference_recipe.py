#!/usr/bin/env python
"""Run the UNMODIFIED reference on the seeded hot-path inputs and write tests/golden/octave_out/*.bin.

    python tests/golden/run_reference_recipe.py [--ref /root/reference] [--out tests/golden/octave_out] [--packets N]

The build image has neither MATLAB nor GNU Octave, so the recipe ``make_octave_goldens.m`` (unchanged -- the same file runs under
real MATLAB / Octave) and every reference function it calls (``interpolate``, ``interpolate_U``, ``SpectralScheme`` /
``RaytracingScheme``, ``ode_symplectic``, ``cg_sw``, ``step_packet``, ``step_packet_xka``, ``g2k`` / ``k2g`` / ``fulspec``,
``read_field`` / ``write_field``) are executed from where they lie under ``--ref`` by ``oracle/minimat``, the MATLAB-subset
interpreter of this repository (itself pinned to numbers real MATLAB produced: ``tests/test_minimat.py``).  Nothing of the
reference is copied or edited; the only things written are the raw fp64 result files, by the reference's own ``write_field.m``.

``--packets N`` truncates the packet list (a copy of the inputs with the first N packets goes to a temporary folder) -- used by
``tests/test_minimat.py`` to re-run a slice of the recipe in seconds and compare it with the committed files.

Run HERE (the container that has /root/reference); the .bin files are what travels.  ``PROVENANCE.json`` next to them records
how they were made and the sha256 of every reference file that was executed.
"""
import argparse
import hashlib
import json
import shutil
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))

from oracle.minimat import Interp          # noqa: E402


def truncated_inputs(src, n, dst):
    """copy of tests/golden/octave_in with only the first n packets (params(8) = n)"""
    dst.mkdir(parents=True, exist_ok=True)
    for f in src.glob("*.bin"):
        a = np.fromfile(f, dtype=np.float64)
        if f.stem in ("x", "y", "k", "l"):
            a = a[:n]
        elif f.stem == "params":
            a = a.copy()
            a[7] = float(n)
        a.tofile(dst / f.name)
    return dst


def run(ref, indir, outdir, quiet=False):
    outdir.mkdir(parents=True, exist_ok=True)
    out = open("/dev/null", "w") if quiet else sys.stdout
    I = Interp(cwd=str(ROOT), out=out)
    I.path.insert(0, str(HERE))                      # make_octave_goldens.m itself
    t0 = time.time()
    I.call("make_octave_goldens", str(ref), str(indir), str(outdir), nargout=0)
    I.close_all()
    executed = sorted(p for p in I.units if str(p).startswith(str(ref)))
    return time.time() - t0, executed


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--in", dest="indir", default=str(HERE / "octave_in"))
    ap.add_argument("--out", default=str(HERE / "octave_out"))
    ap.add_argument("--packets", type=int, default=0)
    a = ap.parse_args()
    ref, indir, outdir = Path(a.ref), Path(a.indir), Path(a.out)
    tmp = None
    if a.packets:
        tmp = Path(tempfile.mkdtemp(prefix="swrt_recipe_in_"))
        indir = truncated_inputs(indir, a.packets, tmp)
    secs, executed = run(ref, indir, outdir)
    if tmp:
        shutil.rmtree(tmp)
    prov = {
        "made_by": "tests/golden/run_reference_recipe.py",
        "executor": "oracle/minimat (MATLAB-subset interpreter of this repository; no MATLAB / Octave in the build image)",
        "recipe": "tests/golden/make_octave_goldens.m (unmodified; also runs under MATLAB / GNU Octave)",
        "packets": a.packets or "all",
        "seconds": round(secs, 1),
        "reference_files_executed": {str(Path(p).relative_to(ref)): hashlib.sha256(Path(p).read_bytes()).hexdigest() for p in executed},
        "outputs": {f.name: hashlib.sha256(f.read_bytes()).hexdigest() for f in sorted(outdir.glob("*.bin"))},
    }
    (outdir / "PROVENANCE.json").write_text(json.dumps(prov, indent=1) + "\n")
    print(f"run_reference_recipe: {len(prov['outputs'])} files in {outdir} ({secs:.1f} s); executed {len(executed)} reference files")


if __name__ == "__main__":
    main()
