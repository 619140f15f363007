"""Build tests/mex_harness/build/libmexharness.so (separate real/imag storage: GNU Octave, MATLAB -R2017b) and
libmexharness_ic.so (interleaved complex: MATLAB -R2018a): the UNMODIFIED matlab/swrt_mex.c + the harness' mx/mex
functions, linked against swraytracing_b200/libswrt.so.  Test infrastructure only."""
from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
OUT = HERE / "build"


def build(force=False):
    OUT.mkdir(exist_ok=True)
    srcs = [ROOT / "matlab" / "swrt_mex.c", HERE / "mxharness.c"]
    deps = srcs + [ROOT / "matlab" / "stub" / "mex.h", ROOT / "include" / "swrt.h"]
    libdir = ROOT / "swraytracing_b200"
    outs = []
    for name, defs in (("libmexharness.so", []), ("libmexharness_ic.so", ["-DMX_HAS_INTERLEAVED_COMPLEX=1"])):
        so = OUT / name
        if force or not so.exists() or so.stat().st_mtime < max(d.stat().st_mtime for d in deps):
            subprocess.check_call(["gcc", "-std=c99", "-O1", "-g", "-Wall", "-Wextra", "-Werror", "-fPIC", "-shared", *defs,
                                   f"-I{ROOT / 'matlab' / 'stub'}", f"-I{ROOT / 'include'}", *map(str, srcs), "-o", str(so),
                                   f"-L{libdir}", "-lswrt", f"-Wl,-rpath,{libdir}", "-Wl,-rpath,$ORIGIN/../../../swraytracing_b200"])
        outs.append(so)
    return outs


if __name__ == "__main__":
    print(build(force=True))
