# A/B: the shipped 8-warp CTA against a 12-warp CTA (three MMA warps per SM sub-partition, 24 accumulator tiles each, 168 registers)
for lib in libswrt.so libswrt_w12.so; do
  echo "== $lib"
  SWRT_LIB=$PWD/swraytracing_b200/$lib python tools/prof_case.py C2 spectral --substeps 16 --reps 3 | tail -1
  SWRT_LIB=$PWD/swraytracing_b200/$lib python tools/prof_case.py C2 spectral --packets 681984 --substeps 16 --reps 3 | tail -1
  SWRT_LIB=$PWD/swraytracing_b200/$lib python tools/prof_case.py C3 spectral --substeps 4 --reps 3 | tail -1
  SWRT_LIB=$PWD/swraytracing_b200/$lib python tools/prof_case.py C4 spectral --packets 1048576 --substeps 2 --reps 3 | tail -1
  SWRT_LIB=$PWD/swraytracing_b200/$lib python tools/prof_case.py C5 spectral --packets 1048576 --substeps 1 --reps 3 | tail -1
done
