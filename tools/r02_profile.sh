set -x
NCU="ncu --set full --clock-control none --import-source on"
python tools/prof_case.py C4 spectral --packets 262144 --substeps 2 --reps 2 | tail -1
$NCU -k regex:spectral_kernel -s 1 -c 1 -o gpurun_out/r02_spec512 -f python tools/prof_case.py C4 spectral --packets 262144 --substeps 2 --reps 2 > gpurun_out/ncu1.log 2>&1
python tools/prof_case.py C5 spectral --packets 262144 --substeps 1 --reps 2 | tail -1
$NCU -k regex:spectral_rk4_kernel -s 1 -c 1 -o gpurun_out/r02_spec_rk4_xka -f python tools/prof_case.py C5 spectral --packets 262144 --substeps 1 --reps 2 > gpurun_out/ncu2.log 2>&1
python tools/prof_case.py C5 lagrange6 --packets 1048576 --substeps 2 --reps 2 | tail -1
$NCU -k regex:lagrange_rk4_kernel -s 1 -c 1 -o gpurun_out/r02_lag_rk4_xka -f python tools/prof_case.py C5 lagrange6 --packets 1048576 --substeps 2 --reps 2 > gpurun_out/ncu3.log 2>&1
python tools/prof_case.py C3 lagrange6 --substeps 16 --reps 2 | tail -1
$NCU -k regex:lagrange_leapfrog_kernel -s 1 -c 1 -o gpurun_out/r02_lag_leapfrog_two -f python tools/prof_case.py C3 lagrange6 --substeps 16 --reps 2 > gpurun_out/ncu4.log 2>&1
python tools/prof_case.py C2 spectral --substeps 16 --reps 2 | tail -1
$NCU -k regex:spectral_kernel -s 1 -c 1 -o gpurun_out/r02_spec128 -f python tools/prof_case.py C2 spectral --substeps 16 --reps 2 > gpurun_out/ncu5.log 2>&1
python bench.py --steps 2 --warmup 3 --side-steps 2 --no-cpu-baseline > gpurun_out/b_plain.json 2> gpurun_out/b_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --side-steps 2 --no-cpu-baseline > gpurun_out/ncu6.log 2>&1
ls -la gpurun_out/*.ncu-rep
