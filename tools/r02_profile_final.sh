set -x
NCU="ncu --set full --clock-control none --import-source on"
python __graft_entry__.py smoke
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k bench_gpu_arm 2>&1 | tail -3
python tools/prof_case.py C4 spectral --packets 262144 --substeps 2 --reps 2 | tail -1
$NCU -k regex:spectral_kernel -s 1 -c 1 -o gpurun_out/r02f_spec512_l2table -f python tools/prof_case.py C4 spectral --packets 262144 --substeps 2 --reps 2 > gpurun_out/ncu1.log 2>&1
python tools/prof_case.py C3 spectral --packets 262144 --substeps 4 --reps 2 | tail -1
$NCU -k regex:spectral_kernel -s 1 -c 1 -o gpurun_out/r02f_spec256_l2table -f python tools/prof_case.py C3 spectral --packets 262144 --substeps 4 --reps 2 > gpurun_out/ncu2.log 2>&1
python tools/prof_case.py C5 spectral --packets 262144 --substeps 1 --reps 2 | tail -1
$NCU -k regex:spectral_rk4_kernel -s 1 -c 1 -o gpurun_out/r02f_spec_rk4_xka_l2table -f python tools/prof_case.py C5 spectral --packets 262144 --substeps 1 --reps 2 > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/r02f*.ncu-rep
