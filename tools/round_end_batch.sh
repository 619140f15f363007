# End-of-round measurement batch (run under gpurun from the repo root); outputs land in gpurun_out/r01c_*
set -x
(time python -m pytest tests/ -x -q -m gpu) > gpurun_out/r01c_pytest_gpu.log 2>&1; tail -3 gpurun_out/r01c_pytest_gpu.log
python bench.py > gpurun_out/r01c_bench_n1.log 2> gpurun_out/r01c_bench_n1.err
python bench.py --workload C3 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r01c_bench_C3.log 2> gpurun_out/r01c_bench_C3.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01c_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01c_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01c_ncu_bench.log 2>&1
python tools/prof_leapfrog.py C2 16 > gpurun_out/r01c_plain_prof.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:spectral_kernel -s 2 -c 1 -f -o gpurun_out/r01c_spectral_leapfrog python tools/prof_leapfrog.py C2 16 > gpurun_out/r01c_ncu_prof.log 2>&1
