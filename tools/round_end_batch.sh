# End-of-round measurement batch (run under gpurun from the repo root); outputs land in gpurun_out/r01c_*
set -x
(time python -m pytest tests/ -x -q -m gpu) > gpurun_out/r01c_pytest_gpu.log 2>&1; tail -3 gpurun_out/r01c_pytest_gpu.log
python bench.py > gpurun_out/r01c_bench_n1.log 2> gpurun_out/r01c_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01c_bench_ref.log 2> gpurun_out/r01c_bench_ref.err
python bench.py --workload C3 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r01c_bench_C3.log 2> gpurun_out/r01c_bench_C3.err
python bench.py --workload C5 --mode nufft --no-lagrange --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r01c_C5_nufft.log 2> gpurun_out/r01c_C5_nufft.err
python bench.py --workload C5 --mode lagrange6 --no-lagrange --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r01c_C5_lagrange6.log 2> gpurun_out/r01c_C5_lagrange6.err
python bench.py --workload C4 --mode lagrange6 --no-lagrange --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r01c_C4_lagrange6.log 2> gpurun_out/r01c_C4_lagrange6.err
python tools/ode23_probe.py > gpurun_out/r01c_ode23_probe.log 2>&1
python tools/lag_probe.py > gpurun_out/r01c_lag_probe.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01c_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01c_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01c_ncu_bench.log 2>&1
