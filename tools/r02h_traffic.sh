# DRAM traffic of the dominant kernel at the bench's OWN launch shapes (bench.py -> roofline.traffic), round 2 end state.
# One --set full capture of the headline launch (C4: 16,777,216 packets, 2 fused sub-steps) and DRAM-only captures of the
# side-config launches.  Run under gpurun on one GPU; the .ncu-rep files come back in gpurun_out/.
set -x
python tools/prof_case.py C4 spectral --substeps 2 --reps 2 | tail -1
ncu --set full --clock-control none --import-source on -k regex:spectral_kernel -s 1 -c 1 -o gpurun_out/r02h_spec512_c4_full -f \
    python tools/prof_case.py C4 spectral --substeps 2 --reps 2 > gpurun_out/ncu_h1.log 2>&1
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"
ncu --metrics $M --clock-control none -k regex:spectral_kernel -s 1 -c 1 --csv --log-file gpurun_out/r02h_traffic_C2.csv \
    python tools/prof_case.py C2 spectral --substeps 16 --reps 2 > gpurun_out/ncu_h2.log 2>&1
ncu --metrics $M --clock-control none -k regex:spectral_kernel -s 1 -c 1 --csv --log-file gpurun_out/r02h_traffic_C3.csv \
    python tools/prof_case.py C3 spectral --substeps 16 --reps 2 > gpurun_out/ncu_h3.log 2>&1
ncu --metrics $M --clock-control none -k regex:spectral_rk4_kernel -s 1 -c 1 --csv --log-file gpurun_out/r02h_traffic_C5.csv \
    python tools/prof_case.py C5 spectral --substeps 2 --reps 2 > gpurun_out/ncu_h5.log 2>&1
ls -la gpurun_out/r02h*
