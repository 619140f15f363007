# where the x twiddles of the dense kernel should live: rotated in registers (1), shared-memory table (auto at <= 256^2),
# global / L2 table (2), shared if it fits else global (3)
for tw in 0 1 2 3; do
  python tools/prof_case.py C2 spectral --substeps 16 --reps 3 --twiddles $tw | tail -1
  python tools/prof_case.py C3 spectral --substeps 4 --reps 3 --twiddles $tw | tail -1
  python tools/prof_case.py C4 spectral --packets 1048576 --substeps 2 --reps 3 --twiddles $tw | tail -1
  python tools/prof_case.py C5 spectral --packets 1048576 --substeps 1 --reps 3 --twiddles $tw | tail -1
done
