python -m pytest tests -x -q -m gpu 2>&1 | tail -3
# needs a library built with the developer overrides: make clean && make EXTRA=-DSWRT_DEV_TUNING
for cfg in "-1 -1 -1 -1" "0 3 4 8" "0 2 4 8" "0 4 8 4" "4000 4 8 4"; do
set -- $cfg
echo "desync=$1 lag=$2 nstages=$3 kc=$4"
if [ "$1" = "-1" ]; then python tools/dev_perf.py 128 65536 16 spec 1,2 2>&1 | grep "mt="; python tools/dev_perf.py 256 1048576 4 spec 1 2>&1 | grep "mt=";
else SWRT_DESYNC_NS=$1 SWRT_LAG=$2 SWRT_NSTAGES=$3 SWRT_KC=$4 python tools/dev_perf.py 128 65536 16 spec 1 2>&1 | grep "mt=1"; fi
done
