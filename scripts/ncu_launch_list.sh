# usage: bash scripts/ncu_launch_list.sh <out.csv> [bench args...]   (under gpurun, one GPU)
set -e
OUT=$1; shift
python bench.py --steps 10 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/$OUT python bench.py --steps 10 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/ncu.log 2>&1 || true
tail -1 gpurun_out/ncu.log | cut -c1-200
