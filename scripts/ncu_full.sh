# usage: bash scripts/ncu_full.sh <kernel-regex> <out-name> [bench args]   (run under gpurun, one GPU)
set -e
RE=$1; OUT=$2; shift; shift
python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/plain_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s ${NCU_SKIP:-6} -c ${NCU_COUNT:-4} -o gpurun_out/$OUT -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ncu_full.log 2>&1 || true
tail -2 gpurun_out/ncu_full.log
