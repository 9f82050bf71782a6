# usage: bash scripts/ncu_full.sh <kernel-regex> <out-name>   (run under gpurun, one GPU)
set -e
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"$1" -s 6 -c 3 -o gpurun_out/$2 -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1 || true
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/$2.ncu-rep
