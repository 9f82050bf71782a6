set -e
python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/ncu.log 2>&1 || true
tail -2 gpurun_out/ncu.log
