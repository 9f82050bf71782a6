#!/bin/bash
# ncu --set full of the tensor-core contractions of scripts/gemm_probe.py; the report stays on the GPU box (it exceeds the
# 64 MiB gpurun_out limit), its raw page (and the SASS source page of launch $5, if given) come back as CSV.
#   usage: scripts/ncu_gemm.sh <tag> <math> <kernel regex> <count> [source-page launch index]
tag=$1; math=${2:-tf32}; rx=${3:-gemm_tc2}; cnt=${4:-9}; src=$5
export B200NN_PROBE_REPS=1
python scripts/gemm_probe.py $math > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$rx -c $cnt -o /tmp/${tag} -f python scripts/gemm_probe.py $math > gpurun_out/${tag}_ncu.log 2>&1
ncu -i /tmp/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
if [ -n "$src" ]; then
  ncu -i /tmp/${tag}.ncu-rep --page source --csv --print-source sass --launch-skip $src --launch-count 1 > gpurun_out/${tag}_src${src}.csv 2>/dev/null
fi
tail -2 gpurun_out/${tag}_ncu.log
