#!/bin/bash
# one `ncu --set full` capture of the same-grid fit kernel (C3 variant): scratch/ncu_fit.sh [n] [case] [tag]
N=${1:-16384}; CASE=${2:-c3}; TAG=${3:-fit}
python scratch/perf_fit.py $N > gpurun_out/perf_${TAG}.txt 2>&1
ncu --set full --import-source on --clock-control none -k regex:"fit_same_grid" -s 3 -c 1 -f -o gpurun_out/ncu_${TAG} python scratch/perf_fit.py $N $CASE > gpurun_out/ncu_${TAG}.log 2>&1
ncu -i gpurun_out/ncu_${TAG}.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_raw.csv 2>/dev/null
cat gpurun_out/perf_${TAG}.txt
