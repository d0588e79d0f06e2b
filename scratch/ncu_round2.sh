#!/bin/bash
# Round-2 profile set (one GPU, never under torchrun): launch lists + `ncu --set full` captures of the hot kernels.
# Outputs go to gpurun_out/ (scratch); scratch/ncu_summary.py turns them into the JSON summaries under profiles/.
set -x
K='regex:downsample|upsample|fit_same|inpaint|norm_'
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-row-band"
# (1) launch lists: every kernel of the timed steps of C2 / C4 / C3 (cold-cache, serialised: compare SHARES)
for w in c2 c4 c3; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 600 --csv --log-file gpurun_out/r02_launches_$w.csv $B --workload $w > gpurun_out/r02_launches_$w.log 2>&1
done
# (2) full captures: one band's kernels of a C2 step and of a C4 step (after the warm-up steps)
timeout 400 ncu --set full --import-source on --clock-control none -k "$K" -s 140 -c 7 -f -o gpurun_out/r02_full_c2 $B --workload c2 > gpurun_out/r02_full_c2.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k "$K" -s 120 -c 6 -f -o gpurun_out/r02_full_c4 $B --workload c4 > gpurun_out/r02_full_c4.log 2>&1
# (3) the same-grid fit kernel (C3 variant, 16384 x 16384) and SrcSpace's reference up-sampler
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"fit_same_grid" -s 3 -c 1 -f -o gpurun_out/r02_full_fit_c3 python scratch/perf_fit.py 16384 c3 > gpurun_out/r02_full_fit_c3.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"fit_same_grid" -s 3 -c 1 -f -o gpurun_out/r02_full_fit_gbo15 python scratch/perf_fit.py 16384 gbo15 > gpurun_out/r02_full_fit_gbo15.log 2>&1
# gpurun brings back at most 64 MiB: keep the raw-metric pages and the per-opcode mix, drop the big reports (the fit one stays)
for r in r02_full_c2 r02_full_c4 r02_full_fit_c3 r02_full_fit_gbo15; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null
done
python scratch/ncu_opmix.py gpurun_out/r02_full_fit_c3.ncu-rep 268435456 > gpurun_out/r02_full_fit_c3_opmix.txt 2>&1
python scratch/ncu_opmix.py gpurun_out/r02_full_fit_gbo15.ncu-rep 268435456 > gpurun_out/r02_full_fit_gbo15_opmix.txt 2>&1
rm -f gpurun_out/r02_full_c2.ncu-rep gpurun_out/r02_full_c4.ncu-rep gpurun_out/r02_full_fit_gbo15.ncu-rep
ls -la gpurun_out/r02_*
