#!/bin/bash
# Runs ON the GPU box (under gpurun): kernel launch list of one bench step, one full capture of a steady-state
# 8-frame launch of the persistent decode kernel, raw CSV exports for tools/ncu_summary.py.  usage: profile_box.sh r1c
R=${1:-r1c}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 6000 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fq3_decode_kernel -s 3 -c 1 -f -o gpurun_out/prof_decode_$R \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$R.log 2>&1
ncu -i gpurun_out/prof_decode_$R.ncu-rep --page raw --csv > gpurun_out/prof_decode_${R}_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
