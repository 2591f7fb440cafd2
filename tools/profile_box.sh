#!/bin/bash
# Runs ON the GPU box (under gpurun, ONE GPU): the ncu evidence kept under profiles/.
#   1. kernel launch list of one bench step (time shares)
#   2. full capture of one steady-state 8-frame launch of the single-sequence persistent decode kernel
#   3. full capture of one launch of the batched decode kernel (32 slots)
#   4. full captures of the tcgen05 GEMM launches of one codec window decode (T=33)
# Only the CSV exports travel back (gpurun_out/ is capped at 64 MiB); the decode kernel's .ncu-rep is kept as well.
# usage: tools/profile_box.sh r2   -> gpurun_out/{launches,prof_decode,prof_batch,prof_gemm}_<tag>*
R=${1:-r2}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference --no-stateful"
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 6000 --csv --log-file gpurun_out/launches_$R.csv \
    $B --batch 0 > gpurun_out/ncu_launch_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fq3_decode_kernel -s 3 -c 1 -f -o gpurun_out/prof_decode_$R \
    $B --batch 0 > gpurun_out/ncu_full_$R.log 2>&1
ncu -i gpurun_out/prof_decode_$R.ncu-rep --page raw --csv > gpurun_out/prof_decode_${R}_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:fq3_decode_batch_kernel -s 2 -c 1 -f -o /tmp/prof_batch_$R \
    python tools/batch_bench.py --batches 32 --frames 16 > gpurun_out/ncu_batch_$R.log 2>&1
ncu -i /tmp/prof_batch_$R.ncu-rep --page raw --csv > gpurun_out/prof_batch_${R}_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:conv_gemm_tc_kernel -s 200 -c 24 -f -o /tmp/prof_gemm_$R \
    python tools/codec_bench3.py --variants tcgen05 --cases 1x33 > gpurun_out/ncu_gemm_$R.log 2>&1
ncu -i /tmp/prof_gemm_$R.ncu-rep --page raw --csv > gpurun_out/prof_gemm_${R}_raw.csv 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out | grep -E "prof_|launches_" | awk '{print $5, $9}'
