#!/bin/bash
# Runs ON the GPU box (under gpurun): end-of-round check -- smoke(), the whole `-m gpu` suite, the default bench line and the
# ncu launch list of one bench step.  Outputs under gpurun_out/ (r2p_*).  usage: bash tools/final_check_box.sh
mkdir -p gpurun_out; python __graft_entry__.py smoke 2>&1 | tail -1; (time timeout 1000 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -14) > gpurun_out/r2p_pytest.log 2>&1; tail -9 gpurun_out/r2p_pytest.log; (timeout 900 python bench.py) > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; tail -2 gpurun_out/r2p_bench.err; python -c "
import json
for l in open('gpurun_out/r2p_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); c=d['config']; print(round(d['value'],2), round(d['e2e']['value'],2), c['ttfa_ms_p50'], c['ttfa_ms_e2e_p50'], c.get('ttfa_ms_e2e_cached_voice_p50'), d['roofline']['frac'], d.get('stateful_codec',{}).get('rtf'), {k:round(v,1) for k,v in d['config4'].items() if k.startswith('rtf')}, d['gpu_reference'].get('rtf'), d['gpu_reference'].get('ttfa_ms_p50'), d['ms_per_step'], d['gpu_launches'])"; ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 6000 --csv --log-file gpurun_out/launches_r2p.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference --no-stateful --batch 0 > gpurun_out/ncu_launch_r2p.log 2>&1; ls -la gpurun_out | awk '{print $5, $9}'
