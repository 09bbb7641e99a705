#!/usr/bin/env bash
# Everything under profiles/ for one round comes from this script (run on the GPU box through gpurun):
#   bash tools/capture_profiles.sh r01e
set -u
R=${1:-rXX}
O=gpurun_out/$R
mkdir -p "$O"
timeout 500 python bench.py --steps 5 --warmup 3 > "$O/bench.json" 2> "$O/bench.err"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > "$O/bench_reference.json" 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$O/launches.csv" \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ar_persistent -s 1 -c 1 -o "$O/ar_full" \
  python tools/prof_ar.py 64 bf16 401 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:igemm_tc -s 32 -c 12 -o "$O/mimi_seanet_full" \
  python tools/prof_mimi.py 2000 bf16_tc 1 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:igemm_tc -c 4 -o "$O/mimi_tr_full" \
  python tools/prof_mimi.py 2000 bf16_tc 1 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:resblock_tc -c 3 -o "$O/mimi_resblock_full" \
  python tools/prof_mimi.py 2000 bf16_tc 1 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:attn_tc -c 1 -o "$O/mimi_attn_full" \
  python tools/prof_mimi.py 2000 bf16_tc 1 1 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$O/mimi_launches.csv" \
  python tools/prof_mimi.py 10000 bf16_tc 1 1 > /dev/null 2>&1
timeout 200 python tools/gpu_stage_timing.py 64:bf16:0 1:fp32:0 > "$O/stage_timing.log" 2>&1
ls -la "$O"
