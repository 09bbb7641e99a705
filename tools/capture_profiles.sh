mkdir -p gpurun_out/r01d
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r01d/bench.json 2> gpurun_out/r01d/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01d/bench_reference.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01d/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ar_persistent -s 1 -c 1 -o gpurun_out/r01d/ar_full python tools/prof_ar.py 64 bf16 401 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:igemm_tc -s 32 -c 12 -o gpurun_out/r01d/mimi_seanet_full python tools/prof_mimi.py 2000 bf16_tc 1 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:igemm_tc -c 4 -o gpurun_out/r01d/mimi_tr_full python tools/prof_mimi.py 2000 bf16_tc 1 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:attn_tc -c 1 -o gpurun_out/r01d/mimi_attn_full python tools/prof_mimi.py 2000 bf16_tc 1 1 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01d/mimi_launches.csv python tools/prof_mimi.py 10000 bf16_tc 1 1 > /dev/null 2>&1
timeout 200 python tools/gpu_stage_timing.py 64:bf16:0 1:fp32:0 > gpurun_out/r01d/stage_timing.log 2>&1
ls -la gpurun_out/r01d; head -c 1500 gpurun_out/r01d/bench.json
