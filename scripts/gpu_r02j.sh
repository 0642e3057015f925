#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_tests.sh r02j tests/test_gpu_segment.py tests/test_gpu_chain.py tests/test_gpu_tree.py > /dev/null
grep -cE "PASSED" gpurun_out/r02j_tests.log; grep -E "FAILED|ERROR|Timeout|^E " gpurun_out/r02j_tests.log | head -20
timeout 300 python bench.py --function segment --steps 5 --warmup 3 --no-cpu > gpurun_out/r02j_bench_segment.json 2> gpurun_out/r02j_bench_segment.err
python -c "
import json
j=json.loads(open('gpurun_out/r02j_bench_segment.json').read().strip().splitlines()[-1]); print(round(j['value']), j['roofline']['frac'], round(j['e2e']['value']), j['kernel_ms_per_step'], j['config'].get('heap_order_replays'))"
timeout 200 python scripts/probe.py --cfg cfg2 --reps 3 > gpurun_out/r02j_probe_cfg2.log 2>&1; tail -1 gpurun_out/r02j_probe_cfg2.log | cut -c1-500
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cs_k_shortestILi3 -s 1 -c 1 -o gpurun_out/r02j_arena_cfg2 \
    python scripts/probe.py --cfg cfg2 --reps 2 > gpurun_out/r02j_ncu_arena.log 2>&1
timeout 300 python scripts/probe.py --cfg cfg5 --nsrc 32768 --distances 5000 --reps 2 > gpurun_out/r02j_probe_cfg5.log 2>&1; tail -1 gpurun_out/r02j_probe_cfg5.log | cut -c1-500
