#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_tests.sh r02f tests/test_gpu_chain.py tests/test_gpu_segment.py tests/test_gpu_slope_transport.py tests/test_gpu_tree.py tests/test_gpu_sharded.py tests/test_gpu_simplest.py > /dev/null
grep -cE "PASSED" gpurun_out/r02f_tests.log; grep -E "FAILED|ERROR|Timeout|^E " gpurun_out/r02f_tests.log | head -40
timeout 400 python bench.py --function segment --steps 5 --warmup 3 --no-cpu > gpurun_out/r02f_bench_segment.json 2> gpurun_out/r02f_bench_segment.err
timeout 300 python bench.py --function simplest --steps 4 --warmup 2 --no-cpu > gpurun_out/r02f_bench_simplest.json 2> gpurun_out/r02f_bench_simplest.err
for f in gpurun_out/r02f_bench_*.json; do echo $f; python -c "
import json,sys
try:
    j=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(j['value']), j['roofline']['frac'], round(j['e2e']['value']), j['kernel_ms_per_step'], j['config'].get('heap_order_replays'), j['roofline']['kernel'])
except Exception as e: print('ERR', e)
"; done
tail -5 gpurun_out/r02f_bench_segment.err
