#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_tests.sh r02n tests/test_gpu_chain.py tests/test_gpu_shortest.py tests/test_gpu_segment.py tests/test_gpu_slope_transport.py > /dev/null
grep -cE "PASSED" gpurun_out/r02n_tests.log; grep -E "FAILED|ERROR|Timeout|^E " gpurun_out/r02n_tests.log | head -20
for fn in shortest segment; do
timeout 300 python bench.py --function $fn --steps 5 --warmup 3 --no-cpu > gpurun_out/r02n_bench_$fn.json 2> gpurun_out/r02n_bench_$fn.err
python -c "
import json
j=json.loads(open('gpurun_out/r02n_bench_$fn.json').read().strip().splitlines()[-1]); print('$fn', round(j['value']), j['roofline']['frac'], round(j['e2e']['value']), j['kernel_ms_per_step'], j['config'].get('heap_order_replays'))"
done
timeout 200 python scripts/probe.py --cfg cfg4 --nsrc 65536 --reps 2 > gpurun_out/r02n_probe.log 2>&1; tail -1 gpurun_out/r02n_probe.log | cut -c1-700
