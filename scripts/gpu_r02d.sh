#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_tests.sh r02d > /dev/null
for fn in segment simplest; do
  timeout 400 python bench.py --function $fn --steps 5 --warmup 3 --no-cpu > gpurun_out/r02d_bench_$fn.json 2> gpurun_out/r02d_bench_$fn.err
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02d_bench_shortest.json 2> gpurun_out/r02d_bench_shortest.err
for v in nopf nocl; do
  CITYSEER_B200_LIB=$PWD/build/lib_$v.so timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu > gpurun_out/r02d_bench_shortest_$v.json 2> gpurun_out/r02d_bench_shortest_$v.err
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cs_k_segment -s 2 -c 1 -o gpurun_out/r02d_segment \
    python bench.py --function segment --steps 1 --warmup 1 --no-cpu > gpurun_out/r02d_ncu_segment.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cs_k_simplest -s 1 -c 1 -o gpurun_out/r02d_simplest \
    python bench.py --function simplest --steps 1 --warmup 1 --no-cpu > gpurun_out/r02d_ncu_simplest.log 2>&1
grep -cE "PASSED" gpurun_out/r02d_tests.log; grep -E "FAILED|ERROR|Timeout" gpurun_out/r02d_tests.log | head -20
for f in gpurun_out/r02d_bench_*.json; do echo $f; python -c "
import json,sys
try:
    j=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(j['value']), j['roofline']['frac'], round(j['e2e']['value']), j['kernel_ms_per_step'], j['config'].get('heap_order_replays'))
except Exception as e: print('ERR', e)
"; done
