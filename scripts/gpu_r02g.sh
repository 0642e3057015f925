#!/bin/bash
mkdir -p gpurun_out
for v in s20 s24 s32; do
  CITYSEER_B200_LIB=$PWD/build/lib_$v.so timeout 300 python bench.py --function segment --steps 4 --warmup 2 --no-cpu > gpurun_out/r02g_bench_segment_$v.json 2> gpurun_out/r02g_bench_segment_$v.err
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cs_k_segment3 -s 1 -c 1 -o gpurun_out/r02g_segment3 \
    python bench.py --function segment --steps 1 --warmup 1 --no-cpu > gpurun_out/r02g_ncu_segment3.log 2>&1
for f in gpurun_out/r02g_bench_*.json; do echo $f; python -c "
import json,sys
try:
    j=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(j['value']), j['roofline']['frac'], round(j['e2e']['value']), j['kernel_ms_per_step'], j['config'].get('heap_order_replays'), j['roofline']['kernel'])
except Exception as e: print('ERR', e)
"; done
