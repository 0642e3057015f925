#!/bin/bash
mkdir -p gpurun_out
for v in a3h512 a4h512 a6h512 a5h256 a6h256 a8h256; do
  CITYSEER_B200_LIB=$PWD/build/lib_$v.so timeout 200 python bench.py --function simplest --steps 4 --warmup 2 --no-cpu > gpurun_out/r02e_bench_simplest_$v.json 2> gpurun_out/r02e_bench_simplest_$v.err
  echo $v; python -c "
import json
try:
    j=json.loads(open('gpurun_out/r02e_bench_simplest_$v.json').read().strip().splitlines()[-1]); print(round(j['value']), j['kernel_ms_per_step'])
except Exception as e: print('ERR', e)
"
done
