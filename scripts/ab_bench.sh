#!/bin/bash
# usage: ab_bench.sh <tag> <function> "variant ..." — bench.py with each build/lib_<variant>.so ("default" = in-tree)
tag=$1; fn=$2; shift 2
mkdir -p gpurun_out
for v in $@; do
  if [ $v = default ]; then unset CITYSEER_B200_LIB; else export CITYSEER_B200_LIB=$PWD/build/lib_$v.so; fi
  timeout 300 python bench.py --function $fn --steps 5 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_${fn}_$v.json 2> gpurun_out/${tag}_bench_${fn}_$v.err
  python -c "
import json
j=json.loads(open('gpurun_out/${tag}_bench_${fn}_$v.json').read().strip().splitlines()[-1]); print('$fn $v', round(j['value']), round(j['roofline']['frac'],4), round(j['e2e']['value']), round(j['kernel_ms_per_step'],2))"
done
