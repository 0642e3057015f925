#!/bin/bash
# round-2 GPU pass: remaining parity file, the three bench lines, chain-kernel variants, ncu captures
mkdir -p gpurun_out
bash scripts/gpu_tests.sh r02c tests/test_gpu_slope_transport.py > /dev/null
for fn in shortest segment simplest; do
  timeout 400 python bench.py --function $fn --steps 5 --warmup 3 > gpurun_out/r02c_bench_$fn.json 2> gpurun_out/r02c_bench_$fn.err
done
# chain-kernel variants on the bench workload (kernel-only, no CPU leg)
for v in w20 st4; do
  CITYSEER_B200_LIB=$PWD/build/lib_$v.so timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu > gpurun_out/r02c_bench_shortest_$v.json 2> gpurun_out/r02c_bench_shortest_$v.err
done
timeout 200 python scripts/probe.py --cfg cfg4 --nsrc 65536 --reps 2 > gpurun_out/r02c_probe.log 2>&1
for k in segment simplest shortest3; do
  fn=$k; [ $k = shortest3 ] && fn=shortest
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:cs_k_$k -s 1 -c 1 -o gpurun_out/r02c_$k \
    python bench.py --function $fn --steps 1 --warmup 1 --no-cpu > gpurun_out/r02c_ncu_$k.log 2>&1
done
cat gpurun_out/r02c_tests.log | tail -5
for f in gpurun_out/r02c_bench_*.json; do echo $f; python -c "
import json,sys
try:
    j=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(j['value']), j['roofline']['frac'], round(j['e2e']['value']), j['kernel_ms_per_step'])
except Exception as e: print('ERR', e)
"; done
