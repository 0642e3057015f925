#!/bin/bash
# round-2 first GPU pass: parity suite, the three bench lines, ncu captures of the segment and angular kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -80 > gpurun_out/r02a_tests.log
for fn in shortest segment simplest; do
  python bench.py --function $fn --steps 5 --warmup 3 > gpurun_out/r02a_bench_$fn.json 2> gpurun_out/r02a_bench_$fn.err
done
ncu --set full --clock-control none --import-source on -k regex:cs_k_segment -s 1 -c 1 -o gpurun_out/r02a_segment \
  python bench.py --function segment --steps 1 --warmup 1 --no-cpu > gpurun_out/r02a_ncu_segment.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cs_k_simplest -s 1 -c 1 -o gpurun_out/r02a_simplest \
  python bench.py --function simplest --steps 1 --warmup 1 --no-cpu > gpurun_out/r02a_ncu_simplest.log 2>&1
tail -3 gpurun_out/r02a_tests.log
