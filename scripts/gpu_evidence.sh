#!/bin/bash
# End-of-round evidence on one B200: parity suite, the three bench lines (+ reference arms), ncu launch lists and full
# captures of the dominant kernel of each function, compute-sanitizer over the new kernels.
# usage: gpu_evidence.sh <tag> [a|b]   (gpurun brings back at most 64 MiB per call: part b = the two large ncu reports)
tag=${1:-r02z}; part=${2:-a}
mkdir -p gpurun_out
if [ "$part" = b ]; then
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:cs_k_shortest3 -s 1 -c 1 -o gpurun_out/${tag}_shortest3 \
      python bench.py --function shortest --steps 1 --warmup 1 --no-cpu > gpurun_out/${tag}_ncu_shortest3.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:cs_k_segment3 -s 1 -c 1 -o gpurun_out/${tag}_segment3 \
      python bench.py --function segment --steps 1 --warmup 1 --no-cpu > gpurun_out/${tag}_ncu_segment3.log 2>&1
  ls -la gpurun_out/${tag}_*.ncu-rep
  exit 0
fi
bash scripts/gpu_tests.sh $tag > /dev/null
grep -cE "PASSED" gpurun_out/${tag}_tests.log; grep -E "FAILED|ERROR|Timeout|^E " gpurun_out/${tag}_tests.log | head -30
for fn in shortest segment simplest; do
  timeout 500 python bench.py --function $fn --steps 5 --warmup 3 > gpurun_out/${tag}_bench_$fn.json 2> gpurun_out/${tag}_bench_$fn.err
  timeout 300 python bench.py --impl reference --function $fn --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference_$fn.json 2> gpurun_out/${tag}_bench_reference_$fn.err
done
for fn in shortest segment simplest; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches_$fn.csv \
    python bench.py --function $fn --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cs_k_simplest -s 1 -c 1 -o gpurun_out/${tag}_simplest \
    python bench.py --function simplest --steps 1 --warmup 1 --no-cpu > gpurun_out/${tag}_ncu_simplest.log 2>&1
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_segment.py tests/test_gpu_simplest.py tests/test_gpu_tree.py \
      "tests/test_gpu_chain.py::test_segment_decomposed_grid" "tests/test_gpu_chain.py::test_segment_regular_grid_is_replayed_in_heap_order" \
      "tests/test_gpu_chain.py::test_decomposed_grid_with_tolerance" tests/test_gpu_slope_transport.py tests/test_gpu_od.py -m gpu -q -x -p no:cacheprovider \
      > gpurun_out/${tag}_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${tag}_sanitizer_$tool.log | tail -3
done
for f in gpurun_out/${tag}_bench_*.json; do echo $f; python -c "
import json,sys
try:
    j=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(j['value']), (j.get('roofline') or {}).get('frac'), round(j['e2e']['value']), j.get('kernel_ms_per_step'))
except Exception as e: print('ERR', e)
"; done
