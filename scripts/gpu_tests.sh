#!/bin/bash
# GPU parity suite, one pytest process per file, every test under a hard timeout (a hung kernel kills only its own file);
# output streams into gpurun_out/ so a killed call still leaves the log.  usage: gpu_tests.sh <tag> [files...]
tag=${1:-r02}; shift
files=${@:-$(ls tests/test_*.py)}
mkdir -p gpurun_out
log=gpurun_out/${tag}_tests.log
: > $log
for f in $files; do
  echo "=== $f" >> $log
  timeout 420 python -u -m pytest $f -m gpu -v --tb=short --timeout=150 --timeout-method=thread -p no:cacheprovider 2>&1 \
    | grep -v "^$" | grep -E "PASSED|FAILED|ERROR|SKIPPED|passed|failed|error|Timeout|assert|Error|^E " | cut -c1-300 >> $log
done
grep -cE "PASSED" $log; grep -E "FAILED|ERROR|Timeout" $log | head -40
