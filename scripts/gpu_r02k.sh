#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_tests.sh r02k > /dev/null
grep -cE "PASSED" gpurun_out/r02k_tests.log; grep -E "FAILED|ERROR|Timeout|^E " gpurun_out/r02k_tests.log | head -20
timeout 200 python scripts/probe.py --cfg cfg2 --reps 3 > gpurun_out/r02k_probe_cfg2.log 2>&1; tail -1 gpurun_out/r02k_probe_cfg2.log | cut -c1-600
timeout 300 python scripts/probe.py --cfg cfg5 --nsrc 32768 --distances 5000 --reps 2 > gpurun_out/r02k_probe_cfg5.log 2>&1; tail -1 gpurun_out/r02k_probe_cfg5.log | cut -c1-600
timeout 300 python scripts/probe.py --cfg cfg4 --fn segment --distances 400,800,1600 --opt kernel=1 --reps 2 > gpurun_out/r02k_probe_seg_node.log 2>&1; tail -1 gpurun_out/r02k_probe_seg_node.log | cut -c1-300
timeout 300 python scripts/probe.py --cfg cfg4 --nsrc 65536 --opt kernel=1 --reps 2 > gpurun_out/r02k_probe_arena_cfg4.log 2>&1; tail -1 gpurun_out/r02k_probe_arena_cfg4.log | cut -c1-600
