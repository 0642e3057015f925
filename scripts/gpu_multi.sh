#!/bin/bash
# usage: gpu_multi.sh N tag — the N-GPU evidence: sharded == single, bench lines of the three functions, cfg5 at 5 / 10 km
N=${1:-2}; tag=${2:-r02m}; only=${3:-all}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$only" = quick ]; then  # the headline bench line and cfg5 only
  timeout 600 $TR --master-port 29711 bench.py --gpus $N --function shortest --steps 5 --warmup 3 > gpurun_out/${tag}_bench_shortest_${N}gpu.json 2> gpurun_out/${tag}_bench_shortest_${N}gpu.err
  tail -c 700 gpurun_out/${tag}_bench_shortest_${N}gpu.json | head -c 400; echo
  timeout 900 $TR --master-port 29720 scripts/cfg5_sharded.py --km 5 10 > gpurun_out/${tag}_cfg5_${N}gpu.log 2>&1; grep workload gpurun_out/${tag}_cfg5_${N}gpu.log
  exit 0
fi
timeout 600 $TR --master-port 29701 scripts/check_sharded.py > gpurun_out/${tag}_check_sharded.log 2>&1; tail -2 gpurun_out/${tag}_check_sharded.log
p=29710
for fn in shortest segment simplest; do
  p=$((p+1))
  timeout 600 $TR --master-port $p bench.py --gpus $N --function $fn --steps 5 --warmup 3 > gpurun_out/${tag}_bench_${fn}_${N}gpu.json 2> gpurun_out/${tag}_bench_${fn}_${N}gpu.err
  tail -c 600 gpurun_out/${tag}_bench_${fn}_${N}gpu.json | head -c 300; echo
done
timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${tag}_pytest_sharded.log 2>&1; tail -2 gpurun_out/${tag}_pytest_sharded.log
timeout 900 $TR --master-port 29720 scripts/cfg5_sharded.py --km 5 10 > gpurun_out/${tag}_cfg5_${N}gpu.log 2>&1; grep workload gpurun_out/${tag}_cfg5_${N}gpu.log
