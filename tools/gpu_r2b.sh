#!/bin/bash
# round 2, call B (N GPUs): the sharded bench line at N = $1 (plus the sharding tests on one of the GPUs)
N=${1:-2}
set -x
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_sharding.py -q) > gpurun_out/pytest_shard.log 2>&1
tail -5 gpurun_out/pytest_shard.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
cat gpurun_out/bench_n$N.json; tail -25 gpurun_out/bench_n$N.err
