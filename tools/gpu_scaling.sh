#!/bin/bash
# one gpurun --gpus 8 call: bench at N = 8 and 4, per-rank timelines of the protocols, PCIe probe
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
cat gpurun_out/bench_n8.json; grep -v Warning gpurun_out/bench_n8.err | tail -5
timeout 600 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --no-configs > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
cat gpurun_out/bench_n4.json; grep -v Warning gpurun_out/bench_n4.err | tail -5
for m in lookback records; do
  LD_TIMELINE=1 timeout 300 $TR --nproc-per-node 8 --master-port 29523 tools/shard_timeline.py $m > /dev/null 2> gpurun_out/timeline_n8_$m.txt
  grep -A12 -- "---- step 3" gpurun_out/timeline_n8_$m.txt | grep "rank [037]\]" | head -40
done
timeout 300 $TR --nproc-per-node 8 --master-port 29524 tools/pcie_probe.py > gpurun_out/pcie_probe_n8.txt 2>&1
grep ranks gpurun_out/pcie_probe_n8.txt
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; nproc >> gpurun_out/topo.txt; numactl -H >> gpurun_out/topo.txt 2>&1; lscpu | head -25 >> gpurun_out/topo.txt
