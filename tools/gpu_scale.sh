#!/bin/bash
# strong-scaling bench + PCIe probe at N ranks (usage: tools/gpu_scale.sh N tag)
N=${1:-2}; tag=${2:-s}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/pcie_probe.py > gpurun_out/pcie_n${N}_$tag.json 2> gpurun_out/pcie_n${N}_$tag.err; echo "pcie rc=$?"; cat gpurun_out/pcie_n${N}_$tag.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_n${N}_$tag.json | cut -c1-3000; tail -5 gpurun_out/bench_n${N}_$tag.err
