#!/bin/bash
# r02l (N GPUs): sharded_check incl. replicate shards, then the bench line at N GPUs
N=${1:-2}
mkdir -p gpurun_out/r02l
cd "$(dirname "$0")/.."
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py 2>&1 | grep -v "^W\|^\[W\|warn\|^\*\|OMP_NUM" | tail -22
bash tools/gpu_scale2.sh $N r02l
