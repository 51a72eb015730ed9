#!/bin/bash
# Multi-GPU session (gpurun --gpus N): bench at N ranks, weak scaling on C2 (the driver's line) and strong
# scaling on C4 (1000 taxa x 1M sites sharded over the ranks); at N=2 also the sharded parity check.
N=${1:-2}; TAG=${2:-r01}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_c2_n${N}_$TAG.json 2> gpurun_out/bench_c2_n${N}_$TAG.err
tail -c 1500 gpurun_out/bench_c2_n${N}_$TAG.json; tail -3 gpurun_out/bench_c2_n${N}_$TAG.err
timeout 1200 $TR bench.py --gpus $N --workload c4 --scaling strong --steps 10 --warmup 3 > gpurun_out/bench_c4_n${N}_$TAG.json 2> gpurun_out/bench_c4_n${N}_$TAG.err
tail -c 1500 gpurun_out/bench_c4_n${N}_$TAG.json; tail -3 gpurun_out/bench_c4_n${N}_$TAG.err
if [ "$N" = "2" ]; then
  timeout 600 $TR tools/sharded_check.py > gpurun_out/sharded_check_$TAG.log 2>&1; tail -12 gpurun_out/sharded_check_$TAG.log
fi
