#!/bin/bash
# GPU session r01j: full parity suite (Fitch + -bb + Sankoff), bench with the -cost section, ncu launch list and
# one full capture of k_sk_scan.
TAG=${1:-r01j}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu_$TAG.log
cat gpurun_out/pytest_gpu_$TAG.log
timeout 300 python bench.py --workload tiny --steps 3 --warmup 3 > gpurun_out/bench_tiny_$TAG.json 2> gpurun_out/bench_tiny_$TAG.err
tail -c 1500 gpurun_out/bench_tiny_$TAG.json; tail -5 gpurun_out/bench_tiny_$TAG.err
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 4000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sk_scan -s 3 -c 1 -f -o gpurun_out/prof_sk_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-bb --no-search > gpurun_out/ncu_full_sk_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_sk_$TAG.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.csv
