#!/bin/bash
# GPU session r02j (same recipe as r01k at the round-2 HEAD): full parity suite, bench (ours incl. -cost / -bb sections + reference arm), ncu launch list and
# full captures of the three dominant kernels (k_spr_scan, k_reps_tc, k_sk_scan).
TAG=${1:-r02j}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu_$TAG.log
cat gpurun_out/pytest_gpu_$TAG.log
timeout 300 python bench.py --workload tiny --steps 3 --warmup 3 > gpurun_out/bench_tiny_$TAG.json 2> gpurun_out/bench_tiny_$TAG.err
tail -c 600 gpurun_out/bench_tiny_$TAG.json; tail -5 gpurun_out/bench_tiny_$TAG.err
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 5000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
tail -c 1200 gpurun_out/bench_ref_$TAG.json; tail -5 gpurun_out/bench_ref_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-bb1000 --no-c4 > gpurun_out/ncu_launches_$TAG.log 2>&1
tail -2 gpurun_out/ncu_launches_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spr_scan -s 3 -c 2 -f -o gpurun_out/prof_scan_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-bb1000 --no-c4 --no-bb --no-cost --no-search > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_reps_tc -s 4 -c 1 -f -o gpurun_out/prof_reps_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-bb1000 --no-c4 --no-cost --no-search > gpurun_out/ncu_full_reps_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_reps_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sk_scan -s 3 -c 1 -f -o gpurun_out/prof_sk_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-bb1000 --no-c4 --no-bb --no-search > gpurun_out/ncu_full_sk_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_sk_$TAG.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.csv
