#!/bin/bash
# One GPU session: parity tests, bench (ours + reference arm), ncu launch list + full captures of the two
# dominant kernels.  Usage (from the repo root on the GPU box):  bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu_$TAG.log
cat gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 6000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
tail -c 1500 gpurun_out/bench_ref_$TAG.json; tail -5 gpurun_out/bench_ref_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1
tail -3 gpurun_out/ncu_launches_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:k_spr_scan -s 3 -c 2 -f -o gpurun_out/prof_scan_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-bb > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:k_reps_tc -s 4 -c 1 -f -o gpurun_out/prof_reps_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_reps_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_reps_$TAG.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.csv
