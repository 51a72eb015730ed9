#!/bin/bash
# GPU session r02p (final session of round 2 at HEAD): full parity suite, bench (ours incl. -cost / -bb / search / bb1000 sections + reference arm),
# ncu launch list of the bench, compute-sanitizer memcheck over the kernels added in this session (k_keep_rows, k_seg_sums, k_prefix_max,
# multi-block k_publish).  The ncu --set full captures of the three dominant kernels are those of r02n (the kernels did not change).
TAG=${1:-r02p}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu_$TAG.log
cat gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 3000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
tail -c 1200 gpurun_out/bench_ref_$TAG.json; tail -5 gpurun_out/bench_ref_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-bb1000 --no-c4 > gpurun_out/ncu_launches_$TAG.log 2>&1
tail -2 gpurun_out/ncu_launches_$TAG.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_bb.py -q -x \
    -k "(prefix_max and 12-300) or (distinct_iter and 12-300 and tensor) or (min_iter1 and 12-300)" > gpurun_out/sanitizer_$TAG.log 2>&1
echo "sanitizer rc=$?"; tail -4 gpurun_out/sanitizer_$TAG.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.csv
