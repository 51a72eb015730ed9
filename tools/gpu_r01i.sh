#!/bin/bash
# GPU session: parity tests, search probe with the library's wall-clock breakdown, bench at C2 and C4.
TAG=${1:-r01i}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu_$TAG.log
cat gpurun_out/pytest_gpu_$TAG.log
MPGPU_PROFILE=1 python tools/search_probe.py c2 3 > gpurun_out/probe_$TAG.log 2>&1
tail -40 gpurun_out/probe_$TAG.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 3000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
timeout 900 python bench.py --workload c4 --steps 10 --warmup 3 --no-bb > gpurun_out/bench_c4_$TAG.json 2> gpurun_out/bench_c4_$TAG.err
tail -c 3000 gpurun_out/bench_c4_$TAG.json; tail -5 gpurun_out/bench_c4_$TAG.err
