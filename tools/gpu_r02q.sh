#!/bin/bash
# GPU session r02q (regression at HEAD after the multi-threaded plan enumeration): full parity suite + the default bench line
TAG=${1:-r02q}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu_$TAG.log
cat gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print('ms_per_step',d['ms_per_step'],'e2e',d['e2e'])
print('search',d['search']['wall_s'],'bb',d['bb']['ms_per_step'],'bb1000',d['bb1000'].get('gpu_search_wall_s'),d['bb1000'].get('identical_outputs'))
"; tail -3 gpurun_out/bench_$TAG.err
