#!/bin/bash
# quick Sankoff iteration: parity tests, -cost bench section, one full ncu capture of k_sk_scan
TAG=${1:-sk}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sankoff.py -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 5 --no-bb --no-cpu-baseline --no-search 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps(d['cost']))" | tee gpurun_out/cost_$TAG.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sk_scan -s 3 -c 1 -f -o gpurun_out/prof_sk_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-bb --no-search > gpurun_out/ncu_full_sk_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_sk_$TAG.log
