#!/bin/bash
# -cost sections of the bench on C2 / C3 / C5 (DNA, AA, 32-state)
TAG=${1:-cost}
mkdir -p gpurun_out
for W in ${WL:-c2 c3 c5}; do
  timeout 900 python bench.py --workload $W --steps 5 --no-search 2>gpurun_out/cost_${W}_$TAG.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'workload': d['config']['workload'], 'fitch_insertions_per_s': d['insertions_per_s'], 'fitch_frac': d['roofline']['frac'], 'cost': d['cost']}))" | tee gpurun_out/cost_${W}_$TAG.json
  tail -3 gpurun_out/cost_${W}_$TAG.err
done
