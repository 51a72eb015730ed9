#!/bin/bash
# A/B on the C2 sweep's e2e time: host threads of the plan enumeration (MPGPU_PLAN_THREADS), pieces
mkdir -p gpurun_out/r02p
run() {
    echo -n "== $*   "
    env "$@" python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-bb --no-cost --no-search --no-c4 --no-bb1000 2>gpurun_out/r02p/err.log | \
        python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('kernel_ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],4))"
}
nproc
for rep in 1 2 3; do
run MPGPU_PLAN_THREADS=1
run MPGPU_PLAN_THREADS=4
run MPGPU_PLAN_THREADS=4 MPGPU_SCAN_PIECES=1
run MPGPU_PLAN_THREADS=3
done
run MPGPU_PLAN_THREADS=6
run MPGPU_PLAN_THREADS=2
