#!/bin/bash
# A/B on the C2 sweep's e2e time: piece boundaries of the scan plan (MPGPU_SCAN_SPLITS, percent of the visits)
mkdir -p gpurun_out/r02p
run() {
    echo "== $*"
    env "$@" python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-bb --no-cost --no-search --no-c4 --no-bb1000 2>/dev/null | \
        python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('kernel_ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],4))"
}
run A=1
run MPGPU_SCAN_SPLITS=22
run MPGPU_SCAN_SPLITS=22,44
run MPGPU_SCAN_SPLITS=22,44,66
run MPGPU_SCAN_SPLITS=22,66
run MPGPU_SCAN_SPLITS=33,66
run MPGPU_SCAN_SPLITS=44
run MPGPU_SCAN_SPLITS=44,88
run MPGPU_SCAN_SPLITS=65
run MPGPU_SCAN_SPLITS=30,65
MPGPU_PROFILE=4 MPGPU_SCAN_SPLITS=22,44,66 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-bb --no-cost --no-search --no-c4 --no-bb1000 2>&1 | grep timeline
