#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/probe
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_ref.py tests/test_gpu_bb.py -m gpu -q -x 2>&1 | tail -3
run() { echo "== $1"; env $1 MPBOOT_GPU_STATS=1 python tools/mpboot_dropin_check.py --cases c1_100x5000 --modes bb --skip-stock --out gpurun_out/probe/x 2>&1 | grep -o "search_wall_s\": [0-9.]*\|allocate calls [0-9]* ([a-z0-9 ,-]*) [0-9.]* s\|SPR searches [0-9]* [0-9.]* s (of which -bb [0-9]* [0-9.]* s)\|RAS trees [0-9]* [0-9.]* s\|computeParsimony on the device [0-9]* [0-9.]* s" | tr '\n' ';'; echo; }
run "A=1"; run "MPGPU_LEVEL_VIEWS=1"; run "A=2"; run "MPGPU_LEVEL_VIEWS=1"
SEARCH_PROBE_REF=0 python tools/search_probe.py c2 3 2>&1 | grep optimize_spr | tail -1
