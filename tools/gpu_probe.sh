#!/bin/bash
# which change slowed the -bb searches of the drop-in on 100 x 5000?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/probe
run() { echo "== $1"; env $1 MPBOOT_GPU_STATS=1 MPGPU_PROFILE=2 python tools/mpboot_dropin_check.py --cases c1_100x5000 --modes bb --skip-stock --out gpurun_out/probe/x 2>&1 | grep -o "search_wall_s\": [0-9.]*\|SPR searches [0-9]* [0-9.]* s (of which -bb [0-9]* [0-9.]* s)"; grep "mpgpu profile" gpurun_out/probe/x/c1_100x5000.bb.gpu.stdout | tr '\n' ';' | cut -c1-900; echo; }
run "A=1"
run "MPGPU_REPS_TMEM_A=0"
run "MPGPU_EAGER_VIEWS=1"
run "MPGPU_NO_LEAN=1"
run "MPGPU_SPLIT_DEPTH=0"
