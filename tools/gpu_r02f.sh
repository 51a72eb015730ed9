#!/bin/bash
# r02f GPU session (1 GPU): full suite at HEAD (first children's up-views in registers, -cost segment bound fix, K9 on
# the device for re-created PLL instances), default bench line, drop-in on C2 -bb 1000 (patched binary only).
mkdir -p gpurun_out/r02f
cd "$(dirname "$0")/.."
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
( time python bench.py > gpurun_out/r02f/bench.json 2> gpurun_out/r02f/bench.err ) 2>&1 | grep real
tail -c 600 gpurun_out/r02f/bench.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r02f/bench.json").read().strip().splitlines()[-1])
print("value %.3e ms %.4f ins/s %.1fM e2e ms %.4f" % (l["value"], l["ms_per_step"], l["insertions_per_s"]/1e6, l["e2e"]["ms_per_step"]))
print("search", json.dumps(l.get("search", {}))[:500])
print("bb1000", json.dumps(l.get("bb1000", {}))[:600])
PY
( time MPBOOT_GPU_STATS=1 MPGPU_PROFILE=2 timeout 1200 python tools/mpboot_dropin_check.py --cases c2_200x100000 --modes bb --skip-stock --out gpurun_out/r02f/x1 ) 2>&1 | cut -c1-1500
grep "mpgpu profile" gpurun_out/r02f/x1/c2_200x100000.bb.gpu.stdout | head -20
grep -n "Iteration 100 \|Iteration 200 \|Optimizing boot\|CPU Time\|Wall-clock" gpurun_out/r02f/x1/c2_200x100000.bb.gpu.stdout | tail -8
cmp gpurun_out/r02f/x1/c2_200x100000.bb.gpu.splits.nex gpurun_out/r02c/x1/c2_200x100000.bb.gpu.splits.nex 2>/dev/null && echo "C2 splits identical to the r02c run (host kernel)"
