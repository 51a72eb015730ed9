#!/bin/bash
# r02c GPU session (1 GPU): full suite, default bench line, ncu full captures of k_spr_scan at 1 and 4 chunks per warp,
# the drop-in on C2 (-bb 1000, patched binary only: the stock run takes hours).
mkdir -p gpurun_out/r02c
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
( time python bench.py > gpurun_out/r02c/bench.json 2> gpurun_out/r02c/bench.err ) 2>&1 | grep real
tail -c 600 gpurun_out/r02c/bench.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r02c/bench.json").read().strip().splitlines()[-1])
print("value %.3e ms %.4f ins/s %.1fM e2e ms %.4f" % (l["value"], l["ms_per_step"], l["insertions_per_s"]/1e6, l["e2e"]["ms_per_step"]))
print("roofline", json.dumps(l["roofline"])[:600])
print("search", json.dumps(l.get("search", {}))[:900])
print("c4", json.dumps(l.get("c4_strong", {}))[:700])
print("bb1000", json.dumps(l.get("bb1000", {}))[:1200])
print("bb.search", json.dumps(l.get("bb", {}).get("search", {}))[:500])
PY
for vw in 1 4; do
MPGPU_SCAN_VW=$vw timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spr_scan -s 4 -c 1 -f -o gpurun_out/r02c/scan_vw$vw python bench.py --steps 3 --warmup 3 --no-bb --no-cost --no-search --no-cpu-baseline --no-c4 --no-bb1000 > gpurun_out/r02c/ncu_vw$vw.log 2>&1
done
( time MPBOOT_GPU_STATS=1 MPGPU_PROFILE=2 timeout 1200 python tools/mpboot_dropin_check.py --cases c2_200x100000 --modes bb --skip-stock --out gpurun_out/r02c/x1 ) 2>&1 | cut -c1-1500
grep "mpgpu profile" gpurun_out/r02c/x1/c2_200x100000.bb.gpu.stdout | head -20
grep -n "Iteration\|replicates done\|CPU Time\|Wall-clock" gpurun_out/r02c/x1/c2_200x100000.bb.gpu.stdout | tail -25
