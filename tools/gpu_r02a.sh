#!/bin/bash
# r02a GPU session: parity of the widened scan kernel (forced 2 and 4 chunks per warp), full suite, scan-kernel variants
# on C2, the drop-in's -bb 1000 run with the library's cumulative host profile.
mkdir -p gpurun_out/r02a
cd "$(dirname "$0")/.."
for vw in 2 4; do
  MPGPU_SCAN_VW=$vw timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bb.py tests/test_gpu_vs_ref.py -m gpu -x -q 2>&1 | tail -3 | sed "s/^/[VW=$vw] /"
done
MPGPU_SCAN_VW=4 MPGPU_SCAN_PF4=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2 | sed "s/^/[VW=4 PF] /"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for cfg in "1 0" "2 0" "4 0" "4 1"; do
  set -- $cfg
  MPGPU_SCAN_VW=$1 MPGPU_SCAN_PF4=$2 python bench.py --steps 20 --warmup 3 --no-bb --no-cost --no-search --no-cpu-baseline > gpurun_out/r02a/bench_vw$1_pf$2.json 2> gpurun_out/r02a/bench_vw$1_pf$2.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/r02a/bench_vw$1_pf$2.json").read().strip().splitlines()[-1])
print("VW=$1 PF4=$2 ms_per_step %.4f ins/s %.1fM e2e_ms %.4f" % (l["ms_per_step"], l["insertions_per_s"]/1e6, l["e2e"]["ms_per_step"]))
PY
done
MPGPU_PROFILE=2 python tools/mpboot_dropin_check.py --cases c1_100x5000 --modes plain,bb --golden tests/golden/mpboot --out gpurun_out/r02a/x1 > gpurun_out/r02a/dropin.log 2>&1
cat gpurun_out/r02a/dropin.log | cut -c1-900
grep "mpgpu profile" gpurun_out/r02a/x1/c1_100x5000.bb.gpu.stdout | head -20
