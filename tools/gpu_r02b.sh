#!/bin/bash
# r02b GPU session: slot-reuse planner (7 stack slots) parity, scan variants on C2, drop-in timing with the known-topology
# table, ncu launch list of the drop-in's -bb run.
mkdir -p gpurun_out/r02b
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sankoff.py tests/test_gpu_bb.py -m gpu -x -q 2>&1 | tail -3
for cfg in "1 0" "2 0" "4 0" "4 1"; do
  set -- $cfg
  MPGPU_SCAN_VW=$1 MPGPU_SCAN_PF4=$2 python bench.py --steps 20 --warmup 3 --no-bb --no-cost --no-search --no-cpu-baseline > gpurun_out/r02b/bench_vw$1_pf$2.json 2> gpurun_out/r02b/bench_vw$1_pf$2.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/r02b/bench_vw$1_pf$2.json").read().strip().splitlines()[-1])
print("VW=$1 PF4=$2 ms_per_step %.4f ins/s %.1fM e2e_ms %.4f" % (l["ms_per_step"], l["insertions_per_s"]/1e6, l["e2e"]["ms_per_step"]))
PY
done
for wpb in 2 8; do
MPGPU_SCAN_VW=2 MPGPU_SCAN_WPB=$wpb python bench.py --steps 20 --warmup 3 --no-bb --no-cost --no-search --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('VW=2 WPB=$wpb ms_per_step %.4f ins/s %.1fM' % (l['ms_per_step'], l['insertions_per_s']/1e6))"
done
MPGPU_PROFILE=2 python tools/mpboot_dropin_check.py --cases c1_100x5000 --modes bb --golden tests/golden/mpboot --out gpurun_out/r02b/x1 > gpurun_out/r02b/dropin.log 2>&1
cat gpurun_out/r02b/dropin.log | cut -c1-900
grep "mpgpu profile" gpurun_out/r02b/x1/c1_100x5000.bb.gpu.stdout | head -20
python tools/mpboot_dropin_check.py --cases c1_17x1998,c1_12x300 --modes bb --golden tests/golden/mpboot --out gpurun_out/r02b/x1 2>&1 | cut -c1-700
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 40000 --csv --log-file gpurun_out/r02b/dropin_launches.csv integration/_bin/mpboot-avx-gpu -s gpurun_out/r02b/x1/c1_100x5000.phy -seed 1 -bb 1000 -pre gpurun_out/r02b/x1/ncu > gpurun_out/r02b/ncu_run.log 2>&1
python tools/launch_summary.py gpurun_out/r02b/dropin_launches.csv
