#!/bin/bash
# quick GPU probe (1 GPU): staged multi-piece plans, number of pieces
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_ref.py -m gpu -q -x 2>&1 | tail -3
B="python bench.py --no-bb --no-cost --no-search --no-cpu-baseline --no-bb1000 --no-c4"
P='import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("bench ms %.4f e2e ms %.4f e2e ins/s %.1fM" % (l["ms_per_step"], l["e2e"]["ms_per_step"], l["e2e"]["insertions_per_s"]/1e6))'
echo "== old (memcpy, 2 pieces)"; MPGPU_NO_LEAN=1 $B 2>/dev/null | python -c "$P"
for p in 2 3 4 6 8 12; do echo "== staged, $p pieces"; MPGPU_SCAN_PIECES=$p $B 2>/dev/null | python -c "$P"; done
