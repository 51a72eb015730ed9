#!/bin/bash
# quick GPU probe (1 GPU): lean batches (staged plan, fused publish), A/B in one call
cd "$(dirname "$0")/.."
echo "== default"; SEARCH_PROBE_REF=1 MPGPU_PROFILE=1 python tools/search_probe.py c2 3 2>&1 | grep "optimize_spr\|identical\|plan+launch\|scan wait\|replay" | tail -5
echo "== split 4"; SEARCH_PROBE_REF=0 MPGPU_SPLIT_DEPTH=4 python tools/search_probe.py c2 3 2>&1 | grep "optimize_spr" | tail -1
echo "== split 0"; SEARCH_PROBE_REF=0 MPGPU_SPLIT_DEPTH=0 python tools/search_probe.py c2 3 2>&1 | grep "optimize_spr" | tail -1
echo "== no lean"; SEARCH_PROBE_REF=0 MPGPU_NO_LEAN=1 python tools/search_probe.py c2 3 2>&1 | grep "optimize_spr" | tail -1
echo "== no lean, split 0, eager"; SEARCH_PROBE_REF=0 MPGPU_NO_LEAN=1 MPGPU_SPLIT_DEPTH=0 MPGPU_EAGER_VIEWS=1 python tools/search_probe.py c2 3 2>&1 | grep "optimize_spr" | tail -1
echo "== default again"; SEARCH_PROBE_REF=0 python tools/search_probe.py c2 3 2>&1 | grep "optimize_spr" | tail -1
python bench.py --no-bb --no-cost --no-search --no-cpu-baseline --no-c4 --no-bb1000 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench ms %.4f ins/s %.1fM e2e ms %.4f' % (l['ms_per_step'], l['insertions_per_s']/1e6, l['e2e']['ms_per_step']))"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_ref.py tests/test_gpu_bb.py -m gpu -q -x 2>&1 | tail -3
