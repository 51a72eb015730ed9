#!/bin/bash
cd "$(dirname "$0")/.."
for w in 16 4; do echo "== MPGPU_WAVE_WARPS=$w"; SEARCH_PROBE_REF=0 MPGPU_WAVE_WARPS=$w MPGPU_PROFILE=1 python tools/search_probe.py c2 3 2>&1 | grep "optimize_spr\|scan wait" | tail -2; done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_ref.py tests/test_gpu_bb.py -m gpu -q -x 2>&1 | tail -3
