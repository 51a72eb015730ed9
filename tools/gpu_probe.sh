#!/bin/bash
cd "$(dirname "$0")/.."
SEARCH_PROBE_REF=1 MPGPU_PROFILE=1 python tools/search_probe.py c2 3 2>&1 | grep "optimize_spr\|identical\|plan+launch\|scan wait\|replay" | tail -5
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_ref.py tests/test_gpu_bb.py tests/test_gpu_sankoff.py -m gpu -q -x 2>&1 | tail -3
