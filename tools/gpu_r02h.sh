#!/bin/bash
# r02h GPU session (1 GPU): lazy views + k_publish in the SPR search: A/B on the C2 search, then the whole suite.
mkdir -p gpurun_out/r02h
cd "$(dirname "$0")/.."
echo "== lazy + publish"; MPGPU_PROFILE=1 python tools/search_probe.py c2 3 2>&1 | grep -v "^$" | tail -22
echo "== eager views"; SEARCH_PROBE_REF=0 MPGPU_EAGER_VIEWS=1 MPGPU_PROFILE=1 python tools/search_probe.py c2 2 2>&1 | tail -9
echo "== lazy, no publish"; SEARCH_PROBE_REF=0 MPGPU_NO_PUBLISH=1 MPGPU_PROFILE=1 python tools/search_probe.py c2 2 2>&1 | tail -9
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15
