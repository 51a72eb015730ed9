#!/bin/bash
# quick GPU probe (1 GPU): first -bb search on a context, lazy vs eager views
cd "$(dirname "$0")/.."
echo "== default"; MPGPU_PROFILE=1 python tools/bb_search_probe.py c2 1 2>&1 | grep -v " 0.000 ms" | tail -16
echo "== eager"; MPGPU_EAGER_VIEWS=1 MPGPU_PROFILE=1 python tools/bb_search_probe.py c2 1 2>&1 | grep -v " 0.000 ms" | tail -16
