#!/bin/bash
# quick GPU probe (1 GPU): stream-K contraction
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_bb.py -m gpu -q -x 2>&1 | tail -12
for sk in 1 0; do echo "== stream_k $sk"; MPGPU_REPS_STREAMK=$sk BB_PROBE_SEARCH=0 timeout 300 python tools/bb_probe.py c2 1000 2>&1 | grep "bb step\|differ\|equal\|check\|Error\|error" | tail -3; done
