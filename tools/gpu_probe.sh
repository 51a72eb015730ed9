#!/bin/bash
# quick GPU probe (1 GPU): k_reps_tc with the A operand in tensor memory
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_bb.py -m gpu -q -x 2>&1 | tail -12
for ta in 1 0; do echo "== tmem_a $ta"; MPGPU_REPS_TMEM_A=$ta BB_PROBE_SEARCH=0 timeout 300 python tools/bb_probe.py c2 1000 2>&1 | grep "bb step\|differ\|equal\|check\|Error\|error" | tail -3; done
