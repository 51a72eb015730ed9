#!/bin/bash
# quick GPU probe (1 GPU): asymmetric Sankoff
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_gpu_sankoff.py -m gpu -q -x -k "asymmetric or preconditions" 2>&1 | tail -15
