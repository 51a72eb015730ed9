#!/bin/bash
# quick GPU probe (1 GPU)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/probe
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30
