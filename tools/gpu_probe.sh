#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_sankoff.py -m gpu -q -x -k "short_off" 2>&1 | tail -12
timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -q -x -k "cost" 2>&1 | tail -5
