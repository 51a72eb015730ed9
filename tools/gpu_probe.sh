#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_bb.py -m gpu -q -x -k "tensor_path_equals_exact and asymmetric" 2>&1 | tail -40
