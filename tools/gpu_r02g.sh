#!/bin/bash
# r02g GPU session (1 GPU): morph drop-in with the IQ-TREE-tree kernel cross-check, then the whole suite without -x.
mkdir -p gpurun_out/r02g
cd "$(dirname "$0")/.."
MPBOOT_GPU_CHECK_K9=1 python tools/mpboot_dropin_check.py --cases morph_16x400 --modes plain --golden tests/golden/mpboot --out gpurun_out/r02g/chk 2>&1 | cut -c1-1500
grep -n "libmpgpu\|ERROR\|differ" gpurun_out/r02g/chk/morph_16x400.plain.gpu.stdout | head
MPBOOT_GPU_HOST_K9=1 python tools/mpboot_dropin_check.py --cases morph_16x400 --modes plain --golden tests/golden/mpboot --out gpurun_out/r02g/chk2 2>&1 | cut -c1-600
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15
