#!/bin/bash
# r02d GPU session (1 GPU): drop-in with the IQ-TREE-tree kernel calls answered by the device (N3), cross-checked
# against the reference's kernel, then timed on C1's largest fixture and on C2.
mkdir -p gpurun_out/r02d
cd "$(dirname "$0")/.."
MPBOOT_GPU_CHECK_K9=1 python tools/mpboot_dropin_check.py --cases c1_17x1998,aa_20x600 --modes plain,bb --golden tests/golden/mpboot --out gpurun_out/r02d/chk 2>&1 | cut -c1-900
MPBOOT_GPU_CHECK_K9=1 python tools/mpboot_dropin_check.py --cases c1_100x5000,morph_16x400 --modes plain --golden tests/golden/mpboot --out gpurun_out/r02d/chk 2>&1 | cut -c1-900
timeout 1500 python -m pytest tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -5
python tools/mpboot_dropin_check.py --cases c1_100x5000 --modes bb --golden tests/golden/mpboot --out gpurun_out/r02d/x1 2>&1 | cut -c1-1200
( time MPGPU_PROFILE=2 timeout 1200 python tools/mpboot_dropin_check.py --cases c2_200x100000 --modes bb --skip-stock --out gpurun_out/r02d/x1 ) 2>&1 | cut -c1-1500
grep "mpgpu profile" gpurun_out/r02d/x1/c2_200x100000.bb.gpu.stdout | head -8
grep -n "Iteration 100 \|Iteration 200 \|Optimizing boot\|CPU Time\|Wall-clock" gpurun_out/r02d/x1/c2_200x100000.bb.gpu.stdout | tail -8
cmp gpurun_out/r02d/x1/c2_200x100000.bb.gpu.splits.nex gpurun_out/r02c/x1/c2_200x100000.bb.gpu.splits.nex 2>/dev/null && echo "C2 splits identical to the r02c run (host kernel)"
