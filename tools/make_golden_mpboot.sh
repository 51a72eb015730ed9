#!/bin/bash
# Golden outputs of the UNMODIFIED reference program for tests/test_gpu_dropin.py: builds integration/_bin/mpboot-avx
# from /root/reference (integration/build.sh stock) if needed, runs it with -seed 1 (plain and -bb 1000) on the seeded
# synthetic alignments of tools/mpboot_dropin_check.py and stores .treefile / .contree / .splits.nex under
# tests/golden/mpboot/.  The reference aborts on MORPH data under -bb, so that case is plain only.
set -euo pipefail
REPO="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
[ -x "$REPO/integration/_bin/mpboot-avx" ] || "$REPO/integration/build.sh" stock
TMP="$(mktemp -d)"
python "$REPO/tools/mpboot_dropin_check.py" --skip-gpu --out "$TMP" --cases c1_12x300,c1_17x1998,aa_20x600,c1_100x5000 --modes plain,bb
python "$REPO/tools/mpboot_dropin_check.py" --skip-gpu --out "$TMP" --cases morph_16x400 --modes plain
python "$REPO/tools/mpboot_dropin_check.py" --skip-gpu --out "$TMP" --cases mulhits_17x1998,topboot_17x1998,mulhits_aa_20x600 --modes bb
python "$REPO/tools/mpboot_dropin_check.py" --skip-gpu --out "$TMP" --cases distinct_20x800,distinct_30x1500,cutoffbt_30x1500,cutoffbt_mulhits_20x800,cutoffbt_distinct_20x800,miniter1_20x800,autovec_30x1500,firstrell_30x1500,firstrell_distinct_30x1500 --modes bb
python "$REPO/tools/mpboot_dropin_check.py" --skip-gpu --out "$TMP" --cases cost_17x1998,costasym_17x1998,costu32_17x1998 --modes plain,bb
mkdir -p "$REPO/tests/golden/mpboot"
for f in "$TMP"/*.stock.treefile "$TMP"/*.stock.contree "$TMP"/*.stock.splits.nex; do
    [ -e "$f" ] || continue
    b="$(basename "$f")"
    case "$b" in morph_16x400.bb.*) continue;; esac
    cp "$f" "$REPO/tests/golden/mpboot/${b/.stock/}"
done
rm -rf "$TMP"
