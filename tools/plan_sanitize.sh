#!/bin/bash
# ThreadSanitizer and AddressSanitizer/UBSan over the multi-threaded plan enumeration (ScanPlanner::add_parallel, spr_host.cpp):
# host code only, no device.  720 self-tests (parallel plan == sequential plan) on trees of 50 / 200 / 600 taxa, 2-8 threads, 1-3 pieces.
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"; REPO="$(dirname "$HERE")"
TMP="$(mktemp -d)"
INC="-I$REPO/mpboot_b200/csrc -I/usr/local/cuda/include"
g++ -std=c++17 -O1 -g -fsanitize=thread $INC "$HERE/plan_sanitizer_harness.cpp" "$REPO/mpboot_b200/csrc/spr_host.cpp" -o "$TMP/tsan" -L/usr/local/cuda/lib64 -lcudart -pthread
g++ -std=c++17 -O1 -g -fsanitize=address,undefined $INC "$HERE/plan_sanitizer_harness.cpp" "$REPO/mpboot_b200/csrc/spr_host.cpp" -o "$TMP/asan" -L/usr/local/cuda/lib64 -lcudart -pthread
echo "tsan: $("$TMP/tsan" 2>&1 | tail -1)"
echo "asan+ubsan: $(ASAN_OPTIONS=detect_leaks=0 "$TMP/asan" 2>&1 | tail -1)"
rm -rf "$TMP"
