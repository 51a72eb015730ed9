#!/bin/bash
cd "$(dirname "$0")/.."
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_plan --csv python bench.py --steps 3 --warmup 3 --no-bb --no-cost --no-search --no-cpu-baseline --no-bb1000 --no-c4 2>/dev/null | grep k_plan | head -4 | awk -F'","' '{print $5, $(NF-1), $NF}' | cut -c1-200
