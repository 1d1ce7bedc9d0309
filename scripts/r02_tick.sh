#!/bin/bash
# refresh tick on one GPU: device time, blocking full upload vs 1 % delta (C4 and C5)
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 300 python scripts/tick_cost.py > gpurun_out/r02r_tick.txt 2>&1; cat gpurun_out/r02r_tick.txt
