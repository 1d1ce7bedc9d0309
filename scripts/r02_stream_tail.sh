#!/bin/bash
# C5 latency tail, 3 x 10 s, with the runtime's slowest-device-call / thread-CPU counters
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 300 python scripts/exp_stream_tail.py > gpurun_out/r02s_stream_tail.txt 2>&1; cat gpurun_out/r02s_stream_tail.txt
