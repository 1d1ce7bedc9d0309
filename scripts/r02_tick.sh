#!/bin/bash
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 300 python scripts/exp_stream_tail.py > gpurun_out/r02s_stream_tail.txt 2>&1; cat gpurun_out/r02s_stream_tail.txt
