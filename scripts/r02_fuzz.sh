#!/bin/bash
# randomised shapes vs the oracle natively and under the sanitizer tools
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 900 python scripts/fuzz_gpu.py 300 11 > gpurun_out/r02p_fuzz.txt 2>&1; echo "fuzz rc=$?" >> gpurun_out/r02p_fuzz.txt
tail -3 gpurun_out/r02p_fuzz.txt
for tool in memcheck racecheck; do
  echo "===== compute-sanitizer --tool $tool python scripts/fuzz_gpu.py 8 3 =====" >> gpurun_out/r02_compute_sanitizer.txt
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/fuzz_gpu.py 8 3 2>&1 | cut -c1-400 | tail -8 >> gpurun_out/r02_compute_sanitizer.txt
done
tail -12 gpurun_out/r02_compute_sanitizer.txt
