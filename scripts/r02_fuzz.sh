#!/bin/bash
# randomised shapes vs the oracle natively, then the sanitizer tools over smoke() and a short fuzz
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 900 python scripts/fuzz_gpu.py 300 11 > gpurun_out/r02p_fuzz.txt 2>&1; echo "fuzz rc=$?" >> gpurun_out/r02p_fuzz.txt
tail -3 gpurun_out/r02p_fuzz.txt
bash scripts/r02_sanitizer.sh > /dev/null 2>&1
for tool in memcheck racecheck; do
  echo "===== compute-sanitizer --tool $tool python scripts/fuzz_gpu.py 8 3 =====" >> gpurun_out/r02_compute_sanitizer.txt
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/fuzz_gpu.py 8 3 2>&1 | cut -c1-400 | tail -8 >> gpurun_out/r02_compute_sanitizer.txt
done
grep -E "=====|SUMMARY|mismatches|smoke ok" gpurun_out/r02_compute_sanitizer.txt | cut -c1-160
