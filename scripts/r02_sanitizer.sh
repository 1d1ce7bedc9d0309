#!/bin/bash
# compute-sanitizer over smoke(): every kernel family (host-buffer pick, direct scan, persistent
# TMA-ring queue kernel with tables in shared memory, model-request kernel, class build)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r02_compute_sanitizer.txt
: > $out
for tool in memcheck racecheck synccheck initcheck; do
  echo "===== compute-sanitizer --tool $tool python __graft_entry__.py smoke =====" >> $out
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python __graft_entry__.py smoke 2>&1 | grep -v "^make" | tail -25 >> $out
done
cat $out
