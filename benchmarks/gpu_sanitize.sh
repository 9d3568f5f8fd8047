#!/bin/bash
# compute-sanitizer over benchmarks/sanitize_frame.py (memcheck, racecheck, synccheck) + the launch-mode tests
set -u
timeout 600 python -m pytest tests/test_gpu_launch_modes.py tests/test_gpu_async.py tests/test_gpu_shared_sort.py -q -m gpu 2>&1 | tail -3
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python benchmarks/sanitize_frame.py > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "sanitizer $tool: exit $?"; tail -3 gpurun_out/sanitizer_$tool.txt
done
