#!/bin/bash
# end-of-round call: full GPU suite, bench line, ncu launch lists + full captures (traffic.json), sanitizers
set -u
bash benchmarks/gpu_call2.sh full ncu
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python benchmarks/sanitize_frame.py > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "sanitizer $tool: exit $?"; tail -3 gpurun_out/sanitizer_$tool.txt
done
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "reference arm: exit $?"; cut -c1-300 gpurun_out/bench_reference.json
