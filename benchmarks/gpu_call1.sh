#!/bin/bash
# Round 2, first GPU call: everything written without a device, in one gpurun command.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash benchmarks/gpu_call1.sh'
set -u
mkdir -p gpurun_out
export SPLAT_SCENE_DIR=/tmp/splat_b200_scenes
python -m splatapult_b200.build > gpurun_out/build.log 2>&1 || { echo "build failed"; tail -20 gpurun_out/build.log; exit 1; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tail -2
# new paths first (fast feedback)
timeout 300 python -m pytest tests/test_gpu_shared_sort.py tests/test_gpu_async.py tests/test_gpu_multigpu_cpp.py -x -q -m gpu > gpurun_out/new_tests.log 2>&1
echo "new tests: exit $?"; tail -12 gpurun_out/new_tests.log
timeout 600 python -m pytest tests/test_gpu_parity_baseline.py -x -q -m gpu > gpurun_out/baseline_parity.log 2>&1
echo "baseline parity: exit $?"; tail -12 gpurun_out/baseline_parity.log
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_parity_baseline.py > gpurun_out/gpu_tests.log 2>&1
echo "gpu suite: exit $?"; tail -6 gpurun_out/gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench: exit $?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    print("headline", d["config"]["workload"][:40], d["value"], d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "exact", d["exact_mode"]["ms_per_step"])
    print("stage_ms", d["stage_ms"]); print("parity", d["parity"]); print("stereo", d["stereo_pair"]); print("cpu", d.get("cpu_baseline"))
    for k, b in d["workloads"].items():
        print(k, b["ms_per_step"], "e2e", b["e2e"]["ms_per_step"], "exact", b["exact_mode"]["ms_per_step"], b["stage_ms"], "parity", b["parity"])
except Exception as e:
    print("no bench line:", e)
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "reference arm: exit $?"; cut -c1-600 gpurun_out/bench_reference.json
# sanitizers on test-size frames: logs are copied to profiles/r2_sanitizer_*.txt
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python benchmarks/sanitize_frame.py > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "sanitizer $tool: exit $?"; tail -4 gpurun_out/sanitizer_$tool.txt
done
