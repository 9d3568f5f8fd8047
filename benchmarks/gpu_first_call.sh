#!/bin/bash
# First GPU call after host-only work: everything that could not be verified without a device, in one gpurun command.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash benchmarks/gpu_first_call.sh'
# Results land in gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
python -m splatapult_b200.build > gpurun_out/build.log 2>&1 || { echo "build failed"; tail -20 gpurun_out/build.log; exit 1; }
# the new path first (fast feedback), then the whole GPU suite
timeout 600 python -m pytest tests/test_gpu_shared_sort.py -x -q -m gpu > gpurun_out/shared_sort.log 2>&1
echo "shared sort tests: exit $?"; tail -15 gpurun_out/shared_sort.log
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/gpu_tests.log 2>&1
echo "gpu suite: exit $?"; tail -5 gpurun_out/gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_1m_deg0.json 2> gpurun_out/bench_1m_deg0.err
echo "bench: exit $?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_1m_deg0.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "exact_mode", "stereo_pair")}, d["e2e"]["ms_per_step"])
except Exception as e:
    print("no bench line:", e)
PY
