#!/bin/bash
# tests + A/B + ncu launch lists (per-kernel times):  gpurun -- 'bash benchmarks/gpu_call3.sh'
set -u
mkdir -p gpurun_out
export SPLAT_SCENE_DIR=/tmp/splat_b200_scenes
python -m splatapult_b200.build > gpurun_out/build.log 2>&1 || { echo "build failed"; tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_parity_baseline.py > gpurun_out/gpu_tests.log 2>&1
echo "gpu suite: exit $?"; tail -5 gpurun_out/gpu_tests.log
timeout 300 python -m pytest tests/test_gpu_parity_baseline.py -q -m gpu -x > gpurun_out/baseline.log 2>&1
echo "baseline parity: exit $?"; tail -3 gpurun_out/baseline.log
bash benchmarks/gpu_ab.sh "" 2>&1 | grep "ms "
for wl in 1m_deg0 6m_deg3_fp16; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 22 --csv --log-file gpurun_out/launches_$wl.csv \
    python bench.py --workload $wl --headline-only --no-parity --no-cpu-baseline --steps 3 --warmup 3 --graph 0 > gpurun_out/ncu_launch_$wl.log 2>&1
  python - gpurun_out/launches_$wl.csv <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
out = []
for r in rows:
    name = r[4].split("(")[0].replace("void ", "").replace("splat::", "")
    out.append(f"{name[:34]}={float(r[-1])/1000:.1f}")
print(sys.argv[1], " ".join(out[:22]))
PY
done
