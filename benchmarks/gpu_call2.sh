#!/bin/bash
# usage: gpurun --timeout 1500 -- 'bash benchmarks/gpu_call2.sh [quick|full] [ncu]'
set -u
mode=${1:-quick}
mkdir -p gpurun_out
export SPLAT_SCENE_DIR=/tmp/splat_b200_scenes
python -m splatapult_b200.build > gpurun_out/build.log 2>&1 || { echo "build failed"; tail -20 gpurun_out/build.log; exit 1; }
if [ "$mode" = "full" ]; then
  timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/gpu_tests.log 2>&1
  echo "gpu suite: exit $?"; tail -8 gpurun_out/gpu_tests.log
else
  timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cull.py tests/test_gpu_shared_sort.py tests/test_gpu_async.py tests/test_gpu_multigpu_cpp.py tests/test_gpu_parity_baseline.py -q -m gpu -x > gpurun_out/gpu_tests.log 2>&1
  echo "gpu tests (subset): exit $?"; tail -8 gpurun_out/gpu_tests.log
fi
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench: exit $?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    print("headline", d["config"]["workload"][:10], round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "exact", round(d["exact_mode"]["ms_per_step"],4), "D", d["config"]["tile_instances"])
    print("  stage_ms", {k: round(v,4) for k,v in d["stage_ms"].items()})
    p = d["parity"]; print("  parity pass", p["pass"], "rec", p["records_bit_exact"], "max", p["max_abs"], "mean", p["mean"], "exact", p.get("exact_mode"))
    print("  stereo", d["stereo_pair"]["ms_resort"], d["stereo_pair"]["ms_shared_sort"], "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"],4))
    for k, b in d["workloads"].items():
        p = b["parity"]
        print(k, "ms", round(b["ms_per_step"],4), "e2e", round(b["e2e"]["ms_per_step"],4), "exact", round(b["exact_mode"]["ms_per_step"],4), "D", b["tile_instances"], {kk: round(v,4) for kk,v in b["stage_ms"].items()})
        print("  parity pass", p["pass"], "rec", p["records_bit_exact"], "max", p["max_abs"], "mean", p["mean"], p.get("fp16_max_rel_vs_round_every_blend"))
except Exception as e:
    print("no bench line:", e)
PY
if [ "${2:-}" = "ncu" ]; then
  for wl in 1m_deg3 6m_deg3_fp16; do
    timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches_$wl.csv \
      python bench.py --workload $wl --headline-only --no-parity --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ncu_launch_$wl.log 2>&1
    echo "ncu launch list $wl: exit $?"
    timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_composite|k_project_keys|k_sort_onesweep|k_emit_sorted' -s 40 -c 10 -f -o gpurun_out/prof_$wl \
      python bench.py --workload $wl --headline-only --no-parity --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ncu_full_$wl.log 2>&1
    echo "ncu full $wl: exit $?"; ls -la gpurun_out/prof_$wl.ncu-rep 2>/dev/null
  done
fi
