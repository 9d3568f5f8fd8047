#!/bin/bash
# quick A/B of bench options: gpurun -- 'bash benchmarks/gpu_ab.sh "<args A>" "<args B>" ...'   (each run: 1m_deg0 and 6m_deg3_fp16, no parity)
set -u
mkdir -p gpurun_out
export SPLAT_SCENE_DIR=/tmp/splat_b200_scenes
python -m splatapult_b200.build > gpurun_out/build.log 2>&1 || { echo "build failed"; tail -20 gpurun_out/build.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cull.py -q -m gpu -x > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests (parity, cull): exit $?"; tail -3 gpurun_out/gpu_tests.log
i=0
for args in "$@"; do
  for wl in 1m_deg0 6m_deg3_fp16; do
    envs=""; opts=""
    for tok in $args; do case "$tok" in [A-Z_]*=*) envs="$envs $tok";; *) opts="$opts $tok";; esac; done   # VAR=value tokens go to the environment
    timeout 300 env $envs python bench.py --workload $wl --headline-only --no-parity --no-cpu-baseline --steps 30 --warmup 5 $opts > gpurun_out/ab_${i}_$wl.json 2> gpurun_out/ab_${i}_$wl.err || tail -3 gpurun_out/ab_${i}_$wl.err
    python - "$wl" "$args" gpurun_out/ab_${i}_$wl.json <<'PY'
import json, sys
wl, args, path = sys.argv[1:4]
try:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    print(f"[{args}] {wl}: ms {d['ms_per_step']:.4f} exact {d['exact_mode']['ms_per_step']:.4f}", {k: round(v, 4) for k, v in d["stage_ms"].items()})
except Exception as e:
    print(f"[{args}] {wl}: no line ({e})")
PY
  done
  i=$((i+1))
done
