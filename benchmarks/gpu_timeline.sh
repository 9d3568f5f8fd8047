#!/bin/bash
# per-tile phase timeline of the onesweep passes inside a frame (diagnostic build on the box only; the shipped .so is untouched)
set -u
mkdir -p gpurun_out
export SPLAT_SCENE_DIR=/tmp/splat_b200_scenes
SPL_BUILD_DEFINES=SPL_SORT_TIMELINE python -m splatapult_b200.build > gpurun_out/build_tl.log 2>&1 || { echo "build failed"; tail -20 gpurun_out/build_tl.log; exit 1; }
for wl in ${1:-1m_deg3}; do
  SPL_SORT_TIMELINE=gpurun_out/tl_$wl timeout 300 python bench.py --workload $wl --headline-only --no-parity --no-cpu-baseline --steps 3 --warmup 3 --graph 0 --pdl 0 > gpurun_out/tl_bench_$wl.json 2> gpurun_out/tl_bench_$wl.err || tail -5 gpurun_out/tl_bench_$wl.err
  python benchmarks/tl_analyze.py gpurun_out/tl_$wl.b*
done
