#!/bin/bash
# usage: gpurun --gpus N --timeout 1500 -- 'bash benchmarks/gpu_multi.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out
export SPLAT_SCENE_DIR=/tmp/splat_b200_scenes
python -m splatapult_b200.build > gpurun_out/build.log 2>&1 || { echo "build failed"; tail -20 gpurun_out/build.log; exit 1; }
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_multigpu_cpp.py -q -m gpu -x > gpurun_out/multigpu_cpp_n$N.log 2>&1
echo "C++ multi-GPU test: exit $?"; tail -4 gpurun_out/multigpu_cpp_n$N.log
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N: exit $?"; grep -c "NCCL INFO" gpurun_out/bench_n$N.err gpurun_out/bench_n$N.json | head; tail -3 gpurun_out/bench_n$N.err | cut -c1-300
python - $N <<'PY'
import json, sys
N = sys.argv[1]
try:
    line = [l for l in open(f"gpurun_out/bench_n{N}.json").read().splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    print("headline", d["config"]["workload"][:10], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), d["e2e"]["d2h_gbs_per_rank"], "bcast ms", d["broadcast_ms"], "numa", d["config"]["numa_node"], "nccl", d["config"]["nccl"])
    p = d["parity"]; print("  parity pass", p["pass"], "ranks", p["ranks_checked"], "max", p["max_abs"], p.get("per_rank_max_abs"))
    print("  launch_modes", d.get("launch_modes"))
    for k, b in d["workloads"].items():
        p = b["parity"]
        print(k, "value", round(b["value"], 1), "ms", round(b["ms_per_step"], 4), "e2e", round(b["e2e"]["value"], 1), b["e2e"]["d2h_gbs_per_rank"], "bcast", b["broadcast_ms"], "views", b.get("views"))
        print("  parity pass", p["pass"], "ranks", p["ranks_checked"], "max", p["max_abs"], p.get("fp16_max_rel_vs_round_every_blend"))
except Exception as e:
    print("no bench line:", e)
PY
