#!/bin/bash
# One gpurun call of round 2: args = tag, then a list of steps (tail, tests, bench, prof, ncu)
set -u
TAG=${1:-r2}; shift
mkdir -p gpurun_out
for step in "$@"; do
  case $step in
    tail)  timeout 300 tools/bin/tail_bench > gpurun_out/${TAG}_tail_bench.txt 2>&1; tail -40 gpurun_out/${TAG}_tail_bench.txt ;;
    tests) timeout 2400 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/${TAG}_pytest_full.log 2>&1; grep -E "^E |Error|passed|failed" gpurun_out/${TAG}_pytest_full.log | head -40 | tee gpurun_out/${TAG}_pytest.log ;;
    tests_all) timeout 3000 python -m pytest tests -m gpu -q --tb=short > gpurun_out/${TAG}_pytest_full.log 2>&1; grep -E "^E  +(Assert|assert)|Error|passed|failed|FAILED|^cfg odometry|colour mismatches" gpurun_out/${TAG}_pytest_full.log | cut -c1-400 | head -60 | tee gpurun_out/${TAG}_pytest.log ;;
    replayab) python tools/replay_ab.py 1 12 2>&1 | tail -30 | cut -c1-300 | tee gpurun_out/${TAG}_replayab.txt ;;
    mg2) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py --vol 256 --frames 10 --voxel-shift 2 2>&1 | grep -E "MGPU_CHECK|mismatch|Error|error|slice" | cut -c1-400 | tee gpurun_out/${TAG}_mg2.log
         timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/mgpu_check.py --vol 256 --frames 8 --voxel-shift 2 --odometry 2 2>&1 | grep -E "MGPU_CHECK|mismatch|Error|error|slice" | cut -c1-400 | tee -a gpurun_out/${TAG}_mg2.log ;;
    mgbench) NG=${NG:-2}; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${TAG}_bench_${NG}gpu.json
         python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_${NG}gpu.json"))
print("streams value", round(d["value"], 1), "zslab", json.dumps(d.get("zslab"))[:900])
PY
         ;;
    newtests) timeout 2400 python -m pytest tests/test_gpu_baseline_configs.py tests/test_shim_builds.py -m gpu -x -q -s 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest_new.log ;;
    smoke) python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log ;;
    bench) python bench.py --steps 200 --warmup 20 2>&1 | tail -1 > gpurun_out/${TAG}_bench_cfg2.json
           python bench.py --steps 150 --warmup 20 --no-cpu-baseline --odometry 2 2>&1 | tail -1 > gpurun_out/${TAG}_bench_cfg3.json
           python - <<PY
import json
for f in ("bench_cfg2", "bench_cfg3"):
    try:
        d = json.load(open("gpurun_out/${TAG}_" + f + ".json"))
        st = {k: round(v["ms"], 4) for k, v in d.get("stages", {}).items()}
        print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), st, d.get("roofline", {}).get("frac"), d.get("roofline", {}).get("avg_launch_ms"), d.get("clocks"))
    except Exception as e:
        print(f, "ERR", e)
PY
           ;;
    benchref) python bench.py --impl reference --steps 100 --warmup 10 2>&1 | tail -1 > gpurun_out/${TAG}_bench_reference_cfg2.json; cat gpurun_out/${TAG}_bench_reference_cfg2.json | cut -c1-300 ;;
    prof)  python tools/icp_prof.py 2>&1 | tail -30 | tee gpurun_out/${TAG}_icp_prof.txt ;;
    stages) for v in 512 1024; do python tools/stage_ab.py 24 $v 0 2>&1 | tail -1; done | tee gpurun_out/${TAG}_stages.txt
            python tools/stage_ab.py 24 512 2 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_stages.txt ;;
    *) echo "unknown step $step" ;;
  esac
done
