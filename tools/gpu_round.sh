#!/bin/bash
# Everything one round needs from ONE single-GPU gpurun call (each call costs >= 3 GPU-minutes of box acquisition, so batch):
#   parity tests, smoke(), the bench lines of configs[1] / configs[2] for both arms, the launch list and the ncu captures.
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'tools/gpu_round.sh r2a'          (results under gpurun_out/<tag>_*)
# Add "quick" as second argument to skip the ncu part.
set -u
TAG=${1:-r}; MODE=${2:-full}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
python bench.py --steps 200 --warmup 20 2>&1 | tail -1 > gpurun_out/${TAG}_bench_cfg2.json
python bench.py --steps 150 --warmup 20 --no-cpu-baseline --odometry 2 2>&1 | tail -1 > gpurun_out/${TAG}_bench_cfg3.json
python bench.py --impl reference --steps 100 --warmup 10 2>&1 | tail -1 > gpurun_out/${TAG}_bench_reference_cfg2.json
python bench.py --impl reference --steps 60 --warmup 6 --odometry 2 2>&1 | tail -1 > gpurun_out/${TAG}_bench_reference_cfg3.json
python - <<PY
import json
for f in ("bench_cfg2", "bench_cfg3", "bench_reference_cfg2", "bench_reference_cfg3"):
    try:
        d = json.load(open("gpurun_out/${TAG}_" + f + ".json"))
        st = {k: round(v["ms"], 4) for k, v in d.get("stages", {}).items()}
        print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), st, d.get("roofline", {}).get("frac"), d.get("clocks"))
    except Exception as e:
        print(f, "ERR", e)
PY
if [ "$MODE" != "quick" ]; then
    tools/ncu_capture.sh $TAG 0
    tools/ncu_capture.sh $TAG 2 rgbd_frame_kernel
fi
