#!/bin/bash
# A/B of tuning knobs on the GPU box (stage timers, CUDA events): usage tools/run_knobs.sh
for b in 4 5; do KT_ICP_BATCH=$b python tools/stage_ab.py 40 512 0 2>&1 | tail -1; done
for b in 4 5; do KT_ICP_BATCH=$b python tools/icp_prof.py 2>&1 | tail -2 | cut -c1-200; done
