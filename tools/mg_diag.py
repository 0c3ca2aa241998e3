#!/usr/bin/env python
"""Print the shared-volume record of a `bench.py --shared-only` line (stdin): frames/s, stage timers, kernel-only timers."""
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])["zslab"]
print(round(d["value"], 1), {k: round(v, 4) for k, v in d["stages_ms_rank0"].items()},
      {k: round(v, 4) for k, v in d.get("kernels_ms_rank0", {}).items() if isinstance(v, float)})
