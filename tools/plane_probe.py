#!/usr/bin/env python
"""Step time of the bench batch for different plane thresholds (VELOCI_MID_DIV: a term gets a bits-only plane when its
df >= span / MID_DIV) and group sizes (VELOCI_GROUP_TILES).  Needs a library built with VELOCI_PROBES=1."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for div in sys.argv[1].split(","):
    for grp in sys.argv[2].split(","):
        env = dict(os.environ, VELOCI_MID_DIV=div, VELOCI_GROUP_TILES=grp)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_step.py")] + sys.argv[3:], env=env, capture_output=True, text=True)
        try:
            r = json.loads(out.stdout.strip().split("\n")[-1])
            print(json.dumps({"mid_div": div, "group_tiles": grp, "step_ms": min(r["step_ms"]), "phase_ms": [round(x, 2) for x in r["phase_ms"]],
                              "planes": r["work"]["planes"], "sparse_entries": r["work"]["sparse_entries"], "plane_items": r["work"]["plane_items"],
                              "general_items": r["work"]["general_items"], "evaluated": r["work"]["plane_evaluated"], "num_hits": r["num_hits"]}), flush=True)
        except Exception as e:
            print("failed", div, grp, out.stderr[-500:], flush=True)
