#!/usr/bin/env python
"""Step time of the bench batch for different seed sweep lengths (VELOCI_SEED_WORDS) and work-counter batches
(VELOCI_ITEM_BATCH), unsharded and on one shard of eight.  Needs a library built with VELOCI_PROBES=1."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for shards in ("1", "8"):
    for words in sys.argv[1].split(","):
        for batch in sys.argv[2].split(","):
            env = dict(os.environ, VELOCI_SEED_WORDS=words, VELOCI_ITEM_BATCH=batch)
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_step.py"), "--shards", shards, "--rank", "3" if shards != "1" else "0"] + sys.argv[3:], env=env, capture_output=True, text=True)
            try:
                r = json.loads(out.stdout.strip().split("\n")[-1])
                print(json.dumps({"shards": shards, "seed_words": words, "item_batch": batch, "step_ms": round(min(r["step_ms"]), 3), "phase_ms": [round(x, 2) for x in r["phase_ms"]],
                                  "evaluated": r["work"]["plane_evaluated"], "num_hits": r["num_hits"]}), flush=True)
            except Exception:
                print("failed", shards, words, batch, out.stderr[-400:], flush=True)
