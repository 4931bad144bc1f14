#!/usr/bin/env python
"""Where the time of Index.search_stream goes: per step, the wait for the planner thread, the GPU work and the fetch."""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
import veloci_b200

docs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
corpus = dict(num_docs=docs, vocab=1_000_000, seed=42, tokens_per_doc=8, zipf_s=1.07)
d = f"/tmp/veloci_b200_bench/idx_d{docs}_v1000000_s42"
if not os.path.exists(os.path.join(d, ".complete")):
    os.makedirs(os.path.dirname(d), exist_ok=True)
    helpers.create_synthetic_index(d, **corpus)
    open(os.path.join(d, ".complete"), "w").write("ok")
reqs = helpers.synthetic_requests(num_queries=10_000, query_kind="or3", levenshtein=1, query_seed=43, edit_prob=0.5, top=10, **corpus)
index = veloci_b200.Index(d)
for _ in range(3):
    b = index.prepare(reqs); b.execute(); b.results_flat(10); b.close()
for label in ("serial", "stream", "stream"):
    t0 = time.perf_counter()
    rows = []
    if label == "serial":
        for _ in range(6):
            a = time.perf_counter(); b = index.prepare(reqs); c = time.perf_counter(); b.execute(); e = time.perf_counter(); b.results_flat(10); b.close(); f = time.perf_counter()
            rows.append((c - a, e - c, f - e))
    else:
        with ThreadPoolExecutor(max_workers=1) as ex:
            def timed_prepare():
                a = time.perf_counter(); b = index.prepare(reqs); return b, time.perf_counter() - a
            ahead = ex.submit(timed_prepare)
            for i in range(6):
                a = time.perf_counter(); b, prep = ahead.result(); c = time.perf_counter()
                ahead = ex.submit(timed_prepare) if i < 5 else None
                b.execute(); e = time.perf_counter(); b.results_flat(10); b.close(); f = time.perf_counter()
                rows.append((c - a, e - c, f - e, prep))
    total = time.perf_counter() - t0
    print(label, "total ms/step %.2f" % (1000 * total / 6))
    for r in rows:
        print("   ", " ".join("%.2f" % (1000 * x) for x in r))
