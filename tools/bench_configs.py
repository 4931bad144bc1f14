#!/usr/bin/env python
"""Side measurements of the BASELINE.json configs that are parity-test cases, not bench lines:
config 3 (AND + phrase + text locality + facets) and config 4 (large dictionary, levenshtein 2).
Each is timed resident (execute only) and end to end on one GPU, next to the CPU oracle on a
sample of the same requests.  Prints one JSON line per config.

    python tools/bench_configs.py [--scale 1.0]
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run(name, corpus, queries, cpu_sample=32):
    import numpy as np

    import helpers
    import veloci_b200

    d = tempfile.mkdtemp(prefix=f"vb200_{name}_")
    t0 = time.time()
    helpers.create_synthetic_index(d, **corpus)
    gen_s = time.time() - t0
    reqs = helpers.synthetic_requests(**queries, **corpus)
    index = veloci_b200.Index(d)
    batch = index.prepare(reqs)
    for _ in range(2):
        batch.execute()
    times = []
    for _ in range(3):
        t1 = time.perf_counter()
        batch.execute()
        times.append(time.perf_counter() - t1)
    phase = batch.phase_ms()
    flat = batch.results_flat(10)
    e2e = []
    for _ in range(3):
        t1 = time.perf_counter()
        b = index.prepare(reqs)
        b.execute()
        b.results_flat(10)
        e2e.append(time.perf_counter() - t1)
        b.close()
    oracle = helpers.Oracle(d)
    cores = os.cpu_count() or 1
    r = oracle.search_batch(reqs[:cpu_sample], threads=cores, k=10)
    ok = int((flat["status"] == 0).sum())
    same = int((r["num_hits"] == flat["num_hits"][:cpu_sample]).sum())
    line = {
        "config": name, "corpus": corpus, "queries": queries, "requests": len(reqs), "requests_ok": ok,
        "resident_requests_per_s": len(reqs) / min(times), "e2e_requests_per_s": len(reqs) / min(e2e[1:]),
        "phase_ms": dict(zip(["fuzzy_match", "group_score", "slice", "plane_eval", "tile_eval", "final_topk"], phase)),
        "paths": batch.path_stats(),
        "cpu_oracle_requests_per_s": cpu_sample / r["seconds"], "cpu_cores": cores, "cpu_sample": cpu_sample,
        "num_hits_equal_on_sample": f"{same}/{cpu_sample}", "index_generation_s": gen_s,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    a = ap.parse_args()
    s = a.scale
    run("config3_and_phrase_locality_facets",
        dict(num_docs=int(2_000_000 * s), vocab=int(200_000 * s), seed=42, tokens_per_doc=8, zipf_s=1.07, tags=1000, text_locality=True, phrase=True),
        dict(num_queries=int(2000 * s), query_kind="and", levenshtein=1, query_seed=44, top=10))
    run("config4_large_dictionary_lev2",
        dict(num_docs=int(5_000_000 * s), vocab=int(5_000_000 * s), seed=42, tokens_per_doc=8, zipf_s=1.07, len_min=4, len_max=16),
        dict(num_queries=int(2000 * s), query_kind="single", levenshtein=2, query_seed=45, edit_prob=0.5, top=10))


if __name__ == "__main__":
    main()
