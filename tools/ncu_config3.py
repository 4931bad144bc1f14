#!/usr/bin/env python
"""BASELINE config 3 (AND + phrase boosts + text locality + facets) at a size one GPU call can capture: the general path's
kernels (list producers, sparse buckets, tile_eval_kernel) under ncu.

    python tools/ncu_config3.py [--scale 0.5]            # plain: must exit 0, prints the step and phase times
    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
        --clock-control none --csv --log-file gpurun_out/config3_launches.csv python tools/ncu_config3.py

Only the last execute lies between cudaProfilerStart / cudaProfilerStop (index open and warm-up are not captured).
"""
import argparse
import ctypes
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.5, help="1.0 = 2M docs / 2000 requests (a fifth of BASELINE config 3)")
    ap.add_argument("--executes", type=int, default=3)
    a = ap.parse_args()
    import helpers
    import veloci_b200

    s = a.scale
    corpus = dict(num_docs=int(2_000_000 * s), vocab=int(200_000 * s), seed=42, tokens_per_doc=8, zipf_s=1.07, tags=1000, text_locality=True, phrase=True)
    d = os.path.join(tempfile.gettempdir(), "vb200_config3_%d" % corpus["num_docs"])  # kept for the run under ncu that follows the plain one
    t0 = time.perf_counter()
    if not os.path.exists(os.path.join(d, "metaData.json")):
        helpers.create_synthetic_index(d, **corpus)
    gen_s = time.perf_counter() - t0
    reqs = helpers.synthetic_requests(num_queries=int(2000 * s), query_kind="and", levenshtein=1, query_seed=44, top=10, **corpus)
    index = veloci_b200.Index(d)
    batch = index.prepare(reqs)
    cudart = ctypes.CDLL("libcudart.so.12")
    ms = []
    for i in range(a.executes):
        if i == a.executes - 1:
            cudart.cudaProfilerStart()
        t = time.perf_counter()
        batch.execute()
        ms.append(1000 * (time.perf_counter() - t))
    cudart.cudaProfilerStop()
    flat = batch.results_flat(10)
    print(json.dumps({"corpus": corpus, "requests": len(reqs), "requests_ok": int((flat["status"] == 0).sum()), "index_generation_s": gen_s, "step_ms": ms,
                      "phase_ms": batch.phase_ms(), "paths": batch.path_stats(), "work": batch.work_stats(), "kernels": batch.profile_execute()}), flush=True)


if __name__ == "__main__":
    main()
