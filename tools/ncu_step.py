#!/usr/bin/env python
"""The bench batch, executed a few times on the (cached) bench index: the command ncu captures.

    python tools/ncu_step.py [--docs 10000000] [--queries 10000] [--executes 3] [--shards 1 --rank 0]

Run it plain first (it must exit 0 and prints the step time), then under
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python tools/ncu_step.py
    ncu --set full --clock-control none --import-source on -k regex:plane_eval_kernel -s 3 -c 1 -o gpurun_out/plane python tools/ncu_step.py
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=10_000_000)
    ap.add_argument("--vocab", type=int, default=1_000_000)
    ap.add_argument("--queries", type=int, default=10_000)
    ap.add_argument("--executes", type=int, default=3)
    ap.add_argument("--shards", type=int, default=1)
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--kind", default="or3")
    ap.add_argument("--lev", type=int, default=1)
    a = ap.parse_args()
    import bench
    import helpers
    import veloci_b200

    d = bench.ensure_index("/tmp/veloci_b200_bench", a.docs, a.vocab, helpers)
    reqs = helpers.synthetic_requests(num_queries=a.queries, query_kind=a.kind, levenshtein=a.lev, query_seed=43, edit_prob=0.5, top=10, **bench.corpus_params(a.docs, a.vocab))
    index = veloci_b200.Index(d, shard_rank=a.rank, n_shards=a.shards)
    batch = index.prepare(reqs)
    ms = []
    for _ in range(a.executes):
        t = time.perf_counter()
        batch.execute()
        ms.append(1000 * (time.perf_counter() - t))
    print(json.dumps({"docs": a.docs, "queries": len(reqs), "shards": a.shards, "rank": a.rank, "step_ms": ms, "phase_ms": batch.phase_ms(), "work": batch.work_stats(),
                      "num_hits": int(batch.results_flat(10)["num_hits"].sum())}), flush=True)


if __name__ == "__main__":
    main()
