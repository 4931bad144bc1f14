#!/usr/bin/env python
"""BASELINE config 5 on one GPU: a 100M-doc synthetic index, of which this process opens shard `rank` of `shards`
(anchor range), and the bench batch of 10k 3-term OR requests evaluated on that shard.  All shards do the same work in
parallel on an 8-GPU box, so the shard's step time is (up to the all-gather and the merge, ~0.3 ms) the step time of the
whole job: requests/s ~= batch / step time.

    python tools/bench_config5_shard.py [--docs 100000000] [--shards 8] [--ranks 0,7]
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
    ap.add_argument("--docs", type=int, default=100_000_000)
    ap.add_argument("--vocab", type=int, default=1_000_000)
    ap.add_argument("--shards", type=int, default=8)
    ap.add_argument("--ranks", default="0")
    ap.add_argument("--queries", type=int, default=10_000)
    a = ap.parse_args()
    import helpers
    import veloci_b200

    corpus = dict(num_docs=a.docs, vocab=a.vocab, seed=42, tokens_per_doc=8, zipf_s=1.07)
    d = f"/tmp/veloci_b200_bench/idx_d{a.docs}_v{a.vocab}_s42"
    t0 = time.time()
    if not os.path.exists(os.path.join(d, ".complete")):
        os.makedirs(os.path.dirname(d), exist_ok=True)
        helpers.create_synthetic_index(d, **corpus)
        open(os.path.join(d, ".complete"), "w").write("ok")
    gen_s = time.time() - t0
    reqs = helpers.synthetic_requests(num_queries=a.queries, query_kind="or3", levenshtein=1, query_seed=43, edit_prob=0.5, top=10, **corpus)
    for rank in [int(r) for r in a.ranks.split(",")]:
        t0 = time.time()
        index = veloci_b200.Index(d, shard_rank=rank, n_shards=a.shards)
        open_s = time.time() - t0
        batch = index.prepare(reqs)
        for _ in range(3):
            batch.execute()
        times = []
        for _ in range(5):
            t1 = time.perf_counter()
            batch.execute()
            times.append(time.perf_counter() - t1)
        e2e = []
        for _ in range(4):
            t1 = time.perf_counter()
            b = index.prepare(reqs)
            b.execute()
            b.results_flat(10)
            e2e.append(time.perf_counter() - t1)
            b.close()
        step = sorted(times)[len(times) // 2]
        print(json.dumps({
            "config": "config5_100M_docs_shard", "docs": a.docs, "vocab": a.vocab, "shards": a.shards, "rank": rank, "requests": len(reqs),
            "index_generation_s": gen_s, "open_s": open_s, "info": index.info(), "step_ms": 1000 * step, "e2e_ms": 1000 * min(e2e[1:]),
            "implied_requests_per_s_all_shards_in_parallel": len(reqs) / step,
            "phase_ms": dict(zip(["fuzzy_match", "group_score", "slice", "plane_eval", "tile_eval", "final_topk"], batch.phase_ms())),
            "paths": batch.path_stats(), "num_hits_local": int(batch.results_flat(10)["num_hits"].sum()),
        }), flush=True)
        batch.close()
        index.close() if hasattr(index, "close") else None


if __name__ == "__main__":
    main()
