#!/usr/bin/env python
"""Step time of the bench batch for different seed-pass settings (VELOCI_SEED_TILES / VELOCI_SEED_LEVEL), unsharded and on
shard 0 of 8."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
import veloci_b200

docs = 10_000_000
corpus = dict(num_docs=docs, vocab=1_000_000, seed=42, tokens_per_doc=8, zipf_s=1.07)
d = f"/tmp/veloci_b200_bench/idx_d{docs}_v1000000_s42"
if not os.path.exists(os.path.join(d, ".complete")):
    os.makedirs(os.path.dirname(d), exist_ok=True)
    helpers.create_synthetic_index(d, **corpus)
    open(os.path.join(d, ".complete"), "w").write("ok")
reqs = helpers.synthetic_requests(num_queries=10_000, query_kind="or3", levenshtein=1, query_seed=43, edit_prob=0.5, top=10, **corpus)
for shards in (1, 8):
    index = veloci_b200.Index(d, shard_rank=0, n_shards=shards)
    batch = index.prepare(reqs)
    base = None
    for tiles, level in ((8, 5), (0, 5), (16, 5), (32, 5), (16, 6), (32, 7), (64, 8), (128, 9), (64, 6), (32, 4), (8, 3)):
        os.environ["VELOCI_SEED_TILES"], os.environ["VELOCI_SEED_LEVEL"] = str(tiles), str(level)
        for _ in range(2):
            batch.execute()
        t = []
        for _ in range(4):
            a = time.perf_counter(); batch.execute(); t.append(time.perf_counter() - a)
        hits = int(batch.results_flat(10)["num_hits"].sum())
        ids = batch.results_flat(10)["ids"]
        if base is None:
            base = (hits, ids.copy())
        same = hits == base[0] and (ids == base[1]).all()
        print(f"shards {shards} seed tiles {tiles:4d} level {level:2d}: {1000 * min(t):7.2f} ms  evaluated {batch.path_stats()['plane_evaluated']:9d}  phases {[round(x, 2) for x in batch.phase_ms()]} same={same}", flush=True)
    batch.close()
    index.close()
