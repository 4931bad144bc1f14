#!/usr/bin/env python
"""Eight anchor-range shards of the bench index emulated on ONE GPU, one after the other, with the threshold exchange
between execute_begin and execute_finish done as an element-wise max over the eight threshold arrays: the per-shard step
time (begin + finish) is what each GPU of an 8-GPU job spends.  Used to choose the seed-pass settings for sharded runs.

    python tools/shard_exchange_probe.py [n_shards]
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
import veloci_b200
from bench import DevArray

n_shards = int(sys.argv[1]) if len(sys.argv) > 1 else 8
docs = 10_000_000
corpus = dict(num_docs=docs, vocab=1_000_000, seed=42, tokens_per_doc=8, zipf_s=1.07)
d = f"/tmp/veloci_b200_bench/idx_d{docs}_v1000000_s42"
if not os.path.exists(os.path.join(d, ".complete")):
    os.makedirs(os.path.dirname(d), exist_ok=True)
    helpers.create_synthetic_index(d, **corpus)
    open(os.path.join(d, ".complete"), "w").write("ok")
reqs = helpers.synthetic_requests(num_queries=10_000, query_kind="or3", levenshtein=1, query_seed=43, edit_prob=0.5, top=10, **corpus)
SIGN = torch.tensor(-(2 ** 63), dtype=torch.int64, device="cuda")
shards = [veloci_b200.Index(d, shard_rank=r, n_shards=n_shards) for r in range(n_shards)]
batches = [ix.prepare(reqs) for ix in shards]


def step():
    tb, tf = [], []
    for b in batches:
        a = time.perf_counter(); b.execute_begin(); tb.append(time.perf_counter() - a)
    taus = []
    for b in batches:
        ptr, cnt = b.thresholds()
        taus.append(torch.as_tensor(DevArray(ptr, cnt), device="cuda"))
    best = torch.stack([t ^ SIGN for t in taus]).max(dim=0).values ^ SIGN
    for t in taus:
        t.copy_(best)
    torch.cuda.synchronize()
    for b in batches:
        a = time.perf_counter(); b.execute_finish(); tf.append(time.perf_counter() - a)
    return tb, tf


for tiles, level in ((8, 5), (4, 5), (2, 5), (1, 5), (2, 4), (4, 4), (1, 3), (2, 6), (4, 6)):
    os.environ["VELOCI_SEED_TILES"], os.environ["VELOCI_SEED_LEVEL"] = str(tiles), str(level)
    step(); step()
    best = None
    for _ in range(3):
        tb, tf = step()
        tot = [x + y for x, y in zip(tb, tf)]
        if best is None or max(tot) < best[0]:
            best = (max(tot), sum(tot) / len(tot), max(tb), max(tf))
    ev = sum(b.path_stats()["plane_evaluated"] for b in batches)
    print(f"{n_shards} shards, seed tiles {tiles} level {level}: slowest shard {1000 * best[0]:.2f} ms (mean {1000 * best[1]:.2f}; begin {1000 * best[2]:.2f} finish {1000 * best[3]:.2f}), evaluated {ev} over all shards", flush=True)
