"""Timing experiment: the bulk plane pass over only the first T tiles of the unsharded bench index (VELOCI_TILE_LIMIT; the
results of such a run are incomplete) next to shard 0 of 8, to see whether the first tiles of a run cost more per item."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers, veloci_b200
docs = 10_000_000
corpus = dict(num_docs=docs, vocab=1_000_000, seed=42, tokens_per_doc=8, zipf_s=1.07)
d = f"/tmp/veloci_b200_bench/idx_d{docs}_v1000000_s42"
if not os.path.exists(os.path.join(d, ".complete")):
    os.makedirs(os.path.dirname(d), exist_ok=True); helpers.create_synthetic_index(d, **corpus); open(os.path.join(d, ".complete"), "w").write("ok")
reqs = helpers.synthetic_requests(num_queries=10_000, query_kind="or3", levenshtein=1, query_seed=43, edit_prob=0.5, top=10, **corpus)
index = veloci_b200.Index(d)
batch = index.prepare(reqs)
for limit in (0, 8, 16, 32, 64, 153, 306, 612, 1221):
    if limit: os.environ["VELOCI_TILE_LIMIT"] = str(limit)
    for _ in range(2): batch.execute()
    t = []
    for _ in range(3):
        a = time.perf_counter(); batch.execute(); t.append(time.perf_counter() - a)
    os.environ["VELOCI_DEBUG"] = "1"; batch.execute(); batch.results_flat(10); os.environ.pop("VELOCI_DEBUG")
    print(f"tiles {limit or 'all'}: {1000*min(t):.2f} ms plane_eval {batch.phase_ms()[3]:.2f} evaluated {batch.path_stats()['plane_evaluated']}", flush=True)
