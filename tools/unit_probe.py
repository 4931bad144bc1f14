"""Step time of the bench batch on shard 0 of 8 and unsharded for different unit sizes of the bulk plane pass (VELOCI_UNIT_ITEMS)."""
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
for shards in (8, 1):
    index = veloci_b200.Index(d, shard_rank=0, n_shards=shards)
    batch = index.prepare(reqs)
    for unit in (0, 128, 256, 512, 768, 1024, 1536, 2048, 3344, 4096):
        if unit: os.environ["VELOCI_UNIT_ITEMS"] = str(unit)
        for _ in range(2): batch.execute()
        t = []
        for _ in range(4):
            a = time.perf_counter(); batch.execute(); t.append(time.perf_counter() - a)
        print(f"shards {shards} unit {unit}: {1000*min(t):.2f} ms phases {[round(x,2) for x in batch.phase_ms()]}", flush=True)
    os.environ.pop("VELOCI_UNIT_ITEMS", None)
    batch.close(); index.close()
