"""Breakdown of vgpu_batch_prepare for the bench batch (VELOCI_DEBUG prints parse + plan / merge / upload) next to the
time the Python binding spends around it."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers, veloci_b200
docs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
corpus = dict(num_docs=docs, vocab=1_000_000, seed=42, tokens_per_doc=8, zipf_s=1.07)
d = f"/tmp/veloci_b200_bench/idx_d{docs}_v1000000_s42"
if not os.path.exists(os.path.join(d, ".complete")):
    os.makedirs(os.path.dirname(d), exist_ok=True); helpers.create_synthetic_index(d, **corpus); open(os.path.join(d, ".complete"), "w").write("ok")
reqs = helpers.synthetic_requests(num_queries=10_000, query_kind="or3", levenshtein=1, query_seed=43, edit_prob=0.5, top=10, **corpus)
index = veloci_b200.Index(d)
for _ in range(3):
    b = index.prepare(reqs); b.execute(); b.close()
os.environ["VELOCI_DEBUG"] = "1"
for _ in range(4):
    a = time.perf_counter(); blob = "\n".join(reqs).encode("utf-8"); j = time.perf_counter(); b = index.prepare(reqs); c = time.perf_counter()
    print(f"python join+encode alone {1000*(j-a):.2f} ms; Index.prepare {1000*(c-j):.2f} ms", flush=True)
    b.close()
