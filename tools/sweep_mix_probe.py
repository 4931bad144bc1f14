"""Timing experiment: the bench step with every swept item forced onto the general sweep (exact per-anchor part count)
instead of the converged-threshold sweep (VELOCI_FORCE_GENERAL_SWEEP): what the two sweeps cost per item."""
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
base = None
for force in (0, 1, 0):
    if force: os.environ["VELOCI_FORCE_GENERAL_SWEEP"] = "1"
    else: os.environ.pop("VELOCI_FORCE_GENERAL_SWEEP", None)
    for _ in range(2): batch.execute()
    t = []
    for _ in range(3):
        a = time.perf_counter(); batch.execute(); t.append(time.perf_counter() - a)
    os.environ["VELOCI_DEBUG"] = "1"; batch.execute(); out = batch.results_flat(10); os.environ.pop("VELOCI_DEBUG")
    if base is None: base = out
    same = (out["ids"] == base["ids"]).all() and (out["num_hits"] == base["num_hits"]).all()
    print(f"force_general {force}: {1000*min(t):.2f} ms plane_eval {batch.phase_ms()[3]:.2f} same={same}", flush=True)
