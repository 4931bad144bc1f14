"""One-off stress: the plane path against the posting path (`Index(..., planes=False)`) on a mid-size corpus, bit for bit,
over several request shapes (with / without boost, k from 1 to 64, skips, single terms, levenshtein 0-2)."""
import json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers, veloci_b200

docs = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
params = dict(num_docs=docs, vocab=docs // 10, seed=99)
d = tempfile.mkdtemp(prefix="vb200_stress_")
helpers.create_synthetic_index(d, **params)
reqs = []
for kind, lev, seed, n in (("or3", 1, 1, 1500), ("or3", 0, 2, 500), ("single", 1, 3, 500), ("single", 2, 4, 300), ("or3", 2, 5, 300)):
    reqs += helpers.synthetic_requests(num_queries=n, query_kind=kind, levenshtein=lev, query_seed=seed, **params)
rng = np.random.default_rng(0)
out = []
for i, r in enumerate(reqs):
    j = json.loads(r)
    if i % 3 == 0:
        j.pop("boost", None)
    if i % 5 == 0:
        j["boost"] = [{"path": "commonness", "boost_fun": ["Log10", "Log2", "Multiply"][i % 3], "param": float(i % 4)}]
    j["top"] = int(rng.choice([1, 3, 10, 10, 10, 25, 64]))
    if i % 7 == 0:
        j["skip"] = int(rng.integers(0, 64 - min(63, j["top"]) + 1))
    out.append(json.dumps(j))
a = veloci_b200.Index(d); ba = a.prepare(out); ba.execute(); ra = ba.results_flat(64); sa = ba.path_stats()
b = veloci_b200.Index(d, planes=False); bb = b.prepare(out); bb.execute(); rb = bb.results_flat(64); sb = bb.path_stats()
print("paths", sa, sb)
assert sa["plane_items"] > 0 and sb["plane_items"] == 0
ok = (ra["status"] == 0).all() and (rb["status"] == 0).all()
same_hits = (ra["num_hits"] == rb["num_hits"]).all()
same_ids = (ra["ids"] == rb["ids"]).all()
same_scores = (ra["scores"].view(np.uint32) == rb["scores"].view(np.uint32)).all()
print(json.dumps({"requests": len(out), "ok": bool(ok), "same_num_hits": bool(same_hits), "same_ids": bool(same_ids), "same_score_bits": bool(same_scores), "hits": int(ra["num_hits"].sum())}))
if not (ok and same_hits and same_ids and same_scores):
    bad = np.nonzero((ra["num_hits"] != rb["num_hits"]) | (ra["ids"] != rb["ids"]).any(axis=1) | (ra["scores"].view(np.uint32) != rb["scores"].view(np.uint32)).any(axis=1))[0]
    print("first differing requests:", bad[:5], [out[i] for i in bad[:2]])
    sys.exit(1)
