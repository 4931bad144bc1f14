#!/usr/bin/env python
"""Side measurements of the BASELINE.json configs that are parity-test cases, not bench lines:
config 3 (AND + phrase + text locality + facets) and config 4 (large dictionary, levenshtein 2).
Each is timed resident (execute only) and end to end on one GPU, next to the CPU oracle on a
sample of the same requests.  Prints one JSON line per config.

    python tools/bench_configs.py [--scale 1.0]
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def jmdict_corpus(n_docs, seed=1):
    """A jmdict-shaped corpus (veloci_bins/src/bin/create_test_index.rs:102-275: kanji[] / kana[] with commonness,
    meanings.ger[] with rank, meanings.eng[]) and the 5-way OR request of benches/bench_jmdict.rs:115-235."""
    import numpy as np

    rng = np.random.default_rng(seed)
    syll = ["ka", "ki", "ku", "ke", "ko", "sa", "shi", "su", "se", "so", "ta", "chi", "tsu", "te", "to", "na", "ni", "nu", "ne", "no", "ma", "mi", "mu", "me", "mo", "ra", "ri", "ru", "re", "ro", "ya", "yu", "yo", "wa"]
    words = ["".join(rng.choice(syll, size=int(rng.integers(2, 6)))) for _ in range(40000)]
    eng = ["".join(rng.choice(list("abcdefghijklmnoprstuw"), size=int(rng.integers(3, 9)))) for _ in range(20000)]
    zipf = lambda pool: pool[min(len(pool) - 1, int(rng.zipf(1.3)) - 1)]
    lines = []
    for i in range(n_docs):
        d = {"ent_seq": str(1000000 + i)}
        if rng.random() < 0.6:
            d["commonness"] = int(rng.integers(1, 5000))
        d["kanji"] = [{"text": zipf(words), **({"commonness": int(rng.integers(1, 500))} if rng.random() < 0.5 else {})} for _ in range(int(rng.integers(0, 3)))]
        d["kana"] = [{"text": zipf(words), **({"commonness": int(rng.integers(1, 500))} if rng.random() < 0.5 else {})} for _ in range(int(rng.integers(1, 3)))]
        d["meanings"] = {"ger": [{"text": " ".join(zipf(words) for _ in range(int(rng.integers(1, 4)))), **({"rank": int(rng.integers(1, 10))} if rng.random() < 0.7 else {})} for _ in range(int(rng.integers(0, 3)))],
                         "eng": [" ".join(zipf(eng) for _ in range(int(rng.integers(1, 4)))) for _ in range(int(rng.integers(1, 3)))]}
        lines.append(json.dumps(d))
    boost = {"boost": {"boost_type": "f32"}}
    config = {"kanji[].text": {"fulltext": {"tokenize": True}}, "kana[].text": {"fulltext": {"tokenize": True}}, "meanings.ger[].text": {"fulltext": {"tokenize": True}},
              "meanings.eng[]": {"fulltext": {"tokenize": True}}, "commonness": boost, "kanji[].commonness": boost, "kana[].commonness": boost, "meanings.ger[].rank": boost}

    def request(term, lev):
        def part(path, boosts, starts_with):
            p = {"terms": [term], "path": path, "levenshtein_distance": lev, "options": {"boost": boosts}}
            if starts_with:
                p["starts_with"] = True
            return {"search": p}
        c1 = {"path": "commonness", "boost_fun": "Log10", "param": 1}
        return json.dumps({"search_req": {"or": {"queries": [
            part("kanji[].text", [c1, {"path": "kanji[].commonness", "boost_fun": "Log10", "param": 1}], True),
            part("kana[].text", [c1, {"path": "kana[].commonness", "boost_fun": "Log10", "param": 1}], True),
            part("kana[].text", [c1, {"path": "kana[].commonness", "boost_fun": "Log10", "param": 1}], True),
            part("meanings.ger[].text", [{"path": "commonness", "boost_fun": "Log10", "param": 0}, {"path": "meanings.ger[].rank", "expression": "10 / $SCORE"}], False),
            part("meanings.eng[]", [c1], False)], "options": {"top": 10, "skip": 0}}}})

    return "\n".join(lines), config, lambda n: [request(zipf(words) if i % 3 else zipf(eng), 1) for i in range(n)]


def run(name, corpus, queries, cpu_sample=32, jmdict=None):
    import numpy as np

    import helpers
    import veloci_b200

    d = tempfile.mkdtemp(prefix=f"vb200_{name}_")
    t0 = time.time()
    if jmdict:
        text, config, make = jmdict_corpus(jmdict["docs"])
        helpers.create_index(d, text, config)
        reqs = make(jmdict["requests"])
    else:
        helpers.create_synthetic_index(d, **corpus)
        reqs = helpers.synthetic_requests(**queries, **corpus)
    gen_s = time.time() - t0
    index = veloci_b200.Index(d)
    if jmdict:  # latency of small batches on the same index
        for nb in (1, 16, 256):
            bb = index.prepare(reqs[:nb])
            for _ in range(2):
                bb.execute()
            tt = []
            for _ in range(5):
                t1 = time.perf_counter()
                bb.execute()
                tt.append(time.perf_counter() - t1)
            print(json.dumps({"config": name, "batch": nb, "execute_ms": 1000 * min(tt)}), flush=True)
            bb.close()
    batch = index.prepare(reqs)
    for _ in range(2):
        batch.execute()
    times = []
    for _ in range(3):
        t1 = time.perf_counter()
        batch.execute()
        times.append(time.perf_counter() - t1)
    phase = batch.phase_ms()
    flat = batch.results_flat(10)
    e2e = []
    for _ in range(3):
        t1 = time.perf_counter()
        b = index.prepare(reqs)
        b.execute()
        b.results_flat(10)
        e2e.append(time.perf_counter() - t1)
        b.close()
    oracle = helpers.Oracle(d)
    cores = os.cpu_count() or 1
    r = oracle.search_batch(reqs[:cpu_sample], threads=cores, k=10)
    ok = int((flat["status"] == 0).sum())
    same = int((r["num_hits"] == flat["num_hits"][:cpu_sample]).sum())
    parity = helpers.batch_parity(flat, r, list(range(cpu_sample)))  # ids + scores (1e-5 rel, ties) + num_hits
    line = {
        "config": name, "corpus": corpus, "queries": queries, "requests": len(reqs), "requests_ok": ok,
        "resident_requests_per_s": len(reqs) / min(times), "e2e_requests_per_s": len(reqs) / min(e2e[1:]),
        "phase_ms": dict(zip(["fuzzy_match", "group_score", "slice", "plane_eval", "tile_eval", "final_topk"], phase)),
        "paths": batch.path_stats(),
        "cpu_oracle_requests_per_s": cpu_sample / r["seconds"], "cpu_cores": cores, "cpu_sample": cpu_sample,
        "num_hits_equal_on_sample": f"{same}/{cpu_sample}", "parity": parity, "work": batch.work_stats(), "index_generation_s": gen_s,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--full", action="store_true", help="BASELINE.json sizes: config 3 on 10M docs / 10k requests, config 4 on a 5M-term dictionary / 10k requests")
    ap.add_argument("--configs", default="1,3,4")
    a = ap.parse_args()
    s = a.scale
    which = set(a.configs.split(","))
    if "1" in which:
        run("config1_jmdict_shape_5way_or_lev1", {"docs": int(166_600 * s)}, {"requests": int(2000 * s)}, jmdict={"docs": int(166_600 * s), "requests": int(2000 * s)})
    s3 = 5.0 if a.full else s
    if "3" in which:
        run("config3_and_phrase_locality_facets",
            dict(num_docs=int(2_000_000 * s3), vocab=int(200_000 * s3), seed=42, tokens_per_doc=8, zipf_s=1.07, tags=1000, text_locality=True, phrase=True),
            dict(num_queries=int(2000 * s3), query_kind="and", levenshtein=1, query_seed=44, top=10), cpu_sample=64)
    if "4" in which:
        run("config4_large_dictionary_lev2",
            dict(num_docs=int(5_000_000 * s), vocab=int(5_000_000 * s), seed=42, tokens_per_doc=8, zipf_s=1.07, len_min=4, len_max=16),
            dict(num_queries=10_000 if a.full else int(2000 * s), query_kind="single", levenshtein=2, query_seed=45, edit_prob=0.5, top=10), cpu_sample=64)


if __name__ == "__main__":
    main()
