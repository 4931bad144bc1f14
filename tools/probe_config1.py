"""Phase times of the jmdict-shaped side config (tools/bench_configs.py) for batches of 16 and 2000 requests."""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_configs, helpers, veloci_b200
text, config, make = bench_configs.jmdict_corpus(166600)
d = tempfile.mkdtemp(prefix="vb200_c1_"); helpers.create_index(d, text, config)
index = veloci_b200.Index(d); reqs = make(2000)
for nb in (16, 2000):
    b = index.prepare(reqs[:nb])
    for _ in range(2): b.execute()
    t = []
    for _ in range(3):
        t1 = time.perf_counter(); b.execute(); t.append(time.perf_counter() - t1)
    print(nb, "ms", 1000 * min(t), [round(x, 2) for x in b.phase_ms()], b.path_stats())
