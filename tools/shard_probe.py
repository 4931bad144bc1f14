"""Times execute_begin / execute_finish of the bench workload on one GPU for an unsharded index and for shard 0 of 8."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, helpers, veloci_b200
class A: pass
args = A(); args.docs=10_000_000; args.vocab=1_000_000; args.queries=10_000; args.cache="/tmp/veloci_b200_bench"
d = bench.ensure_index(args, helpers)
reqs = bench.make_requests(args, helpers)
for shards in (1, 8):
    ix = veloci_b200.Index(d, shard_rank=0, n_shards=shards)
    b = ix.prepare(reqs)
    for _ in range(3): b.execute()
    t=[]; tb=[]; tf=[]
    for _ in range(5):
        t0=time.perf_counter(); b.execute_begin(); t1=time.perf_counter(); b.execute_finish(); t2=time.perf_counter()
        tb.append(t1-t0); tf.append(t2-t1)
    print(shards, "begin ms", 1000*min(tb), "finish ms", 1000*min(tf), b.phase_ms(), b.path_stats(), ix.info())
