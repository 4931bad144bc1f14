#!/usr/bin/env python
"""bench.py -- batched queries/sec of the query-time hit pipeline (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            our CUDA path
    python bench.py --impl reference --gpus N --steps K ...   the CPU search (oracle port) on the host cores

A "step" is one pass of the hot path over one batch of Q synthetic requests
(BASELINE.json configs[1]: 10M-doc Zipfian corpus, 10k batched 3-term OR queries,
levenshtein_distance 1, f32 Log10 boost column).  With N > 1 the same index is
sharded by anchor range over the N GPUs (one process per GPU, torchrun); every
rank evaluates the whole batch against its shard, the shard-local top-k rows are
all-gathered over NCCL and merged on every rank (strong scaling: fixed corpus).

`value`  : requests/s with the prepared batch resident in HBM (vgpu_batch_execute).
`e2e`    : requests/s through vgpu_search_batch-equivalent calls: request JSON on the
           host -> plan -> H2D -> kernels -> D2H of the result rows, every step.
`roofline`: the tile-evaluation kernel (posting expansion + merge + boost + top-k):
           algorithmic bytes of BASELINE.md section 5 / its CUDA-event duration.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "batched_queries_per_sec"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--docs", type=int, default=int(os.environ.get("VELOCI_BENCH_DOCS", 10_000_000)))
    ap.add_argument("--vocab", type=int, default=int(os.environ.get("VELOCI_BENCH_VOCAB", 1_000_000)))
    ap.add_argument("--queries", type=int, default=int(os.environ.get("VELOCI_BENCH_QUERIES", 10_000)))
    ap.add_argument("--cache", default=os.environ.get("VELOCI_BENCH_CACHE", "/tmp/veloci_b200_bench"))
    ap.add_argument("--cpu-sample", type=int, default=0, help="requests per CPU-baseline sample (0: 2 per core, at least 32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def corpus_params(args):
    return dict(num_docs=args.docs, vocab=args.vocab, seed=42, tokens_per_doc=8, zipf_s=1.07)


def ensure_index(args, helpers):
    """The synthetic index directory, generated once per box (rank 0) and reused."""
    d = os.path.join(args.cache, f"idx_d{args.docs}_v{args.vocab}_s42")
    marker = os.path.join(d, ".complete")
    if not os.path.exists(marker):
        os.makedirs(args.cache, exist_ok=True)
        helpers.create_synthetic_index(d, **corpus_params(args))
        open(marker, "w").write("ok")
    return d


def make_requests(args, helpers):
    return helpers.synthetic_requests(num_queries=args.queries, query_kind="or3", levenshtein=1, query_seed=43, edit_prob=0.5, top=10, **corpus_params(args))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device = device
        self.samples = []
        self.source = "nvidia-smi"
        self.stop = threading.Event()
        self.thread = threading.Thread(target=self.run, daemon=True)

    def run(self):
        if self.run_nvml():
            return
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.device)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def run_nvml(self):
        """The same readings through NVML (what nvidia-smi calls), every 10 ms: a 5-step timed region lasts ~0.1 s, too short
        for more than one nvidia-smi process.  False when NVML is not usable (the nvidia-smi loop runs instead)."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.device)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            reasons_of = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        except Exception:
            return False
        self.source = "nvml"
        bits = ((0x8, 2), (0x40, 3), (0x20, 4), (0x4, 5))  # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
        while not self.stop.is_set():
            try:
                row = [str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx), "", "", "", ""]
                mask = reasons_of(h)
                for bit, col in bits:
                    row[col] = "Active" if mask & bit else "Not Active"
                self.samples.append(row)
            except Exception:
                pass
            self.stop.wait(0.01)
        return True

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        sm = [int(s[0]) for s in self.samples if s and s[0].isdigit()]
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": int(statistics.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm), "source": self.source}


def run_reference(args):
    """CPU arm: the reference's search on the host cores (the C++ oracle port: the Rust crate cannot be built in this image)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import helpers
    from veloci_b200 import build

    build.build_index_lib()
    build.build_oracle()
    d = ensure_index(args, helpers)
    reqs = make_requests(args, helpers)
    cores = os.cpu_count() or 1
    sample = args.cpu_sample or max(32, 2 * cores)
    oracle = helpers.Oracle(d)
    times = []
    for step in range(args.warmup + args.steps):
        lo = (step * sample) % max(1, len(reqs) - sample)
        r = oracle.search_batch(reqs[lo:lo + sample], threads=cores, k=10)
        if step >= args.warmup:
            times.append(r["seconds"])
    total = sum(times)
    value = sample * len(times) / total if total > 0 else 0.0
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * total / max(1, len(times)), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"{sample} consecutive requests of the batch per step, one request at a time per thread"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {
        "workload": f"synthetic {args.docs}-doc Zipfian corpus (s=1.07, 8 tokens/doc, vocab {args.vocab}), {args.queries} batched 3-term OR requests, "
                    "levenshtein_distance 1, f32 Log10 boost column, top 10 (BASELINE.json configs[1])",
        "docs": args.docs, "vocab": args.vocab, "batch": args.queries, "terms_per_request": 3, "levenshtein_distance": 1, "boost": "Log10(commonness+1)", "top": 10,
        "cache": "index postings (6 B x ~9 x docs) exceed the 126 MB L2; every step re-streams them",
    }


class DevArray:
    """Zero-copy view of a device pointer for torch (int64 words)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def run_ours(args):
    import numpy as np
    import torch

    import helpers
    import veloci_b200
    from veloci_b200 import build

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        build.build_all()
    if dist:
        dist.barrier()
    if veloci_b200.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device (the CUDA path has no CPU fallback)")
    torch.cuda.set_device(local_rank)

    d = None
    if rank == 0:
        d = ensure_index(args, helpers)
    if dist:
        dist.barrier()
    d = d or ensure_index(args, helpers)
    reqs = make_requests(args, helpers)
    t0 = time.time()
    index = veloci_b200.Index(d, device=local_rank, shard_rank=rank, n_shards=world)
    open_s = time.time() - t0
    info = index.info()

    def sync_all():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    batch = index.prepare(reqs)
    SIGN = -(1 << 63)

    def run_batch(b):
        """One step on this rank's shard.  With several shards the per-request thresholds are shared after the first
        tiles: all-reduce MAX in unsigned order (the sign bit is flipped around the signed reduction)."""
        if not dist:
            b.execute()
            return
        b.execute_begin()
        ptr, cnt = b.thresholds()
        tau = torch.as_tensor(DevArray(ptr, cnt), device="cuda")
        tau.bitwise_xor_(SIGN)
        dist.all_reduce(tau, op=dist.ReduceOp.MAX)
        tau.bitwise_xor_(SIGN)
        torch.cuda.synchronize()
        b.execute_finish()

    def step_resident():
        run_batch(batch)
        if dist:
            keys_ptr, hits_ptr, stride = batch.local_topk()
            n = len(reqs)
            local_keys = torch.as_tensor(DevArray(keys_ptr, n * stride), device="cuda")
            local_hits = torch.as_tensor(DevArray(hits_ptr, n), device="cuda")
            g_keys = torch.empty(world * n * stride, dtype=torch.int64, device="cuda")
            g_hits = torch.empty(world * n, dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(g_keys, local_keys)
            dist.all_gather_into_tensor(g_hits, local_hits)
            torch.cuda.synchronize()
            batch.merge_gathered(g_keys.data_ptr(), g_hits.data_ptr(), world)

    for _ in range(args.warmup):
        step_resident()
    sync_all()
    launches0 = veloci_b200.launch_count()
    step_ms, phase = [], []
    with ClockSampler(local_rank) as clocks:
        t_begin = time.perf_counter()
        for _ in range(args.steps):
            t1 = time.perf_counter()
            step_resident()
            step_ms.append(1000.0 * (time.perf_counter() - t1))
            phase.append(batch.phase_ms())
        sync_all()
        elapsed = time.perf_counter() - t_begin
    launches = veloci_b200.launch_count() - launches0
    traffic = batch.traffic_model()
    flat = batch.results_flat(10)
    n_ok = int((flat["status"] == 0).sum())

    # end to end: host JSON in, host rows out, every step
    e2e_ms = []
    e2e_parts = {"prepare": 0.0, "execute": 0.0, "results": 0.0}
    io = {"h2d": 0, "d2h": 0}
    for i in range(args.warmup + args.steps):
        sync_all()
        t1 = time.perf_counter()
        b = index.prepare(reqs)
        t2 = time.perf_counter()
        run_batch(b)
        t3 = time.perf_counter()
        if dist:
            keys_ptr, hits_ptr, stride = b.local_topk()
            n = len(reqs)
            g_keys = torch.empty(world * n * stride, dtype=torch.int64, device="cuda")
            g_hits = torch.empty(world * n, dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(g_keys, torch.as_tensor(DevArray(keys_ptr, n * stride), device="cuda"))
            dist.all_gather_into_tensor(g_hits, torch.as_tensor(DevArray(hits_ptr, n), device="cuda"))
            torch.cuda.synchronize()
            b.merge_gathered(g_keys.data_ptr(), g_hits.data_ptr(), world)
        out = b.results_flat(10)
        t4 = time.perf_counter()
        dt = 1000.0 * (t4 - t1)
        io = b.io_bytes()
        b.close()
        if i >= args.warmup:
            e2e_ms.append(dt)
            e2e_parts["prepare"] += 1000.0 * (t2 - t1) / args.steps
            e2e_parts["execute"] += 1000.0 * (t3 - t2) / args.steps
            e2e_parts["results"] += 1000.0 * (t4 - t3) / args.steps
    sync_all()

    # The same end-to-end steps through Index.search_stream: the planner thread prepares step k+1 (parse, plan, H2D)
    # while step k is on the GPU and its rows are read back.  Every step still does all of its own work, H2D and D2H
    # inside the timed region.
    def run_e2e(b):
        run_batch(b)
        if dist:
            keys_ptr, hits_ptr, stride = b.local_topk()
            n = len(reqs)
            g_keys = torch.empty(world * n * stride, dtype=torch.int64, device="cuda")
            g_hits = torch.empty(world * n, dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(g_keys, torch.as_tensor(DevArray(keys_ptr, n * stride), device="cuda"))
            dist.all_gather_into_tensor(g_hits, torch.as_tensor(DevArray(hits_ptr, n), device="cuda"))
            torch.cuda.synchronize()
            b.merge_gathered(g_keys.data_ptr(), g_hits.data_ptr(), world)

    for _ in index.search_stream((reqs for _ in range(max(2, args.warmup))), k=10, run=run_e2e):
        pass
    stream_steps = max(8, 2 * args.steps)
    sync_all()
    tp = time.perf_counter()
    stream_hits = 0
    for out in index.search_stream((reqs for _ in range(stream_steps)), k=10, run=run_e2e):
        stream_hits = int(out["num_hits"].sum())
    torch.cuda.synchronize()
    stream_s = time.perf_counter() - tp
    sync_all()

    # max over ranks
    t = torch.tensor([elapsed, sum(e2e_ms) / 1000.0, stream_s], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed, e2e_s, stream_s = float(t[0]), float(t[1]), float(t[2])
    if rank != 0:
        if dist:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    names = ["fuzzy_match", "group_score", "slice", "plane_eval", "tile_eval", "final_topk"]
    phase_ms = {k: statistics.mean(p[i] for p in phase) for i, k in enumerate(names)}
    paths = batch.path_stats()
    # roofline of the dominant kernel: the tile evaluation (plane path + general path), which does the posting expansion,
    # merge, boost and top-k of every request.  Algorithmic bytes = BASELINE.md section 5's fused lower bound
    # sum_t (8 + 6 df_t) + 4 |union| + 8 k per request, summed over the batch.
    tile_ms = phase_ms["plane_eval"] + phase_ms["tile_eval"]
    alg_bytes = traffic["posting_bytes"] + traffic["boost_bytes"] + 8 * 10 * len(reqs)
    achieved = alg_bytes / (tile_ms / 1000.0) / 1e9 if tile_ms > 0 else 0.0
    dominant = "plane_eval_kernel" if phase_ms["plane_eval"] >= phase_ms["tile_eval"] else "tile_eval_kernel"
    ncu_traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_plane_eval_traffic.json")
    if os.path.exists(tpath):
        try:
            t = json.load(open(tpath))
            if t.get("docs") == args.docs and t.get("queries") == args.queries:
                ncu_traffic = t.get("dram_bytes_per_step")
        except Exception:
            pass
    value = len(reqs) * args.steps / elapsed
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * elapsed / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "p50_batch_latency_ms": statistics.median(step_ms),
        "phase_ms": phase_ms,
        "device_ms_per_step": sum(phase_ms.values()),
        "paths": paths,
        "requests_ok": n_ok,
        "index": {"open_s": open_s, "device_bytes": info["device_bytes"], "anchor_range": [info["anchor_lo"], info["anchor_hi"]]},
        "roofline": {
            "kernel": dominant, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "postings_per_launch": traffic["postings"], "union_hits_per_launch": traffic["union_hits"],
            "launch_ms": tile_ms,
            "note": "launch = the plane_eval stages + tile_eval of one step (CUDA events on the library's stream). The algorithmic bytes are those of the posting-list "
                    "formulation (6 B per posting of every matched term + 4 B boost per hit); the plane path answers the same requests from presence bitmaps, boost "
                    "level bitmaps and bound pruning, so it moves far fewer bytes than that (see traffic) and a fraction above 1 is not a measurement error: the "
                    "kernel is bound by shared-memory bit operations, not by HBM (profiles/).",
        },
        "e2e": {"value": len(reqs) * stream_steps / stream_s, "unit": UNIT, "h2d_bytes_per_step": io["h2d"], "d2h_bytes_per_step": io["d2h"],
                "ms_per_step": 1000.0 * stream_s / stream_steps, "steps": stream_steps,
                "mode": "Index.search_stream: host JSON in, host rows out, every step; the planner thread prepares step k+1 (parse, plan, H2D) while "
                        "step k is on the GPU (execute, exchange, D2H)",
                "one_step_at_a_time": {"value": len(reqs) * len(e2e_ms) / e2e_s if e2e_s > 0 else 0.0, "ms_per_step": 1000.0 * e2e_s / max(1, len(e2e_ms)),
                                       "host_ms": e2e_parts},
                "num_hits_last_step": stream_hits},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = args.cpu_sample or max(32, 2 * cores)
        oracle = helpers.Oracle(d)
        r = oracle.search_batch(reqs[:sample], threads=cores, k=10)
        line["cpu_baseline"] = {"value": sample / r["seconds"], "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"first {sample} requests of the batch, one request at a time per thread, {cores} threads"}
        # the sample doubles as a parity spot check of the timed batch
        same = int((r["num_hits"] == flat["num_hits"][:sample]).sum())
        line["cpu_baseline"]["num_hits_equal"] = f"{same}/{sample}"
    print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
