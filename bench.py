#!/usr/bin/env python
"""bench.py -- batched queries/sec of the query-time hit pipeline (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            our CUDA path
    python bench.py --impl reference --gpus N --steps K ...   the CPU search (oracle port) on the host cores

A "step" is one pass of the hot path over one batch of Q synthetic requests.  The headline workload (`value`,
`config.workload`) is BASELINE.json configs[1]: 10M-doc Zipfian corpus, 10k batched 3-term OR requests,
levenshtein_distance 1, f32 Log10 boost column.  The same line carries a `config5` record: the same batch on the
100M-doc corpus of configs[4] (the north-star corpus; its device layout is ~27 GB, so it also runs at N = 1).

With N > 1 (one process per GPU, torchrun) the index is sharded by anchor range over the N GPUs; every rank opens its
shard, joins the library's communicator (vgpu_comm_init: NCCL over NVLink) and then calls vgpu_batch_execute, which is
one collective C call per step: seed pass -> all-reduce(max) of the thresholds -> bulk pass -> all-gather of the local
top-k rows -> merge, on one stream.  Local rank 0 plans each batch and publishes the plan through shared memory
(vgpu_batch_prepare_shared); the other ranks import it.  Strong scaling: the corpus is fixed, each GPU holds 1/N.

`value`   : requests/s with the prepared batch resident in HBM (vgpu_batch_execute), max over ranks of the timed region.
`e2e`     : requests/s through Index.search_stream: request JSON on the host -> parse + plan -> H2D -> kernels
            (-> exchange) -> D2H of the result rows, every step; planning of step k+1 overlaps the GPU work of step k.
`parity`  : ids, scores and hit counts of a sample of the timed batch against the CPU oracle, on the full-size index.
`roofline`: the dominant kernel (plane_eval_kernel) against the bound that actually limits it (instruction issue), its
            measured DRAM traffic, and -- as a separate key -- how it compares with a perfect HBM-streaming implementation
            of the posting-list model.  `kernels` lists every kernel of one step with its time and, for the
            HBM-streaming ones, achieved GB/s of algorithmic bytes.
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
N_SMS = 148  # B200


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
    ap.add_argument("--parity", type=int, default=int(os.environ.get("VELOCI_BENCH_PARITY", 256)), help="requests of the timed batch checked against the CPU oracle (ids, scores, hit counts)")
    ap.add_argument("--config5-docs", type=int, default=int(os.environ.get("VELOCI_BENCH_CONFIG5_DOCS", 100_000_000)), help="docs of the config-5 record (0: skip it)")
    ap.add_argument("--config5-parity", type=int, default=int(os.environ.get("VELOCI_BENCH_CONFIG5_PARITY", 64)))
    ap.add_argument("--cpu-sample", type=int, default=0, help="requests per step of the reference arm (0: 2 per core, at least 32)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle (no cpu_baseline, no parity)")
    return ap.parse_args()


def corpus_params(docs, vocab):
    return dict(num_docs=docs, vocab=vocab, seed=42, tokens_per_doc=8, zipf_s=1.07)


def ensure_index(cache, docs, vocab, helpers):
    """The synthetic index directory, generated once per box and reused."""
    d = os.path.join(cache, f"idx_d{docs}_v{vocab}_s42")
    marker = os.path.join(d, ".complete")
    if not os.path.exists(marker):
        os.makedirs(cache, exist_ok=True)
        helpers.create_synthetic_index(d, **corpus_params(docs, vocab))
        open(marker, "w").write("ok")
    return d


def make_requests(docs, vocab, queries, helpers):
    return helpers.synthetic_requests(num_queries=queries, query_kind="or3", levenshtein=1, query_seed=43, edit_prob=0.5, top=10, **corpus_params(docs, vocab))


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


def workload_config(docs, vocab, queries, tag="BASELINE.json configs[1]"):
    return {
        "workload": f"synthetic {docs}-doc Zipfian corpus (s=1.07, 8 tokens/doc, vocab {vocab}), {queries} batched 3-term OR requests, "
                    f"levenshtein_distance 1, f32 Log10 boost column, top 10 ({tag})",
        "docs": docs, "vocab": vocab, "batch": queries, "terms_per_request": 3, "levenshtein_distance": 1, "boost": "Log10(commonness+1)", "top": 10,
        "cache": "index structures of a shard (planes, level bitmaps, postings: GBs) exceed the 126 MB L2; every step re-streams them",
    }


def run_reference(args):
    """CPU arm: the reference's search on the host cores (the C++ oracle port: the Rust crate cannot be built in this image)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import helpers
    from veloci_b200 import build

    build.build_index_lib()
    build.build_oracle()
    d = ensure_index(args.cache, args.docs, args.vocab, helpers)
    reqs = make_requests(args.docs, args.vocab, args.queries, helpers)
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
        "config": workload_config(args.docs, args.vocab, args.queries),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"{sample} consecutive requests of the batch per step, one request at a time per thread",
                         "note": "C++ restatement of the reference (oracle/), not the Rust build: a small factor slower than the original is likely (full dictionary DP walk "
                                 "instead of FST x DFA, stable sorts)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


class Job:
    """Process-group plumbing of one bench process (rank r of N on one box)."""

    def __init__(self):
        import torch

        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def broadcast_bytes(self, payload, n):
        """`payload` (bytes of length n) from rank 0 to every rank."""
        t = self.torch.zeros(n, dtype=self.torch.uint8, device="cuda")
        if self.rank == 0:
            t.copy_(self.torch.frombuffer(bytearray(payload), dtype=self.torch.uint8))
        if self.dist:
            self.dist.broadcast(t, src=0)
        return bytes(t.cpu().numpy().tobytes())

    def close(self):
        if self.dist:
            self.dist.barrier()
            self.dist.destroy_process_group()


def run_workload(job, args, helpers, veloci_b200, docs, vocab, queries, steps, warmup, parity_n, want_cpu_rate, profile_kernels):
    """Opens (this rank's shard of) the index of `docs` documents and measures the batch on it.  -> record (rank 0) or None."""
    torch = job.torch
    world, rank = job.world, job.rank
    d = ensure_index(args.cache, docs, vocab, helpers) if rank == 0 else None
    job.barrier()
    d = d or ensure_index(args.cache, docs, vocab, helpers)
    reqs = make_requests(docs, vocab, queries, helpers)
    reqs_wire = [r.encode("utf-8") for r in reqs]  # the requests as a server holds them: one JSON text (bytes) per request
    t0 = time.time()
    index = veloci_b200.Index(d, device=job.local_rank, shard_rank=rank, n_shards=world)
    open_s = time.time() - t0
    info = index.info()
    channel = None
    if world > 1:
        uid = job.broadcast_bytes(veloci_b200.comm_unique_id() if rank == 0 else None, 128)
        index.comm_init(uid)
        token = job.broadcast_bytes(os.urandom(8) if rank == 0 else None, 8).hex()
        channel = veloci_b200.PlanChannel(f"/veloci_b200_plan_{token}", rank, world, capacity=64 << 20)

    # ---- resident: the prepared batch stays in HBM, every step is one vgpu_batch_execute (a collective when sharded)
    batch = index.prepare(reqs, channel=channel)
    for _ in range(warmup):
        batch.execute()
    job.barrier()
    launches0 = veloci_b200.launch_count()
    step_ms, phase = [], []
    with ClockSampler(job.local_rank) as clocks:
        t_begin = time.perf_counter()
        for _ in range(steps):
            t1 = time.perf_counter()
            batch.execute()
            step_ms.append(1000.0 * (time.perf_counter() - t1))
            phase.append(batch.phase_ms())
        job.barrier()
        elapsed = time.perf_counter() - t_begin
    launches = veloci_b200.launch_count() - launches0
    traffic = batch.traffic_model()
    work = batch.work_stats()
    flat = batch.results_flat(10)
    n_ok = int((flat["status"] == 0).sum())
    kernels = batch.profile_execute() if profile_kernels else None
    job.barrier()

    # ---- end to end, one step at a time: prepare (parse + plan + H2D) -> execute -> results (D2H), nothing overlapped
    e2e_ms, e2e_parts, io = [], {"prepare": 0.0, "execute": 0.0, "results": 0.0}, {"h2d": 0, "d2h": 0}
    for i in range(warmup + steps):
        job.barrier()
        t1 = time.perf_counter()
        b = index.prepare(reqs_wire, channel=channel)
        t2 = time.perf_counter()
        b.execute()
        t3 = time.perf_counter()
        b.results_flat(10)
        t4 = time.perf_counter()
        io = b.io_bytes()
        b.close()
        if i >= warmup:
            e2e_ms.append(1000.0 * (t4 - t1))
            e2e_parts["prepare"] += 1000.0 * (t2 - t1) / steps
            e2e_parts["execute"] += 1000.0 * (t3 - t2) / steps
            e2e_parts["results"] += 1000.0 * (t4 - t3) / steps
    job.barrier()

    # ---- end to end through Index.search_stream: the planner thread prepares step k+1 (on local rank 0: parse, plan, publish;
    # elsewhere: import; everywhere: H2D) while step k is on the GPU and its rows are read back.  Every step does all of its
    # own work, H2D and D2H inside the timed region.
    for _ in index.search_stream((reqs_wire for _ in range(max(2, warmup))), k=10, channel=channel):
        pass
    stream_steps = max(8, 2 * steps)
    job.barrier()
    tp = time.perf_counter()
    stream_hits = 0
    for out in index.search_stream((reqs_wire for _ in range(stream_steps)), k=10, channel=channel):
        stream_hits = int(out["num_hits"].sum())
    torch.cuda.synchronize()
    stream_s = time.perf_counter() - tp
    job.barrier()

    elapsed, e2e_s, stream_s = job.max_over_ranks([elapsed, sum(e2e_ms) / 1000.0, stream_s])
    rec = None
    if rank == 0:
        names = ["fuzzy_match", "group_score", "slice", "plane_eval", "tile_eval", "final_topk"]
        phase_ms = {k: statistics.mean(p[i] for p in phase) for i, k in enumerate(names)}
        rec = {
            "value": len(reqs) * steps / elapsed, "unit": UNIT, "ms_per_step": 1000.0 * elapsed / steps, "steps": steps, "warmup": warmup,
            "p50_batch_latency_ms": statistics.median(step_ms), "phase_ms": phase_ms, "device_ms_per_step": sum(phase_ms.values()),
            "work": work, "requests_ok": n_ok,
            "index": {"open_s": open_s, "device_bytes": info["device_bytes"], "anchor_range": [info["anchor_lo"], info["anchor_hi"]], "docs": docs},
            "e2e": {"value": len(reqs) * stream_steps / stream_s, "unit": UNIT, "h2d_bytes_per_step": io["h2d"], "d2h_bytes_per_step": io["d2h"],
                    "ms_per_step": 1000.0 * stream_s / stream_steps, "steps": stream_steps,
                    "mode": "Index.search_stream: host JSON in, host rows out, every step; the planner thread prepares step k+1 (parse + plan on local rank 0, "
                            "published to the other ranks through shared memory; H2D on every rank) while step k is on the GPUs (execute incl. exchange, D2H)",
                    "one_step_at_a_time": {"value": len(reqs) * len(e2e_ms) / e2e_s if e2e_s > 0 else 0.0, "ms_per_step": 1000.0 * e2e_s / max(1, len(e2e_ms)), "host_ms": e2e_parts},
                    "num_hits_last_step": stream_hits},
            "gpu_launches": int(launches), "clocks": clocks.summary(), "traffic_model": traffic, "kernels_ms": kernels,
        }
        if not args.no_cpu_baseline and parity_n > 0:
            # the CPU oracle over a sample of the timed batch, on the same full-size index: parity (ids, scores, hit counts
            # of the merged result) and, from the same run, the CPU rate
            cores = os.cpu_count() or 1
            rows = list(range(0, len(reqs), max(1, len(reqs) // parity_n)))[:parity_n]
            oracle = helpers.Oracle(d)
            r = oracle.search_batch([reqs[q] for q in rows], threads=cores, k=10)
            oracle.close()
            rec["parity"] = helpers.batch_parity(flat, r, rows)
            rec["parity"]["against"] = "CPU oracle (oracle/veloci_oracle.cpp) on the same index; ids + scores (1e-5 rel, ties may reorder) + num_hits of the final merged rows"
            if want_cpu_rate:
                rec["cpu_baseline"] = {"value": len(rows) / r["seconds"], "unit": UNIT, "cores": cores, "kind": "port",
                                       "sample": f"{len(rows)} requests spread evenly over the batch, one request at a time per thread, {cores} threads",
                                       "note": "C++ restatement of the reference, not the Rust build: a small factor slower than the original is likely"}
    batch.close()
    if channel:
        job.barrier()
        channel.close()
    if world > 1:
        index.comm_destroy()
    index.close()
    job.barrier()
    return rec


def roofline_of(rec, docs, queries, world):
    """The dominant kernel against its real bound.  plane_eval_kernel answers from bitmaps staged in shared memory: it is
    bound by instruction issue, not by HBM.  Instructions per launch come from the committed ncu capture of this workload
    (profiles/plane_eval_ncu.json; per item when the capture is of another size), the duration is measured live."""
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        pk = json.load(open(peaks_path))
        hbm_peak, clk_mhz, peak_src = float(pk["hbm_gbs"]), float(pk.get("sm_max_mhz", 1965.0)), "MEASURED_PEAKS.json (hbm_gbs; issue roof = 148 SMs x 4 schedulers x sm_max_mhz)"
    else:
        hbm_peak, clk_mhz, peak_src = 6650.0, 1965.0, "fallback (B200_PROFILING.md)"
    if rec["clocks"].get("sm_mhz"):
        clk_mhz = float(rec["clocks"]["sm_mhz"])
    launch_ms = rec["phase_ms"]["plane_eval"]
    items = rec["work"]["plane_item_evals"]
    cap = None
    cpath = os.path.join(ROOT, "profiles", "plane_eval_ncu.json")
    if os.path.exists(cpath):
        try:
            caps = json.load(open(cpath))["captures"]
            caps = [c for c in caps if c.get("round", "") >= "r02n"]  # (an item was one tile before: those figures do not transfer)
            exact = [c for c in caps if c["docs"] == docs and c["queries"] == queries and c.get("n_gpus", 1) == world]
            cap = exact[-1] if exact else (caps[-1] if caps else None)
            estimated = not exact
        except Exception:
            cap = None
    issue_peak = N_SMS * 4 * clk_mhz * 1e6 / 1e9  # G warp-instructions / s
    out = {"kernel": "plane_eval_kernel", "bound": "issue", "unit": "Gwarp-inst/s", "peak": issue_peak, "peak_source": peak_src, "launch_ms": launch_ms,
           "launch": "seed + bulk launch of plane_eval_kernel of one step (CUDA events on the library's stream)", "items_per_launch": items}
    if cap and launch_ms > 0:
        # the capture is of the bulk launch; the live launch time covers seed + bulk: per-item figures x the live item count
        inst = cap["inst_per_item"] * items
        dram = cap["dram_bytes_per_item"] * items
        out.update({"achieved": inst / (launch_ms / 1000.0) / 1e9, "frac": inst / (launch_ms / 1000.0) / 1e9 / issue_peak,
                    "traffic": dram, "hbm_frac_actual": dram / (launch_ms / 1000.0) / 1e9 / hbm_peak, "hbm_peak_gbs": hbm_peak,
                    "warp_instructions_per_launch": inst, "capture": cap.get("capture"), "capture_is_of_this_workload": not estimated,
                    "how": "warp instructions (and DRAM bytes) per (tile, request) item from the ncu capture x items of this launch / live CUDA-event time"})
    else:
        out.update({"achieved": None, "frac": None, "traffic": None})
    tm = rec["traffic_model"]
    model_bytes = tm["posting_bytes"] + tm["boost_bytes"] + 8 * 10 * queries
    tile_ms = rec["phase_ms"]["plane_eval"] + rec["phase_ms"]["tile_eval"]
    out["model_vs_streaming"] = {"ratio": model_bytes / (tile_ms / 1000.0) / 1e9 / hbm_peak if tile_ms > 0 else None, "model_bytes": model_bytes, "ms": tile_ms,
                                 "meaning": "bytes of SURVEY 8(d)'s posting-list model (6 B per posting of every matched term + 4 B boost per hit + 8 B per returned hit) "
                                            "divided by the time of plane_eval + tile_eval and by the HBM peak: above 1 means faster than a perfect HBM-streaming "
                                            "implementation of that model; it is not a bandwidth measurement"}
    # the HBM-streaming kernels of the step: achieved GB/s of algorithmic bytes (8 B read per sparse posting; + 8 B written by the fill)
    if rec.get("kernels_ms"):
        sp = rec["work"]["sparse_entries"]
        rows = []
        total = sum(v["ms"] for v in rec["kernels_ms"].values()) or 1.0
        alg = {"sparse_count": 8 * sp, "sparse_fill": 16 * sp}
        for name, v in sorted(rec["kernels_ms"].items(), key=lambda kv: -kv[1]["ms"]):
            row = {"kernel": name, "launches": v["launches"], "ms": v["ms"], "share": v["ms"] / total}
            if name in alg and v["ms"] > 0:
                row["algorithmic_bytes"] = alg[name]
                row["gbs"] = alg[name] / (v["ms"] / 1000.0) / 1e9
                row["hbm_frac"] = row["gbs"] / hbm_peak
            rows.append(row)
        out["kernels"] = rows
    return out


def run_ours(args):
    import helpers
    import veloci_b200
    from veloci_b200 import build

    # The one JSON line is the only thing that may reach stdout: native libraries (NCCL prints its version banner there)
    # write to file descriptor 1 directly, so descriptor 1 is pointed at stderr for the run and the line goes to a saved copy.
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    job = Job()
    if job.rank == 0:
        build.build_all()
    job.barrier()
    if veloci_b200.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device (the CUDA path has no CPU fallback)")

    main = run_workload(job, args, helpers, veloci_b200, args.docs, args.vocab, args.queries, args.steps, args.warmup, args.parity, True, True)
    c5 = None
    if args.config5_docs > 0:
        try:
            c5 = run_workload(job, args, helpers, veloci_b200, args.config5_docs, args.vocab, args.queries, args.steps, args.warmup, args.config5_parity, False, True)
        except Exception as e:  # the headline line must survive a box that cannot hold the 100M-doc corpus
            c5 = {"error": f"{type(e).__name__}: {e}"} if job.rank == 0 else None
    if job.rank == 0:
        line = {
            "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": job.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.docs, args.vocab, args.queries),
        }
        for k in ("p50_batch_latency_ms", "phase_ms", "device_ms_per_step", "work", "requests_ok", "index", "e2e", "gpu_launches", "clocks", "parity", "cpu_baseline"):
            if k in main:
                line[k] = main[k]
        line["roofline"] = roofline_of(main, args.docs, args.queries, job.world)
        if c5 is not None:
            if "error" not in c5:
                c5["config"] = workload_config(args.config5_docs, args.vocab, args.queries, "BASELINE.json configs[4]: the north-star corpus, sharded by anchor range over the N GPUs")
                c5["roofline"] = roofline_of(c5, args.config5_docs, args.queries, job.world)
                c5.pop("traffic_model", None), c5.pop("kernels_ms", None)
            line["config5"] = c5
        os.write(result_fd, (json.dumps(line) + "\n").encode())
    job.close()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
