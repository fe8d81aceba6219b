#!/usr/bin/env python3
"""bench.py -- query bases/s of the kbo MS hot path (matches/find) on B200.

Workload (BASELINE.json configs[1]): kbo::find of 10,000 synthetic 1 kbp gene queries (1 % SNPs)
against a 5 Mbp reference, k = 31, p = 1e-7; the index (rank words, rank2 words, link words, LCS) is L2 resident.  A "step" is one
pass of the hot path over one batch of 10,000 queries (10^7 query bases).

  value : whole-job throughput with the batch already resident in HBM (kbo_find_batch_device:
          K0 pack -> fused K1+K2b (matching statistics, derandomize, translate; MS only in shared memory) -> K4
          run-length records), independent steps round-robin on `--streams` streams forked from / joined into the
          timing stream, CUDA events on that stream, max over ranks.  Steps rotate through `--batches` distinct
          batches whose total size exceeds L2, so queries always come from HBM while the index stays L2 resident.
          config.single_stream is the same region on ONE stream.
  e2e   : the same metric through the host-buffer C ABI a kbo user calls: kbo_find_batch_submit / kbo_job_wait from
          ONE host thread with `--e2e-depth` batches in flight: pinned host -> device copy of the queries, kernels,
          RLE records and per-query offsets written by the last kernel straight into the caller's pinned buffers.
          Multi-GPU: every rank does this for its own batches; the per-step record counts of all ranks are gathered
          with NCCL at the end, inside the timed region (the records stay on the ranks' hosts).
  roofline     : the fused kernel, algorithmic bytes / its CUDA-event duration, measured in the same
                 configuration as `value`, vs the measured L2 random-sector rate (the index is L2 resident; the
                 HBM copy peak of MEASURED_PEAKS.json is given beside it).
  cpu_baseline : the C++ oracle (restatement of kbo 0.5.1 + sbwt 0.3.4 semantics) running kbo::find
                 on all host cores over a bounded sample of the same workload.

`--impl reference` times that CPU restatement alone (the reference crate cannot be built here: no
Rust toolchain and its MS engine is the un-vendored crate sbwt 0.3.4).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# One hardware work queue per stream: with the default of 8 connections, streams of concurrent host-buffer calls
# alias onto the same queue and wait for each other's copies (measured: e2e 28-30 instead of 43-47 G bases/s when
# that happens).  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

K = 31
P = 1e-7
SECTOR = 32


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json configs[] entry (1-based): 2 = find, 10,000 x 1 kbp vs 5 Mbp (the headline, default); "
                         "5 = find vs an HBM-resident pangenome index with 10 kbp queries; 3 / 4 = call / map of whole "
                         "assemblies against one reference (index built per assembly)")
    ap.add_argument("--ref-len", type=int, default=0, help="reference length (0 = the config's: 5 Mbp, config 5: 1 Gbp per GPU)")
    ap.add_argument("--queries", type=int, default=0)
    ap.add_argument("--query-len", type=int, default=0)
    ap.add_argument("--assemblies", type=int, default=16, help="configs 3 / 4: assemblies per rank")
    ap.add_argument("--blocking-sync", type=int, default=0, help="configs 3 / 4: 1 = host threads sleep while they wait for the GPU")
    ap.add_argument("--asm-reps", type=int, default=3, help="configs 3 / 4: repetitions of every timed region (the fastest counts)")
    ap.add_argument("--asm-threads", type=int, default=4,
                    help="configs 3 / 4: host threads per rank, each taking whole assemblies (kbo-cli style)")
    ap.add_argument("--k", type=int, default=31, help="configs 3 / 4: k of the indexes (k > 32 takes the host builder; the "
                                                      "reference resolves variants only when k - threshold leaves room, e.g. k = 51)")
    ap.add_argument("--snp-rate", type=float, default=0.01, help="SNP rate of the synthetic gene queries (the config's: 0.01)")
    ap.add_argument("--batches", type=int, default=16, help="distinct batches rotated through (16 x 10 MB > L2)")
    ap.add_argument("--chunk-len", type=int, default=0, help="MS chunk length (0 = automatic)")
    ap.add_argument("--ms-flags", type=int, default=0, help="experiment switches (2: K2 instead of K2b)")
    ap.add_argument("--prefix-len", type=int, default=0, help="depth of the prefix-state table (0 = the default, 10)")
    ap.add_argument("--no-prefix-table", action="store_true", help="build the index without the prefix-state table (comparison)")
    ap.add_argument("--no-rank2", action="store_true", help="build the index without the rank2 rows: one base per probe (comparison)")
    ap.add_argument("--no-l2-persist", action="store_true", help="do not mark the index persisting in L2 (comparison)")
    ap.add_argument("--streams", type=int, default=6,
                    help="CUDA streams the device-resident steps are issued round-robin on (independent batches)")
    ap.add_argument("--e2e-depth", type=int, default=6,
                    help="batches one host thread keeps in flight through kbo_find_batch_submit / kbo_job_wait")
    ap.add_argument("--e2e-threads", type=int, default=0,
                    help="> 0: issue the end-to-end steps as synchronous kbo_find_batch calls from this many host threads instead")
    ap.add_argument("--pipeline-parts", type=int, default=0, help="sub-batches of a host-buffer call (0 = automatic)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="target CPU work (core-seconds) of the cpu_baseline sample")
    args = ap.parse_args()
    if args.config == 5:
        args.ref_len = args.ref_len or 1_000_000_000
        args.queries = args.queries or 1000
        args.query_len = args.query_len or 10_000
    else:
        args.ref_len = args.ref_len or 5_000_000
        args.queries = args.queries or 10_000
        args.query_len = args.query_len or 1000
    return args


# ------------------------------------------------------------------------------------ helpers ---
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy peak)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


def committed_traffic():
    """dram bytes per ms_kernel launch from the committed ncu --set full capture, if present."""
    path = os.path.join(ROOT, "profiles", "ms_kernel_traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


def workload(args, rank):
    from kbo_b200 import synth
    ref = synth.random_seq(args.ref_len, synth.SEED_C2_REF)
    batches = []
    for b in range(args.batches):
        concat, offsets = synth.gene_queries(ref, args.queries, args.query_len,
                                             synth.SEED_C2_GENES + 1000 * rank + b, snp=args.snp_rate)
        batches.append(concat)
    return ref, batches, offsets


def cpu_baseline(ref, batches, offsets, args, budget_s):
    """kbo::find with the C++ oracle on all host cores over a bounded sample (rank 0 only)."""
    import oracle_lib as O
    cores = O.hardware_concurrency() or os.cpu_count() or 1
    t0 = time.perf_counter()
    oix = O.OracleIndex([ref.tobytes()], k=K)
    build_s = time.perf_counter() - t0
    nq_total = len(offsets) - 1
    O.find_batch_timed(oix, batches[0], offsets, P, 0, cores)  # warm the caches / thread start-up
    secs, bases, passes = 0.0, 0, 0
    while secs < budget_s / cores and passes < 64:  # budget_s is CPU work (core-seconds)
        t, _, _ = O.find_batch_timed(oix, batches[passes % len(batches)], offsets, P, 0, cores)
        secs += t
        bases += int(offsets[nq_total] - offsets[0])
        passes += 1
    return {"value": bases / secs, "unit": "query bases/s", "cores": cores, "kind": "port",
            "sample": "kbo::find of %d full batches of %d queries (%d bases) in %.2f s wall on %d threads "
                      "(%.1f core-seconds); oracle index build %.1f s excluded; C++ restatement of kbo 0.5.1 + "
                      "sbwt 0.3.4 semantics (the reference crate needs Rust + sbwt, unavailable here)"
                      % (passes, nq_total, bases, secs, cores, secs * cores, build_s)}, oix


# ---------------------------------------------------------------------------------- reference arm ---
def run_reference(args, rank, world):
    if rank != 0:
        return
    import oracle_lib as O
    O.build_oracle()
    ref, batches, offsets = workload(args, 0)
    cores = O.hardware_concurrency() or os.cpu_count() or 1
    oix = O.OracleIndex([ref.tobytes()], k=K)
    # bounded sample per step: sized from a probe so that steps+warmup finish within ~2 minutes
    probe = min(len(offsets) - 1, 50 * cores)
    secs, _, _ = O.find_batch_timed(oix, batches[0], offsets[:probe + 1], P, 0, cores)
    per_step_budget = max(0.5, min(5.0, 100.0 / max(1, args.steps + args.warmup)))
    nq = int(min(len(offsets) - 1, max(probe, probe / max(secs, 1e-6) * per_step_budget)))
    for s in range(args.warmup):
        O.find_batch_timed(oix, batches[s % len(batches)], offsets[:nq + 1], P, 0, cores)
    total = 0.0
    for s in range(args.steps):
        t, _, _ = O.find_batch_timed(oix, batches[s % len(batches)], offsets[:nq + 1], P, 0, cores)
        total += t
    bases = int(offsets[nq] - offsets[0])
    value = bases * args.steps / total
    sample = ("each step = kbo::find of %d of %d queries (%d bases) on %d host threads; C++ restatement of "
              "kbo 0.5.1 + sbwt 0.3.4 semantics" % (nq, len(offsets) - 1, bases, cores))
    line = {"impl": "reference", "metric": "query bases/s (kbo find, whole box)", "value": value,
            "unit": "query bases/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config_dict(args), "impl_detail": {"sample": sample},
            "cpu_baseline": {"value": value, "unit": "query bases/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "query bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def config_dict(args):
    where = ("BASELINE.json configs[4] shape on one GPU per rank: pangenome index HBM-resident, L2-miss-bound rank queries"
             if args.config == 5 else "BASELINE.json configs[1]; index L2-resident")
    c = {"workload": "kbo::find of %d synthetic %d bp gene queries (1%% SNPs) vs a %d bp synthetic reference, k=%d, "
                     "p=%g (%s)" % (args.queries, args.query_len, args.ref_len, K, P, where),
         "queries_per_step": args.queries, "query_len": args.query_len, "ref_len": args.ref_len, "k": K,
         "l2_policy": "inputs larger than L2: steps rotate through %d distinct batches (%d MB of queries); the index "
                      "is L2-resident by the config's design" % (args.batches,
                                                                 args.batches * args.queries * args.query_len // 10**6)}
    return c  # identical for both arms; everything arm-specific goes to the line's top-level "impl_detail"


def partition_cpus(local_rank, local_world):
    """One rank per GPU on one box: give every rank its own slice of the CPUs this process may use (round 1 pinned all
    ranks to the same NUMA-local list, i.e. onto each other).  Returns a note for the JSON line."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
        per = len(cpus) // max(1, local_world)
        if per < 1:
            return "no pinning (%d CPUs for %d ranks)" % (len(cpus), local_world)
        mine = cpus[local_rank * per:(local_rank + 1) * per]
        os.sched_setaffinity(0, mine)
        return "rank pinned to its own %d of %d CPUs" % (len(mine), len(cpus))
    except Exception as ex:
        return "no pinning (%s)" % type(ex).__name__


# --------------------------------------------------------------------------------------- our arm ---
def run_ours(args, rank, local_rank, world):
    import torch
    from kbo_b200 import api, build
    build.build_library()
    api.load_library()
    if not torch.cuda.is_available() or api.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (kbo_b200 has no CPU fallback)")
    dev = local_rank
    torch.cuda.set_device(dev)
    numa_note = partition_cpus(local_rank, int(os.environ.get("LOCAL_WORLD_SIZE", world))) if world > 1 else None
    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # the version banner goes to stdout, ahead of the JSON line
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    api.set_chunk_len(args.chunk_len)
    if args.ms_flags:
        api.set_ms_flags(args.ms_flags)
    if args.no_l2_persist:
        api.set_l2_persist(False)
    if args.no_prefix_table:
        api.set_prefix_table(False)
    if args.prefix_len:
        api.set_prefix_len(args.prefix_len)
    if args.no_rank2:
        api.set_rank2(False)
    if args.pipeline_parts:
        api.set_pipeline_parts(args.pipeline_parts)

    ref, batches, offsets = workload(args, rank)
    hw = os.cpu_count() or 1
    t0 = time.perf_counter()
    index = api.build([ref], api.BuildOpts(k=K, num_threads=min(hw, 16)), device=dev)
    index_build_s = time.perf_counter() - t0
    nq = len(offsets) - 1
    bases_per_step = int(offsets[nq])

    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    d_off = torch.from_numpy(offsets.astype(np.int64)).cuda()
    pinned_in = [torch.from_numpy(b).pin_memory() for b in batches]
    d_in = [p.cuda(non_blocking=True) for p in pinned_in]
    rle_cap = 16 * nq + bases_per_step // 16 + 1024
    d_rle = [torch.empty(rle_cap * 7, dtype=torch.int64, device="cuda") for _ in batches]
    d_rle_off = [torch.empty(nq + 1, dtype=torch.int64, device="cuda") for _ in batches]
    torch.cuda.synchronize()

    workers = [torch.cuda.Stream() for _ in range(max(1, args.streams))]

    def step_device(s, on=None, use=None):
        b = s % len(batches)
        use = use or workers
        st = on if on is not None else use[s % len(use)].cuda_stream
        api.find_device(index, d_in[b].data_ptr(), d_off.data_ptr(), offsets, d_rle[b].data_ptr(), rle_cap,
                        d_rle_off[b].data_ptr(), P, 0, st)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident ---------------------------------------------------------------
    # Steps are independent batches; they are issued round-robin on `--streams` streams that fork from and
    # join into the timing stream, so the CUDA events on that stream bracket exactly the K steps.
    def timed_region(n_steps, use):
        """K steps round-robin on the streams `use`, bracketed by events on the timing stream; returns ms."""
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record(stream)
        for w in use:
            w.wait_event(a)
        for s in range(n_steps):
            step_device(args.warmup + s, use=use)
        for w in use:
            done = torch.cuda.Event()
            done.record(w)
            stream.wait_event(done)
        b.record(stream)
        barrier()
        return a.elapsed_time(b)

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for s in range(max(args.warmup, len(workers))):
        step_device(s)
    barrier()
    sampler = ClockSampler(dev) if rank == 0 else None  # one sampler per box, 100 ms period
    if sampler:
        sampler.start()
    n0 = api.kernel_launch_count()
    ms_total = timed_region(args.steps, workers)
    launches = api.kernel_launch_count() - n0
    ms_max = reduce_max(ms_total)
    value = world * args.steps * bases_per_step / (ms_max * 1e-3)
    # the same steps on ONE stream (no overlap between steps)
    for s in range(args.warmup):
        step_device(s, use=workers[:1])
    ms_single = reduce_max(timed_region(args.steps, workers[:1]))
    # per-kernel durations: the same steps again with CUDA events around K0 / fused K1+K2b of every call, serially on
    # the timing stream (same kernels, same geometry as the timed region: the library derives it from the batch only)
    api.set_kernel_timing(True)
    step_device(0, on=sptr)  # sizes the serial workspace; discarded
    torch.cuda.synchronize()
    api.collect_kernel_times(index, sptr)
    for s in range(args.steps):
        step_device(args.warmup + s, on=sptr)
    torch.cuda.synchronize()
    api.set_kernel_timing(False)
    ksum, kcalls = api.collect_kernel_times(index, sptr)
    lt = torch.tensor([float(launches)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)

    # ---- e2e: host buffers through the public C ABI (H2D + kernels + results written into pinned host memory) -----
    # host batches in memory from the library's own pinned allocator (cudaHostAlloc); measured on this box:
    # copies from it run at 51-55 GB/s, copies from torch's pin_memory() buffers at 11-27 GB/s
    pinned_keep = [api.PinnedBytes(len(b)) for b in batches]
    for pb, b in zip(pinned_keep, batches):
        pb.array[:] = b
    pinned_np = [pb.array for pb in pinned_keep]
    pin_off = api.PinnedBytes(8 * (nq + 1))
    offsets_pinned = pin_off.array.view(np.uint64)
    offsets_pinned[:] = offsets
    depth = max(1, args.e2e_depth)
    n_thr = max(0, args.e2e_threads)
    fbufs = [api.FindBuffers(nq, cap=rle_cap, pinned=True) for _ in range(max(depth, n_thr))]
    counts = np.zeros(max(args.steps, 1), dtype=np.int64)

    def e2e_async(first, count, record):
        """One host thread, `depth` batches in flight: submit, submit, ..., wait for the oldest, submit, ..."""
        inflight = []
        for s in range(first, first + count):
            if len(inflight) == depth:
                s0, job = inflight.pop(0)
                n = job.wait()
                if record:
                    counts[s0 - first] = n
            inflight.append((s, api.find_submit(pinned_np[s % len(batches)], offsets_pinned, index, api.FindOpts(P, 0),
                                                fbufs[s % depth])))
        for s0, job in inflight:
            n = job.wait()
            if record:
                counts[s0 - first] = n

    def e2e_threads(first, count, record):
        def worker(t):
            for s in range(first + t, first + count, n_thr):
                _, n = api.find_csr(pinned_np[s % len(batches)], offsets_pinned, index, api.FindOpts(P, 0), fbufs[t])
                if record:
                    counts[s - first] = n
        ths = [threading.Thread(target=worker, args=(t,)) for t in range(n_thr)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()

    run_e2e = e2e_threads if n_thr else e2e_async
    # warm-up with the same concurrency until every workspace the timed region needs exists at its final size
    run_e2e(0, max(args.warmup, 2 * max(depth, n_thr, 1)) + 8, False)
    barrier()
    w0 = time.perf_counter()
    run_e2e(args.warmup, args.steps, True)
    gathered = None
    if dist is not None:  # final gather of the per-step record counts of every rank (NCCL), inside the timed region
        mine = torch.from_numpy(counts[:args.steps]).cuda()
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        torch.cuda.synchronize()
    e2e_s = time.perf_counter() - w0
    n_rle = int(counts[0])
    e2e_value = world * args.steps * bases_per_step / reduce_max(e2e_s)
    # the copy-in alone, all ranks at once: what the host -> device path of this box allows for this many GPUs
    hin = [torch.from_numpy(a) for a in pinned_np]
    d_tmp = [torch.empty(len(batches[0]), dtype=torch.uint8, device="cuda") for _ in range(4)]
    cstreams = workers[:4] if len(workers) >= 4 else [torch.cuda.Stream() for _ in range(4)]

    def copy_in_only(n):
        for s in range(n):
            with torch.cuda.stream(cstreams[s % 4]):
                d_tmp[s % 4].copy_(hin[s % len(hin)], non_blocking=True)
        torch.cuda.synchronize()
    copy_in_only(4)
    barrier()
    w0 = time.perf_counter()
    copy_in_only(args.steps)
    copy_value = world * args.steps * bases_per_step / reduce_max(time.perf_counter() - w0)
    total_records = int(sum(int(g.sum().item()) for g in gathered)) if gathered is not None else int(counts[:args.steps].sum())
    clocks = sampler.stop() if sampler else None

    # ---- roofline inputs: event counters of one batch (profiling build of the kernel, outside the timing) --
    api.set_profile_counters(True)
    out_host = api.matches_csr(pinned_np[0], offsets, index, P)
    api.set_profile_counters(False)
    cnt = index.ms_counters()
    L = max(cnt["bases_emitted"], 1)
    alg_sectors = (cnt["emit_extend_attempts"] + cnt["emit_extend_split_sector"] + cnt["emit_contractions"] +
                   cnt["emit_contraction_extra_words"])
    alg_bytes = SECTOR * alg_sectors + 2 * L
    k1_ms = ksum["ms"] / max(kcalls, 1)
    achieved = alg_bytes / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else 0.0
    hbm_peak, peak_src = measured_peaks()
    roof = {"kernel": "ms_fused_kernel (K1 matching statistics + K2b derandomize/translate in one kernel)" if args.ms_flags & 16
            else "ms_kernel (K1: matching statistics; the dominant kernel of a find step)",
            "achieved": achieved, "unit": "GB/s", "traffic": committed_traffic(),
            "algorithmic_bytes_per_launch": alg_bytes, "algorithmic_bytes_per_base": alg_bytes / L,
            "algorithmic_bytes_how": "32 B x (rank probes + probes whose two ends fall in different sectors + contractions "
                                     "+ contractions that had to scan), counted by the profiling build of the kernel for "
                                     "the tile's own positions only (repair-pass work included, chunk warm-up and look-ahead "
                                     "excluded), + 2 B per base (1 in, 1 out)",
            "reference_algorithm_bytes_per_base": 51.0,
            "reference_algorithm_note": "SURVEY 8d figure for this workload (one base per probe, one-by-one contraction: "
                                        "1.26 probes + 0.26 contractions per base); the kernel does the same job with fewer "
                                        "probes (two bases per probe, contraction jumps), which LOWERS its own figure",
            "achieved_reference_algorithm_GBps": 51.0 * L / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else 0.0,
            "kernel_ms": {"pack": ksum["pack"] / max(kcalls, 1), "ms": k1_ms,
                          "how": "CUDA events around each kernel over %d serial steps on the launch stream: the configuration "
                                 "of impl_detail.single_stream (chunk length 64); the %d-stream `value` region overlaps steps "
                                 "and runs K1 with longer chunks" % (kcalls, len(workers))},
            "step_ms_single_stream": ms_single / args.steps,
            "events_per_base": {"rank_probes": cnt["emit_extend_attempts"] / L,
                                "contractions": cnt["emit_contractions"] / L,
                                "processed_over_emitted": cnt["bases_processed"] / L}}
    l2_ind = None
    if rank == 0:
        try:
            l2_dep = api.measure_random_sector_rate(8 << 20, True, dev)
            l2_ind = api.measure_random_sector_rate(8 << 20, False, dev)
            hbm_ind = api.measure_random_sector_rate(4 << 30, False, dev)
            roof["random_sector"] = {"l2_dependent_chain_sectors_per_s": l2_dep, "l2_independent_sectors_per_s": l2_ind,
                                     "hbm_independent_sectors_per_s": hbm_ind,
                                     "kernel_algorithmic_sectors_per_s": alg_sectors / (k1_ms * 1e-3) if k1_ms > 0 else 0.0}
        except Exception as ex:  # instrumentation only
            roof["random_sector"] = {"error": str(ex)}
    if l2_ind and index.device_bytes < 100e6:
        roof.update({"bound": "l2", "peak": l2_ind * SECTOR / 1e9,
                     "peak_source": "measured in this run: independent random 32-byte-sector loads over an 8 MB (L2-resident) "
                                    "buffer, %.1f G sectors/s x 32 B (kbo_measure_random_sector_rate); the index (%d MB) is "
                                    "L2 resident" % (l2_ind / 1e9, index.device_bytes >> 20),
                     "hbm_copy_peak": hbm_peak, "hbm_copy_peak_source": peak_src, "frac_of_hbm_copy_peak": achieved / hbm_peak})
    else:
        roof.update({"bound": "hbm", "peak": hbm_peak, "peak_source": peak_src})
    roof["frac"] = achieved / roof["peak"]

    # ---- cpu baseline + a parity spot check against it (rank 0, N = 1) --------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.config == 2:
        cpu, oix = cpu_baseline(ref, batches, offsets, args, args.cpu_seconds)
        nchk = min(nq, 200)
        _, want, _ = oix.matches_batch(batches[0][:int(offsets[nchk])], offsets[:nchk + 1], P, n_threads=cpu["cores"])
        if not np.array_equal(out_host[:int(offsets[nchk])], want):
            raise SystemExit("bench.py: GPU output differs from the oracle on the first %d queries" % nchk)

    if rank == 0:
        cfg = config_dict(args)
        detail = {"index_device_bytes": index.device_bytes, "n_sets": index.n_sets,
                  "prefix_table_depth": args.prefix_len or 10,
                              "index_build_s": round(index_build_s, 3), "rle_records_per_step": n_rle,
                              "streams": len(workers),
                              "single_stream": {"ms_per_step": ms_single / args.steps,
                                                "value": world * args.steps * bases_per_step / (ms_single * 1e-3)},
                              "chunk_len": args.chunk_len or "auto (about 64 positions per lane, tiles a multiple of the SM count)",
                              "parallelism": "replicated index, %d rank(s) x own batches" % world}
        if numa_note:
            detail["host_affinity"] = numa_note
        line = {"metric": "query bases/s (kbo find, whole box)", "value": value, "unit": "query bases/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": cfg, "impl_detail": detail, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "query bases/s",
                        "h2d_bytes_per_step": bases_per_step + 8 * (nq + 1),
                        "d2h_bytes_per_step": 8 * (nq + 1) + 56 * n_rle + 8,
                        "records_total": total_records,
                        "copy_in_only_value": copy_value,
                        "copy_in_only_note": "same steps, only the pinned host -> device copy of the queries, all ranks at "
                                             "once: the ceiling of any host-buffer path on this box at this GPU count",
                        "api": ("kbo_find_batch_submit / kbo_job_wait: pinned host queries in, RLE records + per-query "
                                "offsets written by the device into pinned host buffers; %d steps, one host thread, %d "
                                "batches in flight" % (args.steps, depth)) if not n_thr else
                               ("kbo_find_batch (synchronous) from %d host threads; %d steps" % (n_thr, args.steps)),
                        "gather": "NCCL all_gather of every rank's per-step record counts at the end, inside the timed "
                                  "region; records stay in the ranks' host buffers" if dist is not None else "single rank"},
                "gpu_launches": int(lt.item()), "roofline": roof}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------- configs 3 / 4: whole assemblies ---
def run_assemblies(args, rank, local_rank, world):
    """BASELINE configs[2] (kbo::call) / configs[3] (kbo::map): a stream of mutated 5 Mbp assemblies against one
    reference.  Every assembly gets its own SBWT index (built on the GPU), the reference is streamed through it
    (K0, K1 with intervals, K2b), and the host refinement runs on the device's (d, l, r); kbo::call / kbo::map also
    rebuild the index of the reference per call, as the reference does (lib.rs:553).  A "step" is one assembly; the
    metric counts the streamed reference bases.  Assemblies are sharded over the ranks (no collective)."""
    import torch
    from kbo_b200 import api, build, synth
    build.build_library()
    api.load_library()
    dev = local_rank
    torch.cuda.set_device(dev)
    sync_note = "spin (CUDA default)"
    if args.blocking_sync:
        # host threads that wait for the GPU sleep instead of spinning: with several worker threads per rank and
        # 8 ranks on a 32-core box the spinning waiters take the cores the other workers' host work needs
        import ctypes
        try:
            rc = ctypes.CDLL("libcudart.so.12").cudaSetDeviceFlags(4)  # cudaDeviceScheduleBlockingSync
            sync_note = "blocking (cudaDeviceScheduleBlockingSync, rc %d)" % rc
        except OSError as ex:
            sync_note = "spin (libcudart not found: %s)" % ex
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
        partition_cpus(local_rank, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
    ref = synth.random_seq(args.ref_len, synth.SEED_C2_REF)
    refb = ref.tobytes()
    n_asm = max(1, args.assemblies)
    hw = len(os.sched_getaffinity(0))
    kk = args.k
    bo = api.BuildOpts(k=kk, build_select=True, num_threads=hw)  # (num_threads also threads the gap filling)
    name = "call" if args.config == 3 else "map"

    asms = {i: synth.mutate(ref, 0x6B626F10 + 1000 * rank + i) for i in range(n_asm + 2)}  # synthetic inputs: untimed

    def one(i, ref_index=None):
        asm = asms[i]
        t0 = time.perf_counter()
        ix = api.build([asm], bo, device=dev)
        t1 = time.perf_counter()
        if args.config == 3:
            res = api.call(ix, refb, api.CallOpts(sbwt_build_opts=bo), ref_index=ref_index)
        else:
            res = api.map(refb, ix, api.MapOpts(sbwt_build_opts=bo), ref_index=ref_index)
        t2 = time.perf_counter()
        ix.close()
        return asm, res, t1 - t0, t2 - t1, time.perf_counter() - t2

    def region(n_threads, ref_index=None):
        """n_asm assemblies on n_threads host threads (kbo-cli style: one assembly per worker at a time); returns
        (seconds max over ranks, summed build / run / free seconds, result of assembly 0)."""
        results = [None] * n_asm

        def worker(t):
            torch.cuda.set_device(dev)
            for i in range(t, n_asm, n_threads):
                results[i] = one(i, ref_index)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        w0 = time.perf_counter()
        if n_threads == 1:
            worker(0)
        else:
            ths = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
            for th in ths:
                th.start()
            for th in ths:
                th.join()
        torch.cuda.synchronize()
        total = time.perf_counter() - w0
        t = torch.tensor([total], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sums = [sum(r[j] for r in results) for j in (2, 3, 4)]
        return float(t.item()), sums, (results[0][0], results[0][1])

    n_thr = max(1, args.asm_threads)
    for i in range(2):  # warm-up (allocator pools, pinned staging)
        one(n_asm + i)
    region(n_thr)  # (and once with the timed region's concurrency: every thread's workspaces exist afterwards)

    def best_of(reps, *a):
        """The region `reps` times, the fastest kept: the per-assembly host work (ctypes calls, numpy buffers, the staging
        copies) is exposed to whatever else runs on the box's CPUs, and single regions varied up to 5x between boxes."""
        runs = [region(*a) for _ in range(reps)]
        return min(runs, key=lambda r: r[0])

    reps = max(1, args.asm_reps)
    total_s, (build_s, run_s, free_s), first = best_of(reps, n_thr)
    value = world * n_asm * len(ref) / total_s
    one_s, (build1, run1, free1), first1 = best_of(reps, 1) if n_thr > 1 else (total_s, (build_s, run_s, free_s), first)
    # the index of the reference built once and passed in (kbo_call_with_ref / kbo_map_with_ref) instead of per call
    ref_ix = api.build([ref], bo, device=dev)
    reuse_s, _, first_reuse = best_of(reps, n_thr, ref_ix)
    ref_ix.close()
    if first1[1] != first[1] or first_reuse[1] != first[1]:
        raise SystemExit("bench.py: kbo::%s results differ between the timed regions" % name)
    parity = None
    if rank == 0:  # bit-exact against the oracle on the first assembly
        import oracle_lib as O
        O.build_oracle()
        t0 = time.perf_counter()
        oix = O.OracleIndex([first[0].tobytes()], k=kk)
        if args.config == 3:
            want = oix.call(refb, 1e-7, kk)
            got = [(v.query_pos, v.query_chars, v.ref_chars) for v in first[1]]
        else:
            want = oix.map(refb, build_k=kk)
            got = first[1]
        cpu_s = time.perf_counter() - t0
        if got != want:
            raise SystemExit("bench.py: kbo::%s differs from the oracle on the first assembly" % name)
        parity = {"checked": "assembly 0 bit-exact vs the CPU oracle (%s)" % ("%d variants" % len(want) if args.config == 3
                                                                             else "%d output bytes" % len(want)),
                  "oracle_seconds_index_plus_%s_one_thread" % name: round(cpu_s, 2),
                  "oracle_bases_per_s": len(ref) / cpu_s}
        line = {"metric": "query bases/s (kbo %s, whole box)" % name, "value": value, "unit": "query bases/s",
                "n_gpus": world, "steps": n_asm, "warmup": 2 + n_asm, "ms_per_step": 1e3 * total_s / n_asm,
                "timing": "fastest of %d repetitions of the %d-assembly region (wall clock, max over ranks)" % (reps, n_asm),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": "kbo::%s of %d mutated %d bp synthetic assemblies per rank (1%% SNPs + short indels) "
                                       "against one reference, k=%d, default options (BASELINE.json configs[%d])"
                                       % (name, n_asm, args.ref_len, kk, args.config - 1),
                           "assemblies_per_rank": n_asm, "ref_len": args.ref_len, "k": kk},
                "impl_detail": {"host_threads": n_thr, "host_waits": sync_note,
                                "one_host_thread": {"value": world * n_asm * len(ref) / one_s, "ms_per_assembly": 1e3 * one_s / n_asm,
                                                    "split_ms_per_assembly": {
                                                        "assembly_index_build (GPU builder incl. copy-in)": 1e3 * build1 / n_asm,
                                                        "kbo::%s (MS, fill_gaps on the device, reference-index build, "
                                                        "variant resolution)" % name: 1e3 * run1 / n_asm,
                                                        "index free": 1e3 * free1 / n_asm}},
                                "reference_index_built_once": {"value": world * n_asm * len(ref) / reuse_s,
                                                               "ms_per_assembly": 1e3 * reuse_s / n_asm,
                                                               "api": "kbo_%s_with_ref (the reference rebuilds it per call, "
                                                                      "lib.rs:553; `value` does too)" % name},
                                "parity": parity}}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun like the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.config in (3, 4):
        run_assemblies(args, rank, local_rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
