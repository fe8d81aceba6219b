#!/usr/bin/env python3
"""bench.py -- query bases/s of the kbo MS hot path (matches/find) on B200.

Workload (BASELINE.json configs[1]): kbo::find of 10,000 synthetic 1 kbp gene queries (1 % SNPs)
against a 5 Mbp reference, k = 31, p = 1e-7; the index (30 MB: rank words, link words, LCS) is L2 resident.  A "step" is one
pass of the hot path over one batch of 10,000 queries (10^7 query bases).

  value : whole-job throughput with the batch already resident in HBM (kbo_find_batch_device:
          K0 pack -> K1 matching statistics -> K2b derandomize+translate (masks) -> K4 run-length records),
          independent steps round-robin on `--streams` streams forked from / joined into the timing stream,
          CUDA events on that stream, max over ranks.  (`--tuned-chunk-len N` repeats the region with an explicit
          chunk length for comparison; the library picks it per call from the batch size and the observed overlap.)  Steps rotate through `--batches` distinct batches whose total
          size exceeds L2, so queries always come from HBM while the index stays L2 resident.
  e2e   : the same metric through the host-buffer C ABI call a kbo user makes (kbo_find_batch):
          pinned host -> device copy of the queries, kernels, device -> host copy of the RLE records.
  roofline     : K1 (ms_kernel), algorithmic bytes / its CUDA-event duration vs the
                 measured HBM peak (MEASURED_PEAKS.json); random-sector L2/HBM rates beside it.
  cpu_baseline : the C++ oracle (restatement of kbo 0.5.1 + sbwt 0.3.4 semantics) running kbo::find
                 on all host cores over a bounded sample of the same workload.

`--impl reference` times that CPU restatement alone (the reference crate cannot be built here: no
Rust toolchain and its MS engine is the un-vendored crate sbwt 0.3.4).
Multi-GPU (torchrun, one rank per GPU): the index is replicated, every rank processes its own
batches (weak scaling), no collective on the data path; NCCL only carries the timing reduction.
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
    ap.add_argument("--ref-len", type=int, default=5_000_000)
    ap.add_argument("--queries", type=int, default=10_000)
    ap.add_argument("--query-len", type=int, default=1000)
    ap.add_argument("--batches", type=int, default=16, help="distinct batches rotated through (16 x 10 MB > L2)")
    ap.add_argument("--chunk-len", type=int, default=0, help="MS chunk length (0 = automatic)")
    ap.add_argument("--ms-flags", type=int, default=0, help="experiment switches (2: K2 instead of K2b)")
    ap.add_argument("--no-prefix-table", action="store_true", help="build the index without the prefix-state table (comparison)")
    ap.add_argument("--no-l2-persist", action="store_true", help="do not mark the index persisting in L2 (comparison)")
    ap.add_argument("--tuned-chunk-len", type=int, default=0,
                    help="second timed region with this chunk length (0 = skip); only when --chunk-len is automatic")
    ap.add_argument("--streams", type=int, default=6,
                    help="CUDA streams the device-resident steps are issued round-robin on (independent batches)")
    ap.add_argument("--e2e-threads", type=int, default=6,
                    help="host threads issuing the end-to-end calls concurrently (kbo-cli style per-query threading)")
    ap.add_argument("--pipeline-parts", type=int, default=0, help="sub-batches of a host-buffer call (0 = automatic)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="target CPU work (core-seconds) of the cpu_baseline sample")
    return ap.parse_args()


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
                                          "--format=csv,noheader,nounits", "-lms", "20"],
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
                                             synth.SEED_C2_GENES + 1000 * rank + b)
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
            "config": config_dict(args, sample_note=sample),
            "cpu_baseline": {"value": value, "unit": "query bases/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "query bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def config_dict(args, sample_note=None):
    c = {"workload": "kbo::find of %d synthetic %d bp gene queries (1%% SNPs) vs a %d bp synthetic reference, k=%d, "
                     "p=%g (BASELINE.json configs[1]; index L2-resident)" % (args.queries, args.query_len, args.ref_len,
                                                                           K, P),
         "queries_per_step": args.queries, "query_len": args.query_len, "ref_len": args.ref_len, "k": K,
         "l2_policy": "inputs larger than L2: steps rotate through %d distinct batches (%d MB of queries); the index "
                      "is L2-resident by the config's design" % (args.batches,
                                                                 args.batches * args.queries * args.query_len // 10**6)}
    if sample_note:
        c["sample"] = sample_note
    return c


def pin_to_gpu_numa_node(torch, dev):
    """Multi-GPU boxes: run this rank's host threads (and first-touch its pinned buffers) on the CPUs local to its GPU,
    so that the end-to-end copies do not cross the socket interconnect.  Best effort; returns a note for the JSON line."""
    try:
        bus = torch.cuda.get_device_properties(dev).pci_bus_id
        dom = torch.cuda.get_device_properties(dev).pci_domain_id
        devn = torch.cuda.get_device_properties(dev).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (dom, bus, devn)
        cpus = set()
        for part in open(path).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "rank pinned to the %d CPUs local to its GPU" % len(cpus)
    except Exception as ex:  # topology not exposed (containers): keep the inherited affinity
        return "no NUMA pinning (%s)" % type(ex).__name__
    return None


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
    numa_note = pin_to_gpu_numa_node(torch, dev) if world > 1 else None
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
    rle_cap = 16 * nq + 1024
    d_rle = [torch.empty(rle_cap * 7, dtype=torch.int64, device="cuda") for _ in batches]
    d_rle_off = [torch.empty(nq + 1, dtype=torch.int64, device="cuda") for _ in batches]
    torch.cuda.synchronize()

    workers = [torch.cuda.Stream() for _ in range(max(1, args.streams))]

    def step_device(s, on=None):
        b = s % len(batches)
        st = on if on is not None else workers[s % len(workers)].cuda_stream
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
    def timed_region(n_steps):
        """K steps round-robin on the worker streams, bracketed by events on the timing stream; returns ms."""
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record(stream)
        for w in workers:
            w.wait_event(a)
        for s in range(n_steps):
            step_device(args.warmup + s)
        for w in workers:
            done = torch.cuda.Event()
            done.record(w)
            stream.wait_event(done)
        b.record(stream)
        barrier()
        return a.elapsed_time(b)

    for s in range(max(args.warmup, len(workers))):
        step_device(s)
    barrier()
    sampler = ClockSampler(dev)
    sampler.start()
    n0 = api.kernel_launch_count()
    ms_total = timed_region(args.steps)
    launches = api.kernel_launch_count() - n0
    clocks = sampler.stop()
    # the same region with the chunk length that suits overlapped launches (fewer warm-up bases per chunk; a single
    # launch would be too narrow, the concurrent ones fill the machine).  Reported in config["overlap_tuned"].
    tuned = None
    if args.chunk_len == 0 and args.tuned_chunk_len:
        api.set_chunk_len(args.tuned_chunk_len)
        for s in range(max(args.warmup, len(workers))):
            step_device(s)
        ms_tuned = timed_region(args.steps)
        api.set_chunk_len(0)
        tt = torch.tensor([ms_tuned], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tuned = {"chunk_len": args.tuned_chunk_len, "ms_per_step": float(tt.item()) / args.steps,
                 "value": world * args.steps * bases_per_step / (float(tt.item()) * 1e-3)}
    # per-kernel durations: the same steps again with CUDA events around K0 / K1 / K2 of every call; the
    # library runs these instrumented calls serially (no sub-batch concurrency), so a kernel's elapsed time
    # is its own duration
    api.set_kernel_timing(True)
    step_device(0, on=sptr)  # sizes the serial workspace; discarded
    torch.cuda.synchronize()
    api.collect_kernel_times(index, sptr)
    for s in range(args.steps):
        step_device(args.warmup + s, on=sptr)
    torch.cuda.synchronize()
    api.set_kernel_timing(False)
    ksum, kcalls = api.collect_kernel_times(index, sptr)
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    lt = torch.tensor([float(launches)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    ms_max = float(t.item())
    value = world * args.steps * bases_per_step / (ms_max * 1e-3)

    # ---- e2e: host buffers through the public C ABI call (H2D + kernels + D2H + RLE) -----------
    # host batches in memory from the library's own pinned allocator (cudaHostAlloc); measured on this box:
    # copies from it run at 51-55 GB/s, copies from torch's pin_memory() buffers at 11-27 GB/s
    pinned_keep = [api.PinnedBytes(len(b)) for b in batches]
    for pb, b in zip(pinned_keep, batches):
        pb.array[:] = b
    pinned_np = [pb.array for pb in pinned_keep]
    n_thr = max(1, args.e2e_threads)
    fbufs = [api.FindBuffers(nq, pinned=True) for _ in range(n_thr)]
    n_rle_box = [0] * n_thr

    def e2e_worker(t, first, count):
        # every call copies its batch host->device, runs the kernels and copies the RLE records back
        for s in range(first + t, first + count, n_thr):
            _, n_rle_box[t] = api.find_csr(pinned_np[s % len(batches)], offsets, index, api.FindOpts(P, 0), fbufs[t])

    def run_threads(first, count):
        ths = [threading.Thread(target=e2e_worker, args=(t, first, count)) for t in range(n_thr)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        torch.cuda.synchronize()

    # warm-up with the same concurrency until every workspace the timed region needs exists at its final size
    # (the library creates them lazily, one per concurrent call; measured: the first ~50 calls carry that cost)
    # ... and one untimed pass of the same length: in some runs the first pass after start-up stays 10-25 % slower
    # for its whole length (profiles/README.md: repeated regions in one process settle at 47 G bases/s from the second on)
    run_threads(0, max(max(args.warmup, 16) * n_thr, args.steps))
    barrier()
    e2e_steps = args.steps
    w0 = time.perf_counter()
    run_threads(args.warmup, e2e_steps)
    e2e_s = time.perf_counter() - w0
    n_rle = n_rle_box[0]
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_steps * bases_per_step / float(te.item())

    # ---- roofline inputs: event counters of one batch (profiling build of K1, outside the timing) --
    api.set_profile_counters(True)
    out_host = api.matches_csr(pinned_np[0], offsets, index, P)
    api.set_profile_counters(False)
    cnt = index.ms_counters()
    L = cnt["bases_emitted"]
    alg_bytes = (SECTOR * (cnt["emit_extend_attempts"] + cnt["emit_extend_split_sector"]) +
                 SECTOR * (cnt["emit_contractions"] + cnt["emit_contraction_extra_words"]) + 2 * L)
    k1_ms = ksum["ms"] / max(kcalls, 1)
    achieved = alg_bytes / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else 0.0
    peak, peak_src = measured_peaks()
    roof = {"bound": "hbm", "kernel": "ms_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": committed_traffic(), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg_bytes, "algorithmic_bytes_per_base": alg_bytes / max(L, 1),
            "achieved_whole_step_overlapped": alg_bytes / ((ms_max / args.steps) * 1e-3) / 1e9,
            "kernel_ms": {"pack": ksum["pack"] / max(kcalls, 1), "ms": k1_ms,
                          "derand_translate": ksum["derand_translate"] / max(kcalls, 1),
                          "how": "CUDA events around each kernel over %d serial instrumented steps on the launch "
                                 "stream (the timed `value` region overlaps independent steps on %d streams)" % (kcalls, len(workers))},
            "events_per_base": {"extend_attempts": cnt["emit_extend_attempts"] / max(L, 1),
                                "contractions": cnt["emit_contractions"] / max(L, 1),
                                "warmup_overhead": cnt["bases_processed"] / max(L, 1)},
            "note": "the index is L2-resident in this config, so the HBM-peak fraction is not a ceiling; "
                    "the measured random-sector rates below are the relevant denominators"}
    if rank == 0:
        try:
            l2_dep = api.measure_random_sector_rate(8 << 20, True, dev)
            l2_ind = api.measure_random_sector_rate(8 << 20, False, dev)
            hbm_ind = api.measure_random_sector_rate(4 << 30, False, dev)
            sectors_per_s = (cnt["emit_extend_attempts"] + cnt["emit_extend_split_sector"] + cnt["emit_contractions"] +
                             cnt["emit_contraction_extra_words"]) / (k1_ms * 1e-3)
            roof["random_sector"] = {"l2_dependent_chain_sectors_per_s": l2_dep, "l2_independent_sectors_per_s": l2_ind,
                                     "hbm_independent_sectors_per_s": hbm_ind, "kernel_sectors_per_s": sectors_per_s,
                                     "frac_of_l2_independent": sectors_per_s / l2_ind if l2_ind else None,
                                     "frac_of_l2_dependent_chain": sectors_per_s / l2_dep if l2_dep else None}
        except Exception as ex:  # instrumentation only
            roof["random_sector"] = {"error": str(ex)}

    # ---- cpu baseline + a parity spot check against it (rank 0, N = 1) --------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, oix = cpu_baseline(ref, batches, offsets, args, args.cpu_seconds)
        nchk = min(nq, 200)
        _, want, _ = oix.matches_batch(batches[0][:int(offsets[nchk])], offsets[:nchk + 1], P, n_threads=cpu["cores"])
        if not np.array_equal(out_host[:int(offsets[nchk])], want):
            raise SystemExit("bench.py: GPU output differs from the oracle on the first %d queries" % nchk)

    if rank == 0:
        cfg = config_dict(args)
        cfg.update({"chunk_len": args.chunk_len or "auto (per call, from the batch size and the number of caller streams "
                                                  "seen in the last 8 stream-ordered calls)",
                    "index_device_bytes": index.device_bytes,
                    "n_sets": index.n_sets, "index_build_s_host": round(index_build_s, 2), "rle_records_per_step": n_rle,
                    "streams": len(workers),
                    "parallelism": "replicated index, %d rank(s) x own batches" % world})
        if numa_note:
            cfg["host_affinity"] = numa_note
        if tuned is not None:
            tuned["note"] = ("same timed region with kbo_set_chunk_len(%d): less chunk warm-up work per base; suits "
                             "overlapped launches, not a lone one (K1 alone is slower at this chunk length)" % tuned["chunk_len"])
            cfg["overlap_tuned"] = tuned
        line = {"metric": "query bases/s (kbo find, whole box)", "value": value, "unit": "query bases/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": cfg, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "query bases/s",
                        "h2d_bytes_per_step": bases_per_step + 8 * (nq + 1),
                        "d2h_bytes_per_step": 8 * (nq + 1) + 56 * n_rle,
                        "api": "kbo_find_batch: pinned host queries in, RLE records + per-query offsets out "
                               "(matches and run lengths computed on the device; one stream per call when several host threads call "
                               "concurrently, else up to 4 pipelined sub-batches); "
                               "%d steps issued by %d host threads" % (e2e_steps, n_thr)},
                "gpu_launches": int(lt.item()), "roofline": roof}
        if cpu is not None:
            line["cpu_baseline"] = cpu
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
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
