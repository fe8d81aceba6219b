"""Where does the end-to-end step time go?  (diagnostic; prints one line per variant)"""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kbo_b200 import api, synth
api.load_library()
K, P = 31, 1e-7
ref = synth.random_seq(5_000_000, synth.SEED_C2_REF)
index = api.build([ref], api.BuildOpts(k=K))
batches = []
for b in range(8):
    concat, offsets = synth.gene_queries(ref, 10_000, 1000, synth.SEED_C2_GENES + b)
    batches.append(concat)
nq = len(offsets) - 1
pins = [api.PinnedBytes(len(b)) for b in batches]
for p_, b in zip(pins, batches):
    p_.array[:] = b
pin_off = api.PinnedBytes(8 * (nq + 1)); offp = pin_off.array.view(np.uint64); offp[:] = offsets
STEPS = 60
def run(name, submit, depth):
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        infl = []
        for s in range(STEPS):
            if len(infl) == depth:
                infl.pop(0)()
            infl.append(submit(s))
        for w in infl:
            w()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print("%-46s depth %d: %.1f us/step  %.1f G bases/s" % (name, depth, 1e6 * dt / STEPS, STEPS * 1e7 / dt / 1e9), flush=True)

# host-side cost of one submit and one wait (depth 6, so that waits rarely block)
bufs = [api.FindBuffers(nq, pinned=True) for _ in range(6)]
ts, tw = [], []
infl = []
for s_ in range(120):
    if len(infl) == 6:
        t0 = time.perf_counter(); infl.pop(0).wait(); tw.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    infl.append(api.find_submit(pins[s_ % 8].array, offp, index, api.FindOpts(P, 0), bufs[s_ % 6]))
    ts.append(time.perf_counter() - t0)
for j in infl:
    j.wait()
print("host time per submit: median %.1f us, per wait: median %.1f us" % (1e6 * np.median(ts[20:]), 1e6 * np.median(tw[20:])), flush=True)
import threading
def two_threads(depth):
    def worker(t):
        bufs_t = [api.FindBuffers(nq, pinned=True) for _ in range(depth)]
        infl = []
        for s in range(t, STEPS, 2):
            if len(infl) == depth:
                infl.pop(0).wait()
            infl.append(api.find_submit(pins[s % 8].array, offp, index, api.FindOpts(P, 0), bufs_t[(s // 2) % depth]))
        for j in infl:
            j.wait()
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ths = [threading.Thread(target=worker, args=(t,)) for t in range(2)]
        [th.start() for th in ths]; [th.join() for th in ths]
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("two host threads x depth %d: %.1f us/step  %.1f G bases/s" % (depth, 1e6 * dt / STEPS, STEPS * 1e7 / dt / 1e9), flush=True)
two_threads(2); two_threads(3)
for depth in (4, 6):
    bufs = [api.FindBuffers(nq, pinned=True) for _ in range(depth)]
    def sub(s):
        j = api.find_submit(pins[s % 8].array, offp, index, api.FindOpts(P, 0), bufs[s % depth]); return j.wait
    run("submit/wait, pinned outputs (direct writes)", sub, depth)
for depth in (3,):
    bufs = [api.FindBuffers(nq, pinned=False) for _ in range(depth)]
    def sub(s):
        j = api.find_submit(pins[s % 8].array, offp, index, api.FindOpts(P, 0), bufs[s % depth]); return j.wait
    run("submit/wait, pageable outputs (staged)", sub, depth)
# device-resident: kernels only, and H2D only
streams = [torch.cuda.Stream() for _ in range(4)]
d_in = [torch.empty(len(batches[0]), dtype=torch.uint8, device="cuda") for _ in range(4)]
d_off = torch.from_numpy(offsets.astype(np.int64)).cuda()
d_rle = [torch.empty(90_000 * 7, dtype=torch.int64, device="cuda") for _ in range(4)]
d_ro = [torch.empty(nq + 1, dtype=torch.int64, device="cuda") for _ in range(4)]
hin = [torch.from_numpy(p_.array) for p_ in pins]
def sub_dev(s):
    i = s % 4
    st = streams[i]
    api.find_device(index, d_in[i].data_ptr(), d_off.data_ptr(), offsets, d_rle[i].data_ptr(), 90_000, d_ro[i].data_ptr(), P, 0, st.cuda_stream)
    return st.synchronize
run("find_device only (no copies), 4 streams", sub_dev, 4)
run("find_device only (no copies), 1 in flight", sub_dev, 1)
def sub_h2d(s):
    i = s % 4
    with torch.cuda.stream(streams[i]):
        d_in[i].copy_(hin[s % 8], non_blocking=True)
    return streams[i].synchronize
run("H2D 10 MB only (torch copy from library-pinned)", sub_h2d, 4)
def sub_both(s):
    i = s % 4
    st = streams[i]
    with torch.cuda.stream(st):
        d_in[i].copy_(hin[s % 8], non_blocking=True)
    api.find_device(index, d_in[i].data_ptr(), d_off.data_ptr(), offsets, d_rle[i].data_ptr(), 90_000, d_ro[i].data_ptr(), P, 0, st.cuda_stream)
    return st.synchronize
run("H2D + find_device (records stay on device)", sub_both, 4)
run("H2D + find_device (records stay on device)", sub_both, 2)
