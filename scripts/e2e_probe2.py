"""Which part of a find job costs the time beyond the copy-in?  (diagnostic)"""
import os, sys, time, ctypes as C
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kbo_b200 import api, synth
L = api.load_library()
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
STEPS, CAP = 80, 90_000
def run(name, mk, depth):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter(); infl = []
        for s in range(STEPS):
            if len(infl) == depth: infl.pop(0)()
            infl.append(mk(s))
        for w in infl: w()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("%-58s depth %d: %6.1f us/step %5.1f G bases/s" % (name, depth, 1e6 * dt / STEPS, STEPS * 1e7 / dt / 1e9), flush=True)
def submit_raw(s, rle_ptr, off_ptr, concat_arr=None, offsets_arr=None):
    h = C.c_void_p()
    ca = concat_arr if concat_arr is not None else pins[s % 8].array
    oa = offsets_arr if offsets_arr is not None else offp
    rc = L.kbo_find_batch_submit(index._h, ca.ctypes.data_as(api.u8p), oa.ctypes.data_as(api.u64p), nq, P, 0,
                                 C.cast(rle_ptr, C.POINTER(api.RleC)), CAP, C.cast(off_ptr, api.u64p), C.byref(h))
    assert rc == 0, L.kbo_last_error_message()
    def wait():
        n = C.c_uint64(0); assert L.kbo_job_wait(h, C.byref(n)) == 0
    return wait
D = 6
d_rle = [torch.empty(CAP * 7, dtype=torch.int64, device="cuda") for _ in range(D)]
d_ro = [torch.empty(nq + 1, dtype=torch.int64, device="cuda") for _ in range(D)]
h_rle = [api.PinnedBytes(CAP * 56) for _ in range(D)]
h_ro = [api.PinnedBytes(8 * (nq + 1)) for _ in range(D)]
for depth in (4, 6):
    run("records -> device, offsets -> device", lambda s: submit_raw(s, d_rle[s % depth].data_ptr(), d_ro[s % depth].data_ptr()), depth)
    run("records -> pinned host, offsets -> device", lambda s: submit_raw(s, h_rle[s % depth]._p.value, d_ro[s % depth].data_ptr()), depth)
    run("records -> device, offsets -> pinned host", lambda s: submit_raw(s, d_rle[s % depth].data_ptr(), h_ro[s % depth]._p.value), depth)
    run("records -> pinned host, offsets -> pinned host", lambda s: submit_raw(s, h_rle[s % depth]._p.value, h_ro[s % depth]._p.value), depth)
# the same with a smaller batch (half): is the gap per call or per byte?
half = 5000
offh = api.PinnedBytes(8 * (half + 1)); offh_a = offh.array.view(np.uint64); offh_a[:] = offsets[:half + 1]
def submit_half(s, depth):
    h = C.c_void_p()
    rc = L.kbo_find_batch_submit(index._h, pins[s % 8].array.ctypes.data_as(api.u8p), offh_a.ctypes.data_as(api.u64p), half, P, 0,
                                 C.cast(h_rle[s % depth]._p.value, C.POINTER(api.RleC)), CAP, C.cast(h_ro[s % depth]._p.value, api.u64p), C.byref(h))
    assert rc == 0
    def wait():
        n = C.c_uint64(0); assert L.kbo_job_wait(h, C.byref(n)) == 0
    return wait
run("HALF batches (5 MB), all pinned  [per-step = per 5e6 bases]", lambda s: submit_half(s, 6), 6)
