"""Host-side mirror of the reference crate's public functions for the MS hot path, over the C ABI
of include/kbo_b200.h (ctypes; no torch types cross the boundary).

Names, argument meaning and error behaviour follow tmaklin/kbo 0.5.1:
  build / matches / find / map            src/lib.rs:501, 612, 808, 720
  query_sbwt                              src/index.rs:243
  random_match_threshold, log_rm_max_cdf,
  derandomize_ms_vec                      src/derandomize.rs:127, 91, 269
  translate_ms_vec                        src/translate.rs:263
  run_lengths, run_lengths_gapped,
  relative_to_ref                         src/format.rs:98, 143, 266
Where the reference panics, these raise KboPanic(status, message).

There is no CPU fallback: importing works without the shared library, but the first call raises
if kbo_b200/libkbo_b200.so has not been built (python -m kbo_b200.build) or no CUDA device exists.
"""
import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libkbo_b200.so")
_lib = None

u8p, u32p, u64p, i64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_int64)

STATUS_NAMES = {0: "OK", 1: "EMPTY_INPUT", 2: "BAD_THRESHOLD", 3: "TOO_SHORT", 4: "BAD_K", 5: "K_MISMATCH",
                6: "BAD_PROB", 7: "BAD_ARGUMENT", 8: "CUDA", 9: "OOM", 10: "INDEX_TOO_LARGE", 11: "BUFFER_TOO_SMALL",
                12: "PANIC", 13: "BATCH_TOO_LARGE", 14: "IO", 15: "FORMAT"}

# every symbol include/kbo_b200.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = [
    "kbo_last_error_message", "kbo_device_count", "kbo_default_build_opts", "kbo_alloc_pinned", "kbo_free_pinned",
    "kbo_index_build", "kbo_index_from_parts", "kbo_index_free", "kbo_index_k", "kbo_index_n_kmers",
    "kbo_index_n_sets", "kbo_index_device", "kbo_index_device_bytes", "kbo_index_export_parts",
    "kbo_index_serialize", "kbo_index_load",
    "kbo_index_access_kmer", "kbo_index_search", "kbo_query_sbwt", "kbo_query_sbwt_batch_compact",
    "kbo_log_rm_max_cdf", "kbo_random_match_threshold", "kbo_derandomize_ms_vec", "kbo_translate_ms_vec",
    "kbo_run_lengths_gapped", "kbo_relative_to_ref", "kbo_matches", "kbo_matches_batch", "kbo_matches_batch_device",
    "kbo_find_batch", "kbo_find_batch_submit", "kbo_job_wait", "kbo_find_batch_device", "kbo_ctx_create", "kbo_ctx_free", "kbo_ctx_n_gpus", "kbo_index_set_build", "kbo_index_set_free", "kbo_index_set_get",
    "kbo_matches_batch_multi", "kbo_find_batch_multi", "kbo_map_unrefined", "kbo_call", "kbo_map", "kbo_call_with_ref", "kbo_map_with_ref", "kbo_set_profile_counters", "kbo_get_ms_counters", "kbo_set_chunk_len", "kbo_set_l2_persist", "kbo_set_prefix_table", "kbo_set_prefix_len", "kbo_set_rank2", "kbo_set_refine_threads", "kbo_set_device_refine", "kbo_index_set_tuning", "kbo_set_host_builder", "kbo_set_pipeline_parts", "kbo_set_device_parts", "kbo_set_ms_flags",
    "kbo_kernel_launch_count", "kbo_last_kernel_ms", "kbo_set_kernel_timing", "kbo_collect_kernel_times",
    "kbo_measure_random_sector_rate",
]


class KboPanic(RuntimeError):
    """Raised where the reference implementation would panic (or on a CUDA failure)."""

    def __init__(self, status, message):
        super().__init__("kbo_b200 status %d (%s): %s" % (status, STATUS_NAMES.get(status, "?"), message))
        self.status = status


class BuildOptsC(C.Structure):
    _fields_ = [("k", C.c_uint32), ("add_revcomp", C.c_int32), ("num_threads", C.c_uint32),
                ("prefix_precalc", C.c_uint32), ("build_select", C.c_int32), ("mem_gb", C.c_uint32),
                ("dedup_batches", C.c_int32), ("temp_dir", C.c_char_p)]


class RleC(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("start", "end", "matches", "mismatches", "jumps", "gap_bases", "gap_opens")]


class MsCountersC(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("extend_attempts", "extend_split_sector", "contractions",
                                          "contraction_extra_words", "bases_processed", "bases_emitted",
                                          "emit_extend_attempts", "emit_extend_split_sector", "emit_contractions",
                                          "emit_contraction_extra_words")]


def load_library():
    """Loads kbo_b200/libkbo_b200.so; raises loudly when it is missing (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError("kbo_b200: %s is missing -- build it with `python -m kbo_b200.build` "
                           "(there is no CPU fallback)" % _LIB_PATH)
    # one hardware work queue per stream (the default of 8 makes concurrent calls wait for each other's copies);
    # only takes effect when the CUDA context does not exist yet
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    L = C.CDLL(_LIB_PATH)
    L.kbo_last_error_message.restype = C.c_char_p
    L.kbo_device_count.argtypes = [C.POINTER(C.c_int)]
    L.kbo_default_build_opts.argtypes = [C.POINTER(BuildOptsC)]
    L.kbo_alloc_pinned.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    L.kbo_free_pinned.argtypes = [C.c_void_p]
    L.kbo_index_build.argtypes = [C.POINTER(u8p), u64p, C.c_uint64, C.POINTER(BuildOptsC), C.c_int,
                                  C.POINTER(C.c_void_p)]
    L.kbo_index_from_parts.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.POINTER(u64p), u8p, C.c_int,
                                       C.POINTER(C.c_void_p)]
    L.kbo_index_free.argtypes = [C.c_void_p]
    L.kbo_index_free.restype = None
    L.kbo_index_k.argtypes = [C.c_void_p]
    L.kbo_index_k.restype = C.c_uint32
    for f in ("kbo_index_n_kmers", "kbo_index_n_sets", "kbo_index_device_bytes"):
        getattr(L, f).argtypes = [C.c_void_p]
        getattr(L, f).restype = C.c_uint64
    L.kbo_index_device.argtypes = [C.c_void_p]
    L.kbo_index_export_parts.argtypes = [C.c_void_p, C.POINTER(u64p), u8p, u64p]
    L.kbo_index_serialize.argtypes = [C.c_void_p, C.c_char_p]
    L.kbo_index_load.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    L.kbo_index_access_kmer.argtypes = [C.c_void_p, C.c_uint64, u8p]
    L.kbo_index_search.argtypes = [C.c_void_p, u8p, C.c_uint64, C.POINTER(C.c_int), u64p, u64p]
    L.kbo_query_sbwt.argtypes = [C.c_void_p, u8p, C.c_uint64, u64p, u64p, u64p]
    L.kbo_query_sbwt_batch_compact.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, u8p, u32p, u32p]
    L.kbo_log_rm_max_cdf.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_double)]
    L.kbo_random_match_threshold.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, u64p]
    L.kbo_derandomize_ms_vec.argtypes = [u64p, C.c_uint64, C.c_uint64, C.c_uint64, i64p, C.c_int]
    L.kbo_translate_ms_vec.argtypes = [i64p, C.c_uint64, C.c_uint64, C.c_uint64, u8p, C.c_int]
    L.kbo_run_lengths_gapped.argtypes = [u8p, C.c_uint64, C.c_uint64, C.POINTER(RleC), C.c_uint64, u64p]
    L.kbo_relative_to_ref.argtypes = [u8p, u8p, C.c_uint64, u8p]
    L.kbo_matches.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_double, u8p]
    L.kbo_matches_batch.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, C.c_double, u8p]
    L.kbo_matches_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, u64p, C.c_uint64, C.c_double,
                                           C.c_void_p, C.c_void_p]
    L.kbo_find_batch.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, C.c_double, C.c_uint64, C.POINTER(RleC),
                                 C.c_uint64, u64p]
    L.kbo_find_batch_submit.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, C.c_double, C.c_uint64, C.POINTER(RleC),
                                        C.c_uint64, u64p, C.POINTER(C.c_void_p)]
    L.kbo_job_wait.argtypes = [C.c_void_p, u64p]
    L.kbo_find_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, u64p, C.c_uint64, C.c_double, C.c_uint64,
                                        C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    L.kbo_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]
    L.kbo_ctx_free.argtypes = [C.c_void_p]
    L.kbo_ctx_free.restype = None
    L.kbo_ctx_n_gpus.argtypes = [C.c_void_p]
    L.kbo_index_set_build.argtypes = [C.c_void_p, C.POINTER(u8p), u64p, C.c_uint64, C.POINTER(BuildOptsC),
                                      C.POINTER(C.c_void_p)]
    L.kbo_index_set_free.argtypes = [C.c_void_p]
    L.kbo_index_set_free.restype = None
    L.kbo_index_set_get.argtypes = [C.c_void_p, C.c_int]
    L.kbo_index_set_get.restype = C.c_void_p
    L.kbo_matches_batch_multi.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, C.c_double, u8p]
    L.kbo_find_batch_multi.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, C.c_double, C.c_uint64, C.POINTER(RleC),
                                       C.c_uint64, u64p]
    L.kbo_call.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_double, C.POINTER(BuildOptsC), u64p, u32p, u32p, u8p, u8p,
                           C.c_uint64, C.c_uint64, u64p]
    L.kbo_map.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(BuildOptsC),
                          u8p]
    L.kbo_map_unrefined.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_double, C.c_int, u8p]
    L.kbo_call_with_ref.argtypes = [C.c_void_p, C.c_void_p, u8p, C.c_uint64, C.c_double, u64p, u32p, u32p, u8p, u8p,
                                    C.c_uint64, C.c_uint64, u64p]
    L.kbo_map_with_ref.argtypes = [C.c_void_p, C.c_void_p, u8p, C.c_uint64, C.c_double, C.c_int, C.c_int, C.c_int, u8p]
    L.kbo_set_profile_counters.argtypes = [C.c_int]
    L.kbo_get_ms_counters.argtypes = [C.c_void_p, C.POINTER(MsCountersC)]
    L.kbo_set_chunk_len.argtypes = [C.c_uint32]
    L.kbo_set_l2_persist.argtypes = [C.c_int]
    L.kbo_set_prefix_table.argtypes = [C.c_int]
    L.kbo_set_prefix_len.argtypes = [C.c_uint32]
    L.kbo_set_rank2.argtypes = [C.c_int]
    L.kbo_set_refine_threads.argtypes = [C.c_uint32]
    L.kbo_set_device_refine.argtypes = [C.c_int]
    L.kbo_index_set_tuning.argtypes = [C.c_void_p, C.c_int, C.c_int64]
    L.kbo_set_host_builder.argtypes = [C.c_int]
    L.kbo_set_pipeline_parts.argtypes = [C.c_uint32]
    L.kbo_set_device_parts.argtypes = [C.c_uint32]
    L.kbo_set_ms_flags.argtypes = [C.c_uint32]
    L.kbo_kernel_launch_count.restype = C.c_uint64
    L.kbo_last_kernel_ms.argtypes = [C.c_void_p]
    L.kbo_last_kernel_ms.restype = C.c_float
    L.kbo_set_kernel_timing.argtypes = [C.c_int]
    L.kbo_collect_kernel_times.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), u64p]
    L.kbo_measure_random_sector_rate.argtypes = [C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_double)]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise KboPanic(rc, load_library().kbo_last_error_message().decode())


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty))


def _u8(x):
    if isinstance(x, (bytes, bytearray)):
        return np.frombuffer(bytes(x), dtype=np.uint8)
    if isinstance(x, str):
        return np.frombuffer(x.encode(), dtype=np.uint8)
    return np.ascontiguousarray(x, dtype=np.uint8)


def csr(queries):
    """List of sequences -> (concat uint8 array, offsets uint64 array)."""
    qs = [_u8(q) for q in queries]
    offsets = np.zeros(len(qs) + 1, dtype=np.uint64)
    if qs:
        offsets[1:] = np.cumsum([len(q) for q in qs])
    concat = np.concatenate(qs) if qs else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(concat), offsets


# ------------------------------------------------------------------ option structs (lib.rs:259-466) ---
@dataclass
class BuildOpts:
    k: int = 31
    add_revcomp: bool = False
    num_threads: int = 1
    prefix_precalc: int = 8
    build_select: bool = False
    mem_gb: int = 4
    dedup_batches: bool = False
    temp_dir: Optional[str] = None


@dataclass
class MatchOpts:
    max_error_prob: float = 0.0000001


@dataclass
class FindOpts:
    max_error_prob: float = 0.0000001
    max_gap_len: int = 0


@dataclass
class CallOpts:
    max_error_prob: float = 0.0000001
    sbwt_build_opts: BuildOpts = field(default_factory=lambda: BuildOpts(build_select=True))


@dataclass(frozen=True)
class Variant:
    query_pos: int
    query_chars: bytes
    ref_chars: bytes


@dataclass
class MapOpts:
    max_error_prob: float = 0.0000001
    fill_gaps: bool = True
    call_variants: bool = True
    format: bool = True
    sbwt_build_opts: BuildOpts = field(default_factory=lambda: BuildOpts(build_select=True))


@dataclass(frozen=True)
class RLE:
    start: int
    end: int
    matches: int
    mismatches: int
    jumps: int
    gap_bases: int
    gap_opens: int


# ------------------------------------------------------------------------------------- index ---
class Index:
    """(SbwtIndexVariant::SubsetMatrix, LcsArray) resident on one GPU."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)
        L = load_library()
        self.k = L.kbo_index_k(self._h)
        self.n_kmers = L.kbo_index_n_kmers(self._h)
        self.n_sets = L.kbo_index_n_sets(self._h)
        self.device = L.kbo_index_device(self._h)

    @property
    def device_bytes(self):
        """Device memory of the index."""
        return load_library().kbo_index_device_bytes(self._h) if self._h else 0

    def close(self):
        if self._h:
            load_library().kbo_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def export_parts(self):
        nw = (self.n_sets + 63) // 64
        rows = [np.zeros(nw, dtype=np.uint64) for _ in range(4)]
        lcs = np.zeros(self.n_sets, dtype=np.uint8)
        Cc = np.zeros(4, dtype=np.uint64)
        ptrs = (u64p * 4)(*[_p(r, C.c_uint64) for r in rows])
        _check(load_library().kbo_index_export_parts(self._h, ptrs, _p(lcs, C.c_uint8), _p(Cc, C.c_uint64)))
        return rows, lcs, Cc

    def access_kmer(self, colex):
        out = np.zeros(self.k, dtype=np.uint8)
        _check(load_library().kbo_index_access_kmer(self._h, colex, _p(out, C.c_uint8)))
        return out.tobytes()

    def search(self, pattern):
        p = _u8(pattern)
        found, l, r = C.c_int(0), C.c_uint64(0), C.c_uint64(0)
        _check(load_library().kbo_index_search(self._h, _p(p, C.c_uint8), len(p), C.byref(found), C.byref(l),
                                               C.byref(r)))
        return (l.value, r.value) if found.value else None

    def ms_counters(self):
        c = MsCountersC()
        _check(load_library().kbo_get_ms_counters(self._h, C.byref(c)))
        return {n: int(getattr(c, n)) for n, _ in MsCountersC._fields_}

    def last_kernel_ms(self):
        return float(load_library().kbo_last_kernel_ms(self._h))


def build(seq_data, build_opts=None, device=0):
    """kbo::build (lib.rs:501-506) / index::build_sbwt_from_vecs (index.rs:56-99)."""
    L = load_library()
    o = build_opts or BuildOpts()
    seqs = [_u8(s) for s in seq_data]
    n = len(seqs)
    ptrs = (u8p * max(n, 1))(*[_p(s, C.c_uint8) for s in seqs])
    lens = np.array([len(s) for s in seqs], dtype=np.uint64)
    co = BuildOptsC(o.k, int(o.add_revcomp), o.num_threads, o.prefix_precalc, int(o.build_select), o.mem_gb,
                    int(o.dedup_batches), o.temp_dir.encode() if o.temp_dir else None)
    h = C.c_void_p()
    _check(L.kbo_index_build(ptrs, _p(lens, C.c_uint64), n, C.byref(co), device, C.byref(h)))
    return Index(h.value)


def index_from_parts(k, n_sets, n_kmers, rows, lcs, device=0):
    L = load_library()
    rows = [np.ascontiguousarray(r, dtype=np.uint64) for r in rows]
    lcs = _u8(lcs)
    ptrs = (u64p * 4)(*[_p(r, C.c_uint64) for r in rows])
    h = C.c_void_p()
    _check(L.kbo_index_from_parts(k, n_sets, n_kmers, ptrs, _p(lcs, C.c_uint8), device, C.byref(h)))
    return Index(h.value)


def serialize_sbwt(outfile_prefix, index):
    """index::serialize_sbwt (index.rs:128-153): writes `<prefix>.sbwt` and `<prefix>.lcs` (layout: include/kbo_b200.h)."""
    _check(load_library().kbo_index_serialize(index._h, os.fsencode(outfile_prefix)))


def load_sbwt(index_prefix, device=0):
    """index::load_sbwt (index.rs:195-212): reads `<prefix>.sbwt` / `<prefix>.lcs` written by serialize_sbwt."""
    h = C.c_void_p()
    _check(load_library().kbo_index_load(os.fsencode(index_prefix), device, C.byref(h)))
    return Index(h.value)


# -------------------------------------------------------------------------------- index.rs ---
def query_sbwt(query, index, intervals=True):
    """index::query_sbwt (index.rs:243-256): (d, l, r) arrays, one entry per query base."""
    q = _u8(query)
    n = len(q)
    d = np.zeros(max(n, 1), dtype=np.uint64)
    l = np.zeros(max(n, 1), dtype=np.uint64) if intervals else None
    r = np.zeros(max(n, 1), dtype=np.uint64) if intervals else None
    _check(load_library().kbo_query_sbwt(index._h, _p(q, C.c_uint8), n, _p(d, C.c_uint64),
                                         _p(l, C.c_uint64) if intervals else None,
                                         _p(r, C.c_uint64) if intervals else None))
    return (d[:n], l[:n], r[:n]) if intervals else (d[:n], None, None)


def query_sbwt_batch(queries, index, intervals=True):
    concat, offsets = csr(queries)
    n = len(concat)
    d = np.zeros(max(n, 1), dtype=np.uint8)
    l = np.zeros(max(n, 1), dtype=np.uint32) if intervals else None
    r = np.zeros(max(n, 1), dtype=np.uint32) if intervals else None
    _check(load_library().kbo_query_sbwt_batch_compact(index._h, _p(concat, C.c_uint8), _p(offsets, C.c_uint64),
                                                       len(queries), _p(d, C.c_uint8),
                                                       _p(l, C.c_uint32) if intervals else None,
                                                       _p(r, C.c_uint32) if intervals else None))
    return d[:n], (l[:n] if intervals else None), (r[:n] if intervals else None), offsets


# -------------------------------------------------------------------------- derandomize.rs ---
def log_rm_max_cdf(t, alphabet_size, n_kmers):
    out = C.c_double(0)
    _check(load_library().kbo_log_rm_max_cdf(t, alphabet_size, n_kmers, C.byref(out)))
    return out.value


def random_match_threshold(k, n_kmers, alphabet_size, max_error_prob):
    out = C.c_uint64(0)
    _check(load_library().kbo_random_match_threshold(k, n_kmers, alphabet_size, max_error_prob, C.byref(out)))
    return out.value


def derandomize_ms_vec(noisy_ms, k, threshold, device=0):
    ms = np.ascontiguousarray(noisy_ms, dtype=np.uint64)
    out = np.zeros(max(len(ms), 1), dtype=np.int64)
    _check(load_library().kbo_derandomize_ms_vec(_p(ms, C.c_uint64), len(ms), k, threshold, _p(out, C.c_int64),
                                                 device))
    return out[:len(ms)]


# ---------------------------------------------------------------------------- translate.rs ---
def translate_ms_vec(derand_ms, k, threshold, device=0):
    d = np.ascontiguousarray(derand_ms, dtype=np.int64)
    out = np.zeros(max(len(d), 1), dtype=np.uint8)
    _check(load_library().kbo_translate_ms_vec(_p(d, C.c_int64), len(d), k, threshold, _p(out, C.c_uint8), device))
    return out[:len(d)].tobytes()


# ------------------------------------------------------------------------------- format.rs ---
def run_lengths_gapped(aln, max_gap_len):
    a = _u8(aln)
    cap = len(a) + 1
    buf = (RleC * cap)()
    n = C.c_uint64(0)
    _check(load_library().kbo_run_lengths_gapped(_p(a, C.c_uint8), len(a), max_gap_len, buf, cap, C.byref(n)))
    return [RLE(*[int(getattr(buf[i], f)) for f, _ in RleC._fields_]) for i in range(n.value)]


def run_lengths(aln):
    return run_lengths_gapped(aln, 0)


def relative_to_ref(ref_seq, alignment):
    r, a = _u8(ref_seq), _u8(alignment)
    n = min(len(r), len(a))
    out = np.zeros(max(n, 1), dtype=np.uint8)
    _check(load_library().kbo_relative_to_ref(_p(r, C.c_uint8), _p(a, C.c_uint8), n, _p(out, C.c_uint8)))
    return out[:n].tobytes()


# ---------------------------------------------------------------------------------- lib.rs ---
def matches(query_seq, index, match_opts=None):
    """kbo::matches (lib.rs:612-628) -> alignment characters as bytes ('M', '-', 'X', 'R')."""
    o = match_opts or MatchOpts()
    q = _u8(query_seq)
    out = np.zeros(max(len(q), 1), dtype=np.uint8)
    _check(load_library().kbo_matches(index._h, _p(q, C.c_uint8), len(q), o.max_error_prob, _p(out, C.c_uint8)))
    return out[:len(q)].tobytes()


def matches_batch(queries, index, match_opts=None):
    """matches() for many queries in one launch sequence; returns a list of bytes."""
    o = match_opts or MatchOpts()
    concat, offsets = csr(queries)
    out = matches_csr(concat, offsets, index, o.max_error_prob)
    return [out[int(offsets[i]):int(offsets[i + 1])].tobytes() for i in range(len(queries))]


def matches_csr(concat, offsets, index, max_error_prob=0.0000001, out=None):
    """CSR form used by bench.py: host arrays in, host array out (copies inside the call)."""
    if out is None:
        out = np.zeros(max(len(concat), 1), dtype=np.uint8)
    _check(load_library().kbo_matches_batch(index._h, _p(concat, C.c_uint8), _p(offsets, C.c_uint64),
                                            len(offsets) - 1, max_error_prob, _p(out, C.c_uint8)))
    return out


def find(query_seq, index, find_opts=None):
    """kbo::find (lib.rs:808-821) -> list of RLE."""
    return find_batch([query_seq], index, find_opts)[0]


def find_batch(queries, index, find_opts=None):
    o = find_opts or FindOpts()
    concat, offsets = csr(queries)
    cap = len(concat) + 1
    buf = (RleC * cap)()
    roff = np.zeros(len(queries) + 1, dtype=np.uint64)
    _check(load_library().kbo_find_batch(index._h, _p(concat, C.c_uint8), _p(offsets, C.c_uint64), len(queries),
                                         o.max_error_prob, o.max_gap_len, buf, cap, _p(roff, C.c_uint64)))
    res = []
    for i in range(len(queries)):
        res.append([RLE(*[int(getattr(buf[j], f)) for f, _ in RleC._fields_])
                    for j in range(int(roff[i]), int(roff[i + 1]))])
    return res


class PinnedBytes:
    """Page-locked host bytes from kbo_alloc_pinned (cudaHostAlloc); `.array` is a numpy view.  Host-buffer entry
    points copy from / to such memory asynchronously and at full PCIe rate."""

    def __init__(self, nbytes):
        p = C.c_void_p()
        _check(load_library().kbo_alloc_pinned(int(nbytes), C.byref(p)))
        self._p = p
        self.array = np.ctypeslib.as_array((C.c_uint8 * int(nbytes)).from_address(p.value))

    def __del__(self):
        try:
            if self._p:
                load_library().kbo_free_pinned(self._p)
                self._p = None
        except Exception:
            pass


class FindBuffers:
    """Reusable output buffers for find_csr / find_submit (avoids reallocating per call in a timed loop)."""

    def __init__(self, n_queries, cap=None, pinned=False):
        self.cap = cap or (8 * n_queries + 1024)
        if pinned:  # page-locked outputs: the last kernel of kbo_find_batch writes records and offsets into them directly
            self._pin = PinnedBytes(self.cap * C.sizeof(RleC))
            self.rle = (RleC * self.cap).from_address(self._pin._p.value)
            self._pin_off = PinnedBytes((n_queries + 1) * 8)
            self.rle_offsets = self._pin_off.array.view(np.uint64)
        else:
            self.rle = (RleC * self.cap)()
            self.rle_offsets = np.zeros(n_queries + 1, dtype=np.uint64)


class FindJob:
    """One kbo::find batch in flight (kbo_find_batch_submit); wait() returns the number of RLE records."""

    def __init__(self, handle, buffers, keep):
        self._h, self.buffers, self._keep = handle, buffers, keep

    def wait(self):
        n = C.c_uint64(0)
        h, self._h = self._h, None
        _check(load_library().kbo_job_wait(h, C.byref(n)))
        self._keep = None
        return int(n.value)


def find_submit(concat, offsets, index, find_opts=None, buffers=None):
    """kbo_find_batch_submit: enqueue one CSR batch (ideally page-locked arrays in, FindBuffers(pinned=True) out) and
    return a FindJob; several jobs may be in flight from one host thread."""
    o = find_opts or FindOpts()
    nq = len(offsets) - 1
    buf = buffers or FindBuffers(nq, pinned=True)
    h = C.c_void_p()
    _check(load_library().kbo_find_batch_submit(index._h, _p(concat, C.c_uint8), _p(offsets, C.c_uint64), nq,
                                                o.max_error_prob, o.max_gap_len, buf.rle, buf.cap,
                                                _p(buf.rle_offsets, C.c_uint64), C.byref(h)))
    return FindJob(h, buf, (concat, offsets))


def find_csr(concat, offsets, index, find_opts=None, buffers=None):
    """kbo::find for a CSR batch with host arrays in (e.g. pinned) and RLE records out.
    Returns (FindBuffers, n_rle): RLEs of query i are buffers.rle[rle_offsets[i]:rle_offsets[i+1]]."""
    o = find_opts or FindOpts()
    nq = len(offsets) - 1
    buf = buffers or FindBuffers(nq)
    rc = load_library().kbo_find_batch(index._h, _p(concat, C.c_uint8), _p(offsets, C.c_uint64), nq,
                                       o.max_error_prob, o.max_gap_len, buf.rle, buf.cap, _p(buf.rle_offsets, C.c_uint64))
    if rc == 11:  # KBO_ERR_BUFFER_TOO_SMALL: grow once and retry
        buf = FindBuffers(nq, cap=int(buf.rle_offsets[nq]) + 1024)
        rc = load_library().kbo_find_batch(index._h, _p(concat, C.c_uint8), _p(offsets, C.c_uint64), nq,
                                           o.max_error_prob, o.max_gap_len, buf.rle, buf.cap,
                                           _p(buf.rle_offsets, C.c_uint64))
    _check(rc)
    return buf, int(buf.rle_offsets[nq])


# ------------------------------------------------------------------------------- multi-GPU ---
class Context:
    """kbo_ctx: several GPUs driven by one process (one worker thread per device)."""

    def __init__(self, n_gpus=0, devices=None):
        h = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices) if devices else None
        _check(load_library().kbo_ctx_create(len(devices) if devices else n_gpus, devs, C.byref(h)))
        self._h = h
        self.n_gpus = load_library().kbo_ctx_n_gpus(h)

    def close(self):
        if self._h:
            load_library().kbo_ctx_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def build(self, seq_data, build_opts=None):
        """kbo::build on every device of the context -> IndexSet."""
        o = build_opts or BuildOpts()
        seqs = [_u8(s) for s in seq_data]
        ptrs = (u8p * max(len(seqs), 1))(*[_p(s, C.c_uint8) for s in seqs])
        lens = np.array([len(s) for s in seqs], dtype=np.uint64)
        co = _build_opts_c(o)
        h = C.c_void_p()
        _check(load_library().kbo_index_set_build(self._h, ptrs, _p(lens, C.c_uint64), len(seqs), C.byref(co), C.byref(h)))
        return IndexSet(h, self)


class IndexSet:
    """kbo_index_set: one replica of an index per device of a Context."""

    def __init__(self, handle, ctx):
        self._h, self.ctx = handle, ctx
        first = load_library().kbo_index_set_get(handle, 0)
        self.k = load_library().kbo_index_k(first)
        self.n_sets = load_library().kbo_index_n_sets(first)
        self.n_kmers = load_library().kbo_index_n_kmers(first)

    def close(self):
        if self._h:
            load_library().kbo_index_set_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def matches_csr(self, concat, offsets, max_error_prob=0.0000001, out=None):
        if out is None:
            out = np.zeros(max(len(concat), 1), dtype=np.uint8)
        _check(load_library().kbo_matches_batch_multi(self._h, _p(concat, C.c_uint8), _p(offsets, C.c_uint64),
                                                      len(offsets) - 1, max_error_prob, _p(out, C.c_uint8)))
        return out

    def find_csr(self, concat, offsets, find_opts=None, buffers=None):
        """Returns (FindBuffers, n_rle) like find_csr."""
        o = find_opts or FindOpts()
        nq = len(offsets) - 1
        buf = buffers or FindBuffers(nq, pinned=True)
        rc = load_library().kbo_find_batch_multi(self._h, _p(concat, C.c_uint8), _p(offsets, C.c_uint64), nq,
                                                 o.max_error_prob, o.max_gap_len, buf.rle, buf.cap,
                                                 _p(buf.rle_offsets, C.c_uint64))
        if rc == 11:
            buf = FindBuffers(nq, cap=int(buf.rle_offsets[nq]) + 1024, pinned=True)
            rc = load_library().kbo_find_batch_multi(self._h, _p(concat, C.c_uint8), _p(offsets, C.c_uint64), nq,
                                                     o.max_error_prob, o.max_gap_len, buf.rle, buf.cap,
                                                     _p(buf.rle_offsets, C.c_uint64))
        _check(rc)
        return buf, int(buf.rle_offsets[nq])


def _build_opts_c(o):
    return BuildOptsC(o.k, int(o.add_revcomp), o.num_threads, o.prefix_precalc, int(o.build_select), o.mem_gb,
                      int(o.dedup_batches), o.temp_dir.encode() if o.temp_dir else None)


def call(query_index, ref_seq, call_opts=None, ref_index=None):
    """kbo::call (lib.rs:547-573) -> list of Variant.  ref_index: the index of ref_seq built once by the caller
    (kbo_call_with_ref) instead of per call as the reference does (lib.rs:553); same result."""
    o = call_opts or CallOpts()
    r = _u8(ref_seq)
    L = load_library()
    co = _build_opts_c(o.sbwt_build_opts)
    # output capacity: a first guess, then (KBO_ERR_BUFFER_TOO_SMALL) the worst case
    for cap, capc in ((len(r) // 32 + 4096, len(r) // 4 + 65536), (len(r) + 1, 4 * len(r) + 64)):
        pos = np.empty(cap, dtype=np.uint64)
        ql, rl = np.empty(cap, dtype=np.uint32), np.empty(cap, dtype=np.uint32)
        qc, rc = np.empty(capc, dtype=np.uint8), np.empty(capc, dtype=np.uint8)
        n = C.c_uint64(0)
        tail = (_p(pos, C.c_uint64), _p(ql, C.c_uint32), _p(rl, C.c_uint32), _p(qc, C.c_uint8), _p(rc, C.c_uint8), cap,
                capc, C.byref(n))
        if ref_index is None:
            status = L.kbo_call(query_index._h, _p(r, C.c_uint8), len(r), o.max_error_prob, C.byref(co), *tail)
        else:
            status = L.kbo_call_with_ref(query_index._h, ref_index._h, _p(r, C.c_uint8), len(r), o.max_error_prob, *tail)
        if status != 11:
            break
    _check(status)
    nv = int(n.value)
    qends, rends = np.cumsum(ql[:nv], dtype=np.int64), np.cumsum(rl[:nv], dtype=np.int64)
    qb, rb = qc[:int(qends[-1]) if nv else 0].tobytes(), rc[:int(rends[-1]) if nv else 0].tobytes()
    out, qo, ro = [], 0, 0
    for p_, qe, re in zip(pos[:nv].tolist(), qends.tolist(), rends.tolist()):
        out.append(Variant(p_, qb[qo:qe], rb[ro:re]))
        qo, ro = qe, re
    return out


def map(ref_seq, query_index, map_opts=None, ref_index=None):
    """kbo::map (lib.rs:720-761).  ref_index: see call()."""
    o = map_opts or MapOpts()
    r = _u8(ref_seq)
    out = np.empty(max(len(r), 1), dtype=np.uint8)
    if ref_index is None:
        co = _build_opts_c(o.sbwt_build_opts)
        _check(load_library().kbo_map(query_index._h, _p(r, C.c_uint8), len(r), o.max_error_prob, int(o.fill_gaps),
                                      int(o.call_variants), int(o.format), C.byref(co), _p(out, C.c_uint8)))
    else:
        _check(load_library().kbo_map_with_ref(query_index._h, ref_index._h, _p(r, C.c_uint8), len(r), o.max_error_prob,
                                               int(o.fill_gaps), int(o.call_variants), int(o.format), _p(out, C.c_uint8)))
    return out[:len(r)].tobytes()


def map_unrefined(ref_seq, query_index, max_error_prob=0.0000001, format=True):
    """kbo::map with fill_gaps = call_variants = false (lib.rs:726-738, 756-760), entirely on the device."""
    r = _u8(ref_seq)
    out = np.zeros(max(len(r), 1), dtype=np.uint8)
    _check(load_library().kbo_map_unrefined(query_index._h, _p(r, C.c_uint8), len(r), max_error_prob, int(format),
                                            _p(out, C.c_uint8)))
    return out[:len(r)].tobytes()


# ---------------------------------------------------------------------------- instrumentation ---
def set_profile_counters(enabled):
    _check(load_library().kbo_set_profile_counters(int(enabled)))


def set_chunk_len(chunk_len):
    _check(load_library().kbo_set_chunk_len(int(chunk_len)))


def set_device_parts(parts):
    _check(load_library().kbo_set_device_parts(int(parts)))


def set_pipeline_parts(parts):
    _check(load_library().kbo_set_pipeline_parts(int(parts)))


def set_host_builder(enabled):
    _check(load_library().kbo_set_host_builder(int(enabled)))


def set_prefix_len(p):
    """Depth of the prefix-state table of indexes built afterwards (0 = the default, 10; at most 14)."""
    _check(load_library().kbo_set_prefix_len(int(p)))


def set_prefix_table(enabled):
    """Indexes built after this call do / do not carry the prefix-state table that shortens K1's chunk warm-up."""
    _check(load_library().kbo_set_prefix_table(int(bool(enabled))))


TUNE_CHUNK_LEN, TUNE_PIPELINE_PARTS, TUNE_DEVICE_PARTS, TUNE_MS_FLAGS, TUNE_REFINE_THREADS = range(5)


def set_refine_threads(n):
    """Host threads that bridge gaps in map() (0 = hardware concurrency, at most 16)."""
    _check(load_library().kbo_set_refine_threads(int(n)))


def set_device_refine(enabled):
    """fill_gaps / access_kmer of map() and call() on the device (default) or on the host (comparison runs)."""
    _check(load_library().kbo_set_device_refine(int(bool(enabled))))


def set_index_tuning(index, key, value):
    """A tuning knob for ONE index (value -1: follow the process-wide default again)."""
    _check(load_library().kbo_index_set_tuning(index._h, int(key), int(value)))


def set_rank2(enabled):
    """Indexes built after this call do / do not carry the rank2 rows (two bases per probe in K1)."""
    _check(load_library().kbo_set_rank2(int(bool(enabled))))


def set_l2_persist(enabled):
    """Streams created after this call do / do not mark the index arrays as persisting in L2."""
    _check(load_library().kbo_set_l2_persist(int(bool(enabled))))


def set_ms_flags(flags):
    _check(load_library().kbo_set_ms_flags(int(flags)))


def kernel_launch_count():
    return int(load_library().kbo_kernel_launch_count())


def set_kernel_timing(enabled):
    _check(load_library().kbo_set_kernel_timing(int(enabled)))


def collect_kernel_times(index, stream=0):
    """After synchronising `stream`: ({'pack','ms','derand_translate'} summed ms, number of timed calls)."""
    sums = (C.c_double * 3)()
    n = C.c_uint64(0)
    _check(load_library().kbo_collect_kernel_times(index._h, C.c_void_p(stream), sums, C.byref(n)))
    return {"pack": sums[0], "ms": sums[1], "derand_translate": sums[2]}, int(n.value)


def measure_random_sector_rate(buffer_bytes, dependent, device=0):
    out = C.c_double(0)
    _check(load_library().kbo_measure_random_sector_rate(device, buffer_bytes, int(dependent), C.byref(out)))
    return out.value


def find_device(index, d_concat_ptr, d_offsets_ptr, host_offsets, d_rle_ptr, rle_cap, d_rle_offsets_ptr,
                max_error_prob=0.0000001, max_gap_len=0, stream=0):
    """kbo_find_batch_device: raw device pointers (ints); async on `stream`."""
    off = np.ascontiguousarray(host_offsets, dtype=np.uint64)
    _check(load_library().kbo_find_batch_device(index._h, C.c_void_p(d_concat_ptr), C.c_void_p(d_offsets_ptr),
                                                _p(off, C.c_uint64), len(off) - 1, max_error_prob, max_gap_len,
                                                C.c_void_p(d_rle_ptr), rle_cap, C.c_void_p(d_rle_offsets_ptr),
                                                C.c_void_p(stream)))


def matches_device(index, d_concat_ptr, d_offsets_ptr, host_offsets, d_out_ptr, max_error_prob=0.0000001, stream=0):
    """kbo_matches_batch_device: raw device pointers (ints) on the index's device; async on `stream`."""
    off = np.ascontiguousarray(host_offsets, dtype=np.uint64)
    _check(load_library().kbo_matches_batch_device(index._h, C.c_void_p(d_concat_ptr), C.c_void_p(d_offsets_ptr),
                                                   _p(off, C.c_uint64), len(off) - 1, max_error_prob,
                                                   C.c_void_p(d_out_ptr), C.c_void_p(stream)))


def device_count():
    n = C.c_int(0)
    rc = load_library().kbo_device_count(C.byref(n))
    return n.value if rc == 0 else 0
