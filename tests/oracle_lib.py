"""ctypes front-end of the CPU oracle (oracle/libkbo_oracle.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never by kbo_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ORACLE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
_SO = os.path.join(_ORACLE_DIR, "libkbo_oracle.so")


def build_oracle(force=False):
    srcs = [os.path.join(_ORACLE_DIR, f) for f in ("kbo_oracle.cpp", "kbo_oracle_capi.cpp", "kbo_oracle.hpp")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _ORACLE_DIR, "-s"] + (["-B"] if force else []))
    return _SO


class OraclePanic(Exception):
    """The reference would have panicked on these inputs."""


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(_SO)
        u8p, u64p, i64p, u32p = (C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_int64),
                                 C.POINTER(C.c_uint32))
        L.kbo_oracle_last_error.restype = C.c_char_p
        L.kbo_oracle_build.restype = C.c_void_p
        L.kbo_oracle_build.argtypes = [C.POINTER(u8p), u64p, C.c_uint64, C.c_int, C.c_int]
        L.kbo_oracle_free.argtypes = [C.c_void_p]
        L.kbo_oracle_k.argtypes = [C.c_void_p]
        for f in ("kbo_oracle_n_sets", "kbo_oracle_n_kmers"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_void_p]
        L.kbo_oracle_C.argtypes = [C.c_void_p, u64p]
        L.kbo_oracle_rows.argtypes = [C.c_void_p, u64p, u64p, u64p, u64p]
        L.kbo_oracle_lcs.argtypes = [C.c_void_p, u8p]
        L.kbo_oracle_access_kmer.argtypes = [C.c_void_p, C.c_uint64, u8p]
        L.kbo_oracle_search.argtypes = [C.c_void_p, u8p, C.c_uint64, u64p, u64p]
        L.kbo_oracle_query_sbwt.argtypes = [C.c_void_p, u8p, C.c_uint64, u64p, u64p, u64p]
        L.kbo_oracle_log_rm_max_cdf.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_double)]
        L.kbo_oracle_random_match_threshold.restype = C.c_int64
        L.kbo_oracle_random_match_threshold.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_double]
        L.kbo_oracle_derandomize_ms_val.argtypes = [C.c_uint64, C.c_int64, C.c_uint64, C.c_uint64, i64p]
        L.kbo_oracle_derandomize_ms_vec.argtypes = [u64p, C.c_uint64, C.c_uint64, C.c_uint64, i64p]
        L.kbo_oracle_translate_ms_val.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_uint64, C.c_char_p]
        L.kbo_oracle_translate_ms_vec.argtypes = [i64p, C.c_uint64, C.c_uint64, C.c_uint64, u8p]
        L.kbo_oracle_matches.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_double, u8p]
        L.kbo_oracle_run_lengths_gapped.restype = C.c_int64
        L.kbo_oracle_run_lengths_gapped.argtypes = [u8p, C.c_uint64, C.c_uint64, u64p, C.c_uint64]
        L.kbo_oracle_find.restype = C.c_int64
        L.kbo_oracle_find.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_double, C.c_uint64, u64p, C.c_uint64]
        L.kbo_oracle_relative_to_ref.argtypes = [u8p, C.c_uint64, u8p, C.c_uint64, u8p]
        L.kbo_oracle_call.restype = C.c_int64
        L.kbo_oracle_call.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_double, C.c_int, C.c_int, u64p, u32p, u32p,
                                      u8p, u8p, C.c_uint64, C.c_uint64]
        L.kbo_oracle_call_variants.restype = C.c_int64
        L.kbo_oracle_call_variants.argtypes = [C.c_void_p, C.c_void_p, u8p, C.c_uint64, C.c_double, u64p, u32p, u32p,
                                               u8p, u8p, C.c_uint64, C.c_uint64]
        L.kbo_oracle_add_variants.argtypes = [u8p, C.c_uint64, C.c_uint64, u64p, u32p, u32p, u8p, u8p, u8p]
        L.kbo_oracle_fill_gaps.argtypes = [C.c_void_p, u8p, u8p, C.c_uint64, C.c_uint64, C.c_double, u8p]
        L.kbo_oracle_nearest_unique_context.restype = C.c_int64
        L.kbo_oracle_nearest_unique_context.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_uint64, C.c_uint64, u64p, u8p]
        L.kbo_oracle_left_extend_kmer.restype = C.c_int64
        L.kbo_oracle_left_extend_kmer.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_uint64, u8p, C.c_uint64]
        L.kbo_oracle_left_extend_over_gap.restype = C.c_int64
        L.kbo_oracle_left_extend_over_gap.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                                      C.c_uint64, C.c_uint64, u8p, C.c_uint64]
        L.kbo_oracle_map.restype = C.c_int64
        L.kbo_oracle_map.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, u8p]
        L.kbo_oracle_matches_batch.restype = C.c_double
        L.kbo_oracle_matches_batch.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, C.c_double, u8p, C.c_int, u64p]
        L.kbo_oracle_find_batch.restype = C.c_double
        L.kbo_oracle_find_batch.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, C.c_double, C.c_uint64, C.c_int, u64p,
                                            u64p]
        L.kbo_oracle_hardware_concurrency.restype = C.c_int
        _lib = L
    return _lib


def _p(arr, ty):
    return arr.ctypes.data_as(C.POINTER(ty))


def _u8(x):
    if isinstance(x, (bytes, bytearray)):
        return np.frombuffer(bytes(x), dtype=np.uint8).copy()
    if isinstance(x, str):
        return np.frombuffer(x.encode(), dtype=np.uint8).copy()
    return np.ascontiguousarray(x, dtype=np.uint8)


def _check(rc):
    if rc == -1:
        raise OraclePanic(lib().kbo_oracle_last_error().decode())
    if rc == -2:
        raise RuntimeError("oracle: output buffer too small")
    return rc


class Variant(tuple):
    """(query_pos, query_chars: bytes, ref_chars: bytes)"""


def _unpack_variants(n, pos, ql, rl, qc, rc):
    out, qo, ro = [], 0, 0
    for i in range(n):
        out.append((int(pos[i]), bytes(qc[qo:qo + ql[i]]), bytes(rc[ro:ro + rl[i]])))
        qo += int(ql[i])
        ro += int(rl[i])
    return out


class OracleIndex:
    """CPU oracle of (SbwtIndexVariant::SubsetMatrix, LcsArray) built like index.rs:56-99."""

    def __init__(self, seqs, k=31, add_revcomp=False):
        L = lib()
        self._seqs = [_u8(s) for s in seqs]
        n = len(self._seqs)
        ptrs = (C.POINTER(C.c_uint8) * max(n, 1))(*[_p(s, C.c_uint8) for s in self._seqs])
        lens = np.array([len(s) for s in self._seqs], dtype=np.uint64)
        self.h = L.kbo_oracle_build(ptrs, _p(lens, C.c_uint64), n, k, int(add_revcomp))
        if not self.h:
            raise OraclePanic(L.kbo_oracle_last_error().decode())
        self.k = L.kbo_oracle_k(self.h)
        self.n_sets = L.kbo_oracle_n_sets(self.h)
        self.n_kmers = L.kbo_oracle_n_kmers(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().kbo_oracle_free(self.h)
            self.h = None

    def C(self):
        out = np.zeros(4, dtype=np.uint64)
        lib().kbo_oracle_C(self.h, _p(out, C.c_uint64))
        return out

    def rows(self):
        nw = (self.n_sets + 63) // 64
        rows = [np.zeros(nw, dtype=np.uint64) for _ in range(4)]
        lib().kbo_oracle_rows(self.h, *[_p(r, C.c_uint64) for r in rows])
        return rows

    def lcs(self):
        out = np.zeros(self.n_sets, dtype=np.uint8)
        lib().kbo_oracle_lcs(self.h, _p(out, C.c_uint8))
        return out

    def access_kmer(self, colex):
        out = np.zeros(self.k, dtype=np.uint8)
        _check(lib().kbo_oracle_access_kmer(self.h, colex, _p(out, C.c_uint8)))
        return out.tobytes()

    def search(self, pat):
        p = _u8(pat)
        l, r = C.c_uint64(0), C.c_uint64(0)
        ok = lib().kbo_oracle_search(self.h, _p(p, C.c_uint8), len(p), C.byref(l), C.byref(r))
        return (l.value, r.value) if ok else None

    def query_sbwt(self, q):
        q = _u8(q)
        n = len(q)
        d, l, r = (np.zeros(max(n, 1), dtype=np.uint64) for _ in range(3))
        _check(lib().kbo_oracle_query_sbwt(self.h, _p(q, C.c_uint8), n, _p(d, C.c_uint64), _p(l, C.c_uint64),
                                           _p(r, C.c_uint64)))
        return d[:n], l[:n], r[:n]

    def matches(self, q, max_error_prob=1e-7):
        q = _u8(q)
        out = np.zeros(max(len(q), 1), dtype=np.uint8)
        _check(lib().kbo_oracle_matches(self.h, _p(q, C.c_uint8), len(q), max_error_prob, _p(out, C.c_uint8)))
        return out[:len(q)].tobytes()

    def find(self, q, max_error_prob=1e-7, max_gap_len=0):
        q = _u8(q)
        cap = len(q) + 1
        out = np.zeros(7 * cap, dtype=np.uint64)
        n = _check(lib().kbo_oracle_find(self.h, _p(q, C.c_uint8), len(q), max_error_prob, max_gap_len,
                                         _p(out, C.c_uint64), cap))
        return [tuple(int(x) for x in out[7 * i:7 * i + 7]) for i in range(n)]

    def call(self, ref_seq, max_error_prob=1e-7, build_k=31, build_revcomp=False):
        r = _u8(ref_seq)
        cap, capc = len(r) + 1, 4 * len(r) + 64
        pos = np.zeros(cap, dtype=np.uint64)
        ql, rl = np.zeros(cap, dtype=np.uint32), np.zeros(cap, dtype=np.uint32)
        qc, rc = np.zeros(capc, dtype=np.uint8), np.zeros(capc, dtype=np.uint8)
        n = _check(lib().kbo_oracle_call(self.h, _p(r, C.c_uint8), len(r), max_error_prob, build_k, int(build_revcomp),
                                         _p(pos, C.c_uint64), _p(ql, C.c_uint32), _p(rl, C.c_uint32),
                                         _p(qc, C.c_uint8), _p(rc, C.c_uint8), cap, capc))
        return _unpack_variants(n, pos, ql, rl, qc, rc)

    def fill_gaps(self, translation, ref_seq, threshold, max_error_prob):
        t, r = _u8(translation), _u8(ref_seq)
        out = np.zeros(len(r), dtype=np.uint8)
        _check(lib().kbo_oracle_fill_gaps(self.h, _p(t, C.c_uint8), _p(r, C.c_uint8), len(r), threshold,
                                          max_error_prob, _p(out, C.c_uint8)))
        return out.tobytes()

    def nearest_unique_context(self, ref_seq, start, end):
        r = _u8(ref_seq)
        out = np.zeros(self.k, dtype=np.uint8)
        idx = C.c_uint64(0)
        n = _check(lib().kbo_oracle_nearest_unique_context(self.h, _p(r, C.c_uint8), len(r), start, end,
                                                           C.byref(idx), _p(out, C.c_uint8)))
        return idx.value, out[:n].tobytes()

    def left_extend_kmer(self, kmer, max_ext):
        km = _u8(kmer)
        cap = len(km) + max_ext + 1
        out = np.zeros(cap, dtype=np.uint8)
        n = _check(lib().kbo_oracle_left_extend_kmer(self.h, _p(km, C.c_uint8), len(km), max_ext, _p(out, C.c_uint8),
                                                     cap))
        return out[:n].tobytes()

    def left_extend_over_gap(self, ref_seq, left_req, right_req, gap_start, gap_end, radius):
        r = _u8(ref_seq)
        cap = len(r) + self.k + 1
        out = np.zeros(cap, dtype=np.uint8)
        n = _check(lib().kbo_oracle_left_extend_over_gap(self.h, _p(r, C.c_uint8), len(r), left_req, right_req,
                                                         gap_start, gap_end, radius, _p(out, C.c_uint8), cap))
        return out[:n].tobytes()

    def map(self, ref_seq, max_error_prob=1e-7, fill_gaps=True, call_variants=True, format=True, build_k=31,
            build_revcomp=False):
        r = _u8(ref_seq)
        out = np.zeros(max(len(r), 1), dtype=np.uint8)
        n = _check(lib().kbo_oracle_map(self.h, _p(r, C.c_uint8), len(r), max_error_prob, int(fill_gaps),
                                        int(call_variants), int(format), build_k, int(build_revcomp),
                                        _p(out, C.c_uint8)))
        return out[:n].tobytes()

    def matches_batch(self, concat, offsets, max_error_prob=1e-7, n_threads=1, want_output=True):
        """Returns (seconds, output bytes or None, checksum)."""
        cc = _u8(concat)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        out = np.zeros(len(cc), dtype=np.uint8) if want_output else None
        cs = C.c_uint64(0)
        secs = lib().kbo_oracle_matches_batch(self.h, _p(cc, C.c_uint8), _p(off, C.c_uint64), len(off) - 1,
                                              max_error_prob, _p(out, C.c_uint8) if want_output else None, n_threads,
                                              C.byref(cs))
        if secs < 0:
            raise OraclePanic("matches_batch: a query panicked")
        return secs, out, cs.value


def find_batch_timed(ix, concat, offsets, max_error_prob=1e-7, max_gap_len=0, n_threads=1):
    """kbo::find over a CSR batch with n_threads host threads: (seconds, n_rle, checksum)."""
    cc = _u8(concat)
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    cs, nr = C.c_uint64(0), C.c_uint64(0)
    secs = lib().kbo_oracle_find_batch(ix.h, _p(cc, C.c_uint8), _p(off, C.c_uint64), len(off) - 1, max_error_prob,
                                       max_gap_len, n_threads, C.byref(nr), C.byref(cs))
    if secs < 0:
        raise OraclePanic("find_batch: a query panicked")
    return secs, nr.value, cs.value


def hardware_concurrency():
    return int(lib().kbo_oracle_hardware_concurrency())


def call_variants(ix_ref, ix_query, query, max_error_prob):
    q = _u8(query)
    cap, capc = len(q) + 1, 4 * len(q) + 64
    pos = np.zeros(cap, dtype=np.uint64)
    ql, rl = np.zeros(cap, dtype=np.uint32), np.zeros(cap, dtype=np.uint32)
    qc, rc = np.zeros(capc, dtype=np.uint8), np.zeros(capc, dtype=np.uint8)
    n = _check(lib().kbo_oracle_call_variants(ix_ref.h, ix_query.h, _p(q, C.c_uint8), len(q), max_error_prob,
                                              _p(pos, C.c_uint64), _p(ql, C.c_uint32), _p(rl, C.c_uint32),
                                              _p(qc, C.c_uint8), _p(rc, C.c_uint8), cap, capc))
    return _unpack_variants(n, pos, ql, rl, qc, rc)


def add_variants(translation, variants):
    t = _u8(translation)
    n = len(variants)
    pos = np.array([v[0] for v in variants] + [0], dtype=np.uint64)
    ql = np.array([len(v[1]) for v in variants] + [0], dtype=np.uint32)
    rl = np.array([len(v[2]) for v in variants] + [0], dtype=np.uint32)
    qc = _u8(b"".join(v[1] for v in variants) + b"\0")
    rc = _u8(b"".join(v[2] for v in variants) + b"\0")
    out = np.zeros(len(t), dtype=np.uint8)
    _check(lib().kbo_oracle_add_variants(_p(t, C.c_uint8), len(t), n, _p(pos, C.c_uint64), _p(ql, C.c_uint32),
                                         _p(rl, C.c_uint32), _p(qc, C.c_uint8), _p(rc, C.c_uint8), _p(out, C.c_uint8)))
    return out.tobytes()


def log_rm_max_cdf(t, s, n):
    out = C.c_double(0)
    _check(lib().kbo_oracle_log_rm_max_cdf(t, s, n, C.byref(out)))
    return out.value


def random_match_threshold(k, n_kmers, s, p):
    return _check(lib().kbo_oracle_random_match_threshold(k, n_kmers, s, p))


def derandomize_ms_val(cur, nxt, thr, k):
    out = C.c_int64(0)
    _check(lib().kbo_oracle_derandomize_ms_val(cur, nxt, thr, k, C.byref(out)))
    return out.value


def derandomize_ms_vec(ms, k, thr):
    ms = np.ascontiguousarray(ms, dtype=np.uint64)
    out = np.zeros(max(len(ms), 1), dtype=np.int64)
    _check(lib().kbo_oracle_derandomize_ms_vec(_p(ms, C.c_uint64), len(ms), k, thr, _p(out, C.c_int64)))
    return out[:len(ms)]


def translate_ms_val(cur, nxt, prev, thr):
    buf = C.create_string_buffer(2)
    _check(lib().kbo_oracle_translate_ms_val(cur, nxt, prev, thr, buf))
    return (buf.raw[0:1].decode(), buf.raw[1:2].decode())


def translate_ms_vec(derand, k, thr):
    d = np.ascontiguousarray(derand, dtype=np.int64)
    out = np.zeros(max(len(d), 1), dtype=np.uint8)
    _check(lib().kbo_oracle_translate_ms_vec(_p(d, C.c_int64), len(d), k, thr, _p(out, C.c_uint8)))
    return out[:len(d)].tobytes()


def run_lengths_gapped(aln, max_gap_len):
    a = _u8(aln)
    cap = len(a) + 1
    out = np.zeros(7 * cap, dtype=np.uint64)
    n = _check(lib().kbo_oracle_run_lengths_gapped(_p(a, C.c_uint8), len(a), max_gap_len, _p(out, C.c_uint64), cap))
    return [tuple(int(x) for x in out[7 * i:7 * i + 7]) for i in range(n)]


def run_lengths(aln):
    return run_lengths_gapped(aln, 0)


def relative_to_ref(ref_seq, aln):
    r, a = _u8(ref_seq), _u8(aln)
    out = np.zeros(max(len(r), 1), dtype=np.uint8)
    n = _check(lib().kbo_oracle_relative_to_ref(_p(r, C.c_uint8), len(r), _p(a, C.c_uint8), len(a), _p(out, C.c_uint8)))
    return out[:n].tobytes()
