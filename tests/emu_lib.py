"""ctypes front-end of tests/emu/libkbo_emu.so: the product's kernels compiled for the
CPU through tests/emu/host_emu.hpp, plus the product's host-side index builder.

TEST INFRASTRUCTURE ONLY (CPU check of kernel logic in a GPU-less container).
The shipped library (kbo_b200/libkbo_b200.so) has no CPU path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
_SO = os.path.join(_DIR, "libkbo_emu.so")
_CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kbo_b200", "csrc")


def build_emu():
    srcs = [os.path.join(_DIR, "emu_driver.cpp"), os.path.join(_DIR, "host_emu.hpp")] + [
        os.path.join(_CSRC, f) for f in ("kernels.cuh", "fused.cuh", "refine.cuh", "host_layout.hpp", "sbwt_host.hpp", "sbwt_host.cpp", "refine_host.hpp",
                                        "refine_host.cpp")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if stale:
        subprocess.check_call(["g++", "-O2", "-std=c++20", "-march=x86-64-v3", "-fPIC", "-shared", "-pthread",
                               "-I", _DIR, "-o", _SO, os.path.join(_DIR, "emu_driver.cpp"),
                               os.path.join(_CSRC, "sbwt_host.cpp"), os.path.join(_CSRC, "refine_host.cpp")])
    return _SO


_lib = None
u8p, u64p, i64p, u32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_int64), C.POINTER(C.c_uint32)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_emu())
        L.emu_host_build.restype = C.c_void_p
        L.emu_host_build.argtypes = [C.POINTER(u8p), u64p, C.c_uint64, C.c_uint32, C.c_int, C.c_uint32, C.c_char_p,
                                     C.c_uint64]
        L.emu_from_parts.restype = C.c_void_p
        L.emu_from_parts.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.POINTER(u64p), u8p]
        L.emu_free.argtypes = [C.c_void_p]
        L.emu_n_sets.restype = C.c_uint64
        L.emu_n_sets.argtypes = [C.c_void_p]
        L.emu_n_kmers.restype = C.c_uint64
        L.emu_n_kmers.argtypes = [C.c_void_p]
        L.emu_k.restype = C.c_uint32
        L.emu_k.argtypes = [C.c_void_p]
        L.emu_export.argtypes = [C.c_void_p, u64p, u64p, u64p, u64p, u8p, u64p]
        L.emu_access_kmer.argtypes = [C.c_void_p, C.c_uint64, u8p]
        L.emu_search.argtypes = [C.c_void_p, u8p, C.c_uint64, u64p, u64p]
        L.emu_query_sbwt_batch.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, C.c_uint32, u8p, u32p, u32p, u64p]
        L.emu_pack.restype = C.c_uint64
        L.emu_pack.argtypes = [u8p, u64p, C.c_uint64, u64p, u32p, u32p, u32p]
        L.emu_matches_batch.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, C.c_uint32, C.c_uint32, u8p]
        L.emu_derand_translate_u8.argtypes = [u8p, C.c_uint64, C.c_uint32, C.c_uint32, u8p]
        L.emu_derandomize_general.argtypes = [u64p, C.c_uint64, C.c_uint32, C.c_uint32, i64p]
        L.emu_translate_i64.argtypes = [i64p, C.c_uint64, C.c_uint32, C.c_uint32, u8p]
        L.emu_call.restype = C.c_int64
        L.emu_call.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, u64p, u32p, u32p, u8p, u8p]
        L.emu_map.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_double, C.c_int, C.c_int, C.c_int,
                              C.c_uint32, C.c_int, u8p]
        L.emu_rle_batch.restype = C.c_uint64
        L.emu_rle_batch.argtypes = [u8p, u64p, C.c_uint64, C.c_uint32, u64p, C.c_uint64, u64p]
        L.emu_set_rank2.argtypes = [C.c_int]
        L.emu_set_device_refine.argtypes = [C.c_int]
        L.emu_set_prefix_len.argtypes = [C.c_uint32]
        L.emu_device_gap_count.restype = C.c_uint64
        L.emu_set_fused.argtypes = [C.c_int, C.c_uint32, C.c_int]
        L.emu_fused_launches.restype = C.c_uint64
        L.emu_fused_tiles.restype = C.c_uint64
        L.emu_tail_extension_count.restype = C.c_uint64
        L.emu_find_batch.restype = C.c_uint64
        L.emu_find_batch.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, C.c_uint32, C.c_uint32, u64p, C.c_uint64, u64p]
        _lib = L
    return _lib


def set_device_refine(on):
    """map / call: fill_gaps and access_kmer through the kernels of refine.cuh (as capi.cu does when the index keeps
    its node keys on the device) instead of refine_host.cpp."""
    lib().emu_set_device_refine(int(bool(on)))


def set_prefix_len(p):
    """Depth of the prefix-state table of indexes built afterwards (0 = the default, 10)."""
    lib().emu_set_prefix_len(int(p))


def device_gap_count():
    return int(lib().emu_device_gap_count())


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty))


def _u8(x):
    if isinstance(x, (bytes, bytearray)):
        return np.frombuffer(bytes(x), dtype=np.uint8).copy()
    return np.ascontiguousarray(x, dtype=np.uint8)


def csr(queries):
    qs = [_u8(q) for q in queries]
    offsets = np.zeros(len(qs) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(q) for q in qs])
    concat = np.concatenate(qs) if qs else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(concat), offsets


class EmuIndex:
    def __init__(self, h):
        self.h = h
        L = lib()
        self.k, self.n_sets, self.n_kmers = L.emu_k(h), L.emu_n_sets(h), L.emu_n_kmers(h)

    @classmethod
    def build(cls, seqs, k=31, add_revcomp=False, threads=1):
        L = lib()
        ss = [_u8(s) for s in seqs]
        ptrs = (u8p * max(len(ss), 1))(*[_p(s, C.c_uint8) for s in ss])
        lens = np.array([len(s) for s in ss], dtype=np.uint64)
        err = C.create_string_buffer(256)
        h = L.emu_host_build(ptrs, _p(lens, C.c_uint64), len(ss), k, int(add_revcomp), threads, err, 256)
        if not h:
            raise RuntimeError(err.value.decode())
        return cls(h)

    @classmethod
    def from_parts(cls, k, n_sets, n_kmers, rows, lcs):
        L = lib()
        rows = [np.ascontiguousarray(r, dtype=np.uint64) for r in rows]
        lcs = _u8(lcs)
        ptrs = (u64p * 4)(*[_p(r, C.c_uint64) for r in rows])
        return cls(L.emu_from_parts(k, n_sets, n_kmers, ptrs, _p(lcs, C.c_uint8)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().emu_free(self.h)
            self.h = None

    def export(self):
        nw = (self.n_sets + 63) // 64
        rows = [np.zeros(nw, dtype=np.uint64) for _ in range(4)]
        lcs = np.zeros(self.n_sets, dtype=np.uint8)
        Cc = np.zeros(4, dtype=np.uint64)
        lib().emu_export(self.h, *[_p(r, C.c_uint64) for r in rows], _p(lcs, C.c_uint8), _p(Cc, C.c_uint64))
        return rows, lcs, Cc

    def access_kmer(self, colex):
        out = np.zeros(self.k, dtype=np.uint8)
        lib().emu_access_kmer(self.h, colex, _p(out, C.c_uint8))
        return out.tobytes()

    def search(self, pat):
        p = _u8(pat)
        l, r = C.c_uint64(0), C.c_uint64(0)
        ok = lib().emu_search(self.h, _p(p, C.c_uint8), len(p), C.byref(l), C.byref(r))
        return (l.value, r.value) if ok else None

    def query_sbwt_batch(self, queries, chunk_len=0, intervals=True, counters=False):
        concat, offsets = csr(queries)
        n = len(concat)
        d = np.zeros(n, dtype=np.uint8)
        l = np.zeros(n, dtype=np.uint32) if intervals else None
        r = np.zeros(n, dtype=np.uint32) if intervals else None
        cnt = np.zeros(10, dtype=np.uint64) if counters else None
        lib().emu_query_sbwt_batch(self.h, _p(concat, C.c_uint8), _p(offsets, C.c_uint64), len(queries), chunk_len,
                                   _p(d, C.c_uint8), _p(l, C.c_uint32) if intervals else None,
                                   _p(r, C.c_uint32) if intervals else None, _p(cnt, C.c_uint64) if counters else None)
        return d, l, r, offsets, cnt

    def call(self, ref_seq, thr, build_k, revcomp=False):
        r = _u8(ref_seq)
        cap, capc = len(r) + 1, 4 * len(r) + 64
        pos = np.zeros(cap, dtype=np.uint64)
        ql, rl = np.zeros(cap, dtype=np.uint32), np.zeros(cap, dtype=np.uint32)
        qc, rc = np.zeros(capc, dtype=np.uint8), np.zeros(capc, dtype=np.uint8)
        n = lib().emu_call(self.h, _p(r, C.c_uint8), len(r), thr, build_k, int(revcomp), _p(pos, C.c_uint64),
                           _p(ql, C.c_uint32), _p(rl, C.c_uint32), _p(qc, C.c_uint8), _p(rc, C.c_uint8))
        if n < 0:
            raise RuntimeError("panic")
        out, qo, ro = [], 0, 0
        for i in range(n):
            out.append((int(pos[i]), bytes(qc[qo:qo + ql[i]]), bytes(rc[ro:ro + rl[i]])))
            qo += int(ql[i])
            ro += int(rl[i])
        return out

    def map(self, ref_seq, thr, p, fill_gaps=True, call_variants=True, format=True, build_k=31, revcomp=False,
            call_thr=0):
        r = _u8(ref_seq)
        out = np.zeros(len(r), dtype=np.uint8)
        rc = lib().emu_map(self.h, _p(r, C.c_uint8), len(r), thr, call_thr, p, int(fill_gaps), int(call_variants), int(format),
                           build_k, int(revcomp), _p(out, C.c_uint8))
        if rc < 0:
            raise RuntimeError("panic")
        return out.tobytes()

    def find_batch(self, queries, thr, max_gap_len=0):
        """K0+K1+K2b(masks)+K4; list (per query) of 7-tuples (start, end, matches, mismatches, jumps, gap_bases, gap_opens)."""
        concat, offsets = csr(queries)
        cap = len(concat) + 1
        out = np.zeros(7 * cap, dtype=np.uint64)
        roff = np.zeros(len(queries) + 1, dtype=np.uint64)
        n = lib().emu_find_batch(self.h, _p(concat, C.c_uint8), _p(offsets, C.c_uint64), len(queries), thr, max_gap_len,
                                 _p(out, C.c_uint64), cap, _p(roff, C.c_uint64))
        assert n == roff[-1]
        return [[tuple(int(x) for x in out[7 * j:7 * j + 7]) for j in range(int(roff[i]), int(roff[i + 1]))]
                for i in range(len(queries))]

    def matches_batch(self, queries, thr, chunk_len=0):
        concat, offsets = csr(queries)
        out = np.zeros(len(concat), dtype=np.uint8)
        lib().emu_matches_batch(self.h, _p(concat, C.c_uint8), _p(offsets, C.c_uint64), len(queries), thr, chunk_len,
                                _p(out, C.c_uint8))
        return [out[int(offsets[i]):int(offsets[i + 1])].tobytes() for i in range(len(queries))]


def pack(concat, offsets):
    """K0 on a CSR batch: (pack u64, inv u32, sep u32, wq u32) arrays in padded space."""
    concat = np.ascontiguousarray(concat, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    nq = len(offsets) - 1
    nw = int(lib().emu_pack(_p(concat, C.c_uint8), _p(offsets, C.c_uint64), nq, None, None, None, None))
    pk = np.zeros(nw, dtype=np.uint64)
    iv, sp, wq = (np.zeros(nw, dtype=np.uint32) for _ in range(3))
    lib().emu_pack(_p(concat, C.c_uint8), _p(offsets, C.c_uint64), nq, _p(pk, C.c_uint64), _p(iv, C.c_uint32),
                   _p(sp, C.c_uint32), _p(wq, C.c_uint32))
    return pk, iv, sp, wq


def set_fused(on, chunk=0, sms=4):
    """matches / find through the fused K1 + K2b kernel (fused.cuh); chunk = target positions per lane, sms = emulated SM count."""
    lib().emu_set_fused(int(on), int(chunk), int(sms))


def set_k2_mode(mode):
    """0: product dispatch, 1: always K2, 2: K2b where supported."""
    lib().emu_set_k2_mode(int(mode))


def derand_translate_u8(ms, k, thr):
    m = _u8(ms)
    out = np.zeros(len(m), dtype=np.uint8)
    lib().emu_derand_translate_u8(_p(m, C.c_uint8), len(m), k, thr, _p(out, C.c_uint8))
    return out.tobytes()


def derandomize_general(ms, k, thr):
    m = np.ascontiguousarray(ms, dtype=np.uint64)
    out = np.zeros(len(m), dtype=np.int64)
    lib().emu_derandomize_general(_p(m, C.c_uint64), len(m), k, thr, _p(out, C.c_int64))
    return out


def translate_i64(d, k, thr):
    dd = np.ascontiguousarray(d, dtype=np.int64)
    out = np.zeros(len(dd), dtype=np.uint8)
    lib().emu_translate_i64(_p(dd, C.c_int64), len(dd), k, thr, _p(out, C.c_uint8))
    return out.tobytes()


def rle_batch(alns, max_gap_len):
    """K4 on a list of plain translations; returns a list (per query) of 7-tuples."""
    concat, offsets = csr(alns)
    cap = len(concat) + 1
    out = np.zeros(7 * cap, dtype=np.uint64)
    roff = np.zeros(len(alns) + 1, dtype=np.uint64)
    n = lib().emu_rle_batch(_p(concat, C.c_uint8), _p(offsets, C.c_uint64), len(alns), max_gap_len,
                            _p(out, C.c_uint64), cap, _p(roff, C.c_uint64))
    assert n == roff[-1]
    return [[tuple(int(x) for x in out[7 * j:7 * j + 7]) for j in range(int(roff[i]), int(roff[i + 1]))]
            for i in range(len(alns))]
