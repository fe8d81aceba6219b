"""CPU-side checks of the product's host builder and of the kernel LOGIC (kernels.cuh compiled
through tests/emu/host_emu.hpp) against the oracle.  The real parity tests run the CUDA build on
a B200 (tests/test_gpu_parity.py, -m gpu); these exist so that logic errors are found here first.
"""
import json
import os

import numpy as np
import pytest

import emu_lib as E
import oracle_lib as O
from kbo_b200 import synth

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))
REF_K3 = b"AAAGAACCA-TCAGGGCG"


def rand_seq(n, seed):
    return synth.random_seq(n, seed).tobytes()


def with_ns(seq, seed, rate=0.01):
    a = np.frombuffer(seq, dtype=np.uint8).copy()
    rng = np.random.default_rng(seed)
    a[rng.random(len(a)) < rate] = ord("N")
    return a.tobytes()


def assert_same_index(seqs, k, revcomp=False, threads=1):
    o = O.OracleIndex(seqs, k=k, add_revcomp=revcomp)
    e = E.EmuIndex.build(seqs, k=k, add_revcomp=revcomp, threads=threads)
    assert (e.k, e.n_sets, e.n_kmers) == (o.k, o.n_sets, o.n_kmers)
    rows, lcs, Cc = e.export()
    for a, b in zip(rows, o.rows()):
        assert np.array_equal(a, b)
    assert np.array_equal(lcs, o.lcs())
    assert np.array_equal(Cc, o.C())
    return o, e


# ------------------------------------------------------------- host builder ---
@pytest.mark.parametrize("k", [1, 2, 3, 4, 7, 20, 31, 32, 33, 51, 63, 64])
def test_host_builder_matches_oracle_small(k):
    seqs = [REF_K3, rand_seq(300, 1), b"ACGTNNACGTTTGACCANGGTA" * 3, rand_seq(70, 2)]
    o, e = assert_same_index(seqs, k)
    for i in range(0, o.n_sets, max(1, o.n_sets // 50)):
        assert e.access_kmer(i) == o.access_kmer(i)


def test_host_builder_golden_find_index():
    b = GOLD["lib.rs::doc@779"]
    o, e = assert_same_index([b["gene1"].encode(), b["gene2_rc"].encode()], 31)
    assert (e.n_kmers, e.n_sets) == (1176, 1237)


@pytest.mark.parametrize("k,revcomp,threads", [(31, False, 1), (31, True, 3), (15, False, 4), (47, True, 1)])
def test_host_builder_matches_oracle_medium(k, revcomp, threads):
    ref = with_ns(rand_seq(120_000, 7), 8, rate=0.0005)
    seqs = [ref[:50_000], ref[50_000:], ref[1000:3000]]  # duplicated region
    o, e = assert_same_index(seqs, k, revcomp, threads)
    rng = np.random.default_rng(3)
    for i in rng.integers(0, o.n_sets, size=200).tolist():
        assert e.access_kmer(i) == o.access_kmer(i)
    for s in rng.integers(0, 49_000, size=50).tolist():
        pat = ref[s:s + int(rng.integers(1, k + 1))]
        assert e.search(pat) == o.search(pat)
    assert e.search(b"ACGTN") is None


def test_repetitive_sequences():
    seqs = [b"A" * 200 + b"C" * 5 + b"AC" * 100, b"T" * 64, b"ACGT" * 50]
    for k in (5, 31):
        assert_same_index(seqs, k)


# ----------------------------------------------------------------- K1: MS ---
def check_ms(o, e, queries, chunk_len):
    d, l, r, off, _ = e.query_sbwt_batch(queries, chunk_len=chunk_len)
    for i, q in enumerate(queries):
        od, ol, orr = o.query_sbwt(q)
        a, b = int(off[i]), int(off[i + 1])
        assert np.array_equal(d[a:b].astype(np.uint64), od), (i, chunk_len)
        assert np.array_equal(l[a:b].astype(np.uint64), ol), (i, chunk_len)
        assert np.array_equal(r[a:b].astype(np.uint64), orr), (i, chunk_len)


def test_ms_golden_k3():
    o = O.OracleIndex([REF_K3], k=3)
    e = E.EmuIndex.build([REF_K3], k=3)
    d, l, r, off, _ = e.query_sbwt_batch([b"CAAGCCACTCATTGGGTC"], chunk_len=32)
    assert d.tolist() == [1, 2, 2, 3, 2, 2, 3, 2, 1, 2, 3, 1, 1, 1, 2, 3, 1, 2]  # index.rs:264-274
    check_ms(o, e, [b"CAAGCCACTCATTGGGTC", b"A", b"NNNN", b"GTGACTATGAGGAT"], 32)


@pytest.mark.parametrize("k", [3, 7, 20, 31, 51, 63])
def test_ms_matches_oracle(k):
    ref = rand_seq(30_000, 11)
    asm = synth.mutate(np.frombuffer(ref, dtype=np.uint8), 12).tobytes()
    o = O.OracleIndex([asm], k=k)
    e = E.EmuIndex.build([asm], k=k)
    queries = [ref[:5000], with_ns(ref[5000:9000], 5, 0.02), rand_seq(700, 13), ref[10_000:10_003], b"N",
               ref[20_000:20_257], b"ACGT" * 40 + b"$" + ref[100:400]]
    for chunk_len in (32, 64, 256, 1024):
        check_ms(o, e, queries, chunk_len)


def test_ms_wide_intervals_take_the_scan_path():
    """Two-letter reference: a G/T in the query fails at every depth, so contract_left walks down to depths whose
    intervals are wider than the 4095-node reach of the link array (kernels.cuh LINK_FAR) and falls back to scanning."""
    rng = np.random.default_rng(5)
    ref = np.frombuffer(b"AC", dtype=np.uint8)[rng.integers(0, 2, 30_000)].tobytes()
    o = O.OracleIndex([ref], k=31)
    e = E.EmuIndex.build([ref], k=31)
    q = bytearray(ref[1000:5000])
    for i in rng.integers(0, len(q), 150):
        q[int(i)] = ord("GT"[int(i) & 1])
    q2 = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.choice(4, 3000, p=[0.45, 0.45, 0.05, 0.05])].tobytes()
    check_ms(o, e, [bytes(q), q2, b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAG" * 10], 64)
    d, l, r, off, cnt = e.query_sbwt_batch([bytes(q), q2], chunk_len=64, counters=True)
    assert cnt[3] > 0  # scans past the link reach did happen


def test_ms_prefix_table_shortens_warm_up_only():
    """With the prefix-state table (k >= 16) a chunk steps through at most k-1-10 warm-up bases; (d, l, r) are unchanged."""
    ref = rand_seq(30_000, 61)
    asm = synth.mutate(np.frombuffer(ref, dtype=np.uint8), 62).tobytes()
    o = O.OracleIndex([asm], k=31)
    queries = [ref[:9000], with_ns(ref[9000:12_000], 63, 0.03), rand_seq(800, 64), ref[15_000:15_040], b"ACGTN" * 30]
    processed = {}
    try:
        for on in (0, 1):
            E.lib().emu_set_prefix_table(on)
            e = E.EmuIndex.build([asm], k=31)
            check_ms(o, e, queries, 64)
            processed[on] = e.query_sbwt_batch(queries, chunk_len=64, counters=True)[4][4]
    finally:
        E.lib().emu_set_prefix_table(1)
    assert processed[1] < 0.93 * processed[0]


@pytest.mark.parametrize("k,depth", [(31, 1), (31, 4), (31, 7), (31, 12), (20, 13), (63, 9)])
def test_ms_prefix_table_depths(k, depth):
    """The prefix-state table at other depths than the default 10 (kbo_set_prefix_len): (d, l, r) must not depend on it --
    any depth up to k-1, N-rich queries, tiny queries, separators inside the warm-up window."""
    ref = rand_seq(30_000, 71)
    asm = synth.mutate(np.frombuffer(ref, dtype=np.uint8), 72).tobytes()
    o = O.OracleIndex([asm], k=k)
    queries = [ref[:6000], with_ns(ref[6000:9000], 73, 0.05), rand_seq(900, 74), ref[15_000:15_009], b"ACGTN" * 30,
               b"A", b"NNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNACGT", ref[20_000:20_012], ref[21_000:21_300] + b"$" + ref[50:400]]
    try:
        E.set_prefix_len(depth)
        e = E.EmuIndex.build([asm], k=k)
        for chunk_len in (32, 64, 512):
            check_ms(o, e, queries, chunk_len)
    finally:
        E.set_prefix_len(0)


@pytest.mark.parametrize("k", [3, 7, 31, 63])
def test_ms_two_bases_per_probe(k):
    """Lengths-only MS (no intervals) with and without the rank2 rows in the index: same d as the oracle for every
    chunk length."""
    ref = rand_seq(30_000, 71)
    asm = synth.mutate(np.frombuffer(ref, dtype=np.uint8), 72).tobytes()
    o = O.OracleIndex([asm], k=k)
    queries = [ref[:6000], with_ns(ref[6000:9000], 73, 0.02), rand_seq(900, 74), ref[10_000:10_003], b"N", b"AC",
               ref[20_000:20_257], b"ACGT" * 40 + b"$" + ref[100:400], asm[500:2500]]
    want = [o.query_sbwt(q)[0] for q in queries]
    attempts = {}
    # K1p (flag bit 5): two bases per probe inside K1; same lengths, fewer probes
    E.lib().emu_set_ms_flags(32)
    try:
        e = E.EmuIndex.build([asm], k=k)
        for chunk_len in (32, 64, 96, 1024):
            d, _, _, off, cnt = e.query_sbwt_batch(queries, chunk_len=chunk_len, intervals=False, counters=True)
            for i, w in enumerate(want):
                assert np.array_equal(d[int(off[i]):int(off[i + 1])].astype(np.uint64), w), ("pairs", k, chunk_len, i)
            assert cnt[5] == sum(len(q) + 1 for q in queries)
            attempts[("pairs", chunk_len)] = int(cnt[0])
    finally:
        E.lib().emu_set_ms_flags(0)
    try:
        for on in (1, 0):
            E.lib().emu_set_rank2(on)
            e = E.EmuIndex.build([asm], k=k)
            for chunk_len in (32, 64, 96, 1024):
                d, _, _, off, cnt = e.query_sbwt_batch(queries, chunk_len=chunk_len, intervals=False, counters=True)
                for i, w in enumerate(want):
                    assert np.array_equal(d[int(off[i]):int(off[i + 1])].astype(np.uint64), w), (k, on, chunk_len, i)
                assert cnt[5] == sum(len(q) + 1 for q in queries)
                attempts[(on, chunk_len)] = int(cnt[0])
    finally:
        E.lib().emu_set_rank2(1)
    # (K1 itself probes one base at a time -- the pair probes through rank2 live in the fused kernel, which the
    # matches / find tests below run; here rank2 must simply not change anything)
    assert attempts[(1, 64)] == attempts[(0, 64)]
    if k >= 31:
        assert attempts[("pairs", 64)] < 0.8 * attempts[(1, 64)]


def test_ms_tiny_index_and_counters():
    o = O.OracleIndex([b"ACG"], k=3)
    e = E.EmuIndex.build([b"ACG"], k=3)
    assert o.n_sets < 64
    check_ms(o, e, [b"ACGACGTTACG", b"TTTT"], 32)
    d, l, r, off, cnt = e.query_sbwt_batch([b"ACGACGTTACG" * 30], chunk_len=64, counters=True)
    assert cnt[5] == 330 + 1  # emitted = padded positions (incl. the separator)
    assert cnt[4] >= cnt[5] and cnt[0] >= 330


# -------------------------------------------------- K2: derandomize+translate ---
@pytest.fixture(autouse=True, params=["dispatch", "k2-only", "fused", "fused-small-chunks", "fused-exact",
                                     "fused-exact-small-chunks"])
def k2_mode(request):
    """Tests that reach derandomize+translate run several times: product dispatch of the separate kernels (K2b where it
    applies), K2 alone, and the fused K1 + K2b kernel (fused.cuh) -- its two-pass form and its one-pass form ("exact":
    K1's recurrence in pass A, no repair pass) -- with the default and with a tiny lane chunk (many tiles, tasks that span
    lanes' whole chunks, look-ahead past the shared-memory tile)."""
    name = request.node.originalname or request.node.name  # the function name, without the parameter ids
    touches_k2 = any(t in name for t in ("k2", "matches", "map", "call", "find", "rle"))
    touches_fused = any(t in name for t in ("matches", "find")) and "rle_kernel" not in name
    if request.param == "k2-only" and not touches_k2:
        pytest.skip("does not reach K2")
    if request.param.startswith("fused") and not touches_fused:
        pytest.skip("does not reach the fused kernel")
    E.set_k2_mode(1 if request.param == "k2-only" else 0)
    E.set_fused(request.param.startswith("fused"), 9 if request.param.endswith("small-chunks") else 0, 3)
    E.lib().emu_set_ms_flags(8 if "exact" in request.param else 0)  # bit 3: the one-pass form of the fused kernel
    yield
    E.set_k2_mode(0)
    E.set_fused(False)
    E.lib().emu_set_ms_flags(0)


def valid_ms_vector(rng, n, k, thr):
    """Random vector obeying ms[i+1] <= ms[i] + 1 with long flat / rising runs above and below thr."""
    out = np.zeros(n, dtype=np.int64)
    cur = int(rng.integers(0, k + 1))
    for i in range(n):
        out[i] = cur
        u = rng.random()
        if u < 0.45:
            cur = min(cur + 1, k)
        elif u < 0.80:
            cur = cur
        elif u < 0.9:
            cur = int(rng.integers(0, cur + 1))
        else:
            cur = max(cur - int(rng.integers(1, 4)), 0)
    return out


@pytest.mark.parametrize("k,thr", [(3, 2), (31, 15), (31, 22), (31, 30), (63, 24), (7, 6), (31, 2), (127, 60), (128, 60),
                                   (31, 31), (5, 3)])
def test_k2_on_valid_ms(k, thr):
    rng = np.random.default_rng(k * 100 + thr)
    for n in (3, 4, 31, 32, 33, 511, 512, 513, 1023, 1024, 1025, 2047, 2048, 5000):
        ms = valid_ms_vector(rng, n, k, thr)
        want = O.translate_ms_vec(O.derandomize_ms_vec(ms, k, thr), k, thr)
        got = E.derand_translate_u8(ms.astype(np.uint8), k, thr)
        assert got == want, (n, k, thr)


def test_k2_long_flat_and_noise_runs():
    k, thr = 31, 15
    for ms in ([20] * 3000 + [5, 5, 5], [5] * 2000 + [31] * 40 + [3] * 1500, [20] * 8 + [5, 5, 5],
               [16] * 700 + [17] * 700 + [31] * 5 + [0] * 900, [31] * 2048, [0] * 2048):
        ms = np.array(ms, dtype=np.int64)
        want = O.translate_ms_vec(O.derandomize_ms_vec(ms, k, thr), k, thr)
        assert E.derand_translate_u8(ms.astype(np.uint8), k, thr) == want


# ------------------------------------------------------- K0+K1+K2: matches ---
def test_matches_goldens():
    o = O.OracleIndex([REF_K3], k=3)
    e = E.EmuIndex.build([REF_K3], k=3)
    assert e.matches_batch([b"GTGACTATGAGGAT"], 3, 32) == [b"---------MMM--"]  # lib.rs:600-610
    b = GOLD["lib.rs::doc@779"]
    o = O.OracleIndex([b["gene1"].encode(), b["gene2_rc"].encode()], k=31)
    e = E.EmuIndex.build([b["gene1"].encode(), b["gene2_rc"].encode()], k=31)
    q = b["query"].encode()
    assert e.matches_batch([q], 16, 64) == [o.matches(q)]


@pytest.mark.parametrize("k,p", [(31, 1e-7), (20, 1e-3), (51, 1e-7)])
def test_matches_batch_matches_oracle(k, p):
    ref = np.frombuffer(rand_seq(40_000, 21), dtype=np.uint8)
    o = O.OracleIndex([ref.tobytes()], k=k)
    e = E.EmuIndex.build([ref.tobytes()], k=k)
    thr = O.random_match_threshold(k, o.n_kmers, 4, p)
    genes, off = synth.gene_queries(ref, 40, 1000, 22)
    queries = [genes[int(off[i]):int(off[i + 1])].tobytes() for i in range(40)]
    asm = synth.mutate(ref, 23).tobytes()
    queries += [asm[:7000], rand_seq(1500, 24), with_ns(asm[7000:9000], 25, 0.01), asm[9000:9003], asm[9100:9611],
                asm[10_000:10_512], asm[11_000:11_513]]
    for chunk_len in (64, 512):
        got = e.matches_batch(queries, thr, chunk_len)
        for g_, q in zip(got, queries):
            assert g_ == o.matches(q, p)


@pytest.mark.parametrize("seed", [31, 32])
def test_matches_batch_many_tiny_queries(seed):
    """Separators in almost every word (3 bases is the shortest query the reference accepts, derandomize.rs:276),
    queries around the word / tile sizes."""
    k, p = 31, 1e-7
    ref = np.frombuffer(rand_seq(30_000, 26), dtype=np.uint8)
    o = O.OracleIndex([ref.tobytes()], k=k)
    e = E.EmuIndex.build([ref.tobytes()], k=k)
    thr = O.random_match_threshold(k, o.n_kmers, 4, p)
    asm = synth.mutate(ref, 27).tobytes()
    rng = np.random.default_rng(seed)
    lens = [3, 3, 3, 4, 5, 31, 32, 33, 63, 64, 65, 1023, 1024, 1025, 3, 3, 3, 2047] + \
           [int(x) for x in rng.integers(3, 90, 150)] + [int(x) for x in rng.integers(25, 400, 30)]
    rng.shuffle(lens)
    queries = []
    for n in lens:
        a = int(rng.integers(0, len(asm) - n))
        queries.append(asm[a:a + n])
    got = e.matches_batch(queries, thr, 64)
    for g_, q in zip(got, queries):
        assert g_ == o.matches(q, p)


# -------------------------------------------- standalone derandomize / translate ---
def test_matches_long_plateaus_cross_tiles(request):
    """MS plateaus between the threshold and k (low-complexity runs longer than anything in the index) keep the
    derandomize look-ahead undecided for hundreds of positions: in the fused kernel it runs past the tile's MS bytes
    and continues the recurrence on the fly (MsTail); find on the same batch crosses tiles with open segments."""
    rng = np.random.default_rng(77)
    left, right = rand_seq(3000, 78), rand_seq(3000, 79)
    ref = left + b"C" + b"A" * 25 + b"G" + right + b"T" + b"AC" * 13 + b"G" + rand_seq(500, 80)
    o = O.OracleIndex([ref], k=31)
    e = E.EmuIndex.build([ref], k=31)
    queries = [left[2000:] + b"C" + b"A" * 700 + b"G" + right[:800],
               b"A" * 1500,
               right[100:900] + b"AC" * 400 + right[900:1500],
               left[:1200], b"ACA", b"A" * 90 + b"N" + b"A" * 200]
    concat, offsets = E.csr(queries)
    ext0, launches0 = E.lib().emu_tail_extension_count(), E.lib().emu_fused_launches()
    for p in (1e-7, 0.3):
        thr = O.random_match_threshold(31, o.n_kmers, 4, p)
        assert 2 <= thr < 25
        _, want, _ = o.matches_batch(concat, offsets, p)
        got = b"".join(e.matches_batch(queries, thr))
        assert got == want.tobytes()
        for gap in (0, 3):
            assert e.find_batch(queries, thr, gap) == [o.find(q, p, gap) for q in queries]
    if "fused" in request.node.name:
        assert E.lib().emu_fused_launches() > launches0
        if "small-chunks" in request.node.name:  # (tile ends fall inside the plateaus)
            assert E.lib().emu_tail_extension_count() > ext0  # the look-ahead did leave the tile's MS bytes


@pytest.mark.parametrize("k,thr", [(3, 2), (31, 15), (31, 22), (63, 40), (5, 4)])
def test_general_derandomize_arbitrary_vectors(k, thr):
    rng = np.random.default_rng(k + thr)
    for n in (3, 5, 255, 256, 257, 1023, 1024, 1025, 4097, 10_000):
        for kind in range(3):
            if kind == 0:
                ms = rng.integers(0, k + 1, size=n)
            elif kind == 1:
                ms = np.clip(rng.integers(thr - 2, thr + 4, size=n), 0, k)
            else:
                ms = valid_ms_vector(rng, n, k, thr)
            want = O.derandomize_ms_vec(ms, k, thr)
            got = E.derandomize_general(ms, k, thr)
            assert np.array_equal(got, want), (n, kind)


def test_general_derandomize_golden():
    got = E.derandomize_general([1, 2, 2, 3, 2, 2, 3, 2, 1, 2, 3, 1, 1, 1, 2, 3, 1, 2], 3, 2)
    assert got.tolist() == [0, 1, 2, 3, 1, 2, 3, 0, 1, 2, 3, -1, 0, 1, 2, 3, -1, 0]  # derandomize.rs:373-379


@pytest.mark.parametrize("k,thr", [(3, 2), (31, 15)])
def test_translate_i64_arbitrary_vectors(k, thr):
    rng = np.random.default_rng(99)
    assert E.translate_i64([0, 1, 2, 3, 1, 2, 3, 0, 1, 2, 3, -1, 0, 1, 2, 3, -1, 0], 3, 2) == b"XMMRRMMXMMM--MMM--"
    for n in (3, 4, 5, 100, 3000):
        for _ in range(20):
            d = rng.integers(-5, k + 1, size=n)
            assert E.translate_i64(d, k, thr) == O.translate_ms_vec(d, k, thr)


# ------------------------------------------------------------------- K4: device RLE ---
def test_rle_kernel_goldens():
    aln = b"XMMRRMMXMMM--MMM--"
    assert E.rle_batch([aln], 0) == [[(0, 11, 9, 2, 1, 0, 0), (13, 16, 3, 0, 0, 0, 0)]]  # format.rs:77-95
    assert E.rle_batch([aln], 3) == [[(0, 16, 12, 2, 1, 2, 1)]]                            # format.rs:122-140
    big = GOLD["format.rs::run_lengths"]["input"].encode()                                 # format.rs:295-330
    assert E.rle_batch([big], 0) == [[(5, 33, 28, 0, 0, 0, 0), (81, 207, 126, 0, 0, 0, 0), (372, 423, 51, 0, 0, 0, 0),
                                      (487, 512, 25, 0, 0, 0, 0)]]


def test_rle_kernel_random_plain_alignments():
    rng = np.random.default_rng(17)
    for weights in ([8, 1, 1, 0.3], [2, 6, 1, 1], [1, 1, 1, 1]):
        p = np.array(weights, dtype=float) / sum(weights)
        for gap in (0, 1, 2, 5, 40):
            alns = []
            for n in (1, 2, 3, 31, 32, 33, 64, 65, 100, 1000, 1500):
                a = np.frombuffer(b"M-XR", dtype=np.uint8)[rng.choice(4, size=n, p=p)].copy()
                if a[0] == ord("R"):
                    a[0] = ord("M")  # format.rs:176 panics on a leading 'R'
                alns.append(a.tobytes())
            got = E.rle_batch(alns, gap)
            for g_, a in zip(got, alns):
                assert g_ == O.run_lengths_gapped(a, gap), (gap, a[:80])


@pytest.mark.parametrize("gap", [0, 3, 25, 100_000])
def test_find_batch_matches_oracle(gap):
    """K0+K1+K2b(masks)+K4 (or K2+chars_to_masks in k2-only mode) against kbo::find of the oracle."""
    k, p = 31, 1e-7
    ref = np.frombuffer(rand_seq(40_000, 41), dtype=np.uint8)
    o = O.OracleIndex([ref.tobytes()], k=k)
    e = E.EmuIndex.build([ref.tobytes()], k=k)
    thr = O.random_match_threshold(k, o.n_kmers, 4, p)
    genes, off = synth.gene_queries(ref, 30, 1000, 42)
    queries = [genes[int(off[i]):int(off[i + 1])].tobytes() for i in range(30)]
    asm = synth.mutate(ref, 43).tobytes()
    queries += [asm[:7000], rand_seq(1500, 44), with_ns(asm[7000:9000], 45, 0.01), asm[9000:9003], asm[9100:9611],
                asm[10_000:10_040] + rand_seq(30, 46) + asm[10_070:10_200] + rand_seq(300, 47) + asm[10_500:10_600],
                rand_seq(40, 48) + asm[11_000:11_513] + rand_seq(33, 49)]
    rng = np.random.default_rng(50)
    for n in rng.integers(3, 120, 60):
        a = int(rng.integers(0, len(asm) - 200))
        queries.append(asm[a:a + int(n)])
    got = e.find_batch(queries, thr, gap)
    for i, (g_, q) in enumerate(zip(got, queries)):
        assert g_ == o.find(q, p, gap), (gap, i)


# ---------------------------------------------------------------------------
# K0 alone: arbitrary bytes, empty queries, misaligned batch start
# ---------------------------------------------------------------------------
def _pack_model(concat, offsets):
    """Per-position model of the packed layout (kernels.cuh QueryView)."""
    nq = len(offsets) - 1
    off0 = int(offsets[0])
    lens = [int(offsets[i + 1] - offsets[i]) for i in range(nq)]
    Lp = sum(lens) + nq
    code, inv, sep, nsep_before = [], [], [], []
    seen = 0
    for i in range(nq):
        for b in concat[int(offsets[i]):int(offsets[i + 1])]:
            c = {65: 0, 67: 1, 71: 2, 84: 3}.get(int(b))
            code.append(c or 0); inv.append(c is None); sep.append(False); nsep_before.append(seen)
        code.append(0); inv.append(True); sep.append(True); nsep_before.append(seen)
        seen += 1
    assert len(code) == Lp and off0 >= 0
    return code, inv, sep, nsep_before


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_pack_kernel_matches_position_model(seed):
    rng = np.random.default_rng(seed)
    lens = [0, 1, 7, 8, 9, 31, 32, 33, 0, 0, 64, 5, 100, 257, 3, 0] + [int(x) for x in rng.integers(0, 70, 40)]
    rng.shuffle(lens)
    body = rng.integers(0, 256, sum(lens), dtype=np.uint8)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, len(body))]
    body = np.where(rng.random(len(body)) < 0.7, acgt, body).astype(np.uint8)
    lead = seed  # the batch starts `lead` bytes into the buffer
    concat = np.concatenate([np.zeros(lead, dtype=np.uint8), body])
    offsets = np.concatenate([[lead], lead + np.cumsum(lens)]).astype(np.uint64)
    pk, iv, sp, wq = E.pack(concat, offsets)
    code, inv, sep, nsb = _pack_model(concat, offsets)
    Lp = len(code)
    for pp in range(len(pk) * 32):
        w, j = pp >> 5, pp & 31
        got = (int(pk[w]) >> (2 * j)) & 3, bool((int(iv[w]) >> j) & 1), bool((int(sp[w]) >> j) & 1)
        want = (code[pp], inv[pp], sep[pp]) if pp < Lp else (0, True, True)
        assert got == want, (pp, got, want)
    for w in range(len(pk)):
        pp = 32 * w
        assert int(wq[w]) == (nsb[pp] if pp < Lp else len(lens))
