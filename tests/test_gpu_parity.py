"""Parity tests proper: the CUDA path, called through the C ABI (kbo_b200.api over
libkbo_b200.so), against the CPU oracle and the reference's golden vectors, bit-exact.
Needs a B200:  python -m pytest tests -m gpu
"""
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from kbo_b200 import api, build, synth

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))
REF_K3 = b"AAAGAACCA-TCAGGGCG"


@pytest.fixture(scope="module", autouse=True)
def _lib():
    build.build_library()
    api.load_library()
    assert api.device_count() >= 1, "no CUDA device: the gpu tests must not pass on a fallback"
    api.set_chunk_len(0)
    yield
    api.set_chunk_len(0)


@pytest.fixture(autouse=True, params=["separate", "fused", "fused-exact", "pairs"])
def kernel_path(request):
    """Tests that reach matches / find run several times: K1 followed by K2b, the fused K1 + K2b kernel of fused.cuh
    (kbo_set_ms_flags bit 4: MS bytes only in shared memory, TMA-staged queries) in its two-pass form (two bases per
    probe + repair pass) and its one-pass form (bit 3: K1's recurrence), and K1p (bit 5)."""
    name = request.node.originalname or request.node.name
    if request.param != "separate":
        if not any(t in name for t in ("matches", "find", "config2", "randomised", "tiny")):
            pytest.skip("does not reach the fused kernel / K1p")
        api.set_ms_flags({"fused": 16, "fused-exact": 24, "pairs": 32}[request.param])
    yield
    api.set_ms_flags(0)


def g(block, var):
    return GOLD[block][var].encode()


def rand_seq(n, seed):
    return synth.random_seq(n, seed).tobytes()


def with_ns(seq, seed, rate=0.01):
    a = np.frombuffer(seq, dtype=np.uint8).copy()
    rng = np.random.default_rng(seed)
    a[rng.random(len(a)) < rate] = ord("N")
    return a.tobytes()


def rle_tuples(rl):
    return [(r.start, r.end, r.matches, r.mismatches, r.jumps, r.gap_bases, r.gap_opens) for r in rl]


# ---------------------------------------------------------------- goldens through the ABI ---
def test_golden_query_sbwt_k3():
    ix = api.build([REF_K3], api.BuildOpts(k=3))  # index.rs:264-274
    d, l, r = api.query_sbwt(b"CAAGCCACTCATTGGGTC", ix)
    assert d.tolist() == [1, 2, 2, 3, 2, 2, 3, 2, 1, 2, 3, 1, 1, 1, 2, 3, 1, 2]
    assert (ix.n_kmers, ix.n_sets) == (13, 16)


def test_golden_matches_k3():
    ix = api.build([REF_K3], api.BuildOpts(k=3))  # lib.rs:600-610
    assert api.matches(b"GTGACTATGAGGAT", ix) == b"---------MMM--"


def test_golden_map_unrefined_k7():
    b = "lib.rs::doc@664"  # lib.rs:670-689, :698-717
    ix = api.build([g(b, "query")], api.BuildOpts(k=7, build_select=True))
    o = api.MapOpts(max_error_prob=0.1, fill_gaps=False, call_variants=False)
    assert api.map(g(b, "reference"), ix, o) == g(b, "expected")
    o.format = False
    assert api.map(g(b, "reference"), ix, o) == g("lib.rs::doc@692", "expected")


def test_golden_find_k31():
    b = "lib.rs::doc@779"  # lib.rs:786-806
    ix = api.build([g(b, "gene1"), g(b, "gene2_rc")], api.BuildOpts(k=31))
    assert (ix.n_kmers, ix.n_sets) == (1176, 1237)
    got = api.find(g(b, "query"), ix, api.FindOpts(max_gap_len=50))
    assert rle_tuples(got) == [(0, 513, 512, 1, 0, 0, 0), (593, 1340, 709, 0, 0, 38, 3)]


def test_golden_derandomize_translate_standalone():
    ms = [1, 2, 2, 3, 2, 2, 3, 2, 1, 2, 3, 1, 1, 1, 2, 3, 1, 2]  # derandomize.rs:373-379
    d = api.derandomize_ms_vec(ms, 3, 2)
    assert d.tolist() == [0, 1, 2, 3, 1, 2, 3, 0, 1, 2, 3, -1, 0, 1, 2, 3, -1, 0]
    assert api.translate_ms_vec(d, 3, 2) == b"XMMRRMMXMMM--MMM--"  # translate.rs:501-515
    assert api.translate_ms_vec([1, 2, 3, 1, 2, 3, 3, 3, 3, 1, 2, 3], 3, 2) == b"MMRRMMMMRRMM"  # :518-532
    # format.rs:229-247 chain, k=4 thr=3
    ms4 = [1, 2, 3, 4, 1, 2, 3, 3, 3, 4, 4, 4, 4, 3, 1, 2, 3, 4, 4, 4, 4]
    d4 = api.derandomize_ms_vec(ms4, 4, 3)
    assert d4.tolist() == [1, 2, 3, 4, -1, 0, 1, 2, 3, 4, 4, 4, 4, 0, 1, 2, 3, 4, 4, 4, 4]
    assert api.translate_ms_vec(d4, 4, 3) == b"MMMM--MMMMMMMXMMMMMMM"


def test_golden_intervals_pin_unique_kmers():
    # gap_filling.rs:535-564: the unique interval at position 16 decodes to CAGACAGCT (k=9)
    b = "gap_filling.rs::nearest_unique_context"
    ix = api.build([g(b, "query")], api.BuildOpts(k=9, build_select=True))
    d, l, r = api.query_sbwt(g(b, "reference"), ix)
    idx = 16
    while r[idx] - l[idx] != 1:
        idx -= 1
    assert idx == 16 and ix.access_kmer(int(l[idx])) == b"CAGACAGCT"
    # gap_filling.rs:567-600: search + access_kmer
    b = "gap_filling.rs::left_extend_kmer"
    ix = api.build([g(b, "sequence")], api.BuildOpts(k=6, build_select=True))
    lo, hi = ix.search(g(b, "query"))
    assert hi - lo == 1 and ix.access_kmer(lo) == b"GACTGC"


# --------------------------------------------------------------------- index vs oracle ---
@pytest.mark.parametrize("k,revcomp", [(3, False), (31, False), (31, True), (51, False), (64, False)])
def test_index_build_matches_oracle(k, revcomp):
    ref = with_ns(rand_seq(60_000, 7), 8, rate=0.0005)
    seqs = [ref[:30_000], ref[30_000:], ref[500:900]]
    o = O.OracleIndex(seqs, k=k, add_revcomp=revcomp)
    ix = api.build(seqs, api.BuildOpts(k=k, add_revcomp=revcomp, num_threads=2))
    assert (ix.k, ix.n_sets, ix.n_kmers) == (o.k, o.n_sets, o.n_kmers)
    rows, lcs, Cc = ix.export_parts()
    for a, b_ in zip(rows, o.rows()):
        assert np.array_equal(a, b_)
    assert np.array_equal(lcs, o.lcs()) and np.array_equal(Cc, o.C())


@pytest.mark.parametrize("k,revcomp", [(2, False), (5, True), (16, False), (32, False), (32, True), (33, False), (33, True),
                                       (51, True), (63, False), (64, True)])
def test_gpu_and_host_builders_agree_with_oracle(k, revcomp):
    """Every k <= 64 is built on the device (index_build.cuh: 64-bit keys up to 32, 128-bit keys above); the host builder
    is forced for comparison."""
    ref = with_ns(rand_seq(40_000, 17), 18, rate=0.001)
    seqs = [ref[:25_000], ref[25_000:], b"ACGT" * 20, b"AC", ref[100:160]]
    o = O.OracleIndex(seqs, k=k, add_revcomp=revcomp)
    parts = []
    for host in (False, True):
        api.set_host_builder(host)
        try:
            ix = api.build(seqs, api.BuildOpts(k=k, add_revcomp=revcomp))
        finally:
            api.set_host_builder(False)
        assert (ix.k, ix.n_sets, ix.n_kmers) == (o.k, o.n_sets, o.n_kmers), host
        rows, lcs, Cc = ix.export_parts()
        for a, b_ in zip(rows, o.rows()):
            assert np.array_equal(a, b_), host
        assert np.array_equal(lcs, o.lcs()) and np.array_equal(Cc, o.C()), host
        q = synth.mutate(np.frombuffer(ref.replace(b"N", b"A"), dtype=np.uint8), 19).tobytes()[:5000]
        parts.append(api.query_sbwt(q, ix))
        od, ol, orr = o.query_sbwt(q)
        assert np.array_equal(parts[-1][0], od) and np.array_equal(parts[-1][1], ol)
    assert all(np.array_equal(a, b_) for a, b_ in zip(parts[0], parts[1]))


@pytest.mark.parametrize("k", [9, 31, 40])
def test_access_kmer_with_and_without_stored_nodes(k):
    """build_select keeps the sorted nodes on the host (O(1) access_kmer); without it access_kmer walks the
    incoming edges back.  Both must decode every node like the oracle, dummies included."""
    seqs = [rand_seq(3000, 5), b"ACGTTGCA" * 10, rand_seq(90, 6)]
    o = O.OracleIndex(seqs, k=k)
    fast = api.build(seqs, api.BuildOpts(k=k, build_select=True))
    slow = api.build(seqs, api.BuildOpts(k=k, build_select=False))
    for i in list(range(0, o.n_sets, 7)) + [o.n_sets - 1]:
        want = o.access_kmer(i)
        assert fast.access_kmer(i) == want and slow.access_kmer(i) == want, i
    with pytest.raises(api.KboPanic):
        fast.access_kmer(o.n_sets)


def test_index_build_degenerate_inputs():
    for seqs, k in (([b"AC"], 31), ([b"NNNN", b""], 5), ([b"ACGTACGTAC"], 10)):
        o = O.OracleIndex(seqs, k=k)
        ix = api.build(seqs, api.BuildOpts(k=k))
        assert (ix.n_sets, ix.n_kmers) == (o.n_sets, o.n_kmers)
        rows, lcs, Cc = ix.export_parts()
        assert np.array_equal(lcs, o.lcs()) and all(np.array_equal(a, b_) for a, b_ in zip(rows, o.rows()))


def test_index_from_parts_roundtrip():
    ref = rand_seq(20_000, 3)
    o = O.OracleIndex([ref], k=31)
    ix = api.index_from_parts(31, o.n_sets, o.n_kmers, o.rows(), o.lcs())
    q = synth.mutate(np.frombuffer(ref, dtype=np.uint8), 4).tobytes()[:6000]
    d, l, r = api.query_sbwt(q, ix)
    od, ol, orr = o.query_sbwt(q)
    assert np.array_equal(d, od) and np.array_equal(l, ol) and np.array_equal(r, orr)
    for i in (0, 1, 17, o.n_sets - 1):
        assert ix.access_kmer(i) == o.access_kmer(i)


@pytest.mark.parametrize("k,revcomp", [(3, False), (31, True), (51, False)])
def test_serialize_load_roundtrip(tmp_path, k, revcomp):
    """index::serialize_sbwt / load_sbwt (index.rs:128-212): the reference pins the files by a round trip
    (index.rs:277-296: loaded == built, LCS equal).  Here: the loaded index has the same arrays, answers query_sbwt,
    matches and find like the built one (and like the oracle), and the files have exactly the documented layout."""
    from test_cabi_host import write_index_files
    if k == 3:
        ref = b"AAAGAACCA-TCAGGGCG"  # the doctest input of index.rs:117-126
        q = b"AAAGAACCATCAGGGCGTTGA"
    else:
        ref = rand_seq(30_000, 41)
        q = with_ns(synth.mutate(np.frombuffer(ref, dtype=np.uint8), 42).tobytes()[:9000], 43, 0.01)
    opts = api.BuildOpts(k=k, add_revcomp=revcomp, build_select=True)
    ix = api.build([ref], opts)
    prefix = str(tmp_path / "serialized_index")
    api.serialize_sbwt(prefix, ix)
    ld = api.load_sbwt(prefix)
    assert (ld.k, ld.n_sets, ld.n_kmers) == (ix.k, ix.n_sets, ix.n_kmers)
    rows, lcs, Cc = ix.export_parts()
    rows2, lcs2, Cc2 = ld.export_parts()
    assert all(np.array_equal(a, b) for a, b in zip(rows, rows2)) and np.array_equal(lcs, lcs2) and np.array_equal(Cc, Cc2)
    o = O.OracleIndex([ref], k=k, add_revcomp=revcomp)
    od, ol, orr = o.query_sbwt(q)
    for index in (ix, ld):
        d, l, r = api.query_sbwt(q, index)
        assert np.array_equal(d, od) and np.array_equal(l, ol) and np.array_equal(r, orr)
    if k > 3:
        assert api.matches(q, ld) == api.matches(q, ix) == o.matches(q)
        assert api.find(q, ld) == api.find(q, ix) and len(api.find(q, ix)) > 0
    for i in (0, 1, ix.n_sets // 2, ix.n_sets - 1):
        assert ld.access_kmer(i) == ix.access_kmer(i) == o.access_kmer(i)
    # the bytes on disk are the documented layout, written here from the ORACLE's arrays
    write_index_files(prefix + "_expect", k, o.n_sets, o.n_kmers, o.rows(), o.lcs())
    for ext in (".sbwt", ".lcs"):
        assert open(prefix + ext, "rb").read() == open(prefix + "_expect" + ext, "rb").read(), ext
    # a second generation is byte-identical
    api.serialize_sbwt(prefix + "_2", ld)
    for ext in (".sbwt", ".lcs"):
        assert open(prefix + ext, "rb").read() == open(prefix + "_2" + ext, "rb").read(), ext


# ------------------------------------------------------------------------ MS vs oracle ---
@pytest.mark.parametrize("k", [3, 7, 20, 31, 51, 63])
def test_query_sbwt_matches_oracle(k):
    ref = rand_seq(50_000, 11)
    asm = synth.mutate(np.frombuffer(ref, dtype=np.uint8), 12).tobytes()
    o = O.OracleIndex([asm], k=k)
    ix = api.build([asm], api.BuildOpts(k=k))
    queries = [ref[:20_000], with_ns(ref[20_000:26_000], 5, 0.02), rand_seq(3000, 13), ref[30_000:30_003], b"N",
               ref[31_000:31_257], b"ACGT" * 40 + b"$" + ref[100:400], b"acgtacgt" + ref[40_000:40_100]]
    for chunk_len in (32, 128, 1024):
        api.set_chunk_len(chunk_len)
        d, l, r, off = api.query_sbwt_batch(queries, ix)
        for i, q in enumerate(queries):
            od, ol, orr = o.query_sbwt(q)
            a, b_ = int(off[i]), int(off[i + 1])
            assert np.array_equal(d[a:b_].astype(np.uint64), od), (k, chunk_len, i)
            assert np.array_equal(l[a:b_].astype(np.uint64), ol), (k, chunk_len, i)
            assert np.array_equal(r[a:b_].astype(np.uint64), orr), (k, chunk_len, i)
    api.set_chunk_len(0)
    with pytest.raises(api.KboPanic) as e:  # index.rs:248
        api.query_sbwt(b"", ix)
    assert e.value.status == 1


def test_prefix_table_and_wide_intervals():
    """(d, l, r) with and without the prefix-state table; a two-letter reference drives contractions into intervals
    wider than the link reach (kernels.cuh LINK_FAR), i.e. through the scanning fallback."""
    rng = np.random.default_rng(5)
    two = np.frombuffer(b"AC", dtype=np.uint8)[rng.integers(0, 2, 60_000)].tobytes()
    q = bytearray(two[1000:9000])
    for i in rng.integers(0, len(q), 300):
        q[int(i)] = ord("GT"[int(i) & 1])
    ref = rand_seq(50_000, 11)
    asm = synth.mutate(np.frombuffer(ref, dtype=np.uint8), 12).tobytes()
    cases = [(two, [bytes(q), b"A" * 500 + b"G" + b"C" * 300]),
             (asm, [ref[:20_000], with_ns(ref[20_000:26_000], 5, 0.02), rand_seq(3000, 13), b"ACGTN" * 50])]
    try:
        for on in (False, True):
            api.set_prefix_table(on)
            for text, queries in cases:
                o = O.OracleIndex([text], k=31)
                ix = api.build([text], api.BuildOpts(k=31))
                d, l, r, off = api.query_sbwt_batch(queries, ix)
                for i, qq in enumerate(queries):
                    od, ol, orr = o.query_sbwt(qq)
                    a, b_ = int(off[i]), int(off[i + 1])
                    assert np.array_equal(d[a:b_].astype(np.uint64), od), (on, i)
                    assert np.array_equal(l[a:b_].astype(np.uint64), ol), (on, i)
                    assert np.array_equal(r[a:b_].astype(np.uint64), orr), (on, i)
    finally:
        api.set_prefix_table(True)


@pytest.mark.parametrize("depth", [1, 5, 11, 12, 13, 14])
def test_prefix_table_depths(depth):
    """The prefix-state table at explicit depths (kbo_set_prefix_len; a chunk's warm-up starts from its entry).
    (d, l, r) and matches must not depend on the depth."""
    ref = rand_seq(120_000, 21)
    asm = synth.mutate(np.frombuffer(ref, dtype=np.uint8), 22).tobytes()
    o = O.OracleIndex([asm], k=31)
    queries = [ref[:30_000], with_ns(ref[30_000:36_000], 5, 0.03), rand_seq(3000, 23), b"ACGTN" * 50, b"A",
               ref[40_000:40_011], b"N" * 40 + ref[100:160], ref[50_000:50_700] + b"-" + ref[9:500]]
    api.set_prefix_len(depth)
    try:
        ix = api.build([asm], api.BuildOpts(k=31))
    finally:
        api.set_prefix_len(0)
    d, l, r, off = api.query_sbwt_batch(queries, ix)
    for i, qq in enumerate(queries):
        od, ol, orr = o.query_sbwt(qq)
        a, b_ = int(off[i]), int(off[i + 1])
        assert np.array_equal(d[a:b_].astype(np.uint64), od), i
        assert np.array_equal(l[a:b_].astype(np.uint64), ol), i
        assert np.array_equal(r[a:b_].astype(np.uint64), orr), i
    long_q = [q for q in queries if len(q) > 2]
    got = api.matches_batch(long_q, ix)
    for i, qq in enumerate(long_q):
        assert got[i] == o.matches(qq), i


def test_ms_invariants_large():
    """Size-independent properties at a larger size: chunking never changes the result, MS grows by
    at most one per base, intervals are non-empty and inside [0, n_sets]."""
    ref = synth.random_seq(600_000, 31)
    asm = synth.mutate(ref, 32)
    ix = api.build([asm.tobytes()], api.BuildOpts(k=31, num_threads=4))
    outs = []
    for chunk_len in (64, 256, 4096):
        api.set_chunk_len(chunk_len)
        outs.append(api.query_sbwt_batch([ref.tobytes()], ix))
    api.set_chunk_len(0)
    d, l, r, _ = outs[0]
    for d2, l2, r2, _ in outs[1:]:
        assert np.array_equal(d, d2) and np.array_equal(l, l2) and np.array_equal(r, r2)
    di = d.astype(np.int64)
    assert (di[1:] <= di[:-1] + 1).all() and di.max() == 31
    assert (l < r).all() and (r <= ix.n_sets).all()
    assert ((d > 0) | ((l == 0) & (r == ix.n_sets))).all()


# ------------------------------------------------------------------- matches vs oracle ---
@pytest.mark.parametrize("k,p", [(31, 1e-7), (20, 1e-3), (51, 1e-7), (7, 0.1)])
def test_matches_batch_matches_oracle(k, p):
    ref = synth.random_seq(80_000, 21)
    o = O.OracleIndex([ref.tobytes()], k=k)
    ix = api.build([ref.tobytes()], api.BuildOpts(k=k))
    genes, off = synth.gene_queries(ref, 200, 1000, 22)
    queries = [genes[int(off[i]):int(off[i + 1])].tobytes() for i in range(200)]
    asm = synth.mutate(ref, 23).tobytes()
    queries += [asm[:30_000], rand_seq(5000, 24), with_ns(asm[30_000:34_000], 25, 0.01), asm[40_000:40_003],
                asm[41_000:41_511], asm[42_000:42_512], asm[43_000:43_513], b"A" * 3000, b"AC" * 2000]
    want = [o.matches(q, p) for q in queries]
    try:
        for chunk_len, flags in ((64, 0), (0, 0), (0, 2)):  # flags bit1: K2 instead of the bit-parallel K2b
            api.set_chunk_len(chunk_len)
            api.set_ms_flags(flags)
            got = api.matches_batch(queries, ix, api.MatchOpts(max_error_prob=p))
            for i, (g_, w_) in enumerate(zip(got, want)):
                assert g_ == w_, (k, p, chunk_len, flags, i)
    finally:
        api.set_chunk_len(0)
        api.set_ms_flags(0)


def test_matches_many_tiny_queries():
    """Separators in almost every 32-position word; queries around the word and tile sizes."""
    k, p = 31, 1e-7
    ref = synth.random_seq(60_000, 26)
    o = O.OracleIndex([ref.tobytes()], k=k)
    ix = api.build([ref.tobytes()], api.BuildOpts(k=k))
    asm = synth.mutate(ref, 27).tobytes()
    rng = np.random.default_rng(31)
    lens = [3, 3, 3, 4, 5, 31, 32, 33, 63, 64, 65, 1023, 1024, 1025, 3, 3, 3, 2047] + \
           [int(x) for x in rng.integers(3, 90, 3000)] + [int(x) for x in rng.integers(25, 400, 300)]
    rng.shuffle(lens)
    queries = []
    for n in lens:
        a = int(rng.integers(0, len(asm) - n))
        queries.append(asm[a:a + n])
    got = api.matches_batch(queries, ix, api.MatchOpts(max_error_prob=p))
    for i, (g_, q) in enumerate(zip(got, queries)):
        assert g_ == o.matches(q, p), i


def test_matches_preconditions():
    ix = api.build([rand_seq(5000, 1)], api.BuildOpts(k=31))
    for q, status in [(b"", 1), (b"AC", 3)]:  # index.rs:248, derandomize.rs:276
        with pytest.raises(api.KboPanic) as e:
            api.matches(q, ix)
        assert e.value.status == status
    with pytest.raises(api.KboPanic) as e:   # derandomize.rs:136-137
        api.matches(b"ACGTACGT", ix, api.MatchOpts(max_error_prob=0.0))
    assert e.value.status == 6
    tiny = api.build([b"ACGT"], api.BuildOpts(k=2))
    with pytest.raises(api.KboPanic) as e:   # threshold 1 -> derandomize.rs:275
        api.matches(b"ACGTACGT", tiny, api.MatchOpts(max_error_prob=0.9999))
    assert e.value.status == 2


def test_matches_config2_shape_checksum():
    """BASELINE config 2 shape at 1/10 scale through the batch API: every query equals the oracle."""
    ref = synth.random_seq(500_000, synth.SEED_C2_REF)
    o = O.OracleIndex([ref.tobytes()], k=31)
    ix = api.build([ref.tobytes()], api.BuildOpts(k=31, num_threads=4))
    concat, off = synth.gene_queries(ref, 1000, 1000, synth.SEED_C2_GENES)
    got = api.matches_csr(concat, off, ix)
    _, want, _ = o.matches_batch(concat, off, n_threads=4)
    assert np.array_equal(got[:len(concat)], want)
    frac_m = (got[:len(concat)] == ord("M")).mean()
    assert 0.5 < frac_m < 1.0


def test_config2_full_size_bit_exact():
    """BASELINE.json configs[1] at FULL size (10,000 x 1 kbp vs 5 Mbp, k = 31): every alignment character and
    every RLE record equals the oracle's; plus size-independent properties of the MS vector."""
    ref = synth.random_seq(5_000_000, synth.SEED_C2_REF)
    o = O.OracleIndex([ref.tobytes()], k=31)
    ix = api.build([ref], api.BuildOpts(k=31))
    assert (ix.n_sets, ix.n_kmers) == (o.n_sets, o.n_kmers)
    concat, off = synth.gene_queries(ref, 10_000, 1000, synth.SEED_C2_GENES)
    got = api.matches_csr(concat, off, ix)
    _, want, _ = o.matches_batch(concat, off, n_threads=16)
    assert np.array_equal(got[:len(concat)], want)
    buf, n = api.find_csr(concat, off, ix)
    secs, n_want, _ = O.find_batch_timed(o, concat, off, 1e-7, 0, 16)
    assert n == n_want
    for q in range(0, 10_000, 501):
        a, b_ = int(buf.rle_offsets[q]), int(buf.rle_offsets[q + 1])
        got_q = [tuple(int(getattr(buf.rle[j], f)) for f, _ in api.RleC._fields_) for j in range(a, b_)]
        assert got_q == o.find(concat[int(off[q]):int(off[q + 1])].tobytes()), q
    d, _, _, _ = api.query_sbwt_batch([concat[:2_000_000].tobytes()], ix, intervals=False)
    di = d.astype(np.int64)
    assert (di[1:] <= di[:-1] + 1).all() and di.max() == 31


def test_config1_full_size_map_bit_exact():
    """BASELINE.json configs[0] at FULL size: kbo::map of a 4.6 Mbp reference against the index of a mutated
    4.6 Mbp assembly (1 % SNPs + short indels), k = 31, MapOpts::default(): output identical to the oracle."""
    ref = synth.random_seq(4_600_000, synth.SEED_C1_REF)
    asm = synth.mutate(ref, synth.SEED_C1_ASM)
    o = O.OracleIndex([asm.tobytes()], k=31)
    ix = api.build([asm], api.BuildOpts(k=31, build_select=True))
    assert (ix.n_sets, ix.n_kmers) == (o.n_sets, o.n_kmers)
    r = ref.tobytes()
    want = o.map(r, build_k=31)
    got = api.map(r, ix, api.MapOpts())
    assert got == want
    frac_gap = got.count(b"-") / len(got)
    assert 0.005 < frac_gap < 0.2
    # the unformatted translation as well (fill_gaps on, variants on)
    assert api.map(r, ix, api.MapOpts(format=False)) == o.map(r, format=False, build_k=31)
    # gaps bridged on 8 host threads (sbwt_build_opts.num_threads): same result
    threaded = api.MapOpts(sbwt_build_opts=api.BuildOpts(k=31, build_select=True, num_threads=8))
    assert api.map(r, ix, threaded) == want
    # fill_gaps / access_kmer on the host (refine_host.cpp) instead of the device (refine.cuh): same result
    api.set_device_refine(False)
    try:
        assert api.map(r, ix, api.MapOpts()) == want
    finally:
        api.set_device_refine(True)


def test_hbm_resident_index_regime():
    """BASELINE config 5's regime: an index well beyond L2 (120 Mbp: ~1.2 GB on the device, rank probes miss L2).
    Exact parity with an oracle index of the SAME 120 Mbp text (about a minute and ~6 GB of host memory): d, l and r
    of index::query_sbwt, kbo::matches and kbo::find, plus chunking independence."""
    big = synth.random_seq(120_000_000, 81)
    ix = api.build([big], api.BuildOpts(k=31))
    assert ix.device_bytes > 1_000_000_000
    o = O.OracleIndex([big.tobytes()], k=31)
    assert (ix.n_sets, ix.n_kmers) == (o.n_sets, o.n_kmers)
    # queries from all over the text: 1 % SNPs, some with indels and N's, one unrelated
    rng = np.random.default_rng(82)
    queries = []
    for i in range(60):
        a0 = int(rng.integers(0, len(big) - 10_000))
        q = big[a0:a0 + 10_000]
        q = synth.mutate(q, 83 + i) if i % 3 else synth.mutate(q, 83 + i, indel=0.0)
        queries.append(with_ns(q.tobytes(), 84 + i, 0.001) if i % 5 == 0 else q.tobytes())
    queries.append(rand_seq(5000, 85))
    concat, off = api.csr(queries)
    d, l, r, _ = api.query_sbwt_batch(queries, ix)
    for i in (0, 7, 30, 60):
        od, ol, orr = o.query_sbwt(queries[i])
        sl = slice(int(off[i]), int(off[i + 1]))
        assert np.array_equal(d[sl], od) and np.array_equal(l[sl], ol) and np.array_equal(r[sl], orr), i
    _, want, _ = o.matches_batch(concat, off, n_threads=8)
    outs = []
    for chunk_len in (64, 512):
        api.set_chunk_len(chunk_len)
        outs.append(api.matches_csr(concat, off, ix).copy())
    api.set_chunk_len(0)
    assert np.array_equal(outs[0], outs[1])
    assert np.array_equal(outs[0][:len(concat)], want)
    api.set_ms_flags(16)  # the fused K1 + K2b kernel on the same index
    try:
        assert np.array_equal(api.matches_csr(concat, off, ix)[:len(concat)], want)
    finally:
        api.set_ms_flags(0)
    got = api.find_batch(queries[:12], ix, api.FindOpts(max_gap_len=10))
    for i in range(12):
        assert rle_tuples(got[i]) == o.find(queries[i], 1e-7, 10), i


def test_pipelined_sub_batches_match_oracle():
    """Batches above 2 MB are cut into sub-batches on separate streams (copy/compute overlap): the result and
    the RLE offsets must be identical to the unsplit computation."""
    ref = synth.random_seq(300_000, 71)
    o = O.OracleIndex([ref.tobytes()], k=31)
    ix = api.build([ref.tobytes()], api.BuildOpts(k=31))
    concat, off = synth.gene_queries(ref, 7000, 1000, 72, snp=0.02)  # 7 MB -> 3 parts
    # ragged lengths so that part borders are not multiples of anything
    lens = np.full(7000, 1000, dtype=np.int64)
    lens[::7] = 333
    pieces = [concat[int(off[i]):int(off[i]) + int(lens[i])] for i in range(7000)]
    concat = np.concatenate(pieces)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    got = api.matches_csr(concat, off, ix)
    _, want, _ = o.matches_batch(concat, off, n_threads=8)
    assert np.array_equal(got[:len(concat)], want)
    buf, n = api.find_csr(concat, off, ix, api.FindOpts(max_gap_len=10))
    assert n == int(buf.rle_offsets[-1]) and (np.diff(buf.rle_offsets.astype(np.int64)) >= 0).all()
    for q in list(range(0, 7000, 97)) + [6999]:
        a, b_ = int(buf.rle_offsets[q]), int(buf.rle_offsets[q + 1])
        got_q = [tuple(int(getattr(buf.rle[j], f)) for f, _ in api.RleC._fields_) for j in range(a, b_)]
        assert got_q == o.find(concat[int(off[q]):int(off[q + 1])].tobytes(), max_gap_len=10), q
    small = api.FindBuffers(7000, cap=10)  # capacity error still reports the true count
    with pytest.raises(api.KboPanic) as e:
        api._check(api.load_library().kbo_find_batch(ix._h, api._p(concat, api.C.c_uint8), api._p(off, api.C.c_uint64),
                                                     7000, 1e-7, 10, small.rle, small.cap,
                                                     api._p(small.rle_offsets, api.C.c_uint64)))
    assert e.value.status == 11 and int(small.rle_offsets[-1]) == n


# ----------------------------------------------------- standalone derandomize / translate ---
def valid_ms_vector(rng, n, k):
    out = np.zeros(n, dtype=np.int64)
    cur = int(rng.integers(0, k + 1))
    for i in range(n):
        out[i] = cur
        u = rng.random()
        if u < 0.45:
            cur = min(cur + 1, k)
        elif u < 0.80:
            pass
        elif u < 0.9:
            cur = int(rng.integers(0, cur + 1))
        else:
            cur = max(cur - int(rng.integers(1, 4)), 0)
    return out


@pytest.mark.parametrize("k,thr", [(3, 2), (31, 15), (31, 22), (63, 40), (3, 3)])
def test_derandomize_translate_arbitrary_vectors(k, thr):
    rng = np.random.default_rng(k + thr)
    for n in (3, 255, 1024, 1025, 20_000):
        for kind in range(3):
            ms = (rng.integers(0, k + 1, size=n) if kind == 0 else
                  np.clip(rng.integers(thr - 2, thr + 4, size=n), 0, k) if kind == 1 else valid_ms_vector(rng, n, k))
            want = O.derandomize_ms_vec(ms, k, thr)
            got = api.derandomize_ms_vec(ms, k, thr)
            assert np.array_equal(got, want), (n, kind)
            assert api.translate_ms_vec(got, k, thr) == O.translate_ms_vec(want, k, thr)
        d = rng.integers(-5, k + 1, size=n)
        assert api.translate_ms_vec(d, k, thr) == O.translate_ms_vec(d, k, thr)
    with pytest.raises(api.KboPanic) as e:  # derandomize.rs:229
        api.derandomize_ms_vec([1, 2, k + 1], k, thr)
    assert e.value.status == 7


def test_find_batch_matches_oracle():
    ref = synth.random_seq(100_000, 41)
    o = O.OracleIndex([ref.tobytes()], k=31)
    ix = api.build([ref.tobytes()], api.BuildOpts(k=31))
    asm = synth.mutate(ref, 42, snp=0.02, indel=0.002).tobytes()
    queries = [asm[i * 3000:(i + 1) * 3000] for i in range(20)] + [rand_seq(2000, 43)]
    for gap in (0, 50):
        got = api.find_batch(queries, ix, api.FindOpts(max_gap_len=gap))
        for g_, q in zip(got, queries):
            assert rle_tuples(g_) == o.find(q, max_gap_len=gap)


def test_find_mixed_query_sizes_and_gap_lengths():
    """Thousands of tiny queries (separators in most words), queries around the word / tile / block sizes of K2b and
    K4, one long query with inserted unrelated stretches (long '-' runs), K2 as well as K2b, several max_gap_len."""
    k, p = 31, 1e-7
    ref = synth.random_seq(400_000, 51)
    o = O.OracleIndex([ref.tobytes()], k=k)
    ix = api.build([ref.tobytes()], api.BuildOpts(k=k))
    asm = synth.mutate(ref, 52, snp=0.02, indel=0.002).tobytes()
    rng = np.random.default_rng(53)
    lens = [3, 4, 31, 32, 33, 1023, 1024, 1025, 8191, 8192, 8193] + [int(x) for x in rng.integers(3, 120, 4000)] + \
           [int(x) for x in rng.integers(200, 3000, 60)]
    rng.shuffle(lens)
    queries = []
    for n in lens:
        a = int(rng.integers(0, len(asm) - n))
        queries.append(asm[a:a + n])
    long_q = bytearray(asm[100_000:300_000])
    for a, n in ((5_000, 40), (20_000, 700), (60_000, 33), (61_000, 5_000), (150_000, 1)):
        long_q[a:a + n] = rand_seq(n, a)
    queries.append(bytes(long_q))
    try:
        for gap, flags in ((0, 0), (7, 0), (100, 0), (10_000, 0), (7, 2)):  # flags 2: K2 + chars_to_masks
            api.set_ms_flags(flags)
            got = api.find_batch(queries, ix, api.FindOpts(max_error_prob=p, max_gap_len=gap))
            for i, (g_, q) in enumerate(zip(got, queries)):
                assert rle_tuples(g_) == o.find(q, p, gap), (gap, flags, i, len(q))
    finally:
        api.set_ms_flags(0)


@pytest.mark.parametrize("seed", list(range(8)))
def test_randomised_small_configurations(seed):
    """Random k, reference sizes, error probabilities and query mixes (tiny indexes, k below the prefix-table limit,
    thresholds equal to k -> K2, N-rich and unrelated queries): query_sbwt (d, l, r), matches and find vs the oracle."""
    rng = np.random.default_rng(1000 + seed)
    for _ in range(4):
        k = int(rng.integers(3, 33))
        n_ref = int(rng.choice([k + 5, 200, 1500, 20_000]))
        ref = synth.random_seq(max(n_ref, k + 1), int(rng.integers(1 << 30))).tobytes()
        revcomp = bool(rng.integers(2))
        o = O.OracleIndex([ref], k=k, add_revcomp=revcomp)
        ix = api.build([ref], api.BuildOpts(k=k, add_revcomp=revcomp))
        assert (ix.n_sets, ix.n_kmers) == (o.n_sets, o.n_kmers)
        asm = synth.mutate(np.frombuffer(ref, dtype=np.uint8), int(rng.integers(1 << 30)), snp=0.03, indel=0.003).tobytes()
        queries = []
        for _q in range(int(rng.integers(1, 25))):
            n = int(rng.integers(3, min(len(asm), 3000) + 1))
            a = int(rng.integers(0, len(asm) - n + 1))
            q = asm[a:a + n]
            u = rng.random()
            if u < 0.2:
                q = with_ns(q, int(rng.integers(1 << 30)), 0.05)
            elif u < 0.3:
                q = rand_seq(n, int(rng.integers(1 << 30)))
            queries.append(q)
        d, l, r, off = api.query_sbwt_batch(queries, ix)
        for i, q in enumerate(queries):
            od, ol, orr = o.query_sbwt(q)
            a, b_ = int(off[i]), int(off[i + 1])
            assert np.array_equal(d[a:b_].astype(np.uint64), od), (k, i)
            assert np.array_equal(l[a:b_].astype(np.uint64), ol) and np.array_equal(r[a:b_].astype(np.uint64), orr), (k, i)
        p = float(rng.choice([1e-7, 1e-3, 0.05]))
        try:
            want = [o.matches(q, p) for q in queries]
        except O.OraclePanic:
            continue  # e.g. threshold <= 1 for this (k, n_kmers, p): the reference panics, nothing to compare
        got = api.matches_batch(queries, ix, api.MatchOpts(max_error_prob=p))
        assert got == want, (k, p)
        gap = int(rng.choice([0, 3, 50]))
        try:
            want_f = [o.find(q, p, gap) for q in queries]
        except O.OraclePanic:
            continue  # format.rs:176 panics on an alignment that starts with 'R'
        got_f = api.find_batch(queries, ix, api.FindOpts(max_error_prob=p, max_gap_len=gap))
        assert [rle_tuples(g_) for g_ in got_f] == want_f, (k, p, gap)


def test_device_pointer_entry_points_match_host_entry_points():
    """kbo_matches_batch_device / kbo_find_batch_device on torch-owned device memory and stream."""
    import torch
    ref = synth.random_seq(100_000, 61)
    o = O.OracleIndex([ref.tobytes()], k=31)
    ix = api.build([ref.tobytes()], api.BuildOpts(k=31))
    concat, off = synth.gene_queries(ref, 300, 700, 62, snp=0.03)
    d_in = torch.from_numpy(concat).cuda()
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    d_out = torch.zeros(len(concat) + 16, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        api.matches_device(ix, d_in.data_ptr(), d_off.data_ptr(), off, d_out.data_ptr(), 1e-7, stream.cuda_stream)
        cap = 4096
        d_rle = torch.zeros(cap * 7, dtype=torch.int64, device="cuda")
        d_roff = torch.zeros(len(off), dtype=torch.int64, device="cuda")
        api.find_device(ix, d_in.data_ptr(), d_off.data_ptr(), off, d_rle.data_ptr(), cap, d_roff.data_ptr(), 1e-7, 25,
                        stream.cuda_stream)
    stream.synchronize()
    _, want, _ = o.matches_batch(concat, off, n_threads=4)
    assert np.array_equal(d_out.cpu().numpy()[:len(concat)], want)
    roff = d_roff.cpu().numpy()
    rle = d_rle.cpu().numpy().reshape(-1, 7)
    assert roff[-1] <= cap
    for q in range(300):
        got = [tuple(int(x) for x in rle[j]) for j in range(int(roff[q]), int(roff[q + 1]))]
        assert got == o.find(concat[int(off[q]):int(off[q + 1])].tobytes(), max_gap_len=25)
    # the same batch round-robin on six streams: the library then picks a longer chunk per call (more overlap
    # expected); results must not depend on it
    streams = [torch.cuda.Stream() for _ in range(6)]
    outs = [(torch.zeros(cap * 7, dtype=torch.int64, device="cuda"), torch.zeros(len(off), dtype=torch.int64, device="cuda"),
             torch.zeros(len(concat) + 16, dtype=torch.uint8, device="cuda")) for _ in range(18)]
    for i, (r_, ro_, ch_) in enumerate(outs):
        st = streams[i % 6]
        api.find_device(ix, d_in.data_ptr(), d_off.data_ptr(), off, r_.data_ptr(), cap, ro_.data_ptr(), 1e-7, 25, st.cuda_stream)
        api.matches_device(ix, d_in.data_ptr(), d_off.data_ptr(), off, ch_.data_ptr(), 1e-7, st.cuda_stream)
    torch.cuda.synchronize()
    for r_, ro_, ch_ in outs:
        assert torch.equal(ro_, d_roff) and torch.equal(ch_, d_out)
        assert torch.equal(r_[:int(roff[-1]) * 7], d_rle[:int(roff[-1]) * 7])


# ------------------------------------------------------ call / map with refinement ---
def test_golden_call_and_map_full():
    b = "lib.rs::doc@519"  # lib.rs:526-545
    ix = api.build([g(b, "query")], api.BuildOpts(k=20, build_select=True))
    got = api.call(ix, g(b, "reference"), api.CallOpts(0.001, api.BuildOpts(k=20, build_select=True)))
    assert [(v.query_pos, v.query_chars, v.ref_chars) for v in got] == \
        [(22, bytes([65, 71, 71]), b""), (42, bytes([84]), bytes([67])), (60, b"", bytes([67]))]
    ix3 = api.build([REF_K3], api.BuildOpts(k=3, build_select=True))  # lib.rs:647-661
    o = api.MapOpts(sbwt_build_opts=api.BuildOpts(k=3, build_select=True))
    assert list(api.map(b"GTGACTATGAGGAT", ix3, o)) == [45, 45, 45, 45, 45, 45, 45, 45, 45, 65, 71, 71, 45, 45]
    with pytest.raises(api.KboPanic) as e:  # lib.rs:729
        api.map(b"GTGACTATGAGGAT", ix3, api.MapOpts())
    assert e.value.status == 5
    b = "gap_filling.rs::fill_gaps_default_build_opts"  # gap_filling.rs:892-922 (threshold from the index)
    ixg = api.build([g(b, "query")], api.BuildOpts(build_select=True))
    got = api.map(g(b, "reference"), ixg, api.MapOpts(fill_gaps=True, call_variants=False, format=False))
    assert got == g(b, "expected")


@pytest.mark.parametrize("k,p,seed", [(31, 1e-7, 1), (51, 1e-7, 3), (63, 1e-8, 4)])
def test_map_and_call_match_oracle(k, p, seed):
    ref = synth.random_seq(200_000, 100 + seed)
    asm = synth.mutate(ref, 200 + seed, snp=0.01, indel=0.001)
    o = O.OracleIndex([asm.tobytes()], k=k)
    ix = api.build([asm.tobytes()], api.BuildOpts(k=k, build_select=True, num_threads=4))
    r = ref.tobytes()
    bo = api.BuildOpts(k=k, build_select=True)
    want_vars = o.call(r, max_error_prob=p, build_k=k)
    assert len(want_vars) > 1000 or k < 51
    got = api.call(ix, r, api.CallOpts(p, bo))
    assert [(v.query_pos, v.query_chars, v.ref_chars) for v in got] == want_vars
    for fill, callv, fmt in ((True, True, True), (True, False, False), (False, True, False)):
        want = o.map(r, max_error_prob=p, fill_gaps=fill, call_variants=callv, format=fmt, build_k=k)
        assert api.map(r, ix, api.MapOpts(p, fill, callv, fmt, bo)) == want, (fill, callv, fmt)
    # the index of the reference built once by the caller instead of per call (kbo_call_with_ref / kbo_map_with_ref)
    rix = api.build([r], bo)
    got = api.call(ix, r, api.CallOpts(p, bo), ref_index=rix)
    assert [(v.query_pos, v.query_chars, v.ref_chars) for v in got] == want_vars
    assert api.map(r, ix, api.MapOpts(p, True, True, True, bo), ref_index=rix) == o.map(r, max_error_prob=p, build_k=k)
    api.set_device_refine(False)  # the host versions of fill_gaps / access_kmer (the index mirror is read back lazily)
    try:
        got = api.call(ix, r, api.CallOpts(p, bo))
        assert [(v.query_pos, v.query_chars, v.ref_chars) for v in got] == want_vars
        assert api.map(r, ix, api.MapOpts(p, True, True, True, bo)) == o.map(r, max_error_prob=p, build_k=k)
    finally:
        api.set_device_refine(True)


GF = "gap_filling.rs::"


@pytest.mark.parametrize("device_refine", [True, False])
@pytest.mark.parametrize("name,k,p", [
    ("fill_gaps_with_clustered_changes_k51", 51, 0.0000001), ("fill_gaps_default_build_opts", 31, 0.0000001)])
def test_fill_gaps_goldens(name, k, p, device_refine):
    """gap_filling.rs:641-922 goldens whose threshold is the index's own (the C ABI takes no threshold), through
    kbo_map with fill_gaps only, on the device (refine.cuh) and on the host (refine_host.cpp)."""
    b = GF + name
    ix = api.build([g(b, "query")], api.BuildOpts(k=k, build_select=True))
    if name.endswith("k51") and api.random_match_threshold(k, ix.n_kmers, 4, p) != 23:
        pytest.skip("the golden was made with threshold 23")
    api.set_device_refine(device_refine)
    try:
        got = api.map(g(b, "reference"), ix, api.MapOpts(p, True, False, False, api.BuildOpts(k=k, build_select=True)))
    finally:
        api.set_device_refine(True)
    assert got == g(b, "expected")


def test_counters_and_launch_count():
    ref = synth.random_seq(50_000, 51)
    ix = api.build([ref.tobytes()], api.BuildOpts(k=31))
    api.set_profile_counters(True)
    n0 = api.kernel_launch_count()
    q = synth.mutate(ref, 52).tobytes()[:40_000]
    api.matches(q, ix)
    api.set_profile_counters(False)
    assert api.kernel_launch_count() - n0 == 3  # pack, MS, derandomize+translate (K1 and K2b as separate kernels)
    c = ix.ms_counters()
    assert c["bases_emitted"] == len(q) + 1
    assert c["bases_processed"] >= c["bases_emitted"]
    assert c["extend_attempts"] >= len(q) * 0.9
    assert ix.last_kernel_ms() > 0


# ------------------------------------------------------------------------ multi-GPU context ---
@pytest.mark.parametrize("devices", [[0], [0, 0, 0], "all"])
def test_multi_gpu_context_matches_and_find(devices):
    """kbo_ctx / kbo_index_set: the batch is sharded over the workers and gathered into the caller's buffers; results
    equal the oracle's (and the single-device calls').  [0, 0, 0] runs three workers on one GPU, so the sharding and the
    gather of the record counts are exercised on a one-GPU box too; "all" uses every visible GPU."""
    if devices == "all":
        if api.device_count() < 2:
            pytest.skip("needs at least two GPUs")
        devices = list(range(api.device_count()))
    ref = rand_seq(300_000, 901)
    ctx = api.Context(devices=devices)
    iset = ctx.build([ref], api.BuildOpts(k=31))
    single = api.build([ref], api.BuildOpts(k=31))
    o = O.OracleIndex([ref], k=31)
    assert (iset.n_sets, iset.n_kmers) == (o.n_sets, o.n_kmers)
    rng = np.random.default_rng(902)
    queries = []
    for i in range(400):
        a = int(rng.integers(0, len(ref) - 3000))
        q = synth.mutate(np.frombuffer(ref[a:a + int(rng.integers(3, 3000))], dtype=np.uint8), 903 + i).tobytes()
        queries.append(with_ns(q, 904 + i, 0.002) if len(q) > 2 else ref[a:a + 3])
    concat, offsets = api.csr(queries)
    _, want, _ = o.matches_batch(concat, offsets, n_threads=4)
    got = iset.matches_csr(concat, offsets)
    assert np.array_equal(got[:len(concat)], want)
    for gap in (0, 25):
        for pinned in (True, False):
            buf, n = iset.find_csr(concat, offsets, api.FindOpts(max_gap_len=gap), api.FindBuffers(len(queries), pinned=pinned))
            assert n == int(buf.rle_offsets[len(queries)])
            for i in (0, 1, 7, 133, 134, 265, 266, 399):
                got_i = [tuple(int(getattr(buf.rle[j], f)) for f, _ in api.RleC._fields_)
                         for j in range(int(buf.rle_offsets[i]), int(buf.rle_offsets[i + 1]))]
                assert got_i == o.find(queries[i], 1e-7, gap), (gap, pinned, i)
            # the gathered records are exactly the single-device call's
            sbuf, sn = api.find_csr(concat, offsets, single, api.FindOpts(max_gap_len=gap))
            assert sn == n and np.array_equal(sbuf.rle_offsets, buf.rle_offsets)
            assert bytes(memoryview(sbuf.rle))[:56 * n] == bytes(memoryview(buf.rle))[:56 * n]
    iset.close()
    ctx.close()


def test_find_many_records_exceeds_relay_estimate():
    """Page-locked outputs: records go through a device buffer sized from an estimate and are copied out densely; a query
    with far more segments than the estimate (short matches separated by junk) takes the rewrite path in kbo_job_wait."""
    ref = rand_seq(400_000, 951)
    rng = np.random.default_rng(952)
    pieces = []
    for i in range(6000):
        a = int(rng.integers(0, len(ref) - 40))
        pieces.append(ref[a:a + 33])
        pieces.append(rand_seq(4, 953 + i))
    q = b"".join(pieces)
    ix = api.build([ref], api.BuildOpts(k=31))
    o = O.OracleIndex([ref], k=31)
    want = o.find(q, 1e-7, 0)
    assert len(want) > 4 * 1 + len(q) // 64 + 1024  # beyond the device-side estimate
    concat, offsets = api.csr([q])
    buf, n = api.find_csr(concat, offsets, ix, api.FindOpts(), api.FindBuffers(1, cap=len(q), pinned=True))
    assert n == len(want)
    got = [tuple(int(getattr(buf.rle[j], f)) for f, _ in api.RleC._fields_) for j in range(n)]
    assert got == want
    job = api.find_submit(concat, offsets, ix, api.FindOpts(), api.FindBuffers(1, cap=len(q), pinned=True))
    assert job.wait() == len(want)


@pytest.mark.parametrize("fmt", [True, False])
def test_map_unrefined_on_device_matches_oracle(fmt):
    """kbo_map_unrefined: translate and relative_to_ref (format.rs:266-287) on the device; also what kbo_map does when both
    refinements are switched off."""
    ref = rand_seq(120_000, 961)
    asm = synth.mutate(np.frombuffer(ref, dtype=np.uint8), 962).tobytes()
    q = with_ns(ref, 963, 0.001)
    ix = api.build([asm], api.BuildOpts(k=31, build_select=True))
    o = O.OracleIndex([asm], k=31)
    want = o.map(q, fill_gaps=False, call_variants=False, format=fmt)
    assert api.map_unrefined(q, ix, format=fmt) == want
    assert api.map(q, ix, api.MapOpts(fill_gaps=False, call_variants=False, format=fmt)) == want


def test_per_index_tuning_knobs():
    """kbo_index_set_tuning: two indexes in one process run with different chunk lengths / kernel paths; results are the
    oracle's for both, and -1 returns a knob to the process-wide default."""
    ref = rand_seq(150_000, 971)
    a, b = api.build([ref], api.BuildOpts(k=31)), api.build([ref], api.BuildOpts(k=31))
    o = O.OracleIndex([ref], k=31)
    concat, off = synth.gene_queries(np.frombuffer(ref, dtype=np.uint8), 300, 700, 972, snp=0.02)
    _, want, _ = o.matches_batch(concat, off, n_threads=4)
    api.set_index_tuning(a, api.TUNE_CHUNK_LEN, 96)
    api.set_index_tuning(a, api.TUNE_MS_FLAGS, 16)       # fused kernel on index a only
    api.set_index_tuning(b, api.TUNE_PIPELINE_PARTS, 3)  # three sub-batches per host call on index b only
    assert np.array_equal(api.matches_csr(concat, off, a)[:len(concat)], want)
    assert np.array_equal(api.matches_csr(concat, off, b)[:len(concat)], want)
    n0 = api.kernel_launch_count()
    api.matches_csr(concat, off, a)
    assert api.kernel_launch_count() - n0 == 2           # K0 + the fused kernel
    api.set_index_tuning(a, api.TUNE_MS_FLAGS, -1)
    n0 = api.kernel_launch_count()
    api.matches_csr(concat, off, a)
    assert api.kernel_launch_count() - n0 == 3           # K0, K1, K2b again
