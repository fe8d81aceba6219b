"""Pins the CPU oracle (oracle/) against every golden vector the reference's own
unit tests and doctests hold for the MS -> derandomize -> translate path and its
callers.  CPU only.

Sequence literals come from tests/golden/reference_vectors.json (extracted from
/root/reference by tests/golden/extract_reference_goldens.py); scalar expectations
are transcribed here with the reference file:line they come from.
"""
import json
import math
import os

import numpy as np
import pytest

import oracle_lib as O

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))
REF_K3 = b"AAAGAACCA-TCAGGGCG"  # index.rs:265 / lib.rs:601 (nested vec literal, transcribed by hand)


def g(block, var):
    return GOLD[block][var].encode()


# ---------------------------------------------------------------- index / MS ---
def test_worked_example_k3_structure():
    # SURVEY 8c worked example: k=3, n_kmers=13, n_sets=16, P in colex order
    ix = O.OracleIndex([REF_K3], k=3)
    assert ix.n_kmers == 13 and ix.n_sets == 16
    P = [ix.access_kmer(i).decode() for i in range(16)]
    assert P == ["$$$", "AAA", "GAA", "CCA", "TCA", "AGA", "AAC", "ACC", "GGC", "$TC", "AAG", "CAG", "GCG", "AGG",
                 "GGG", "$$T"]


def test_build_and_query_sbwt():
    # index.rs:264-274 (and doctest index.rs:229-240)
    ix = O.OracleIndex([REF_K3], k=3)
    d, l, r = ix.query_sbwt(g("index.rs::build_and_query_sbwt", "query"))
    assert d.tolist() == [1, 2, 2, 3, 2, 2, 3, 2, 1, 2, 3, 1, 1, 1, 2, 3, 1, 2]


def test_query_sbwt_empty_panics():
    ix = O.OracleIndex([REF_K3], k=3)
    with pytest.raises(O.OraclePanic):  # index.rs:248
        ix.query_sbwt(b"")


# -------------------------------------------------------------- derandomize ---
def test_log_rm_max_cdf():
    # derandomize.rs:298-304
    expected = [-1306319.1078024083, -318761.2492719044, -79220.9269610741, -19776.1823255263, -4942.2344281681,
                -1235.4454790664, -308.8543003470, -77.2131332649, -19.3032557026, -4.8258121998, -1.2064529421,
                -0.3016132288, -0.0754033068, -0.0188508267, -0.0047127067, -0.0011781767, -0.0002945442,
                -0.0000736360, -0.0000184090, -0.0000046023, -0.0000011506, -0.0000002876, -0.0000000719,
                -0.0000000180, -0.0000000045, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]
    for t in range(1, 32):
        assert abs(O.log_rm_max_cdf(t, 4, 20240921) - expected[t - 1]) < 1e-8


def test_random_match_threshold():
    # derandomize.rs:307-314
    expected = [15, 18, 22, 25, 28]
    for i in range(1, 6):
        assert O.random_match_threshold(31, 20240921, 4, math.pow(0.01, float(i))) == expected[i - 1]


def test_random_match_threshold_asserts():
    for args in [(0, 1, 4, 0.1), (31, 0, 4, 0.1), (31, 1, 0, 0.1), (31, 1, 4, 1.5), (31, 1, 4, 0.0)]:
        with pytest.raises(O.OraclePanic):  # derandomize.rs:133-137
            O.random_match_threshold(*args)


def test_derandomize_ms_val_cases():
    assert O.derandomize_ms_val(3, 3, 2, 3) == 3    # derandomize.rs:317-328
    assert O.derandomize_ms_val(2, -1, 2, 3) == -2  # :331-342
    assert O.derandomize_ms_val(3, -1, 2, 3) == 3   # :345-356
    assert O.derandomize_ms_val(3, -1, 2, 4) == 3   # :359-370


def test_derandomize_ms_vec():
    # derandomize.rs:373-379
    got = O.derandomize_ms_vec([1, 2, 2, 3, 2, 2, 3, 2, 1, 2, 3, 1, 1, 1, 2, 3, 1, 2], 3, 2)
    assert got.tolist() == [0, 1, 2, 3, 1, 2, 3, 0, 1, 2, 3, -1, 0, 1, 2, 3, -1, 0]


def test_derandomize_asserts():
    with pytest.raises(O.OraclePanic):  # derandomize.rs:275 threshold > 1
        O.derandomize_ms_vec([1, 2, 3], 3, 1)
    with pytest.raises(O.OraclePanic):  # derandomize.rs:276 len > 2
        O.derandomize_ms_vec([1, 2], 3, 2)


# ---------------------------------------------------------------- translate ---
def test_translate_ms_val_cases():
    assert O.translate_ms_val(3, 1, 2, 2) == ("R", "R")   # translate.rs:396-410
    assert O.translate_ms_val(3, 1, 3, 2) == ("R", "R")   # :413-427
    assert O.translate_ms_val(0, 1, 3, 2) == ("X", " ")   # :430-444, :447-464
    assert O.translate_ms_val(-1, 0, 3, 2) == ("-", " ")  # :467-481
    assert O.translate_ms_val(1, 2, 3, 2) == ("M", " ")   # :484-498


def test_translate_ms_vec():
    # translate.rs:501-515
    got = O.translate_ms_vec([0, 1, 2, 3, 1, 2, 3, 0, 1, 2, 3, -1, 0, 1, 2, 3, -1, 0], 3, 2)
    assert got == g("translate.rs::translate_ms_vec", "expected") == b"XMMRRMMXMMM--MMM--"


def test_translate_ms_vec_with_recombination():
    # translate.rs:518-532
    got = O.translate_ms_vec([1, 2, 3, 1, 2, 3, 3, 3, 3, 1, 2, 3], 3, 2)
    assert got == b"MMRRMMMMRRMM"


def test_format_doc_chain_k4():
    # format.rs:229-247: derand MS -> translate -> relative_to_ref, k=4 thr=3
    derand = [1, 2, 3, 4, -1, 0, 1, 2, 3, 4, 4, 4, 4, 0, 1, 2, 3, 4, 4, 4, 4]
    ms = [1, 2, 3, 4, 1, 2, 3, 3, 3, 4, 4, 4, 4, 3, 1, 2, 3, 4, 4, 4, 4]
    assert O.derandomize_ms_vec(ms, 4, 3).tolist() == derand  # vectors in the comment block format.rs:237-238
    tr = O.translate_ms_vec(derand, 4, 3)
    assert tr == b"MMMM--MMMMMMMXMMMMMMM"
    assert O.relative_to_ref(g("format.rs::doc@225", "reference"), tr) == g("format.rs::doc@225", "expected")


def test_relative_to_ref_refined():
    # format.rs:251-263
    b = "format.rs::doc@250"
    assert O.relative_to_ref(g(b, "reference"), g(b, "refined")) == g(b, "expected")


# ------------------------------------------------------------------- format ---
def test_run_lengths_doc():
    # format.rs:77-95
    got = O.run_lengths(g("format.rs::doc@77", "input"))
    assert got == [(0, 11, 9, 2, 1, 0, 0), (13, 16, 3, 0, 0, 0, 0)]


def test_run_lengths_gapped_doc():
    # format.rs:122-140
    got = O.run_lengths_gapped(g("format.rs::doc@122", "input"), 3)
    assert got == [(0, 16, 12, 2, 1, 2, 1)]


def test_run_lengths_512():
    # format.rs:295-330
    got = O.run_lengths(g("format.rs::run_lengths", "input"))
    assert got == [(5, 33, 28, 0, 0, 0, 0), (81, 207, 126, 0, 0, 0, 0), (372, 423, 51, 0, 0, 0, 0),
                   (487, 512, 25, 0, 0, 0, 0)]


def test_run_lengths_leading_R_panics():
    with pytest.raises(O.OraclePanic):  # format.rs:176 aln[i - 1] with i == 0
        O.run_lengths(b"RRMM")


# ---------------------------------------------------------------- lib.rs API ---
def test_matches_doc_k3():
    # lib.rs:600-610 (threshold degenerates to k = 3)
    ix = O.OracleIndex([REF_K3], k=3)
    assert O.random_match_threshold(3, ix.n_kmers, 4, 1e-7) == 3
    assert ix.matches(g("lib.rs::doc@594", "query")) == b"---------MMM--"


def test_map_doc_full_k3():
    # lib.rs:647-661
    ix = O.OracleIndex([REF_K3], k=3)
    got = ix.map(g("lib.rs::doc@641", "reference"), build_k=3)
    assert list(got) == [45, 45, 45, 45, 45, 45, 45, 45, 45, 65, 71, 71, 45, 45]


def test_map_doc_no_refinement_k7():
    # lib.rs:670-689 and :698-717
    b = "lib.rs::doc@664"
    ix = O.OracleIndex([g(b, "query")], k=7)
    got = ix.map(g(b, "reference"), max_error_prob=0.1, fill_gaps=False, call_variants=False, build_k=7)
    assert got == g(b, "expected")
    b2 = "lib.rs::doc@692"
    got = ix.map(g(b2, "reference"), max_error_prob=0.1, fill_gaps=False, call_variants=False, format=False, build_k=7)
    assert got == g(b2, "expected")


def test_map_k_mismatch_panics():
    ix = O.OracleIndex([REF_K3], k=3)
    with pytest.raises(O.OraclePanic):  # lib.rs:729
        ix.map(b"GTGACTATGAGGAT", build_k=31)


def test_find_doc_k31():
    # lib.rs:786-806; SURVEY 4: n_kmers 1176, n_sets 1237, threshold 16
    b = "lib.rs::doc@779"
    ix = O.OracleIndex([g(b, "gene1"), g(b, "gene2_rc")], k=31)
    assert (ix.n_kmers, ix.n_sets) == (1176, 1237)
    assert O.random_match_threshold(31, ix.n_kmers, 4, 1e-7) == 16
    got = ix.find(g(b, "query"), max_gap_len=50)
    assert got == [(0, 513, 512, 1, 0, 0, 0), (593, 1340, 709, 0, 0, 38, 3)]


def test_call_doc_k20():
    # lib.rs:526-545
    b = "lib.rs::doc@519"
    ix = O.OracleIndex([g(b, "query")], k=20)
    got = ix.call(g(b, "reference"), max_error_prob=0.001, build_k=20)
    assert got == [(22, bytes([65, 71, 71]), b""), (42, bytes([84]), bytes([67])), (60, b"", bytes([67]))]


# ---------------------------------------------------------- variant_calling ---
def run_variant_calling(query, reference, k, p):
    # variant_calling.rs:304-309
    ix_ref = O.OracleIndex([reference], k=k)
    ix_query = O.OracleIndex([query], k=k)
    return O.call_variants(ix_ref, ix_query, query, p)


VC = "variant_calling.rs::"


@pytest.mark.parametrize("name,k,expected", [
    ("test_single_base_substitution", 20, [(49, b"T", b"A")]),                    # :312-321
    ("test_multi_base_substitution", 30, [(29, b"GCG", b"AA")]),                  # :324-335
    ("test_multi_base_insertion_non_overlap_case", 30, [(29, b"GCG", b"")]),      # :338-347
    ("test_multi_base_insertion_overlap_case", 30, [(31, b"AAAA", b"")]),         # :350-359
    ("test_single_base_insertion_non_overlap_case", 20, [(50, b"G", b"")]),       # :362-373
    ("test_single_base_insertion_overlap_case", 20, [(50, b"A", b"")]),           # :376-387
    ("test_single_base_deletion_non_overlap_case", 20, [(50, b"", b"G")]),        # :390-401
    ("test_single_base_deletion_overlap_case", 20, [(51, b"", b"T")]),            # :404-415
    ("test_multi_base_deletion_non_overlap_case", 30, [(29, b"", b"GCG")]),       # :418-427
    ("test_multi_base_deletion_overlap_case", 30, [(31, b"", b"AAAA")]),          # :430-438
    ("test_variants_in_same_query", 20, [(24, b"", b"G"), (41, b"C", b"T"), (59, b"C", b"")]),  # :441-454
])
def test_variant_calling_cases(name, k, expected):
    got = run_variant_calling(g(VC + name, "query"), g(VC + name, "reference"), k, 0.001)
    assert got == expected


def test_call_variants_doc():
    # variant_calling.rs:227-246
    b = VC + "doc@220"
    got = run_variant_calling(g(b, "query"), g(b, "reference"), 20, 0.001)
    assert got == [(22, b"", b"AGG"), (39, b"C", b"T"), (57, b"C", b"")]


class XorShift128Plus:
    """`random` crate 0.14 `Default` source (Xorshift128+), used by variant_calling.rs:467."""

    def __init__(self, seed):
        self.s = [seed[0] & (2 ** 64 - 1), seed[1] & (2 ** 64 - 1)]

    def read_u64(self):
        M = 2 ** 64 - 1
        x, y = self.s
        self.s[0] = y
        x ^= (x << 23) & M
        x ^= x >> 17
        x ^= y ^ (y >> 26)
        self.s[1] = x
        return (x + y) & M


def test_long_generated_testcase():
    # variant_calling.rs:467-553: 100 000 bases, a planted variant every 25 bases, k=63, p=1e-8;
    # every call must equal the planted variant at the same index.
    rng = XorShift128Plus([123412, 121232])
    nt = lambda: b"ACGT"[rng.read_u64() % 4]
    reference, query, true_variants = bytearray(), bytearray(), []
    n, spacing, k, p = 100_000, 25, 63, 1e-8
    for i in range(n):
        if i > spacing and i < n - spacing and i % spacing == 0:
            qlen, rlen = rng.read_u64() % 4, rng.read_u64() % 4
            while qlen == 0 and rlen == 0:
                qlen, rlen = rng.read_u64() % 4, rng.read_u64() % 4
            qv = bytearray(nt() for _ in range(qlen))
            rv = bytearray(nt() for _ in range(rlen))
            while qv and rv and (qv[0] == rv[0] or qv[-1] == rv[-1]):
                qv[-1] = nt()
                qv[0] = nt()
            true_variants.append((len(query), bytes(qv), bytes(rv)))
            reference += rv
            query += qv
            ins = rv if (not qv and rv) else (qv if (qv and not rv) else None)
            if ins is not None:
                c = nt()
                while c == ins[0] or c == ins[-1]:
                    c = nt()
                query.append(c)
                reference.append(c)
        else:
            c = nt()
            query.append(c)
            reference.append(c)
    calls = run_variant_calling(bytes(query), bytes(reference), k, p)
    n_correct = sum(1 for a, b in zip(calls, true_variants) if a == b)
    assert len(calls) > 3000
    assert n_correct == len(calls)


# ---------------------------------------------------- translate::add_variants ---
@pytest.mark.parametrize("name", ["add_variants", "add_variants_multi_base_substitution",
                                  "add_variants_multi_base_substitution_all_same",
                                  "add_variants_clustered_substitutions", "doc@312"])
def test_add_variants(name):
    # translate.rs:535-676 (tests) and :318-347 (doctest): k=20, threshold=10, p=0.001
    b = "translate.rs::" + name
    k, thr = 20, 10
    ix = O.OracleIndex([g(b, "query")], k=k)
    d, _, _ = ix.query_sbwt(g(b, "reference"))
    tr = O.translate_ms_vec(O.derandomize_ms_vec(d, k, thr), k, thr)
    variants = ix.call(g(b, "reference"), max_error_prob=0.001, build_k=k)
    assert O.add_variants(tr, variants) == g(b, "expected")


# --------------------------------------------------------------- gap_filling ---
GF = "gap_filling.rs::"


def test_nearest_unique_context():
    # gap_filling.rs:535-564
    b = GF + "nearest_unique_context"
    ix = O.OracleIndex([g(b, "query")], k=9)
    assert ix.nearest_unique_context(g(b, "reference"), 11, 16) == (16, b"CAGACAGCT")
    # doctest gap_filling.rs:106-124
    b = GF + "doc@91"
    ix = O.OracleIndex([g(b, "query")], k=7)
    assert ix.nearest_unique_context(g(b, "reference"), 8, 14) == (12, b"AGGCTGC")


def test_left_extend_kmer():
    # gap_filling.rs:567-600: search + access_kmer + left_extend_kmer, k=6
    b = GF + "left_extend_kmer"
    ix = O.OracleIndex([g(b, "sequence")], k=6)
    l, r = ix.search(g(b, "query"))
    kmer = ix.access_kmer(l)
    assert ix.left_extend_kmer(kmer, 8) == g(b, "expected")
    # doctest gap_filling.rs:185-203, k=7
    b = GF + "doc@170"
    ix = O.OracleIndex([g(b, "sequence")], k=7)
    assert ix.left_extend_kmer(g(b, "kmer"), 5) == g(b, "expected")


def test_left_extend_over_gap():
    # gap_filling.rs:603-638, k=5
    b = GF + "left_extend_over_gap"
    ix = O.OracleIndex([g(b, "query")], k=5)
    assert ix.left_extend_over_gap(g(b, "reference"), 3, 3, 4, 7, 4) == g(b, "expected")
    # doctest gap_filling.rs:274-292, k=9
    b = GF + "doc@259"
    ix = O.OracleIndex([g(b, "query")], k=9)
    assert ix.left_extend_over_gap(g(b, "reference"), 4, 4, 5, 12, 6) == g(b, "expected")


@pytest.mark.parametrize("name,k,thr,p", [
    ("fill_gaps", 7, 3, 0.001),                                  # gap_filling.rs:641-683
    ("fill_gaps_with_clustered_changes", 9, 3, 0.001),           # :686-727
    ("fill_gaps_with_clustered_changes2", 9, 3, 0.001),          # :730-771
    ("fill_gaps_left_extend_short", 9, 3, 0.001),                # :774-815
    ("fill_gaps_left_extend_long", 9, 4, 0.001),                 # :818-858
    ("doc@401", 9, 4, 0.001),                                    # doctest :418-441
    ("fill_gaps_with_clustered_changes_k51", 51, 23, 0.0000001),  # :861-889
    ("fill_gaps_default_build_opts", 31, None, 0.0000001),       # :892-922 (threshold from the index)
])
def test_fill_gaps(name, k, thr, p):
    b = GF + name
    ix = O.OracleIndex([g(b, "query")], k=k)
    if thr is None:
        thr = O.random_match_threshold(k, ix.n_kmers, 4, p)
        assert thr == 15  # "31(default opts, thr 15)"
    d, _, _ = ix.query_sbwt(g(b, "reference"))
    tr = O.translate_ms_vec(O.derandomize_ms_vec(d, k, thr), k, thr)
    assert ix.fill_gaps(tr, g(b, "reference"), thr, p) == g(b, "expected")
