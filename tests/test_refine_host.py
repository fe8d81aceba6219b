"""CPU checks of the product's refinement layer -- the host version (kbo_b200/csrc/refine_host.cpp: call_variants,
fill_gaps, add_variants) and the device version (kbo_b200/csrc/refine.cuh: gap_list / fill_gaps / access_kmers kernels,
run through the CUDA-on-CPU shim) -- driven by emulated-kernel matching statistics, against the reference's golden
vectors and the oracle.  The same entry points run on the GPU in tests/test_gpu_parity.py."""
import json
import os

import numpy as np
import pytest

import emu_lib as E
import oracle_lib as O
from kbo_b200 import synth

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))
REF_K3 = b"AAAGAACCA-TCAGGGCG"


def g(block, var):
    return GOLD[block][var].encode()


@pytest.fixture(params=["host", "device"], autouse=True)
def refine_mode(request):
    E.set_device_refine(request.param == "device")
    yield request.param
    E.set_device_refine(False)


def thr_of(e, p):
    return O.random_match_threshold(e.k, e.n_kmers, 4, p)


def test_call_doc_k20():
    b = "lib.rs::doc@519"  # lib.rs:526-545
    e = E.EmuIndex.build([g(b, "query")], k=20)
    got = e.call(g(b, "reference"), thr_of(e, 0.001), 20)
    assert got == [(22, bytes([65, 71, 71]), b""), (42, bytes([84]), bytes([67])), (60, b"", bytes([67]))]


VC = "variant_calling.rs::"


@pytest.mark.parametrize("name,k,expected", [
    ("test_single_base_substitution", 20, [(49, b"T", b"A")]),
    ("test_multi_base_substitution", 30, [(29, b"GCG", b"AA")]),
    ("test_multi_base_insertion_non_overlap_case", 30, [(29, b"GCG", b"")]),
    ("test_multi_base_insertion_overlap_case", 30, [(31, b"AAAA", b"")]),
    ("test_single_base_insertion_non_overlap_case", 20, [(50, b"G", b"")]),
    ("test_single_base_insertion_overlap_case", 20, [(50, b"A", b"")]),
    ("test_single_base_deletion_non_overlap_case", 20, [(50, b"", b"G")]),
    ("test_single_base_deletion_overlap_case", 20, [(51, b"", b"T")]),
    ("test_multi_base_deletion_non_overlap_case", 30, [(29, b"", b"GCG")]),
    ("test_multi_base_deletion_overlap_case", 30, [(31, b"", b"AAAA")]),
    ("test_variants_in_same_query", 20, [(24, b"", b"G"), (41, b"C", b"T"), (59, b"C", b"")]),
])
def test_variant_calling_cases(name, k, expected):
    # variant_calling.rs:304-454: call_variants(sbwt_ref = index(reference), sbwt_query = index(query), query)
    # == kbo::call(index(reference), query) with the roles the reference's helper uses
    query, reference = g(VC + name, "query"), g(VC + name, "reference")
    e = E.EmuIndex.build([reference], k=k)
    assert e.call(query, thr_of(e, 0.001), k) == expected


def test_map_doc_full_k3():
    e = E.EmuIndex.build([REF_K3], k=3)  # lib.rs:647-661
    got = e.map(g("lib.rs::doc@641", "reference"), thr_of(e, 1e-7), 1e-7, build_k=3)
    assert list(got) == [45, 45, 45, 45, 45, 45, 45, 45, 45, 65, 71, 71, 45, 45]


GF = "gap_filling.rs::"


@pytest.mark.parametrize("name,k,thr,p", [
    ("fill_gaps", 7, 3, 0.001), ("fill_gaps_with_clustered_changes", 9, 3, 0.001),
    ("fill_gaps_with_clustered_changes2", 9, 3, 0.001), ("fill_gaps_left_extend_short", 9, 3, 0.001),
    ("fill_gaps_left_extend_long", 9, 4, 0.001), ("doc@401", 9, 4, 0.001),
    ("fill_gaps_with_clustered_changes_k51", 51, 23, 0.0000001), ("fill_gaps_default_build_opts", 31, None, 0.0000001),
])
def test_fill_gaps_goldens(name, k, thr, p, refine_mode):
    # gap_filling.rs:641-922: map(fill_gaps only, unformatted) with the test's threshold
    b = GF + name
    e = E.EmuIndex.build([g(b, "query")], k=k)
    if thr is None:
        thr = thr_of(e, p)
    n0 = E.device_gap_count()
    got = e.map(g(b, "reference"), thr, p, fill_gaps=True, call_variants=False, format=False, build_k=k)
    assert got == g(b, "expected")
    assert (E.device_gap_count() > n0) == (refine_mode == "device")  # the kernels really ran (every golden has a gap)


@pytest.mark.parametrize("name", ["add_variants", "add_variants_multi_base_substitution",
                                  "add_variants_multi_base_substitution_all_same",
                                  "add_variants_clustered_substitutions", "doc@312"])
def test_add_variants_goldens(name):
    # translate.rs:535-676: translation with threshold 10, variants called with p = 0.001
    b = "translate.rs::" + name
    e = E.EmuIndex.build([g(b, "query")], k=20)
    got = e.map(g(b, "reference"), 10, 0.001, fill_gaps=False, call_variants=True, format=False, build_k=20,
                call_thr=thr_of(e, 0.001))
    assert got == g(b, "expected")


@pytest.mark.parametrize("k,p,seed", [(31, 1e-7, 1), (20, 1e-3, 2), (51, 1e-7, 3), (63, 1e-8, 4)])
def test_map_and_call_match_oracle_on_synthetic(k, p, seed):
    ref = synth.random_seq(60_000, 100 + seed)
    asm = synth.mutate(ref, 200 + seed, snp=0.01, indel=0.001)
    o = O.OracleIndex([asm.tobytes()], k=k)
    e = E.EmuIndex.build([asm.tobytes()], k=k)
    thr = thr_of(e, p)
    r = ref.tobytes()
    want_vars = o.call(r, max_error_prob=p, build_k=k)
    # a variant resolves only if k - d - 1 >= d bases precede it in the k-mer (d = threshold): with 60 kbp indexes
    # k = 31 (d = 19) and k = 20 (d = 12) give no calls in the reference either; k = 51 and 63 do
    assert len(want_vars) > 100 or k < 51
    assert e.call(r, thr, k) == want_vars
    for fill, callv, fmt in ((True, True, True), (True, False, False), (False, True, False), (False, False, True)):
        want = o.map(r, max_error_prob=p, fill_gaps=fill, call_variants=callv, format=fmt, build_k=k)
        assert e.map(r, thr, p, fill_gaps=fill, call_variants=callv, format=fmt, build_k=k) == want, (fill, callv, fmt)
