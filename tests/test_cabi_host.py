"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol that
include/kbo_b200.h declares, its host-side functions reproduce the reference's golden vectors, and
compute entry points FAIL LOUDLY without a GPU (no CPU fallback)."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

from kbo_b200 import api, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build_library()
    return api.load_library()


def header_functions():
    src = open(os.path.join(ROOT, "include", "kbo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kbo_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert sorted(api.EXPORTED_SYMBOLS) == names


def test_no_torch_types_in_header():
    src = open(os.path.join(ROOT, "include", "kbo_b200.h")).read()
    assert "torch" not in src.lower().replace("no c++/torch types", "") and "at::" not in src and "std::" not in src


def test_log_rm_max_cdf_golden(lib):
    # derandomize.rs:298-304
    expected = [-1306319.1078024083, -318761.2492719044, -79220.9269610741, -19776.1823255263, -4942.2344281681,
                -1235.4454790664, -308.8543003470, -77.2131332649, -19.3032557026, -4.8258121998, -1.2064529421,
                -0.3016132288, -0.0754033068, -0.0188508267, -0.0047127067, -0.0011781767, -0.0002945442,
                -0.0000736360, -0.0000184090, -0.0000046023, -0.0000011506, -0.0000002876, -0.0000000719,
                -0.0000000180, -0.0000000045, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]
    for t in range(1, 32):
        assert abs(api.log_rm_max_cdf(t, 4, 20240921) - expected[t - 1]) < 1e-8


def test_random_match_threshold_golden(lib):
    # derandomize.rs:307-314
    for i, want in enumerate([15, 18, 22, 25, 28], start=1):
        assert api.random_match_threshold(31, 20240921, 4, math.pow(0.01, float(i))) == want
    assert api.random_match_threshold(3, 13, 4, 1e-7) == 3      # lib.rs:600-610 degenerate case
    assert api.random_match_threshold(31, 1176, 4, 1e-7) == 16  # lib.rs:786-806


def test_threshold_precondition_codes(lib):
    for args, status in [((0, 1, 4, 0.1), 4), ((31, 0, 4, 0.1), 7), ((31, 1, 0, 0.1), 7), ((31, 1, 4, 1.5), 6),
                         ((31, 1, 4, 0.0), 6)]:
        with pytest.raises(api.KboPanic) as e:
            api.random_match_threshold(*args)
        assert e.value.status == status


def test_run_lengths_goldens(lib):
    t = lambda rl: [(r.start, r.end, r.matches, r.mismatches, r.jumps, r.gap_bases, r.gap_opens) for r in rl]
    aln = b"XMMRRMMXMMM--MMM--"
    assert t(api.run_lengths(aln)) == [(0, 11, 9, 2, 1, 0, 0), (13, 16, 3, 0, 0, 0, 0)]  # format.rs:77-95
    assert t(api.run_lengths_gapped(aln, 3)) == [(0, 16, 12, 2, 1, 2, 1)]                  # format.rs:122-140
    with pytest.raises(api.KboPanic) as e:                                                 # format.rs:176
        api.run_lengths(b"RRMM")
    assert e.value.status == 12


def test_run_lengths_vs_oracle_random(lib):
    import oracle_lib as O
    rng = np.random.default_rng(5)
    alphabet = np.frombuffer(b"MMMMMM--XRIDACGT ", dtype=np.uint8)
    for n in (1, 2, 17, 300, 2000):
        for gap in (0, 1, 3, 50):
            for _ in range(20):
                a = alphabet[rng.integers(0, len(alphabet), size=n)].tobytes()
                try:
                    want = O.run_lengths_gapped(a, gap)
                except O.OraclePanic:
                    with pytest.raises(api.KboPanic):
                        api.run_lengths_gapped(a, gap)
                    continue
                got = api.run_lengths_gapped(a, gap)
                assert [(r.start, r.end, r.matches, r.mismatches, r.jumps, r.gap_bases, r.gap_opens)
                        for r in got] == want


def test_relative_to_ref_goldens(lib):
    # format.rs:251-263
    assert api.relative_to_ref(b"AAAGAACCATCAGGGCG", b"CMMR--RMMMMMMMM--") == b"CAAG--CCATCAGGG--"
    assert api.relative_to_ref(b"TTGATTGGCTGGGCAGAGCTG", b"MMMM--MMMMMMMXMMMMMMM") == b"TTGA--GGCTGGG-AGAGCTG"


def test_relative_to_ref_all_bytes_and_lengths(lib):
    """format.rs:270-286 on every alignment byte value and on lengths around the 16-byte vector width."""
    rng = np.random.default_rng(7)
    for n in (1, 15, 16, 17, 31, 32, 33, 1000, 4099):
        ref = rng.integers(0, 256, n, dtype=np.uint8)
        aln = rng.integers(0, 256, n, dtype=np.uint8)
        mix = rng.random(n) < 0.7
        aln[mix] = rng.choice(np.frombuffer(b"MRIXD-ACGTN", dtype=np.uint8), int(mix.sum()))
        want = bytes(r if a in b"MRI" else (ord("-") if a in b"XD-" else a) for r, a in zip(ref.tolist(), aln.tolist()))
        assert api.relative_to_ref(ref.tobytes(), aln.tobytes()) == want, n


def test_compute_fails_loudly_without_gpu(lib):
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(api.KboPanic) as e:
        api.build([b"ACGTACGTACGT"], api.BuildOpts(k=3))
    assert e.value.status == 8  # KBO_ERR_CUDA: there is no CPU fallback
    with pytest.raises(api.KboPanic) as e:
        api.derandomize_ms_vec([1, 2, 3, 3], 3, 2)
    assert e.value.status == 8
    with pytest.raises(api.KboPanic) as e:
        api.translate_ms_vec([1, 2, 3, 3], 3, 2)
    assert e.value.status == 8


def test_argument_validation_before_any_device_work(lib):
    with pytest.raises(api.KboPanic) as e:
        api.derandomize_ms_vec([1, 2, 3], 3, 1)
    assert e.value.status == 2  # derandomize.rs:275
    with pytest.raises(api.KboPanic) as e:
        api.derandomize_ms_vec([1, 2], 3, 2)
    assert e.value.status == 3  # derandomize.rs:276
    with pytest.raises(api.KboPanic) as e:
        api.translate_ms_vec([1, 2], 3, 2)
    assert e.value.status == 3  # translate.rs:270
    with pytest.raises(api.KboPanic) as e:
        api.build([], api.BuildOpts(k=3))
    assert e.value.status == 1  # index.rs:60
    with pytest.raises(api.KboPanic) as e:
        api.build([b"ACGT"], api.BuildOpts(k=65))
    assert e.value.status == 4


def test_index_from_parts_validates_its_arrays(lib):
    """kbo_index_from_parts rejects arrays K1 could not walk safely: LCS[0] != 0, an LCS value >= k,
    rows whose set bits are not n_sets - 1 -- before any device work, so also without a GPU."""
    import oracle_lib as O
    o = O.OracleIndex([b"AAAGAACCA-TCAGGGCG"], k=3)
    rows, lcs = o.rows(), o.lcs()
    bad = lcs.copy(); bad[0] = 1
    with pytest.raises(api.KboPanic) as e:
        api.index_from_parts(3, o.n_sets, o.n_kmers, rows, bad)
    assert e.value.status == 7 and "LCS" in str(e.value)
    bad = lcs.copy(); bad[5] = 3
    with pytest.raises(api.KboPanic) as e:
        api.index_from_parts(3, o.n_sets, o.n_kmers, rows, bad)
    assert e.value.status == 7
    rows_bad = [r.copy() for r in rows]; rows_bad[0][0] ^= np.uint64(1 << 7)
    with pytest.raises(api.KboPanic) as e:
        api.index_from_parts(3, o.n_sets, o.n_kmers, rows_bad, lcs)
    assert e.value.status == 7 and "set bits" in str(e.value)
    with pytest.raises(api.KboPanic) as e:
        api.index_from_parts(128, o.n_sets, o.n_kmers, rows, lcs)
    assert e.value.status == 4


def test_new_entry_points_check_their_arguments(lib):
    h = C.c_void_p()
    assert lib.kbo_job_wait(None, None) == 7
    assert lib.kbo_find_batch_submit(None, None, None, 0, 1e-7, 0, None, 0, None, C.byref(h)) == 7
    assert lib.kbo_index_set_tuning(None, 0, 1) == 7
    assert lib.kbo_find_batch_multi(None, None, None, 0, 1e-7, 0, None, 0, None) == 7
    assert lib.kbo_matches_batch_multi(None, None, None, 0, 1e-7, None) == 7
    assert lib.kbo_ctx_n_gpus(None) == 0
    assert lib.kbo_index_set_get(None, 0) is None
    if api.device_count() == 0:
        assert lib.kbo_ctx_create(1, None, C.byref(h)) == 8  # no device: KBO_ERR_CUDA, no fallback


def _fnv1a(data):
    h = 1469598103934665603
    for b in data:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def write_index_files(prefix, k, n_sets, n_kmers, rows, lcs):
    """The documented layout of `<prefix>.sbwt` / `<prefix>.lcs` (include/kbo_b200.h), written independently of the library."""
    import struct
    body = struct.pack("<Q", 12) + b"SubsetMatrix" + b"KBOB200\0" + struct.pack("<IIQQ", 1, k, n_sets, n_kmers)
    for r in rows:
        body += np.ascontiguousarray(r[:(n_sets + 63) // 64], dtype="<u8").tobytes()
    open(prefix + ".sbwt", "wb").write(body + struct.pack("<Q", _fnv1a(body)))
    body = b"KBOB200\0" + struct.pack("<IIQ", 1, k, n_sets) + np.asarray(lcs, dtype=np.uint8).tobytes()
    open(prefix + ".lcs", "wb").write(body + struct.pack("<Q", _fnv1a(body)))


def test_load_sbwt_validates_the_files_before_any_device_work(lib, tmp_path):
    """index::load_sbwt (index.rs:195-212): a missing file is the reference's panic (KBO_ERR_IO), anything that is not
    an index file of this library is refused (KBO_ERR_FORMAT) -- checked without a GPU; a well-formed pair gets as far
    as the upload (KBO_ERR_CUDA here)."""
    import oracle_lib as O
    o = O.OracleIndex([b"AAAGAACCA-TCAGGGCG"], k=3)
    prefix = str(tmp_path / "idx")
    with pytest.raises(api.KboPanic) as e:
        api.load_sbwt(prefix)
    assert e.value.status == 14 and "Expected SBWT at" in str(e.value)
    write_index_files(prefix, 3, o.n_sets, o.n_kmers, o.rows(), o.lcs())
    good_sbwt, good_lcs = open(prefix + ".sbwt", "rb").read(), open(prefix + ".lcs", "rb").read()

    def expect_format(sbwt=None, lcs=None, word=None):
        open(prefix + ".sbwt", "wb").write(good_sbwt if sbwt is None else sbwt)
        open(prefix + ".lcs", "wb").write(good_lcs if lcs is None else lcs)
        with pytest.raises(api.KboPanic) as e:
            api.load_sbwt(prefix)
        assert e.value.status == 15, str(e.value)
        if word:
            assert word in str(e.value)

    expect_format(sbwt=b"\x05\0\0\0\0\0\0\0Other" + good_sbwt[20:], word="SubsetMatrix")
    # the variant header of the reference followed by a body this library did not write (e.g. the sbwt crate's)
    expect_format(sbwt=good_sbwt[:20] + b"\x13\0\0\0\0\0\0\0" + good_sbwt[28:], word="not written by this library")
    expect_format(sbwt=good_sbwt[:-9], word="truncated")
    flipped = bytearray(good_sbwt); flipped[60] ^= 1
    expect_format(sbwt=bytes(flipped), word="checksum")
    expect_format(lcs=good_lcs[:-3])
    flipped = bytearray(good_lcs); flipped[26] ^= 1
    expect_format(lcs=bytes(flipped), word="checksum")
    o2 = O.OracleIndex([b"AAAGAACCA-TCAGGGCGTTTT"], k=3)
    write_index_files(prefix + "2", 3, o2.n_sets, o2.n_kmers, o2.rows(), o2.lcs())
    expect_format(lcs=open(prefix + "2.lcs", "rb").read(), word="does not belong")
    os.remove(prefix + ".lcs")
    open(prefix + ".sbwt", "wb").write(good_sbwt)
    with pytest.raises(api.KboPanic) as e:
        api.load_sbwt(prefix)
    assert e.value.status == 14 and "Expected LCS array at" in str(e.value)
    open(prefix + ".lcs", "wb").write(good_lcs)
    if api.device_count() == 0:
        with pytest.raises(api.KboPanic) as e:
            api.load_sbwt(prefix)
        assert e.value.status == 8  # the files are fine; there is no device to upload to and no CPU fallback
    assert lib.kbo_index_serialize(None, b"x") == 7 and lib.kbo_index_load(None, 0, None) == 7


def test_integration_doc_lists_every_entry_point():
    """INTEGRATION.md shows the `-sys` extern block a maintainer of the reference would add: it has to name every
    function the header declares."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in header_functions() if ("fn " + n) not in doc]
    assert not missing, missing


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header has to compile as C99 (what cgo / bindgen / ctypes-style tools consume)."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = tmp_path / "hdr.c"
    src.write_text('#include "kbo_b200.h"\nint main(void) { return KBO_ERR_FORMAT == 15 ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-c", str(src), "-o", str(tmp_path / "hdr.o")])
