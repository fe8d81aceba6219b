"""Deterministic synthetic sequences for tests and bench.py (SURVEY.md section 8d workloads).

Bases are drawn with numpy's PCG64 (stable stream for a given seed); seeds follow SURVEY 8d.
"""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)

SEED_C1_REF = 0x6B626F01
SEED_C1_ASM = 0x6B626F02
SEED_C2_REF = 0x6B626F03
SEED_C2_GENES = 0x6B626F04


def _rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def random_codes(n, seed):
    return _rng(seed).integers(0, 4, size=n, dtype=np.uint8)


def random_seq(n, seed):
    """n iid uniform bases as an ASCII uint8 array."""
    return ACGT[random_codes(n, seed)]


def mutate(seq, seed, snp=0.01, indel=0.0005, max_indel=5):
    """SNPs with prob `snp` per base; with prob `indel` delete 1..max_indel bases; with prob `indel`
    insert 1..max_indel iid bases after the base.  `seq` is an ASCII uint8 array of ACGT."""
    rng = _rng(seed)
    n = len(seq)
    lut = np.full(256, 255, dtype=np.uint8)
    lut[ACGT] = np.arange(4, dtype=np.uint8)
    codes = lut[seq]
    u = rng.random(n)
    snp_mask = u < snp
    codes = codes.copy()
    codes[snp_mask] = (codes[snp_mask] + rng.integers(1, 4, size=int(snp_mask.sum()), dtype=np.uint8)) & 3
    keep = np.ones(n, dtype=bool)
    if indel > 0:
        v = rng.random(n)
        del_starts = np.nonzero(v < indel)[0]
        del_lens = rng.integers(1, max_indel + 1, size=len(del_starts))
        for s, ln in zip(del_starts.tolist(), del_lens.tolist()):
            keep[s:s + ln] = False
        w = rng.random(n)
        ins_after = np.nonzero(w < indel)[0]
        ins_lens = rng.integers(1, max_indel + 1, size=len(ins_after))
        counts = np.ones(n, dtype=np.int64)
        counts[~keep] = 0
        extra = np.zeros(n, dtype=np.int64)
        extra[ins_after] = ins_lens
        total = counts + extra
        out = np.repeat(codes, total)
        # positions produced by an insertion: every repeat after the first kept copy (or all, if deleted)
        ends = np.cumsum(total)
        starts = ends - total
        ins_pos = []
        for i, ln in zip(ins_after.tolist(), ins_lens.tolist()):
            first = starts[i] + counts[i]
            ins_pos.append(np.arange(first, first + ln))
        if ins_pos:
            ins_pos = np.concatenate(ins_pos)
            out[ins_pos] = rng.integers(0, 4, size=len(ins_pos), dtype=np.uint8)
        codes = out
    return ACGT[codes]


def gene_queries(ref, n_queries, length, seed, snp=0.01):
    """n_queries substrings of `ref` of `length` bases at uniform positions, each with `snp` SNPs.
    Returns (concat uint8 array, offsets uint64 array)."""
    rng = _rng(seed)
    starts = rng.integers(0, len(ref) - length + 1, size=n_queries)
    idx = (starts[:, None] + np.arange(length)[None, :]).reshape(-1)
    concat = ref[idx].copy()
    lut = np.full(256, 255, dtype=np.uint8)
    lut[ACGT] = np.arange(4, dtype=np.uint8)
    m = rng.random(len(concat)) < snp
    c = lut[concat[m]]
    concat[m] = ACGT[(c + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) & 3]
    offsets = (np.arange(n_queries + 1, dtype=np.uint64) * np.uint64(length)).astype(np.uint64)
    return concat, offsets
