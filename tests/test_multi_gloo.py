"""N > 1 host logic on CPU: world_size-2 (and 3) gloo runs of kbo_b200.multi with the oracle standing in
for the per-rank GPU compute.  The result gathered on rank 0 must equal the single-process result."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmpfile):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import oracle_lib as O
    from kbo_b200 import multi, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref = synth.random_seq(30_000, 5)
    oix = O.OracleIndex([ref.tobytes()], k=31)  # "replicated index"
    concat, offsets = synth.gene_queries(ref, 37, 400, 6)
    # ragged batch: drop the tail of some queries
    lens = np.diff(offsets).astype(np.int64)
    lens[::5] = 123
    pieces = [concat[int(offsets[i]):int(offsets[i]) + int(lens[i])] for i in range(len(lens))]
    concat = np.concatenate(pieces)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)

    def compute(c, o):
        _, out, _ = oix.matches_batch(c, o, n_threads=1)
        return out

    got = multi.matches_sharded(concat, offsets, compute, rank, world)
    if rank == 0:
        want = compute(concat, offsets)
        assert np.array_equal(got, want)
        open(tmpfile, "w").write("ok")
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_matches_equals_single_process(world, tmp_path):
    out = tmp_path / "ok.txt"
    mp.spawn(_worker, args=(world, 29611 + world, str(out)), nprocs=world, join=True)
    assert out.read_text() == "ok"


def test_partition_by_bases():
    sys.path.insert(0, ROOT)
    from kbo_b200 import multi
    off = np.array([0, 10, 20, 30, 40, 1000, 1010], dtype=np.uint64)
    for world in (1, 2, 3, 4, 8, 16):
        parts = multi.partition_by_bases(off, world)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == 6
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        assert all(a <= b for a, b in parts)
    eq = np.arange(0, 1001, 100, dtype=np.uint64)
    assert multi.partition_by_bases(eq, 2) == [(0, 5), (5, 10)]
    assert multi.partition_by_bases(eq, 5) == [(0, 2), (2, 4), (4, 6), (6, 8), (8, 10)]
