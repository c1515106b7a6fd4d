"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: the entry
partition and the single all-reduce that completes the data vector."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from triumvirate_b200 import dist as tdist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dim = 210
        gen = np.random.default_rng(99)                     # same "full" result on every rank
        full_raw = gen.normal(size=dim) + 1j * gen.normal(size=dim)
        full_shot = gen.normal(size=dim) + 1j * gen.normal(size=dim)
        mine = tdist.local_entries("full", (0, 0), 20, rank, world)
        part = {"bk_raw": np.zeros(dim, complex), "bk_shot": np.zeros(dim, complex),
                "k1_eff": np.arange(dim, dtype=float)}
        part["bk_raw"][mine] = full_raw[mine]
        part["bk_shot"][mine] = full_shot[mine]
        out = tdist.allreduce_result(part, "bispec")
        ok = (np.array_equal(out["bk_raw"], full_raw) and np.array_equal(out["bk_shot"], full_shot)
              and np.array_equal(out["k1_eff"], part["k1_eff"]))
        q.put((rank, bool(ok), len(mine)))
    finally:
        dist.destroy_process_group()


def _pairs_of(form, degrees, nb, idx_bin=0):
    """(row, col) of every data-vector entry (make_data_vector, src/threept.cpp)."""
    if form == "full" and degrees[0] == degrees[1]:
        return [(a, b) for a in range(nb) for b in range(a, nb)]
    if form == "full":
        return [(a, b) for a in range(nb) for b in range(nb)]
    if form == "diag":
        return [(b, b) for b in range(nb)]
    if form == "row":
        return [(idx_bin, b) for b in range(nb)]
    off = abs(idx_bin)
    return [(i, i + off) if idx_bin >= 0 else (i + off, i) for i in range(nb - off)]


def test_partition_covers_every_entry_once():
    from triumvirate_b200 import dist as tdist
    cases = [("full", (0, 0), 1), ("full", (0, 0), 4), ("full", (0, 0), 20), ("full", (0, 0), 40),
             ("full", (2, 0), 20), ("diag", (2, 0), 20), ("row", (1, 1), 20), ("off-diag", (0, 0), 10)]
    for form, deg, nb in cases:
        idx_bin = 3 if form in ("row", "off-diag") else 0
        pairs = _pairs_of(form, deg, nb, idx_bin)
        dim = len(pairs)
        for world in (1, 2, 3, 8):
            own = tdist.owners(form, deg, nb, world, idx_bin=idx_bin)
            assert len(own) == dim and own.min() >= 0 and own.max() < world
            seen = np.concatenate([tdist.local_entries(form, deg, nb, r, world, idx_bin=idx_bin)
                                   for r in range(world)])
            assert sorted(seen) == list(range(dim))
            if form != "full":
                # one (or two) fields per entry: the largest share is as small as it can be
                sizes = np.bincount(own, minlength=world)
                assert sizes.max() == -(-dim // world)


def test_partition_is_compact_in_the_pair_matrix():
    """A rank needs the shell field of every bin that appears in its entries, and those
    transforms are what the pair phase costs: the shares are compact blocks, equal in
    FIELDS (the round-robin split of the first version needed all 40 on every one of 8
    ranks; equal-pair cuts of the serpentine order still left 15 ... 26)."""
    from triumvirate_b200 import dist as tdist
    nb = 40
    pairs = np.array(_pairs_of("full", (0, 0), nb))
    for world, cap in ((8, 19), (7, 20), (4, 25), (2, 40)):
        own = tdist.owners("full", (0, 0), nb, world)
        assert own.max() == world - 1
        fields = [len(set(pairs[own == r].ravel())) for r in range(world)]
        assert max(fields) <= cap, (world, fields)
    own2 = tdist.owners("full", (2, 0), 20, 4)       # 20 x 20 entries, distinct row/col fields
    pairs2 = np.array(_pairs_of("full", (2, 0), 20))
    for r in range(4):
        mine = pairs2[own2 == r]
        assert len(set(mine[:, 0])) + len(set(mine[:, 1])) <= 20


def test_pack_unpack_roundtrip():
    from triumvirate_b200 import dist as tdist
    gen = np.random.default_rng(1)
    out = {"zeta_raw": gen.normal(size=7) + 1j * gen.normal(size=7),
           "zeta_shot": gen.normal(size=7) + 1j * gen.normal(size=7), "r1_eff": np.ones(7)}
    back = tdist.unpack(tdist.pack(out, "3pcf"), out, "3pcf")
    assert np.array_equal(back["zeta_raw"], out["zeta_raw"])
    assert np.array_equal(back["zeta_shot"], out["zeta_shot"])


@pytest.mark.timeout(180)
def test_allreduce_of_partials_is_bit_identical_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res), res
    assert sum(r[2] for r in res) == 210
