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
        mine = tdist.local_entries(dim, rank, world)
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


def test_partition_covers_every_entry_once():
    from triumvirate_b200 import dist as tdist
    for dim in (1, 4, 10, 210, 820):
        for world in (1, 2, 3, 8):
            seen = np.concatenate([tdist.local_entries(dim, r, world) for r in range(world)])
            assert sorted(seen) == list(range(dim))
            sizes = [len(tdist.local_entries(dim, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1          # balanced
            for r in range(world):
                assert all(tdist.owner_of(i, world) == r for i in tdist.local_entries(dim, r, world))


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
