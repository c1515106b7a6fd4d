"""Multi-GPU path behind the C API: the NCCL communicator bound inside
libtrvb.so (one process per GPU) and the single-process multi-device mode (one host
thread per GPU, TRV_GPU_MAXNUM honoured).  Tests that need two GPUs skip on a one-GPU
box; the NCCL binding itself is exercised with a one-rank communicator everywhere."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def core():
    from triumvirate_b200 import core
    if core.gpu_count() < 1:
        pytest.fail("no CUDA device visible: GPU tests cannot run (no CPU fallback)")
    return core


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _case():
    gen = np.random.default_rng(314)
    L, ng = 700., 64
    pos = gen.uniform(0., L, size=(3, 20000))
    kw = dict(boxsize=L, ngrid=ng, assignment="pcs", degrees=(0, 0, 0), form="full",
              bin_range=(0.02, 0.2), num_bins=6, norm_factor=1., deterministic=True)
    return pos, kw


def test_nccl_binding_one_rank_communicator(core):
    """dlopen(libnccl.so.2), ncclGetUniqueId, ncclCommInitRank, ncclAllReduce through
    trv_comm_* / trvb_allreduce with a single rank: the sum of one vector is itself."""
    ident = core.comm_unique_id()
    assert len(ident) == 128
    core.comm_init(1, 0, ident)
    try:
        assert core.comm_size() == 1
        buf = np.arange(1000, dtype=np.float64) * 0.5 - 3.
        out = core.allreduce(buf.copy())
        assert np.array_equal(out, buf)
    finally:
        core.comm_finalize()
    assert core.comm_size() == 1


def _nccl_worker(rank, world, port, q):
    os.environ["TRV_GPU_DEVICE"] = str(rank)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    from triumvirate_b200 import core, dist as tdist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert tdist.init_comm() == world
        pos, kw = _case()
        out = tdist.threept("bispec", "sim", pos_d=pos, **kw)
        q.put((rank, out["bk_raw"], out["bk_shot"], out["nmodes_1"]))
        core.comm_finalize()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_processes_nccl_allreduce_inside_the_call(core):
    """One process per GPU, NCCL communicator attached through the C API: every rank
    returns the complete data vector, bit-identical to the single-GPU result."""
    if core.gpu_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    pos, kw = _case()
    full = core.threept("bispec", "sim", pos_d=pos, **kw)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, raw, shot, nm in res:
        assert raw.tobytes() == full["bk_raw"].tobytes(), rank
        assert shot.tobytes() == full["bk_shot"].tobytes(), rank
        assert np.array_equal(nm, full["nmodes_1"])


def _dist_case():
    gen = np.random.default_rng(2718)
    L, ng = 1000., 128
    pos = gen.uniform(0., L, size=(3, 300000))
    pos[0, :20000] = gen.uniform(0., 3. * L / ng, size=20000)      # crowd the slab boundaries
    pos[0, 20000:40000] = gen.uniform(L / 2 - 8., L / 2 + 8., size=20000)
    kw = dict(boxsize=L, ngrid=ng, assignment="pcs", degrees=(0, 0, 0), form="full",
              bin_range=(0.01, 0.1), num_bins=9, norm_factor=1.)
    return pos, kw


def _dist_worker(rank, world, port, q):
    os.environ["TRV_GPU_DEVICE"] = str(rank)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    from triumvirate_b200 import core, dist as tdist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert tdist.init_comm() == world
        pos, kw = _dist_case()
        res = {}
        for assignment in ("pcs", "tsc", "cic"):
            for flag in ("1", "0"):
                os.environ["TRV_NO_DIST_MESH"] = flag
                before = core.dmesh_call_count()
                out = tdist.threept("bispec", "sim", pos_d=pos, **dict(kw, assignment=assignment))
                assert core.dmesh_call_count() - before == (1 if flag == "0" else 0)
                res[(assignment, flag)] = (out["bk_raw"], out["bk_shot"])
        q.put((rank, res))
        core.comm_finalize()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_processes_distributed_mesh_phase(core):
    """NCCL communicator attached, grid extents that split over the ranks: assignment and the
    full-grid transforms run on x-slabs / k_y-slabs (trvb_dmesh_*: windowed assignment, 2-D
    transforms, all-to-all, 1-D transform; low-|k| modes gathered; xi(r) by the inverse
    route with the radial histogram summed over the ranks).  Every rank returns the complete
    data vector, equal to the one-GPU result and to the replicated-mesh run to round-off."""
    if core.gpu_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    pos, kw = _dist_case()
    full = {a: core.threept("bispec", "sim", pos_d=pos, **dict(kw, assignment=a))
            for a in ("pcs", "tsc", "cic")}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out in res:
        for (assignment, flag), (raw, shot) in out.items():
            ref = full[assignment]
            for got, want in ((raw, ref["bk_raw"]), (shot, ref["bk_shot"])):
                err = np.max(np.abs(got - want)) / np.max(np.abs(want))
                assert err < 1.e-11, (rank, assignment, flag, err)


def test_single_process_spreads_over_the_visible_gpus(core, monkeypatch):
    """No environment pin, several GPUs: trv::compute_* deals the entries to one host
    thread per GPU and sums the shares on the host -- same bits as one GPU."""
    if core.gpu_count() < 2:
        pytest.skip("needs two GPUs")
    for name in ("TRV_GPU_DEVICE", "LOCAL_RANK"):
        monkeypatch.delenv(name, raising=False)
    pos, kw = _case()
    monkeypatch.setenv("TRV_GPU_MULTI", "0")
    assert core.multi_device_count(kw["ngrid"]) == 1
    one = core.threept("bispec", "sim", pos_d=pos, **kw)
    monkeypatch.setenv("TRV_GPU_MULTI", "1")
    assert core.multi_device_count(kw["ngrid"]) == core.gpu_count()
    many = core.threept("bispec", "sim", pos_d=pos, **kw)
    monkeypatch.setenv("TRV_GPU_MAXNUM", "2")
    assert core.multi_device_count(kw["ngrid"]) == 2
    two = core.threept("bispec", "sim", pos_d=pos, **kw)
    for out in (many, two):
        assert out["bk_raw"].tobytes() == one["bk_raw"].tobytes()
        assert out["bk_shot"].tobytes() == one["bk_shot"].tobytes()
    # 3PCF and a survey-type call through the same runner
    kw3 = dict(kw, bin_range=(30., 200.), num_bins=5)
    monkeypatch.setenv("TRV_GPU_MULTI", "0")
    a = core.threept("3pcf", "sim", pos_d=pos, **kw3)
    monkeypatch.setenv("TRV_GPU_MULTI", "1")
    b = core.threept("3pcf", "sim", pos_d=pos, **kw3)
    assert b["zeta_raw"].tobytes() == a["zeta_raw"].tobytes()
    assert b["zeta_shot"].tobytes() == a["zeta_shot"].tobytes()


def test_one_gpu_box_stays_on_one_device(core, monkeypatch):
    for name in ("TRV_GPU_DEVICE", "LOCAL_RANK", "TRV_GPU_MULTI"):
        monkeypatch.delenv(name, raising=False)
    assert core.multi_device_count(64) == 1            # small meshes never spread
    monkeypatch.setenv("LOCAL_RANK", "0")
    assert core.multi_device_count(512) == 1           # one process per GPU: pinned
